"""Device-side timing and clock sampling helpers (the reference only has host wall-clock `system.record_time`)."""
import shutil
import statistics
import subprocess
import threading
import time

import torch


class DeviceTimer:
    """CUDA-event stopwatch on the current stream."""

    def __init__(self):
        self._start = torch.cuda.Event(enable_timing=True)
        self._stop = torch.cuda.Event(enable_timing=True)

    def start(self):
        self._start.record()

    def stop(self):
        self._stop.record()

    def elapsed_ms(self) -> float:
        self._stop.synchronize()
        return self._start.elapsed_time(self._stop)


class ClockSampler:
    """Samples SM clocks / throttle reasons with nvidia-smi while a timed region runs."""

    FIELDS = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
              'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
              'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index: int = 0, period_ms: int = 20):
        self.gpu_index, self.period_ms = gpu_index, period_ms
        self.samples = []
        self._proc = None
        self._thread = None

    def start(self):
        if shutil.which('nvidia-smi') is None:
            return self
        try:
            self._proc = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu_index), '--query-gpu=' + self.FIELDS,
                                           '--format=csv,noheader,nounits', '-lms', str(self.period_ms)],
                                          stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:  # noqa
            self._proc = None
            return self

        def pump():
            for line in self._proc.stdout:
                parts = [p.strip() for p in line.split(',')]
                if len(parts) >= 7:
                    self.samples.append((time.time(), parts))
        self._thread = threading.Thread(target=pump, daemon=True)
        self._thread.start()
        return self

    def stop(self):
        if self._proc is not None:
            self._proc.terminate()
            try:
                self._proc.wait(timeout=2)
            except Exception:  # noqa
                self._proc.kill()
        return self

    def summary(self, t0: float = 0.0, t1: float = float('inf')) -> dict:
        rows = [p for (t, p) in self.samples if t0 <= t <= t1] or [p for (_, p) in self.samples]
        if not rows:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        def num(v):
            try:
                return float(v)
            except Exception:  # noqa
                return None
        sm = [num(r[0]) for r in rows if num(r[0]) is not None]
        reasons = []
        for i, name in enumerate(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap')):
            if any(r[3 + i].lower().startswith('active') for r in rows):
                reasons.append(name)
        power = [num(r[2]) for r in rows if num(r[2]) is not None]
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': num(rows[0][1]), 'reasons': reasons,
                'power_w_max': max(power) if power else None, 'samples': len(rows)}

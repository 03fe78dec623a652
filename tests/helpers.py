"""Shared test utilities: subprocess launchers for the helloworld driver and multi-process (gloo) workers."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_PORT = [29611]


def next_port():
    _PORT[0] += 1
    return _PORT[0] + (os.getpid() % 200)


def run_helloworld(nproc=1, device='cpu', extra=(), timeout=900, env=None):
    """Run tutel_b200.examples.helloworld and return the list of printed losses (floats)."""
    e = dict(os.environ)
    e['PYTHONPATH'] = ROOT + os.pathsep + e.get('PYTHONPATH', '')
    e.setdefault('OMP_NUM_THREADS', '4')
    if env:
        e.update(env)
    if nproc == 1:
        cmd = [sys.executable, '-m', 'tutel_b200.examples.helloworld']
    else:
        cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=%d' % nproc,
               '--master-addr', '127.0.0.1', '--master-port', str(next_port()), '-m', 'tutel_b200.examples.helloworld']
    cmd += ['--device', device] + [str(x) for x in extra]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=e, cwd=ROOT)
    losses = [float(m.group(1)) for m in re.finditer(r'loss = ([-0-9.einfa]+),', p.stdout)]
    assert losses, 'no losses printed:\n%s\n%s' % (p.stdout[-2000:], p.stderr[-3000:])
    return losses


def run_workers(script_body: str, nproc: int, timeout=600, env=None, backend='gloo'):
    """Run `script_body` (python source, sees RANK/WORLD_SIZE and a ready `dist` import) on nproc processes."""
    import tempfile
    e = dict(os.environ)
    e['PYTHONPATH'] = ROOT + os.pathsep + e.get('PYTHONPATH', '')
    e.setdefault('OMP_NUM_THREADS', '2')
    e['TUTEL_TEST_BACKEND'] = backend
    if env:
        e.update(env)
    with tempfile.NamedTemporaryFile('w', suffix='.py', delete=False) as f:
        f.write('import os, sys, torch\nimport torch.distributed as dist\n' + script_body)
        path = f.name
    try:
        cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=%d' % nproc,
               '--master-addr', '127.0.0.1', '--master-port', str(next_port()), path]
        p = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=e, cwd=ROOT)
        assert p.returncode == 0, 'workers failed:\n%s\n%s' % (p.stdout[-3000:], p.stderr[-5000:])
        return p.stdout
    finally:
        os.unlink(path)

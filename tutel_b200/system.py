"""Session helpers: process bootstrap, host-side timers, a process-global record cache, tensor files, NUMA pinning.

API parity with tutel/system.py:10-104 (``init_data_model_parallel``, ``get_local_session``, ``record_time``,
``save`` / ``load``, ``apply_rank_size_from_pattern``, ``cache()`` / ``LocalCache``,
``init_affinity_at_program_beginning``); the implementation is organised around one session record per process.
"""
import atexit
import logging
import os
import sys
import time

TUTEL_CUDA_SANDBOX = int(os.environ.get('TUTEL_CUDA_SANDBOX', 0))
_NODE_ROOT = '/sys/devices/system/node'


class _Session:
    """What this process knows about its distributed session."""
    env = None
    teardown_registered = False


# ---------------------------------------------------------------------------------------------------------------------
# NUMA affinity
# ---------------------------------------------------------------------------------------------------------------------
def _numbered(entries, prefix):
    return sorted(int(e[len(prefix):]) for e in entries if e.startswith(prefix) and e[len(prefix):].isdigit())


def _numa_cpu_sets():
    """CPU ids of every NUMA node, in node order (empty when the sysfs tree is not there)."""
    if not os.path.isdir(_NODE_ROOT):
        return []
    return [_numbered(os.listdir(os.path.join(_NODE_ROOT, 'node%d' % n)), 'cpu') for n in _numbered(os.listdir(_NODE_ROOT), 'node')]


def init_affinity_at_program_beginning():
    """Bind the process to the NUMA node that belongs to its local rank; ``NUMA_TYPE`` consecutive local ranks share one
    node, ``NUMA_TYPE<=0`` or the one-GPU sandbox (``TUTEL_CUDA_SANDBOX``) disable the binding."""
    if TUTEL_CUDA_SANDBOX:
        return
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    ranks_per_node = int(os.environ.get('NUMA_TYPE', '1'))
    if ranks_per_node <= 0:
        return
    try:
        cpu_sets = [c for c in _numa_cpu_sets() if c]
        if not cpu_sets:
            raise RuntimeError('no NUMA topology under %s' % _NODE_ROOT)
        chosen = (local_rank // ranks_per_node) % len(cpu_sets)
        os.sched_setaffinity(0, cpu_sets[chosen])
        logging.info('LOCAL_RANK %d is bound to NUMA node %d (of %d)', local_rank, chosen, len(cpu_sets))
    except Exception as ex:  # noqa
        if local_rank == 0:
            logging.warning('Failed to set NUMA status: %s', ex)


# ---------------------------------------------------------------------------------------------------------------------
# session
# ---------------------------------------------------------------------------------------------------------------------
def _teardown():
    for stream in (sys.stdout, sys.stderr):
        try:
            stream.flush()
        except Exception:  # noqa
            pass
    try:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()
    except Exception:  # noqa
        pass


def init_data_model_parallel(group_count=1, backend='nccl'):
    """Initialise torch.distributed (torchrun / OpenMPI environment, or a single process) and build the
    data-parallel x model-parallel group mesh; the returned record also carries ``is_cuda``."""
    from . import net
    env = net.create_groups_from_world(group_count=group_count, include_init=backend)
    env.is_cuda = env.local_device.type == 'cuda'
    logging.critical('Registering device global rank %s: data_rank = %s, model_rank = %s',
                     env.global_rank, env.data_rank, env.model_rank)
    _Session.env = env
    init_data_model_parallel.default_env = env       # attribute kept for code that reads it like the reference does
    if not _Session.teardown_registered:
        _Session.teardown_registered = True
        atexit.register(_teardown)
    return env


def get_local_session():
    if _Session.env is None:
        raise Exception('Current session is not initialized with: system.init_data_model_parallel(). '
                        'Please try with: system.record_time(is_cuda=False)')
    return _Session.env


def record_time(is_cuda=None):
    """Host wall clock, after a device synchronize when the session runs on CUDA (reference semantics; device-side
    timing lives in :mod:`tutel_b200.utils.timers`)."""
    on_gpu = get_local_session().is_cuda if is_cuda is None else bool(is_cuda)
    if on_gpu:
        import torch
        torch.cuda.synchronize()
    return time.time()


# ---------------------------------------------------------------------------------------------------------------------
# process-global record cache (the fairseq integration parks per-layer (num_tokens, l_aux) pairs here)
# ---------------------------------------------------------------------------------------------------------------------
class LocalCache:
    _records = {}

    @classmethod
    def reset(cls):
        cls._records = {}

    @classmethod
    def set(cls, key, val):
        cls._records[key] = val

    @classmethod
    def get(cls, key=None):
        """The record stored under ``key``; every record (in insertion order) when the key is absent / omitted."""
        if key in cls._records:
            return cls._records[key]
        return list(cls._records.values())


def cache():
    return LocalCache


# ---------------------------------------------------------------------------------------------------------------------
# files
# ---------------------------------------------------------------------------------------------------------------------
def save(t, path):
    """Write a tensor as a ``.npy`` file."""
    import numpy
    numpy.save(path, t.detach().cpu().numpy())


def load(path, device=None):
    import numpy
    import torch
    return torch.as_tensor(numpy.load(path)).to(device) if device is not None else torch.as_tensor(numpy.load(path))


def apply_rank_size_from_pattern(filename, rank, size, create_dir=True):
    """``'ckpt/{rank}-of-{size}.pt' -> 'ckpt/3-of-8.pt'`` (and make sure the directory exists)."""
    if '{rank}' not in filename:
        logging.warning('Keyword `{rank}` is not found in file pattern: %s, which may cause collision in file access.', filename)
    resolved = filename.replace('{size}', str(size)).replace('{rank}', str(rank))
    parent = os.path.dirname(resolved)
    if create_dir and parent:
        os.makedirs(parent, exist_ok=True)
    return resolved

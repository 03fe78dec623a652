"""Session helpers (mirrors tutel/system.py:10-104): process bootstrap, timers, caches, tensor save/load."""
import atexit
import logging
import os
import re
import sys
import time

TUTEL_CUDA_SANDBOX = int(os.environ.get('TUTEL_CUDA_SANDBOX', 0))


def init_affinity_at_program_beginning():
    """Pin the process to the NUMA node of its local rank (``NUMA_TYPE`` ranks share a node)."""
    if TUTEL_CUDA_SANDBOX:
        return
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    try:
        numa_type = int(os.environ.get('NUMA_TYPE', '1'))
        if numa_type <= 0:
            return
        root = '/sys/devices/system/node'
        nodes = sorted(int(n[4:]) for n in os.listdir(root) if re.fullmatch(r'node[0-9]+', n))
        cpus = [sorted(int(c[3:]) for c in os.listdir('%s/node%d' % (root, n)) if re.fullmatch(r'cpu[0-9]+', c)) for n in nodes]
        sel = (local_rank // numa_type) % len(nodes)
        os.sched_setaffinity(0, cpus[sel])
        logging.info('LOCAL_RANK %d is bound to NUMA node %d (of %d)' % (local_rank, sel, len(nodes)))
    except Exception as ex:  # noqa
        if local_rank == 0:
            logging.warning('Failed to set NUMA status: %s' % ex)


def init_data_model_parallel(group_count=1, backend='nccl'):
    from . import net
    result = net.create_groups_from_world(group_count=group_count, include_init=backend)
    result.is_cuda = (result.local_device.type == 'cuda')
    logging.critical('Registering device global rank %s: data_rank = %s, model_rank = %s' % (result.global_rank, result.data_rank, result.model_rank))
    init_data_model_parallel.default_env = result

    def on_quit():
        sys.stdout.flush()
        sys.stderr.flush()
        try:
            import torch.distributed as dist
            if dist.is_initialized():
                dist.destroy_process_group()
        except Exception:  # noqa
            pass

    if not getattr(init_data_model_parallel, '_atexit', False):
        atexit.register(on_quit)
        init_data_model_parallel._atexit = True
    return result


class LocalCache:
    _CACHE = dict()

    @staticmethod
    def reset():
        LocalCache._CACHE = dict()

    @staticmethod
    def set(key, val):
        LocalCache._CACHE[key] = val

    @staticmethod
    def get(key=None):
        if key not in LocalCache._CACHE:
            return list(LocalCache._CACHE.values())
        return LocalCache._CACHE[key]


def cache():
    return LocalCache


def get_local_session():
    if not hasattr(init_data_model_parallel, 'default_env'):
        raise Exception('Current session is not initialized with: system.init_data_model_parallel(). Please try with: system.record_time(is_cuda=False)')
    return init_data_model_parallel.default_env


def record_time(is_cuda=None):
    """Host wall clock after a device synchronize (reference semantics; use utils.timers for device timing)."""
    is_cuda = is_cuda if is_cuda is not None else get_local_session().is_cuda
    if is_cuda:
        import torch
        torch.cuda.synchronize()
    return time.time()


def save(t, path):
    import numpy as np
    np.save(path, t.detach().cpu().numpy())


def load(path, device=None):
    import numpy as np
    import torch
    return torch.tensor(np.load(path), device=device)


def apply_rank_size_from_pattern(filename, rank, size, create_dir=True):
    if not re.search(r'\{rank\}', filename):
        logging.warning('Keyword `{rank}` is not found in file pattern: %s, which may cause collision in file access.' % filename)
    filename = re.sub(r'\{rank\}', str(rank), re.sub(r'\{size\}', str(size), filename))
    if create_dir:
        filedir = os.path.dirname(filename)
        if filedir:
            os.makedirs(filedir, exist_ok=True)
    return filename

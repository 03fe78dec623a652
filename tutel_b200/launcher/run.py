"""mpiexec -> torchrun bridge:  mpiexec -host a,b python -m tutel_b200.launcher.run [-m] prog args...

Each MPI rank is one *node*; it replaces itself with ``torch.distributed.run`` spawning ``LOCAL_SIZE`` workers through
:mod:`tutel_b200.launcher.execl` (reference: tutel/launcher/run.py:6-35).
"""
import os
import sys


def main():
    argv = sys.argv[1:]
    if not argv:
        raise SystemExit('usage: python -m tutel_b200.launcher.run [-m] <program> [args...]')
    env = os.environ
    try:
        import torch
        default_local = max(torch.cuda.device_count(), 1)
    except Exception:  # noqa
        default_local = 1
    local_size = int(env.get('LOCAL_SIZE', default_local))
    nnodes = int(env.get('OMPI_COMM_WORLD_SIZE', 1))
    node_rank = int(env.get('OMPI_COMM_WORLD_RANK', 0))
    env['LOCAL_SIZE'] = str(local_size)
    env.setdefault('MASTER_ADDR', '127.0.0.1')
    env.setdefault('MASTER_PORT', '23232')
    env.setdefault('OMP_NUM_THREADS', str(max(1, (os.cpu_count() or 1) // local_size)))
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nproc_per_node=%d' % local_size, '--nnodes=%d' % nnodes,
           '--node_rank=%d' % node_rank, '--master_addr=%s' % env['MASTER_ADDR'], '--master_port=%s' % env['MASTER_PORT'],
           '-m', 'tutel_b200.launcher.execl'] + argv
    os.execvpe(cmd[0], cmd, env)


if __name__ == '__main__':
    main()

"""Per-worker shim: NUMA binding + optional one-GPU sandbox, then exec the user program
(reference: tutel/launcher/execl.py:8-42).

    TUTEL_CUDA_SANDBOX=2  each worker only sees its own GPU (CUDA_VISIBLE_DEVICES=<local rank>)
    NUMA_TYPE / numactl   when `numactl` exists the worker is bound to the NUMA node of its local rank
"""
import os
import shutil
import sys


def main():
    argv = sys.argv[1:]
    if not argv:
        raise SystemExit('usage: python -m tutel_b200.launcher.execl [-m] <program> [args...]')
    env = os.environ
    local_rank = int(env.get('LOCAL_RANK', 0))
    local_size = int(env.get('LOCAL_SIZE', 1))
    if int(env.get('TUTEL_CUDA_SANDBOX', 0)) == 2:
        env['CUDA_VISIBLE_DEVICES'] = str(local_rank)
    # `-m module args..` runs a module with this interpreter; anything else is a complete command line that is exec'd
    # as given (`python3 train.py ..`, a shell script, ..) - the reference's contract (tutel/launcher/execl.py:36-41)
    cmd = [sys.executable, '-m'] + argv[1:] if argv[0] == '-m' else list(argv)
    numactl = shutil.which('numactl')
    if numactl and int(env.get('NUMA_TYPE', '1')) > 0:
        try:
            nodes = len([n for n in os.listdir('/sys/devices/system/node') if n.startswith('node') and n[4:].isdigit()])
        except Exception:  # noqa
            nodes = 1
        if nodes > 1:
            node = local_rank * nodes // max(local_size, 1)
            cmd = [numactl, '--cpunodebind=%d' % node] + cmd
    os.execvpe(cmd[0], cmd, env)


if __name__ == '__main__':
    main()

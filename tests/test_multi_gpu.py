"""Multi-GPU tests (NVLink P2P collectives, fused dispatch/combine engine) - need >= 2 GPUs."""
import os
import subprocess
import sys

import pytest
import torch

from helpers import ROOT, next_port

pytestmark = pytest.mark.gpu


def _ngpu():
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _run(which, nproc, **extra_env):
    env = dict(os.environ, TUTEL_B200_SPIN_TIMEOUT_SEC=os.environ.get('TUTEL_B200_SPIN_TIMEOUT_SEC', '30'), **extra_env)
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=%d' % nproc, '--master-addr',
           '127.0.0.1', '--master-port', str(next_port()), os.path.join(ROOT, 'tests', 'workers', 'p2p_worker.py'), which]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT, env=env)
    assert p.returncode == 0 and 'WORKER_OK' in p.stdout, p.stdout[-3000:] + p.stderr[-3000:]


@pytest.mark.skipif(_ngpu() < 2, reason='needs 2 GPUs')
def test_p2p_collectives_match_nccl():
    _run('coll', 2)


@pytest.mark.skipif(_ngpu() < 2, reason='needs 2 GPUs')
def test_fused_engine_matches_nccl_path():
    _run('fused', 2)


@pytest.mark.skipif(_ngpu() < 2, reason='needs 2 GPUs')
def test_fused_engine_matches_fp32_torch_oracle():
    """Fused bf16 engine vs NCCL + torch.matmul in fp32 on identical weights (small shapes and the flagship shape)."""
    _run('oracle', 2)


@pytest.mark.skipif(_ngpu() < 2, reason='needs 2 GPUs')
def test_fp8_experts_in_the_fused_engine():
    _run('fp8', 2)


@pytest.mark.skipif(_ngpu() < 2, reason='needs 2 GPUs')
def test_deep_stack_shares_one_buffer_ring():
    _run('deep', 2)


@pytest.mark.skipif(_ngpu() < 2, reason='needs 2 GPUs')
def test_data_vs_model_parallel_and_overlap_equivalence_fp32():
    _run('equiv', 2)


@pytest.mark.skipif(_ngpu() < 4, reason='needs 4 GPUs')
def test_subgroup_transports():
    _run('sub', 4)


@pytest.mark.skipif(_ngpu() < 4, reason='needs 4 GPUs')
def test_hierarchical_all_to_all_native():
    _run('2dh', 4, LOCAL_SIZE='2')


@pytest.mark.skipif(_ngpu() < 2, reason='needs 2 GPUs')
def test_dead_peer_is_diagnosed_not_hung():
    """Fault injection: rank 1 skips one push collective; the other rank must report a peer-wait timeout and fail fast."""
    env = dict(os.environ, TUTEL_B200_FAULT='skip_push:rank=1:call=2', TUTEL_B200_SPIN_TIMEOUT_SEC='5')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2', '--master-addr', '127.0.0.1',
           '--master-port', str(next_port()), os.path.join(ROOT, 'tests', 'workers', 'p2p_worker.py'), 'fault']
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert 'FIRST_OK' in p.stdout and p.returncode != 0
    assert 'timeout' in p.stdout + p.stderr

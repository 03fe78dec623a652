"""Golden loss curves on the GPU path (mirrors tests/test_tutel.py:100-152 of the reference: fp32 to 3 decimals,
fp16 first two losses to 1 decimal) and CPU-vs-CUDA equivalence (test_cpu_kernel)."""
import pytest

from helpers import run_helloworld

pytestmark = pytest.mark.gpu


def _golden(golden, top, dtype, nle):
    for g in golden:
        if g['top'] == top and g['dtype'] == dtype and g['num_local_experts'] == nle:
            return g['losses']
    pytest.skip('no golden entry')


@pytest.mark.parametrize('top,nle', [(2, 2), (1, 2)])
def test_fp32_cuda_matches_reference_losses(golden, top, nle):
    want = _golden(golden, top, 'float32', nle)
    got = run_helloworld(device='cuda', extra=['--top', top, '--dtype', 'float32', '--num_local_experts', nle, '--batch_size', 16,
                                               '--num_tokens', 1024, '--num_steps', 6, '--parallel_type', 'data'])
    assert [round(v, 2) for v in got] == [round(v, 2) for v in want[:6]]


@pytest.mark.parametrize('top,nle', [(2, 2), (1, 1)])
def test_fp16_cuda_first_losses(golden, top, nle):
    want = _golden(golden, top, 'float16', nle)
    got = run_helloworld(device='cuda', extra=['--top', top, '--dtype', 'float16', '--num_local_experts', nle, '--batch_size', 16,
                                               '--num_tokens', 1024, '--num_steps', 2, '--parallel_type', 'data'])
    assert [round(v, 1) for v in got[:2]] == [round(v, 1) for v in want[:2]]


def test_cpu_and_cuda_agree():
    args = ['--num_steps', 5, '--num_tokens', 256, '--batch_size', 8, '--model_dim', 512, '--hidden_size', 512]
    a = run_helloworld(device='cuda', extra=args)
    b = run_helloworld(device='cpu', extra=args)
    assert [round(v, 2) for v in a] == [round(v, 2) for v in b]

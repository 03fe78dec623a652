"""Golden loss curves of the reference (tests/test_baseline.json) reproduced by the new framework on CPU.

Mirrors the reference's end-to-end strategy (tests/test_tutel.py:94-152): the helloworld driver is run as a
subprocess and its printed losses are compared with the recorded curve (3 decimals for fp32/fp64).
"""
import pytest

from helpers import run_helloworld


def _golden(golden, top, dtype, nle):
    for g in golden:
        if g['top'] == top and g['dtype'] == dtype and g['num_local_experts'] == nle:
            return g['losses']
    pytest.skip('no golden entry')


@pytest.mark.parametrize('top,nle', [(2, 2), (1, 1)])
def test_fp32_matches_reference_losses(golden, top, nle):
    want = _golden(golden, top, 'float32', nle)
    got = run_helloworld(extra=['--top', top, '--dtype', 'float32', '--num_local_experts', nle, '--hidden_size', 2048,
                                '--batch_size', 16, '--num_tokens', 1024, '--num_steps', 3, '--parallel_type', 'data'])
    assert [round(v, 3) for v in got] == [round(v, 3) for v in want[:3]]


def test_fp64_matches_reference_losses(golden):
    want = _golden(golden, 2, 'float64', 2)
    got = run_helloworld(extra=['--top', 2, '--dtype', 'float64', '--num_local_experts', 2, '--batch_size', 1,
                                '--num_tokens', 1024, '--num_steps', 4, '--parallel_type', 'data'])
    assert [round(v, 3) for v in got] == [round(v, 3) for v in want[:4]]

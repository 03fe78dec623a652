"""MX block-scaled fp8 kernels (csrc/gemm_mx.cu) against their PyTorch definition (tutel_b200/ops/mx.py)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip('needs a B200')


@pytest.mark.parametrize('shape', [(1, 128, 256), (2, 300, 512), (3, 64, 1024)])
@pytest.mark.parametrize('dtype', [torch.bfloat16, torch.float16])
def test_mx_quantize_kernel_matches_definition(shape, dtype):
    _need_gpu()
    from tutel_b200.ops import mx
    torch.manual_seed(0)
    x = (torch.randn(*shape, device='cuda') * torch.exp(2 * torch.randn(shape[0], shape[1], 1, device='cuda'))).to(dtype)
    x[0, 0, :40] = 0
    q, sf = mx.mx_quantize(x)
    rq, rsf = mx.mx_quantize_reference(x)
    assert torch.equal(q.view(torch.uint8), rq.view(torch.uint8))
    assert torch.equal(sf, rsf)


@pytest.mark.parametrize('G,M,N,K,bn', [(1, 128, 128, 128, 128), (2, 200, 384, 512, 128), (1, 256, 512, 1024, 256),
                                        (2, 1000, 1024, 2048, 0)])
def test_mx_gemm_is_exact_on_exactly_representable_operands(G, M, N, K, bn):
    """Small integers x powers of two: every product and partial sum is exact in fp32, so the tensor-core result must
    equal the fp32 matmul of the dequantised operands bit for bit (after the bf16 rounding of the output) - this pins
    the scale layout (row -> TMEM lane / column, K block -> byte) and the descriptors."""
    _need_gpu()
    from tutel_b200.ops import mx
    g = torch.Generator().manual_seed(M + N)
    a = torch.randint(-3, 4, (G, M, K), generator=g).float().cuda()
    b = torch.randint(-3, 4, (G, N, K), generator=g).float().cuda()
    ea = torch.randint(-2, 3, (G, M, K // 32), generator=g, dtype=torch.int32).cuda()
    eb = torch.randint(-2, 3, (G, N, K // 32), generator=g, dtype=torch.int32).cuda()
    aq, bq = a.to(torch.float8_e4m3fn), b.to(torch.float8_e4m3fn)
    sa, sb = mx.pack_scales(ea), mx.pack_scales(eb)
    ref = torch.matmul(mx.mx_dequantize(aq, sa), mx.mx_dequantize(bq, sb).transpose(1, 2))
    y = mx.mx_gemm(aq, sa, bq, sb, block_n=bn)
    assert torch.equal(y.float(), ref.to(torch.bfloat16).float())


def test_mx_ffn_close_to_bf16_ffn():
    _need_gpu()
    from tutel_b200.ops import mx
    torch.manual_seed(2)
    E, C, M, H = 2, 256, 512, 1024
    x = torch.randn(E, C, M, device='cuda', dtype=torch.bfloat16)
    w1 = (torch.randn(E, H, M, device='cuda') * M ** -0.5).to(torch.bfloat16)
    w2 = (torch.randn(E, H, M, device='cuda') * H ** -0.5).to(torch.bfloat16)
    y = mx.mx_ffn(x, w1, w2).float()
    ref = torch.matmul(torch.relu(torch.matmul(x.float(), w1.float().transpose(1, 2))), w2.float())
    assert float((y - ref).norm() / ref.norm()) < 0.06

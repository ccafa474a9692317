"""GPU: the tcgen05 GEMM against torch fp32 matmul on the same bf16-rounded operands (fp32 accumulation =>
tolerance is accumulation-order noise only: rtol 1e-4 / atol 1e-4 * sqrt(K)-scaled)."""
import pytest
import torch

from se3et_b200.ops.gemm import bmm_bf16, linear_bf16

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("m,n,k", [
    (128, 64, 64), (1000, 32, 64), (4260, 256, 1024), (257, 16, 32), (5000, 128, 1152), (300, 768, 256),
    (1, 256, 256), (20000, 64, 2304), (710, 1024, 512), (333, 512, 1536), (129, 256, 9216),
])
def test_linear_matches_torch(m, n, k):
    g = torch.Generator(device=DEV).manual_seed(m * 7 + n + k)
    a = torch.randn(m, k, device=DEV, generator=g).bfloat16()
    w = (torch.randn(n, k, device=DEV, generator=g) / k ** 0.5).bfloat16()
    bias = torch.randn(n, device=DEV, generator=g)
    ref = a.float() @ w.float().t() + bias
    of, ob = linear_bf16(a, w, bias, out_f32=True, out_bf16=True)
    torch.cuda.synchronize()
    assert torch.allclose(of, ref, rtol=1e-4, atol=2e-4), (of - ref).abs().max().item()
    assert torch.allclose(ob.float(), ref, rtol=1e-2, atol=1e-2)
    of2, _ = linear_bf16(a, w, None, alpha=0.5, relu=True)
    assert torch.allclose(of2, torch.relu(0.5 * (a.float() @ w.float().t())), rtol=1e-4, atol=2e-4)


def test_strided_a_rows():
    g = torch.Generator(device=DEV).manual_seed(1)
    big = torch.randn(500, 256, device=DEV, generator=g).bfloat16()
    a = big[:, 64:192]  # pitch 256, K = 128
    w = torch.randn(64, 128, device=DEV, generator=g).bfloat16()
    of, _ = linear_bf16(a, w)
    assert torch.allclose(of, a.float() @ w.float().t(), rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize("b,m,n,k", [(5, 410, 32, 256), (3, 128, 64, 64), (7, 300, 32, 256)])
def test_batched_matches_torch(b, m, n, k):
    g = torch.Generator(device=DEV).manual_seed(b + m)
    a = torch.randn(b, m, k, device=DEV, generator=g).bfloat16()
    w = torch.randn(b, n, k, device=DEV, generator=g).bfloat16()
    of, _ = bmm_bf16(a, w, alpha=0.125)
    ref = 0.125 * torch.einsum("bmk,bnk->bmn", a.float(), w.float())
    assert torch.allclose(of, ref, rtol=1e-4, atol=1e-3), (of - ref).abs().max().item()

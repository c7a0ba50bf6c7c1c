"""GPU: the tcgen05 building block (bf16 operands, fp32 accumulation in TMEM) against torch."""
import pytest
import torch

from helpers import rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M,K,N,act", [(128, 16, 16, 0), (128, 64, 256, 0), (300, 272, 256, 1), (77, 289, 256, 1),
                                       (2000, 257, 64, 0), (129, 129, 240, 1)])
def test_tc_linear_matches_bf16_reference(M, K, N, act):
    from keypoint_diffusion_b200 import ops, pack
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(M + K + N)
    x = torch.randn(M, K, generator=g)
    w = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g)
    r = torch.randn(M, N, generator=g)
    # reference on the bf16-rounded operands, accumulated in fp64
    xr, wr = x.to(torch.bfloat16).double(), w.to(torch.bfloat16).double()
    ref = xr @ wr.t() + b.double()
    if act:
        ref = torch.nn.functional.silu(ref)
    ref = ref + r.double()
    y = ops.tc_linear(x.to(dev), pack.pack_tc_weight(w).to(dev), N, b.to(dev), r.to(dev), act)
    torch.cuda.synchronize()
    err = rel_err(y.cpu(), ref)
    print(f"tc_linear M={M} K={K} N={N}: rel_err vs bf16-operand reference {err:.2e}; "
          f"vs fp32 {rel_err(y.cpu(), (x.double() @ w.double().t() + b.double()) if not act else ref):.2e}")
    assert err < 2e-5

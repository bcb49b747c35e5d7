"""tcgen05 / TMEM / cp.async.bulk building blocks (csrc/umma.cuh) against torch on bf16-rounded operands."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _arrange(X):
    """[R,K] -> bf16 core-matrix order [R/8][K/8][8][8] (the K-major no-swizzle UMMA operand layout)."""
    R, K = X.shape
    return X.to(torch.bfloat16).reshape(R // 8, 8, K // 8, 8).permute(0, 2, 1, 3).contiguous()


@pytest.mark.parametrize("mode", [0, 1, 2, 3])
@pytest.mark.parametrize("N,K", [(32, 64), (16, 256), (64, 128)])
def test_umma_selftest(mode, N, K):
    from cyclevae_vc_b200._lib import check, ptr
    from tests.native.hooks import load
    lib = load()
    g = torch.Generator().manual_seed(N * 1000 + K + mode)
    A = torch.randn(128, K, generator=g).cuda()
    B = torch.randn(N, K, generator=g).cuda()
    D = torch.full((128, N), float("nan"), device="cuda")
    ref = A.to(torch.bfloat16).double() @ B.to(torch.bfloat16).double().t()
    if mode >= 2:   # 2: bulk-copied operands, SS MMA; 3: A staged smem -> TMEM (tcgen05.cp), TS MMA
        a, b = _arrange(A), _arrange(B)
        check(lib.cvb_selftest_umma(mode, N, K, a.data_ptr(), b.data_ptr(), ptr(D), torch.cuda.current_stream().cuda_stream))
    else:
        check(lib.cvb_selftest_umma(mode, N, K, ptr(A), ptr(B), ptr(D), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    err = (D.double() - ref).abs().max().item()
    assert err < 1e-3, f"mode {mode} N {N} K {K}: max err {err}; D[0,:4]={D[0,:4].tolist()} ref={ref[0,:4].tolist()}"

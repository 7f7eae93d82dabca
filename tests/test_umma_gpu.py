"""GPU: the tcgen05/TMEM GEMM building block (csrc/umma.cuh) against torch on bf16-rounded operands."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("K,N", [(16, 16), (32, 32), (48, 32), (64, 64), (208, 64), (32, 48), (80, 16)])
def test_umma_gemm_matches_torch(K, N):
    from panogrf_b200 import _lib
    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(K * 1000 + N)
    A = torch.randn(128, K, device="cuda", generator=g)
    W = torch.randn(N, K, device="cuda", generator=g)
    out = torch.zeros(128, N, device="cuda")
    rc = lib.pgrf_umma_selftest(_lib.ptr(A), _lib.ptr(W), _lib.ptr(out), K, N, 0, _lib.stream_ptr())
    _lib.check(rc, "pgrf_umma_selftest")
    torch.cuda.synchronize()
    ref = A.bfloat16().float() @ W.bfloat16().float().t()
    err = float((out - ref).abs().max())
    assert err < 1e-3 * max(1.0, float(ref.abs().max())), f"K={K} N={N} max err {err}"


@pytest.mark.parametrize("K,N", [(16, 16), (32, 32), (80, 64), (240, 64)])
def test_umma_a_operand_in_tmem(K, N):
    """tcgen05.mma with the A operand read from TENSOR MEMORY (written by the row-owning threads with tcgen05.st: lane = row,
    two bf16 per 32-bit column, 8 columns per K=16 step) — the building block of the next MLP-kernel layout."""
    from panogrf_b200 import _lib
    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(K * 1000 + N + 7)
    A = torch.randn(128, K, device="cuda", generator=g)
    W = torch.randn(N, K, device="cuda", generator=g)
    out = torch.zeros(128, N, device="cuda")
    _lib.check(lib.pgrf_umma_selftest(_lib.ptr(A), _lib.ptr(W), _lib.ptr(out), K, N, 2, _lib.stream_ptr()), "pgrf_umma_selftest")
    torch.cuda.synchronize()
    ref = A.bfloat16().float() @ W.bfloat16().float().t()
    err = float((out - ref).abs().max())
    assert err < 1e-3 * max(1.0, float(ref.abs().max())), f"K={K} N={N} max err {err}"

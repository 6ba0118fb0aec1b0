"""tcgen05 GEMM mainloop self-tests (C ABI `hcomp_gemm_bf16`): every operand-major combination the
backward uses, full tiles, ragged edges (TMA zero fill) and split-K reduction."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _both_kernel_families(cta_pair_mode):
    """every GEMM test runs on the cta_group::2 (CTA pair) kernels and on the 1-CTA ones (fixture in conftest.py)"""
    yield


def _operands(M, N, K, a_mn, b_mn, seed):
    g = torch.Generator(device='cpu').manual_seed(seed)
    a = torch.randn(M, K, generator=g).to(torch.bfloat16)
    b = torch.randn(K, N, generator=g).to(torch.bfloat16)
    ref = a.double() @ b.double()
    a_st = (a.t().contiguous() if a_mn else a.contiguous()).cuda()      # [K,M] or [M,K]
    b_st = (b.contiguous() if b_mn else b.t().contiguous()).cuda()      # [K,N] or [N,K]
    return a_st, b_st, ref


@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 1), (1, 0)])
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (256, 512, 256), (304, 264, 200), (1000, 768, 520)])
def test_gemm_matches_fp64(M, N, K, a_mn, b_mn):
    from pipnet_b200 import ops
    a, b, ref = _operands(M, N, K, a_mn, b_mn, seed=M + N + K)
    out = ops.gemm_bf16(a, b, M, N, K, a_mn, b_mn, out_mode=1)
    torch.cuda.synchronize()
    err = (out.double().cpu() - ref).abs().max().item()
    assert err <= 2e-4 * max(1.0, ref.abs().max().item()), f'max abs err {err}'


def test_gemm_bf16_output_dx_shape():
    from pipnet_b200 import ops
    M, N, K = 1352, 768, 256
    a, b, ref = _operands(M, N, K, 0, 1, seed=3)
    out = ops.gemm_bf16(a, b, M, N, K, 0, 1, out_mode=0)
    torch.cuda.synchronize()
    torch.testing.assert_close(out.double().cpu(), ref, rtol=1e-2, atol=1e-1)


@pytest.mark.parametrize("splits", [1, 3, 0])
def test_gemm_splitk_red(splits):
    from pipnet_b200 import ops
    M, N, K = 256, 768, 64 * 37
    a, b, ref = _operands(M, N, K, 1, 1, seed=5)
    out = ops.gemm_bf16(a, b, M, N, K, 1, 1, out_mode=2, splits=splits)
    torch.cuda.synchronize()
    err = (out.double().cpu() - ref).abs().max().item()
    assert err <= 1e-3 * max(1.0, ref.abs().max().item()), f'max abs err {err}'


@pytest.mark.parametrize("out_mode", [0, 1, 2])
@pytest.mark.parametrize("M,N,K", [(1352, 768, 504), (136, 264, 72)])
def test_gemm_writes_exactly_its_output(M, N, K, out_mode):
    """guard rows around the output (sanitizer stand-in): ragged M / K (compact dZ: K = 504) must neither leave holes nor
    write past row M -- TMA-store epilogue (bf16), plain fp32 stores and the split-K red.add path."""
    import ctypes as C
    from pipnet_b200._cabi import call, ptr
    a_mn, b_mn = (0, 1) if out_mode != 2 else (1, 1)
    a, b, ref = _operands(M, N, K, a_mn, b_mn, seed=9)
    dt = torch.bfloat16 if out_mode == 0 else torch.float32
    guard = 300
    big = torch.full((M + guard, N), float('nan'), device='cuda', dtype=dt)
    if out_mode == 2:
        big[:M].zero_()                                    # the reduction path accumulates into its output
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    call('hcomp_gemm_bf16', ptr(a), ptr(b), M, N, K, a_mn, b_mn, out_mode, 0 if out_mode == 2 else 1, ptr(big), C.c_longlong(N), stream)
    torch.cuda.synchronize()
    assert torch.isnan(big[M:].float()).all(), "GEMM wrote past its last row"
    got = big[:M].double().cpu()
    assert not torch.isnan(got).any()
    tol = 1e-2 if out_mode == 0 else 2e-4
    assert (got - ref).abs().max().item() <= tol * max(1.0, ref.abs().max().item()) + (1e-1 if out_mode == 0 else 0)

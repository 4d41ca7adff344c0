"""GPU unit test pinning the tcgen05 descriptor encodings of csrc/umma.cuh: a single-CTA
C[128,N] = A[128,K] @ B[N,K]^T through each operand layout, against torch fp32 matmul of the
same bf16 inputs (fp32 accumulation -> tolerance 1e-3 relative to |A||B| row norms)."""
import ctypes

import pytest

torch = pytest.importorskip("torch")

CASES = [
    # mode, N, K
    (0, 256, 64), (0, 256, 256), (0, 64, 128), (0, 32, 64), (0, 16, 64),
    (1, 64, 32), (1, 64, 16), (1, 64, 48), (1, 64, 80), (1, 256, 64), (1, 32, 128),
    (2, 256, 64), (2, 256, 256), (2, 32, 128),
]


@pytest.mark.gpu
@pytest.mark.parametrize("mode,N,K", CASES)
def test_umma_layouts(metrpo_lib, mode, N, K):
    lib = metrpo_lib.load_dev()
    g = torch.Generator(device="cpu").manual_seed(1000 * mode + N + K)
    A = torch.randn(128, K, generator=g).to(torch.bfloat16).cuda()
    B = torch.randn(N, K, generator=g).to(torch.bfloat16).cuda()
    C = torch.full((128, N), float("nan"), device="cuda", dtype=torch.float32)
    st = lib.metrpo_selftest_umma(mode, N, K, 1, metrpo_lib.ptr(A), metrpo_lib.ptr(B),
                                  metrpo_lib.ptr(C), None, metrpo_lib.stream_ptr())
    metrpo_lib.check_dev(st, "selftest")
    torch.cuda.synchronize()
    ref = A.float() @ B.float().t()
    err = (C - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= 1e-3 * scale + 1e-4, (mode, N, K, err, scale)


@pytest.mark.gpu
def test_umma_selftest_rejects_bad_shapes(metrpo_lib):
    lib = metrpo_lib.load_dev()
    st = lib.metrpo_selftest_umma(0, 24, 64, 1, None, None, None, None, None)
    assert st == -1
    assert b"N" in lib.metrpo_last_error()

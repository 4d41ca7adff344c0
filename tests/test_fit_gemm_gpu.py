"""The ensemble fit's batched TF32 tcgen05 GEMM (csrc/fit_gemm.cuh) against a float64 torch product:
all four operand major-ness combinations (K-major / MN-major A and B, read in place), the three fused
epilogues (plain, bias + ReLU, ReLU mask), ragged M / N / K (TMA zero-fill and store clipping), both
tile shapes (N <= 128 -> 128 x 128, else 256 x 256) and several models per launch.  Goes through the
development library's C entry point (include/metrpo_dev.h); the product path calls the same kernel
from csrc/fit_kernels.cu (tests/test_fit_gpu.py).

Tolerance: TF32 keeps 10 mantissa bits and the tensor core TRUNCATES raw fp32 operands: each operand
loses a relative amount uniform in [0, 2^-10), so a product of N(0,1) operands carries an error of rms
~1.05e-3 with random sign, a K-deep dot product one of rms 1.05e-3 * sqrt(K), and the maximum over the
~10^6 outputs of a case sits near 5 sigma.  Asserted: rms error <= 1.5e-3 * sqrt(K), max <= 8e-3 * sqrt(K)
(measured 4.1e-3 ... 5.0e-3 * sqrt(K)); a wrong descriptor / swizzle shows up as errors of order sqrt(K)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _run(M, N, Kd, models, a_mn, b_mn, epi, seed=0):
    from me_trpo_b200 import lib as L
    dev = L.load_dev()
    g = torch.Generator(device="cuda").manual_seed(seed)
    A = torch.randn(models, M, Kd, device="cuda", generator=g)
    B = torch.randn(models, N, Kd, device="cuda", generator=g)
    As = A.transpose(1, 2).contiguous() if a_mn else A.contiguous()
    Bs = B.transpose(1, 2).contiguous() if b_mn else B.contiguous()
    C = torch.full((models, M, N), float("nan"), device="cuda")
    bias = torch.randn(models, N, device="cuda", generator=g)
    aux = torch.randn(models, M, N, device="cuda", generator=g)
    L.check_dev(dev.metrpo_dev_gemm_tf32(M, N, Kd, models, As.data_ptr(), M if a_mn else Kd, M * Kd, a_mn,
                                         Bs.data_ptr(), N if b_mn else Kd, N * Kd, b_mn, C.data_ptr(), N, M * N,
                                         epi, bias.data_ptr(), N, aux.data_ptr(), N, M * N, None,
                                         torch.cuda.current_stream().cuda_stream), "dev_gemm_tf32")
    torch.cuda.synchronize()
    ref = torch.matmul(A.double(), B.double().transpose(1, 2))
    if epi == 1:
        ref = torch.relu(ref + bias[:, None, :].double())
    if epi == 2:
        ref = torch.where(aux > 0, ref, torch.zeros_like(ref))
    assert not torch.isnan(C).any(), "output not fully written"
    diff = C.double() - ref
    err, rms = diff.abs().max().item(), diff.pow(2).mean().sqrt().item()
    assert err <= 8e-3 * np.sqrt(Kd), (M, N, Kd, a_mn, b_mn, epi, err)
    assert rms <= 1.5e-3 * np.sqrt(Kd), (M, N, Kd, a_mn, b_mn, epi, rms)
    if epi == 2:   # masked entries are exact zeros
        assert (C[aux <= 0] == 0).all()


@pytest.mark.parametrize("a_mn", [0, 1])
@pytest.mark.parametrize("b_mn", [0, 1])
@pytest.mark.parametrize("shape", [(200, 96, 72), (1000, 1024, 1024), (384, 320, 1000)],
                         ids=["small-ragged", "fit-shape", "two-tiles-ragged-k"])
def test_gemm_all_major_combinations(shape, a_mn, b_mn):
    M, N, Kd = shape
    if a_mn:
        M = (M + 31) // 32 * 32      # MN-major operands need a multiple of 32 along M / N (H and the padded
    if b_mn:
        N = (N + 31) // 32 * 32      # widths of the fit always are)
    _run(M, N, Kd, 2, a_mn, b_mn, 0)


@pytest.mark.parametrize("epi", [1, 2])
@pytest.mark.parametrize("shape", [(1000, 1024, 1024, 0, 1), (1000, 1024, 32, 0, 0), (500, 64, 256, 0, 1)],
                         ids=["forward", "one-k-block", "narrow-tile"])
def test_gemm_fused_epilogues(shape, epi):
    M, N, Kd, a_mn, b_mn = shape
    _run(M, N, Kd, 3, a_mn, b_mn, epi, seed=epi)


def test_gemm_five_models_wgrad_shape():
    _run(1024, 1024, 1000, 5, 1, 1, 0)     # dW1 = H0^T dH1: both operands MN-major, ragged reduction


def test_gemm_random_shapes_fuzz():
    """40 random (M, N, K, models, major-ness, epilogue) cases -- odd M, N not a multiple of 4 or 32,
    K tails, single-row problems -- through tools/gemm_fuzz.py (also checks that nothing is written
    beyond the clipped output)."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "gemm_fuzz.py"), "7", "40"], capture_output=True,
                         text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "failures: 0" in out.stdout, "\n".join(l for l in out.stdout.splitlines() if not l.startswith("ok"))

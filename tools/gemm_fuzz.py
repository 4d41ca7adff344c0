"""GPU dev probe: random-shape fuzz of the fit's TF32 tcgen05 GEMM against a float64 product."""
import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from me_trpo_b200 import lib as L
dev = L.load_dev()
rng = np.random.RandomState(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
bad = 0
for case in range(int(sys.argv[2]) if len(sys.argv) > 2 else 60):
    a_mn, b_mn = int(rng.randint(2)), int(rng.randint(2))
    epi = int(rng.randint(3))
    models = int(rng.randint(1, 4))
    M = int(rng.randint(1, 700)); N = int(rng.randint(1, 600)); Kd = int(rng.randint(1, 300)) * 4   # ld % 4 == 0 for K-major
    if a_mn: M = max(32, M // 32 * 32)
    if b_mn: N = max(32, N // 32 * 32)
    if not a_mn or not b_mn: pass
    # C's row pitch must be a multiple of 4 floats (TMA store): pad ldc
    ldc = (N + 3) // 4 * 4 + 4 * int(rng.randint(2))
    g = torch.Generator(device="cuda").manual_seed(case)
    A = torch.randn(models, M, Kd, device="cuda", generator=g); B = torch.randn(models, N, Kd, device="cuda", generator=g)
    As = A.transpose(1, 2).contiguous() if a_mn else A.contiguous()
    Bs = B.transpose(1, 2).contiguous() if b_mn else B.contiguous()
    C = torch.full((models, M, ldc), float("nan"), device="cuda")
    bias = torch.randn(models, N, device="cuda", generator=g)
    aux = torch.randn(models, M, ldc, device="cuda", generator=g)
    st = dev.metrpo_dev_gemm_tf32(M, N, Kd, models, As.data_ptr(), M if a_mn else Kd, M * Kd, a_mn, Bs.data_ptr(),
                                  N if b_mn else Kd, N * Kd, b_mn, C.data_ptr(), ldc, M * ldc, epi, bias.data_ptr(), N,
                                  aux.data_ptr(), ldc, M * ldc, None, torch.cuda.current_stream().cuda_stream)
    if st != 0:
        print("case", case, (M, N, Kd, models, a_mn, b_mn, epi), "launch error:", dev.metrpo_last_error().decode()); bad += 1; continue
    torch.cuda.synchronize()
    ref = torch.matmul(A.double(), B.double().transpose(1, 2))
    if epi == 1: ref = torch.relu(ref + bias[:, None, :].double())
    if epi == 2: ref = torch.where(aux[:, :, :N] > 0, ref, torch.zeros_like(ref))
    got = C[:, :, :N].double()
    # the TMA store clips columns at 16-byte granularity: up to round_up(N, 4) may be written (zeros)
    n4 = (N + 3) // 4 * 4
    pad_untouched = bool(torch.isnan(C[:, :, n4:]).all().item()) if ldc > n4 else True
    if n4 > N and epi != 1: pad_untouched = pad_untouched and bool((C[:, :, N:n4] == 0).all().item())
    err = (got - ref).abs().max().item() if not torch.isnan(got).any() else float("nan")
    ok = err <= 8e-3 * np.sqrt(Kd) + 1e-6 and pad_untouched
    if not ok:
        bad += 1
    print(("ok  " if ok else "BAD ") + json.dumps(dict(case=case, M=M, N=N, Kd=Kd, models=models, a_mn=a_mn, b_mn=b_mn, epi=epi,
                                                         err=err, tol=8e-3 * float(np.sqrt(Kd)), pad_untouched=pad_untouched)), flush=True)
print("failures:", bad)

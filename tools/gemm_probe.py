"""GPU dev probe: the fit's batched TF32 tcgen05 GEMM (csrc/fit_gemm.cuh) against torch, all four
operand major-ness combinations and the three epilogues, plus timing at the fit's shapes."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from me_trpo_b200 import lib as L

dev = L.load_dev()
torch.manual_seed(0)
torch.backends.cuda.matmul.allow_tf32 = False


def tf32_round(x):
    # round-to-nearest-even on the 13 dropped mantissa bits is what torch/cuBLAS do on conversion;
    # the tensor core itself truncates -- compare against both bounds below
    return x


def run(M, N, Kd, models, a_mn, b_mn, epi, reps=0):
    A = torch.randn(models, M, Kd, device="cuda")
    B = torch.randn(models, N, Kd, device="cuda")
    As = A.transpose(1, 2).contiguous() if a_mn else A.contiguous()
    Bs = B.transpose(1, 2).contiguous() if b_mn else B.contiguous()
    C = torch.full((models, M, N), float("nan"), device="cuda")
    bias = torch.randn(models, N, device="cuda")
    aux = torch.randn(models, M, N, device="cuda")
    lda = M if a_mn else Kd
    ldb = N if b_mn else Kd
    def call():
        L.check_dev(dev.metrpo_dev_gemm_tf32(M, N, Kd, models, As.data_ptr(), lda, M * Kd, a_mn, Bs.data_ptr(), ldb,
                                             N * Kd, b_mn, C.data_ptr(), N, M * N, epi, bias.data_ptr(), N,
                                             aux.data_ptr(), N, M * N, None, torch.cuda.current_stream().cuda_stream), "gemm")
    call(); torch.cuda.synchronize()
    ref = torch.matmul(A.double(), B.double().transpose(1, 2))
    if epi == 1: ref = torch.relu(ref + bias[:, None, :].double())
    if epi == 2: ref = torch.where(aux > 0, ref, torch.zeros_like(ref))
    err = (C.double() - ref).abs().max().item()
    scale = ref.abs().max().item()
    out = dict(M=M, N=N, Kd=Kd, models=models, a_mn=a_mn, b_mn=b_mn, epi=epi, max_abs_err=err, ref_max=scale,
               nan=int(torch.isnan(C).sum().item()))
    if reps:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3): call()
        e0.record()
        for _ in range(reps): call()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / reps * 1e3
        out["us"] = us; out["tflops"] = 2.0 * M * N * Kd * models / us / 1e6
        # cuBLAS TF32 on the same problem
        torch.backends.cuda.matmul.allow_tf32 = True
        Bt = B.transpose(1, 2).contiguous()
        for _ in range(3): torch.bmm(A, Bt)
        e0.record()
        for _ in range(reps): torch.bmm(A, Bt)
        e1.record(); torch.cuda.synchronize()
        out["cublas_tf32_us"] = e0.elapsed_time(e1) / reps * 1e3
        torch.backends.cuda.matmul.allow_tf32 = False
    print(json.dumps(out), flush=True)
    return out

res = []
# small correctness cases first: ragged M, K tails, N < tile
for a_mn in (0, 1):
    for b_mn in (0, 1):
        res.append(run(200 if not a_mn else 224, 96, 72, 2, a_mn, b_mn, 0))
for epi in (1, 2):
    res.append(run(1000, 1024, 1024, 2, 0, 1, epi))
res.append(run(1024, 1024, 1000, 2, 1, 1, 0))
# the fit's three shapes, timed (K = 5 models, batch 1000, H = 1024)
res.append(run(1000, 1024, 1024, 5, 0, 1, 1, reps=20))   # forward
res.append(run(1024, 1024, 1000, 5, 1, 1, 0, reps=20))   # wgrad
res.append(run(1000, 1024, 1024, 5, 0, 0, 2, reps=20))   # dgrad
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "gemm_probe.json"), "w"), indent=1)

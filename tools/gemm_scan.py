"""GPU dev probe: fit GEMM time vs reduction depth (separates the per-K-block cost from the fixed cost)."""
import json, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from me_trpo_b200 import lib as L
dev = L.load_dev()
def run(M, N, Kd, models, a_mn, b_mn, epi, reps=20):
    A = torch.randn(models, M, Kd, device="cuda"); B = torch.randn(models, N, Kd, device="cuda")
    As = A.transpose(1, 2).contiguous() if a_mn else A; Bs = B.transpose(1, 2).contiguous() if b_mn else B
    C = torch.empty(models, M, N, device="cuda"); bias = torch.randn(models, N, device="cuda"); aux = torch.randn(models, M, N, device="cuda")
    def call():
        L.check_dev(dev.metrpo_dev_gemm_tf32(M, N, Kd, models, As.data_ptr(), M if a_mn else Kd, M * Kd, a_mn, Bs.data_ptr(),
                                             N if b_mn else Kd, N * Kd, b_mn, C.data_ptr(), N, M * N, epi, bias.data_ptr(), N,
                                             aux.data_ptr(), N, M * N, None, torch.cuda.current_stream().cuda_stream), "gemm")
    for _ in range(3): call()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): call()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    print(json.dumps(dict(M=M, N=N, Kd=Kd, models=models, a_mn=a_mn, b_mn=b_mn, epi=epi, us=round(us, 2),
                          tflops=round(2.0 * M * N * Kd * models / us / 1e6, 1))), flush=True)
for models in (5, 9):
    for (a_mn, b_mn) in ((0, 0), (0, 1), (1, 1)):
        for Kd in (32, 1024, 2048, 4096):
            run(1024, 1024, Kd, models, a_mn, b_mn, 0)
for epi in (1, 2):
    for Kd in (32, 1024):
        run(1024, 1024, Kd, 5, 0, 0, epi)

"""GPU dev probe: in-kernel phase stamps (ns) of CTA (0,0,0) of the fit GEMM."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from me_trpo_b200 import lib as L
dev = L.load_dev()
def run(M, N, Kd, models, a_mn, b_mn, epi):
    A = torch.randn(models, M, Kd, device="cuda"); B = torch.randn(models, N, Kd, device="cuda")
    As = A.transpose(1, 2).contiguous() if a_mn else A; Bs = B.transpose(1, 2).contiguous() if b_mn else B
    C = torch.empty(models, M, N, device="cuda"); bias = torch.randn(models, N, device="cuda"); aux = torch.randn(models, M, N, device="cuda")
    dbg = torch.zeros(8, dtype=torch.int64, device="cuda")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for it in range(4):
        e0.record()
        L.check_dev(dev.metrpo_dev_gemm_tf32(M, N, Kd, models, As.data_ptr(), M if a_mn else Kd, M * Kd, a_mn, Bs.data_ptr(),
                                             N if b_mn else Kd, N * Kd, b_mn, C.data_ptr(), N, M * N, epi, bias.data_ptr(), N,
                                             aux.data_ptr(), N, M * N, dbg.data_ptr(), torch.cuda.current_stream().cuda_stream), "gemm")
        e1.record(); torch.cuda.synchronize()
    t = dbg.cpu().numpy()
    print("M %d N %d Kd %d models %d epi %d: event %.1f us | setup %.2f | mainloop %.2f | epilogue issue %.2f | store drain %.2f | teardown %.2f us || group0: ld+transform %.2f, fence %.2f" % (
        M, N, Kd, models, epi, e0.elapsed_time(e1) * 1e3, (t[1] - t[0]) / 1e3, (t[2] - t[1]) / 1e3, (t[3] - t[2]) / 1e3, (t[4] - t[3]) / 1e3, (t[5] - t[4]) / 1e3, (t[6] - t[2]) / 1e3, (t[7] - t[6]) / 1e3))
for Kd in (32, 1024):
    for epi in (0, 1, 2):
        run(1024, 1024, Kd, 5, 0, 1, epi)
run(1000, 32, 1024, 5, 0, 1, 0)
run(1024, 32, 1000, 5, 1, 1, 0)

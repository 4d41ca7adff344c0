"""dev: dump the first TMA stage of the fit GEMM for an MN-major B and decode where elements landed"""
import os, sys, torch, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from me_trpo_b200 import lib as L
dev = L.load_dev()
M, N, Kd = 128, 256, 32
a_mn, b_mn = int(sys.argv[1]), int(sys.argv[2])
# A(m,k) = m*1000 + k ; B(n,k) = -(n*1000 + k)  (exact in fp32)
A = (torch.arange(M, device="cuda")[:, None] * 1000 + torch.arange(Kd, device="cuda")[None, :]).float()
B = -(torch.arange(N, device="cuda")[:, None] * 1000 + torch.arange(Kd, device="cuda")[None, :]).float()
As = A.t().contiguous() if a_mn else A.contiguous()
Bs = B.t().contiguous() if b_mn else B.contiguous()
C = torch.zeros(M, N, device="cuda")
dbg = torch.zeros(48 * 1024 // 4, device="cuda")
L.check_dev(dev.metrpo_dev_gemm_tf32(M, N, Kd, 1, As.data_ptr(), M if a_mn else Kd, M * Kd, a_mn, Bs.data_ptr(),
                                     N if b_mn else Kd, N * Kd, b_mn, C.data_ptr(), N, M * N, 0, None, 0, None, 0, 0,
                                     dbg.data_ptr(), torch.cuda.current_stream().cuda_stream), "gemm")
torch.cuda.synchronize()
d = dbg.cpu().numpy()
sa, sb = d[:4096], d[4096:]
def show(name, t, n):
    print(name, "first 3 rows of 32 floats (128 B each):")
    for r in range(3): print("  row", r, t[r * 32:(r + 1) * 32][:12])
    print("  row 8:", t[8 * 32:9 * 32][:8], " row 32:", t[32 * 32:33 * 32][:8])
    print("  nonzero count", int((t != 0).sum()), "of", t.size)
show("A stage", sa, 128); show("B stage", sb, 256)
ref = A.double() @ B.double().t()
print("C err", (C.double() - ref).abs().max().item(), "ref max", ref.abs().max().item(), "C nonzero", int((C != 0).sum().item()))

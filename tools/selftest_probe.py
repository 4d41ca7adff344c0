"""GPU probe (dev tool): runs the UMMA self-test for every layout and prints per-layout
tensor-pipe cycle counts.  Usage on the GPU box: python tools/selftest_probe.py"""
import ctypes, os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = ctypes.CDLL(os.path.join(ROOT, "me_trpo_b200", "libmetrpo_dev.so"))
lib.metrpo_last_error.restype = ctypes.c_char_p
vp = ctypes.c_void_p
lib.metrpo_selftest_umma.argtypes = [ctypes.c_int] * 4 + [vp] * 5
out = []
for mode, N, K in [(0, 256, 64), (0, 256, 256), (0, 64, 128), (0, 32, 64), (0, 16, 64),
                   (1, 64, 32), (1, 64, 16), (1, 64, 48), (1, 64, 80), (1, 256, 64), (1, 32, 128),
                   (2, 256, 64), (2, 256, 256), (2, 32, 128)]:
    g = torch.Generator().manual_seed(mode * 1000 + N + K)
    A = torch.randn(128, K, generator=g).to(torch.bfloat16).cuda()
    B = torch.randn(N, K, generator=g).to(torch.bfloat16).cuda()
    C = torch.full((128, N), float("nan"), device="cuda")
    cyc = torch.zeros(1, dtype=torch.int64, device="cuda")
    st = lib.metrpo_selftest_umma(mode, N, K, 1, vp(A.data_ptr()), vp(B.data_ptr()), vp(C.data_ptr()),
                                  vp(cyc.data_ptr()), None)
    try:
        torch.cuda.synchronize()
    except Exception as e:
        print("CUDA error", mode, N, K, e); sys.exit(1)
    ref = A.float() @ B.float().t()
    err = (C - ref).abs().max().item()
    rec = dict(mode=mode, N=N, K=K, status=st, max_err=err, ref_max=ref.abs().max().item(),
               nan=int(torch.isnan(C).sum().item()))
    # throughput: 256 back-to-back K loops
    reps = 256
    st = lib.metrpo_selftest_umma(mode, N, K, reps, vp(A.data_ptr()), vp(B.data_ptr()), vp(C.data_ptr()),
                                  vp(cyc.data_ptr()), None)
    torch.cuda.synchronize()
    c = cyc.item()
    rec["cycles_per_mma"] = c / (reps * K / 16)
    rec["ideal_cycles_per_mma"] = 128 * N / 256
    out.append(rec)
    print(json.dumps(rec), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "selftest_probe.json"), "w"), indent=1)

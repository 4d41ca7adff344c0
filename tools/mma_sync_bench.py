"""GPU dev probe: throughput of the legacy mma.sync path on sm_100a (cycles per MMA per SM)."""
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from me_trpo_b200 import lib as L
lib = L.load_dev()
out = torch.zeros(2, dtype=torch.int64, device="cuda")
res = []
for kind, name, macs in ((0, "m16n8k8.tf32", 1024), (1, "m16n8k16.bf16", 2048)):
    for warps in (1, 4, 8, 16):
        reps = 2000
        L.check_dev(lib.metrpo_bench_mma_sync(kind, warps, reps, L.ptr(out), None), "bench")
        torch.cuda.synchronize()
        cyc = out.tolist()[0]
        n = reps * 8 * warps
        r = dict(kind=name, warps=warps, cycles_per_mma_per_sm=cyc / n, mac_per_clk_per_sm=n * macs / cyc,
                 cycles_per_mma_per_warp=cyc / (reps * 8))
        res.append(r); print(json.dumps(r), flush=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "mma_sync_bench.json"), "w"), indent=1)

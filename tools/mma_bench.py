"""GPU dev probe: cycles per tcgen05.mma vs N / sync primitives interleaved in the issue loop."""
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from me_trpo_b200 import lib as L
lib = L.load_dev()
out = torch.zeros(2, dtype=torch.int64, device="cuda")
res = []
def run(ts, N, two, a_col, d_col, wait_each, reps=512):
    L.check_dev(lib.metrpo_bench_mma(ts, N, reps, two, a_col, d_col, wait_each, L.ptr(out), None), "bench")
    torch.cuda.synchronize()
    a, b = out.tolist()
    r = dict(ts=ts, N=N, sync=wait_each, issue_cyc_per_4mma=a / reps, total_cyc_per_4mma=b / reps)
    res.append(r); print(json.dumps(r), flush=True)
for N in (32, 128, 256):
    for w in (0, 1, 2, 3, 4, 7, 16):
        run(1, N, 0, 448, 0, w)
# A from shared memory (SS form) and alternating accumulators, for comparison
for N in (32, 128, 256):
    run(0, N, 0, 448, 0, 0)
    run(1, N, 1, 448, 0, 0)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "mma_bench3.json"), "w"), indent=1)

"""Tabulate gpurun_out/trace.json (written by tools/trace_probe.py)."""
import json, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tr = json.load(open(os.path.join(ROOT, "gpurun_out", "trace.json")))
def tab(evs):
    d = {}
    for c, t in evs: d.setdefault(c, []).append(t)
    return d
P, M, E = [tab(tr[k]) for k in ("producer", "mma", "epilogue")]
t0 = min(min(v) for v in M.values())
def first(d, c, i=0):
    return d[c][i] - t0 if c in d and len(d[c]) > i else None
print("MMA: zwait", first(M, 0x1000), "zready", first(M, 0x1001), "acc2commit", first(M, 0x1002), "| 2nd:", first(M, 0x1000, 1), first(M, 0x1001, 1), first(M, 0x1002, 1))
print("EPI: stepstart", first(E, 0x1000), "zarrive", first(E, 0x1001), "acc2wait", first(E, 0x1002), "acc2ok", first(E, 0x1003), "xchg", first(E, 0x1004), "reward", first(E, 0x1005), "| next", first(E, 0x1000, 1), first(E, 0x1001, 1))
ser = [(c, first(E, c)) for c in (0x1000, 0x1010, 0x1011, 0x1012, 0x1001)]
print("EPI begin-step: start/policy done/eps done/zscratch done/zarrive:", ser)
fin = [(hex(c), first(E, c)) for c in (0x1002, 0x1003, 0x1020, 0x1021, 0x1022, 0x1023, 0x1004, 0x1024, 0x1005)]
print("EPI finish-step:", fin)
rng = range(0, 64) if len(sys.argv) < 2 else range(int(sys.argv[1]), int(sys.argv[2]))
print(" g | MMA: start waited issued | EPI: start acc0ok ld h0free stored arrived | PROD: start emptyok | drain: start acc1ok done")
for g in rng:
    m = [first(M, c | g) for c in (0x100, 0x300, 0x400)]
    e = [first(E, c | g) for c in (0x100, 0x200, 0x300, 0x400, 0x500, 0x600)]
    pr = [first(P, c | g) for c in (0x100, 0x200)]
    dr = [first(E, c | g) for c in (0x700, 0x800, 0x900)]
    print(g, m, e, pr, dr if dr[0] is not None else "")

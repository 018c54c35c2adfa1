"""Device-resident timing of config-2 epoch shards (240 / G epochs, G = 1, 2, 4, 8): the strong-scaling
proxy on one GPU (development aid, not the bench).  `python tools/shard_time.py [once]`; `once` runs every
shard size a single time after warm-up (for an ncu launch list)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import grates_b200 as gb

N, d = 96, 0.5
grid = gb.GeographicGrid(d, d)
plan = gb.get_plan(grid, N, "ewh")
once = len(sys.argv) > 1 and sys.argv[1] == "once"
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
res = {}
for E in (30, 60, 120, 240):
    x = torch.randn(E, N + 1, N + 1, dtype=torch.float64, device="cuda") * 1e-6
    out = torch.empty(E, plan.nlat, plan.nlon, dtype=torch.float64, device="cuda")
    for _ in range(1 if once else 5):
        plan.synthesis(x, out=out)
    torch.cuda.synchronize()
    if once:
        plan.synthesis(x, out=out)
        torch.cuda.synchronize()
        continue
    reps = 50
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        flush.zero_()
        a.record(); plan.synthesis(x, out=out); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    plan.set_profiling(reps)
    for _ in range(reps):
        flush.zero_()
        plan.synthesis(x, out=out)
    st = plan.stage_times(reps).mean(axis=0)
    plan.set_profiling(0)
    res[E] = {"best_ms": ts[0], "median_ms": ts[reps // 2], "mean_ms": float(np.mean(ts)),
              "stages_ms": [round(float(v), 5) for v in st]}
    print(E, res[E], flush=True)
if not once:
    r = {E: res[240]["median_ms"] / res[E]["median_ms"] for E in res}
    print("strong-scaling proxy (t240 / tE):", {k: round(v, 2) for k, v in r.items()})
    print(json.dumps(res))

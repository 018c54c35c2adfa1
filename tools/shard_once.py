"""One config-2 shard of E epochs, run a few times (ncu target; development aid).  python tools/shard_once.py E [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import grates_b200 as gb

E = int(sys.argv[1]) if len(sys.argv) > 1 else 30
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
N, d = 96, 0.5
plan = gb.get_plan(gb.GeographicGrid(d, d), N, "ewh")
x = torch.randn(E, N + 1, N + 1, dtype=torch.float64, device="cuda") * 1e-6
out = torch.empty(E, plan.nlat, plan.nlon, dtype=torch.float64, device="cuda")
for _ in range(reps):
    plan.synthesis(x, out=out)
torch.cuda.synchronize()
print("done")

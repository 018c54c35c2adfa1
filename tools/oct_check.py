"""Octant stage 2 (eight-fold longitude symmetry) against the quadrant kernel (GB_S2_QUADRANT=1) and the oracle, then
timing (development aid).  python tools/oct_check.py [time]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import grates_b200 as gb
from oracle import sh_oracle as orc

def err(a, b):
    return float((a - b).abs().max() / b.abs().max())

for (N, dlon, dlat, Es) in ((96, 0.5, 0.5, (1, 3, 30)), (40, 2.5, 2.0, (5,)), (20, 7.5, 5.0, (2,)), (60, 1.0, 1.0, (4,)), (120, 0.25, 0.5, (2,))):
    grid = gb.GeographicGrid(dlon, dlat)
    plan = gb.get_plan(grid, N, "ewh")
    info = gb._lib.load()
    for E in Es:
        x = torch.randn(E, N + 1, N + 1, dtype=torch.float64, device="cuda") * 1e-6
        os.environ.pop("GB_S2_QUADRANT", None)
        a = plan.synthesis(x).clone()
        os.environ["GB_S2_QUADRANT"] = "1"
        b = plan.synthesis(x).clone()
        os.environ.pop("GB_S2_QUADRANT", None)
        og = orc.geographic_grid(dlon, dlat)
        ref = torch.as_tensor(orc.synthesis(x[0].cpu().numpy(), og, "ewh")).reshape(plan.nlat, plan.nlon).cuda()
        print("N=%d %gx%g nlon=%d E=%d: octant vs quadrant %.2e, octant vs oracle %.2e, quadrant vs oracle %.2e" % (
            N, dlon, dlat, plan.nlon, E, err(a, b), err(a[0], ref), err(b[0], ref)), flush=True)
        assert err(a, b) < 1e-12 and err(a[0], ref) < 1e-12
if len(sys.argv) > 1:
    os.system("%s %s" % (sys.executable, os.path.join(os.path.dirname(os.path.abspath(__file__)), "shard_time.py")))
    os.environ["GB_S2_QUADRANT"] = "1"
    os.system("GB_S2_QUADRANT=1 %s %s" % (sys.executable, os.path.join(os.path.dirname(os.path.abspath(__file__)), "shard_time.py")))
print("ok")

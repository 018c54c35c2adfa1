"""One pass over every kernel family at BASELINE sizes (for the ncu launch list; development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import grates_b200 as gb
from oracle import sh_oracle as orc

which = sys.argv[1:] or ["c2", "c3", "c4", "c5"]
reps = 2
if "c2" in which:
    grid = gb.GeographicGrid(0.5, 0.5); plan = gb.get_plan(grid, 96, "ewh")
    x = torch.randn(240, 97, 97, dtype=torch.float64, device="cuda") * 1e-6
    out = torch.empty(240, plan.nlat, plan.nlon, dtype=torch.float64, device="cuda")
    for _ in range(reps): plan.synthesis(x, out=out)
    torch.cuda.synchronize()
if "c3" in which:
    grid = gb.GeographicGrid(0.25, 0.25); plan = gb.get_plan(grid, 180, "ewh")
    plan.set_analysis(0, grid.area.reshape(plan.nlat, plan.nlon))
    x = torch.randn(120, 181, 181, dtype=torch.float64, device="cuda") * 1e-6
    v = plan.synthesis(x); back = torch.empty_like(x)
    for _ in range(reps): plan.analysis(v, out=back)
    torch.cuda.synchronize()
if "c4" in which:
    grid = gb.GeographicGrid(0.5, 0.5); plan = gb.get_plan(grid, 96, "ewh")
    sigma = torch.as_tensor(orc.synthetic_covariance(96)).cuda()
    out = torch.empty((plan.nlat, plan.nlon), dtype=torch.float64, device="cuda")
    for _ in range(reps): plan.covariance_propagation(sigma, 0, out=out)
    torch.cuda.synchronize()
if "c5" in which:
    blocks = orc.synthetic_filter_blocks(120); flt = gb.OrderWiseFilter(blocks)
    x = torch.randn(500, 121, 121, dtype=torch.float64, device="cuda") * 1e-6
    y = torch.empty_like(x)
    for _ in range(reps): flt.filter_batch(x, out=y)
    torch.cuda.synchronize()
if "pts" in which:
    rng = np.random.default_rng(0)
    P = 41000
    pg = gb.IrregularGrid(rng.uniform(-np.pi, np.pi, P), np.arcsin(rng.uniform(-1, 1, P)))
    pp = gb.get_points_plan(pg, 96, "ewh")
    sigma = torch.as_tensor(orc.synthetic_covariance(96)).cuda()
    x = torch.randn(240, 97, 97, dtype=torch.float64, device="cuda") * 1e-6
    for _ in range(reps):
        pp.covariance_propagation(sigma, 0)
        pp.synthesis(x)
    torch.cuda.synchronize()
print("done")

"""Covariance propagation: folded against unfolded (GB_COV_NO_FOLD=1), mirrored row blocks against the full grid, both
against the CPU oracle; then config-4 timing (development aid).  python tools/cov_check.py [time]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import grates_b200 as gb
from oracle import sh_oracle as orc

def err(a, b):
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))

for (N, d, nmin) in ((24, 5.0, 0), (30, 4.0, 2), (40, 3.0, 5), (17, 9.0, 3)):
    grid = gb.GeographicGrid(d, d); og = orc.geographic_grid(d, d)
    sigma = orc.synthetic_covariance(N, rank=20)[nmin * nmin:, nmin * nmin:]
    ref = orc.covariance_propagation(sigma, og, nmin, N, "ewh") ** 2
    plan = gb.get_plan(grid, N, "ewh")
    s = torch.as_tensor(np.ascontiguousarray(sigma)).cuda()
    for sym in (True, False):
        os.environ.pop("GB_COV_NO_FOLD", None)
        v = plan.covariance_propagation(s, nmin, take_sqrt=False, symmetric=sym).cpu().numpy()
        os.environ["GB_COV_NO_FOLD"] = "1"
        v0 = plan.covariance_propagation(s, nmin, take_sqrt=False, symmetric=sym).cpu().numpy()
        os.environ.pop("GB_COV_NO_FOLD", None)
        nl = plan.nlat
        h = nl // 2
        a, b = 1, max(2, h - 2)
        vm = plan.covariance_propagation(s, nmin, a, b - a, take_sqrt=False, symmetric=sym, mirrored=True).cpu().numpy()
        full = v.reshape(nl, -1)
        em = max(err(vm[:b - a], full[a:b]), err(vm[b - a:], full[nl - b:nl - a]))
        vb = plan.covariance_propagation(s, nmin, 3, 5, take_sqrt=False, symmetric=sym).cpu().numpy()
        print("N=%d nmin=%d nlat=%d sym=%d: folded %.2e  unfolded %.2e  mirrored block %.2e  plain block %.2e" % (
            N, nmin, nl, sym, err(v.ravel(), ref), err(v0.ravel(), ref), em, err(vb, full[3:8])), flush=True)
        assert err(v.ravel(), ref) < 1e-12 and err(v0.ravel(), ref) < 1e-12 and em < 1e-13 and err(vb, full[3:8]) < 1e-13
# a Gauss grid (odd / even parallels) and an odd number of parallels
for nl in (10, 11):
    grid = gb.GaussGrid(nl); og = orc.gauss_grid(nl)
    sigma = orc.synthetic_covariance(8, rank=8)
    ref = orc.covariance_propagation(sigma, og, 0, 8, "geoid")
    std = grid.covariance_propagation(sigma, 0, 8, "geoid")
    print("gauss", nl, err(std, ref)); assert err(std, ref) < 1e-12
if len(sys.argv) > 1:
    N = 96
    grid = gb.GeographicGrid(0.5, 0.5); plan = gb.get_plan(grid, N, "ewh")
    sigma = torch.as_tensor(orc.synthetic_covariance(N)).cuda()
    out = torch.empty((plan.nlat, plan.nlon), dtype=torch.float64, device="cuda")
    for env in ({}, {"GB_COV_NO_FOLD": "1"}):
        os.environ.pop("GB_COV_NO_FOLD", None); os.environ.update(env)
        for _ in range(2): plan.covariance_propagation(sigma, 0, out=out)
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(5)]
        for a, b in ev:
            a.record(); plan.covariance_propagation(sigma, 0, out=out); b.record()
        torch.cuda.synchronize()
        print("config 4", env, "ms:", sorted(round(a.elapsed_time(b), 3) for a, b in ev), flush=True)
    os.environ.pop("GB_COV_NO_FOLD", None)
print("ok")

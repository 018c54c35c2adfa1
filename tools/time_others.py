"""Device timings of the non-headline rows: analysis (config 3), covariance propagation
(config 4), order-wise filter + synthesis (config 5).  Development aid."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import grates_b200 as gb
from oracle import sh_oracle as orc

def ev_time(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts), sorted(ts)[len(ts)//2]

def analysis(N=180, d=0.25, E=120):
    grid = gb.GeographicGrid(d, d)
    plan = gb.get_plan(grid, N, "ewh")
    t0 = time.perf_counter(); plan.set_analysis(0, grid.area.reshape(plan.nlat, plan.nlon)); t_ops = time.perf_counter() - t0
    x = torch.as_tensor(np.stack([orc.synthetic_coefficients(N, e) for e in range(E)])).cuda()
    v = plan.synthesis(x)
    out = torch.empty_like(x)
    best, med = ev_time(lambda: plan.analysis(v, out=out))
    err = float((out - x).abs().max() / x.abs().max())
    L = N + 1
    fl = 2.0 * (2 * L - 1) * E * plan.nlat * plan.nlon + 2.0 * E * plan.nlat * L * L
    print(f"analysis N={N} d={d} E={E}: {best:.3f} ms ({fl/best/1e9:.2f} TF algorithmic), operator build {t_ops:.2f} s host, round-trip err {err:.2e}")

def covprop(N=96, d=0.5):
    grid = gb.GeographicGrid(d, d)
    plan = gb.get_plan(grid, N, "ewh")
    sigma = torch.as_tensor(orc.synthetic_covariance(N)).cuda()
    out = torch.empty((plan.nlat, plan.nlon), dtype=torch.float64, device="cuda")
    best, med = ev_time(lambda: plan.covariance_propagation(sigma, 0, out=out), reps=3, warm=1)
    K = (N + 1) ** 2; P = plan.nlat * plan.nlon
    print(f"covprop N={N} d={d}: {best:.2f} ms; contract 2PK^2 = {2.0*P*K*K/1e12:.1f} TF -> {2.0*P*K*K/best/1e9:.0f} TF algorithmic; "
          f"executed ~{(2.0*plan.nlat*K*K + 2.0*P*(2*N+2)**2)/1e9:.1f} GF -> {(2.0*plan.nlat*K*K + 2.0*P*(2*N+2)**2)/best/1e9:.2f} TF; {P/best*1e3:.3e} points/s")
    # spot check 2 parallels against the oracle
    og = orc.geographic_grid(d, d)
    ref = orc.covariance_propagation(sigma.cpu().numpy(), og, 0, N, "ewh", rows=[3, 200])
    got = out[[3, 200]].cpu().numpy()
    print("   parity (2 parallels):", float(np.abs(got - ref).max() / np.abs(ref).max()))

def filt(N=120, E=500, d=0.25):
    blocks = orc.synthetic_filter_blocks(N)
    flt = gb.OrderWiseFilter(blocks)
    x = torch.as_tensor(np.stack([orc.synthetic_coefficients(N, e) for e in range(E)])).cuda()
    y = torch.empty_like(x)
    best, med = ev_time(lambda: flt.filter_batch(x, out=y))
    byt = 2 * x.numel() * 8 + sum(b.size for b in blocks) * 8
    print(f"filter N={N} E={E}: {best:.3f} ms, {byt/best/1e6:.0f} GB/s algorithmic")
    grid = gb.GeographicGrid(d, d)
    plan = gb.get_plan(grid, N, "ewh")
    out = torch.empty((E, plan.nlat, plan.nlon), dtype=torch.float64, device="cuda")
    best, med = ev_time(lambda: plan.synthesis(flt.filter_batch(x, out=y), out=out), reps=3, warm=1)
    print(f"filter+synthesis N={N} E={E} d={d}: {best:.3f} ms -> {E*plan.nlat*plan.nlon/best/1e6:.2f} Gpt.ep/s")

def points(N=96, E=240, npts=41000):
    rng = np.random.default_rng(1)
    lon, lat = rng.uniform(-np.pi, np.pi, npts), np.arcsin(rng.uniform(-1, 1, npts))
    pp = gb.get_points_plan(gb.IrregularGrid(lon, lat), N, "ewh")
    x = torch.as_tensor(np.stack([orc.synthetic_coefficients(N, e) for e in range(E)])).cuda()
    out = torch.empty((E, npts), dtype=torch.float64, device="cuda")
    fl = 2.0 * npts * (N + 1) ** 2 * E
    best, _ = ev_time(lambda: pp.synthesis(x, out=out), reps=3, warm=1)
    print(f"points synthesis (GEMM) N={N} E={E} points={npts}: {best:.3f} ms, {fl/best/1e9:.2f} TF")
    os.environ["GB_POINTS_SIMPLE"] = "1"
    best2, _ = ev_time(lambda: pp.synthesis(x, out=out), reps=2, warm=1)
    del os.environ["GB_POINTS_SIMPLE"]
    print(f"points synthesis (per-point kernel): {best2:.3f} ms, {fl/best2/1e9:.2f} TF")


if __name__ == "__main__":
    which = sys.argv[1:] or ["analysis", "covprop", "filter"]
    if "analysis" in which: analysis()
    if "covprop" in which: covprop()
    if "filter" in which: filt()
    if "points" in which: points()

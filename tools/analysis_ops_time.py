"""First-call cost of the analysis (operator build) with device-built and host-built latitude operators, and their
agreement, at config 3 (development aid).  python tools/analysis_ops_time.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import grates_b200 as gb

for (N, d) in ((96, 0.5), (180, 0.25)):
    grid = gb.GeographicGrid(d, d)
    res = {}
    for tag, env in (("device", {}), ("host", {"GB_ANALYSIS_HOST_OPERATORS": "1"})):
        gb.clear_plan_cache()
        os.environ.pop("GB_ANALYSIS_HOST_OPERATORS", None); os.environ.update(env)
        plan = gb.get_plan(grid, N, "ewh")
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        plan.set_analysis(0, grid.area.reshape(plan.nlat, plan.nlon))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        x = torch.randn(4, N + 1, N + 1, dtype=torch.float64, device="cuda", generator=torch.Generator(device="cuda").manual_seed(3)).tril() * 1e-6
        v = plan.synthesis(x)
        back = plan.analysis(v)
        res[tag] = (dt, back.clone(), float((back - x).abs().max() / x.abs().max()))
        print("N=%d %s: operator build %.3f s, round trip %.2e" % (N, tag, dt, res[tag][2]), flush=True)
    os.environ.pop("GB_ANALYSIS_HOST_OPERATORS", None)
    print("   device vs host operators: %.2e" % float((res["device"][1] - res["host"][1]).abs().max() / res["host"][1].abs().max()))

"""Quick device-resident timing of the synthesis kernels (development aid, not the bench)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import grates_b200 as gb

def run(N, d, E, reps=10):
    grid = gb.GeographicGrid(d, d)
    plan = gb.get_plan(grid, N, "ewh")
    x = torch.randn(E, N + 1, N + 1, dtype=torch.float64, device="cuda") * 1e-6
    out = torch.empty(E, plan.nlat, plan.nlon, dtype=torch.float64, device="cuda")
    for _ in range(3):
        plan.synthesis(x, out=out)
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); plan.synthesis(x, out=out); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    plan.set_profiling(reps)
    for _ in range(reps):
        plan.synthesis(x, out=out)
    st = plan.stage_times(reps).mean(axis=0)
    plan.set_profiling(0)
    print("   stages (pack, stage1, stage2) ms:", np.round(st, 4))
    L = N + 1
    flops = 2.0 * E * plan.nlat * L * L + 2.0 * (2 * L - 1) * E * plan.nlat * plan.nlon
    best, med = min(ts), sorted(ts)[len(ts) // 2]
    print(f"N={N} d={d} E={E}: best {best:.3f} ms med {med:.3f} ms  -> {flops/best/1e9:.2f} TF algorithmic, "
          f"{E*plan.nlat*plan.nlon/best/1e6:.3f} Gpt.ep/s")

if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "c2":
        run(96, 0.5, 240, reps=2)
    else:
        run(96, 0.5, 240)
        run(60, 1.0, 1)
        run(180, 0.25, 120)

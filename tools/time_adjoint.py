"""Timing of the point-set adjoint (RadialBasisFunctions.to_potential_coefficients) on the GPU."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import grates_b200 as gb
from oracle import sh_oracle as orc

def run(N, P, E):
    rng = np.random.default_rng(1)
    lon = rng.uniform(-np.pi, np.pi, P); lat = np.arcsin(rng.uniform(-1, 1, P))
    K = np.ones((N + 1, N + 1))
    rbf = gb.RadialBasisFunctions(gb.IrregularGrid(lon, lat), K, 0, N)
    v = torch.as_tensor(rng.standard_normal((E, P))).cuda()
    plan = rbf._points_plan()
    for _ in range(2): plan.adjoint(v)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(5): out = plan.adjoint(v)
    ev[1].record(); torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / 5
    t0 = time.perf_counter()
    n_cpu = min(P, 2048)
    ref = orc.radial_basis_to_coefficients(K, v[0, :n_cpu].cpu().numpy(), lon[:n_cpu], lat[:n_cpu], N)
    cpu_s = (time.perf_counter() - t0) * P / n_cpu
    return {"N": N, "points": P, "epochs": E, "ms": ms, "tflops": 2.0 * P * (N + 1) ** 2 * E / ms / 1e9,
            "cpu_s_per_epoch_extrapolated": cpu_s}

if __name__ == "__main__":
    cfgs = ((96, 40962, 1), (96, 40962, 240), (120, 163842, 12))
    if len(sys.argv) > 1:
        cfgs = (tuple(int(x) for x in sys.argv[1:4]),)
    for cfg in cfgs:
        print(json.dumps(run(*cfg)), flush=True)

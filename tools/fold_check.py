"""Deviation of the equator-folded stage 1 from the oracle for decaying and white spectra (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import grates_b200 as gb
from oracle import sh_oracle as orc

def err(a, b): return float(np.abs(a - b).max() / np.abs(b).max())
for N, d in ((96, 0.5), (180, 0.25), (120, 0.25)):
    grid, og = gb.GeographicGrid(d, d), orc.geographic_grid(d, d)
    rng = np.random.default_rng(N)
    kaula = orc.synthetic_coefficients(N, 3)
    white = rng.standard_normal((N + 1, N + 1))
    zonal = np.zeros((N + 1, N + 1)); zonal[:, 0] = rng.standard_normal(N + 1)
    for name, anm in (("kaula", kaula), ("white", white), ("white zonal", zonal)):
        for kernel in ("ewh", "potential"):
            ref = orc.synthesis(anm, og, kernel)
            out = gb.to_grid_batch(anm[None], grid, kernel)[0]
            rows = np.abs(out - ref).max(axis=1) / np.abs(ref).max()
            print(f"N={N} {d}deg {name:12s} {kernel:9s} max-normalised {err(out, ref):.2e}  worst row {int(rows.argmax())} of {rows.size}", flush=True)

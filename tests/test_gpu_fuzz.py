"""Seeded random shapes through the CUDA path against the CPU oracle: ragged grids (parallel / meridian counts
that are not multiples of the tile sizes, meridian counts with and without the four-fold symmetry), small and odd
epoch counts, all kernels, min_degree > 0.  Sizes are kept small enough for the oracle."""
import numpy as np
import pytest

from conftest import maxnorm_err

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

KERNELS = ("ewh", "obp", "potential", "geoid", "surface_density", "anomaly", "deformation", "uplift")
STEPS = (1.5, 2.0, 2.5, 3.0, 4.0, 4.5, 5.0, 6.0, 7.5, 9.0, 10.0, 12.0)
TOL = 1e-12


@pytest.fixture(scope="module")
def gb():
    import grates_b200
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return grates_b200


@pytest.fixture(scope="module")
def orc():
    from oracle import sh_oracle
    return sh_oracle


@pytest.mark.parametrize("seed", range(int(__import__("os").environ.get("GB_FUZZ_SEEDS", "10"))))
def test_random_synthesis_analysis_filter(gb, orc, seed):
    rng = np.random.default_rng(100 + seed)
    dlon, dlat = float(rng.choice(STEPS)), float(rng.choice(STEPS))
    nlon, nlat = int(360 / dlon), int(180 / dlat)
    N = int(rng.integers(1, min(36, nlat - 1)))
    E = int(rng.integers(1, 10))
    kernel = str(rng.choice(KERNELS))
    grid, og = gb.GeographicGrid(dlon, dlat), orc.geographic_grid(dlon, dlat)
    anm = np.stack([orc.synthetic_coefficients(N, 50 * seed + e) for e in range(E)])
    anm[:, 0, 0] = rng.standard_normal(E) * 1e-6
    ref = np.stack([orc.synthesis(a, og, kernel) for a in anm])
    out = gb.to_grid_batch(anm, grid, kernel)
    assert out.shape == ref.shape and maxnorm_err(out, ref) < TOL, (dlon, dlat, N, E, kernel)
    # isotropic filter fused into the synthesis
    w = gb.Butterworth(int(rng.integers(1, 5)), float(rng.integers(3, 30))).degree_weights(N)
    fused = gb.to_grid_batch(anm, grid, kernel, degree_weights=w)
    ref_f = np.stack([orc.synthesis(orc.degreewise_filter(a, w), og, kernel) for a in anm])
    assert maxnorm_err(fused, ref_f) < TOL
    # order-wise filter of a larger degree, truncated to the field's degree (filter.py:182-187)
    blocks = orc.synthetic_filter_blocks(N + int(rng.integers(0, 4)), seed=seed)
    filt = gb.OrderWiseFilter(blocks).filter_batch(anm)
    assert maxnorm_err(filt, np.stack([orc.orderwise_filter(blocks, a) for a in anm])) < 1e-13
    # analysis where the grid resolves the degree
    if 2 * N < nlon and N + 1 <= nlat and kernel in ("ewh", "potential", "geoid", "obp"):
        nmin = int(rng.integers(0, min(3, N) + 1))
        back = gb.analysis_batch(ref, grid, nmin, N, kernel)
        want = orc.analysis_separable(ref, og, nmin, N, kernel)
        # normalised by the size of the input coefficients: a band that holds no signal (only C00 set, nmin = N = 1)
        # comes back as rounding noise around zero from both sides
        scale = max(np.abs(want).max(), np.abs(anm).max())
        assert np.abs(back - want).max() / scale < 1e-10, (dlon, dlat, N, nmin, kernel)


@pytest.mark.parametrize("seed", range(int(__import__("os").environ.get("GB_FUZZ_SEEDS", "10"))))
def test_random_epoch_counts_are_batch_independent(gb, orc, seed):
    """The Legendre stage picks its tiling from the batch width (64-, 80-, 120- and 240-column items, CTAs spread over
    the SMs for narrow launches), the Fourier stage cuts its last round of tiles by the batch height: epoch counts from
    11 to 130 on random grids (folded and not, with and without the longitude symmetries) must give, bit for bit, the
    rows that any sub-batch gives, and agree with the oracle."""
    rng = np.random.default_rng(700 + seed)
    dlon, dlat = float(rng.choice(STEPS)), float(rng.choice(STEPS))
    nlat = int(180 / dlat)
    N = int(rng.integers(1, min(40, nlat - 1)))
    E = int(rng.choice((11, 20, 32, 33, 40, 41, 48, 60, 61, 80, 81, int(rng.integers(11, 131)))))
    kernel = str(rng.choice(KERNELS))
    grid, og = gb.GeographicGrid(dlon, dlat), orc.geographic_grid(dlon, dlat)
    base = np.stack([orc.synthetic_coefficients(N, 7 * seed + e) for e in range(5)])
    scale = rng.uniform(0.5, 1.5, E)
    anm = base[np.arange(E) % 5] * scale[:, None, None]
    x = torch.as_tensor(anm).cuda()
    out = gb.to_grid_batch(x, grid, kernel)
    for _ in range(3):
        lo = int(rng.integers(0, E))
        hi = int(rng.integers(lo + 1, E + 1))
        assert torch.equal(gb.to_grid_batch(x[lo:hi].contiguous(), grid, kernel), out[lo:hi]), (dlon, dlat, N, E, lo, hi)
    e = int(rng.integers(0, E))
    assert maxnorm_err(out[e].cpu().numpy(), orc.synthesis(anm[e], og, kernel)) < TOL, (dlon, dlat, N, E, kernel)


@pytest.mark.parametrize("seed", range(max(5, int(__import__("os").environ.get("GB_FUZZ_SEEDS", "10")) // 2)))
def test_random_covariance_and_statistics(gb, orc, seed):
    rng = np.random.default_rng(300 + seed)
    dlon, dlat = float(rng.choice(STEPS[3:])), float(rng.choice(STEPS[3:]))
    N = int(rng.integers(2, 15))
    nmin = int(rng.integers(0, 3))
    kernel = str(rng.choice(("ewh", "potential", "geoid")))
    grid, og = gb.GeographicGrid(dlon, dlat), orc.geographic_grid(dlon, dlat)
    sigma = orc.synthetic_covariance(N, rank=int(rng.integers(4, 20)), seed=seed)[nmin * nmin:, nmin * nmin:]
    ref = orc.covariance_propagation(sigma, og, nmin, N, kernel)
    std = grid.covariance_propagation(sigma, nmin, N, kernel)
    assert maxnorm_err(std, ref) < TOL, (dlon, dlat, N, nmin, kernel)
    plan = gb.get_plan(grid, N, kernel)
    s = torch.as_tensor(sigma).cuda()
    cut = int(rng.integers(1, plan.nlat))
    parts = torch.cat([plan.covariance_propagation(s, nmin, 0, cut), plan.covariance_propagation(s, nmin, cut, plan.nlat - cut)])
    assert maxnorm_err(parts.cpu().numpy().ravel(), ref) < TOL
    masks = rng.uniform(size=(3, grid.point_count)) < 0.4
    var = gb.basin_variances(sigma, grid, masks, nmin, N, kernel, take_sqrt=False)
    assert maxnorm_err(var, orc.basin_variances(sigma, og, masks, nmin, N, kernel)) < 1e-11
    vals = rng.standard_normal((4, grid.point_count))
    st = gb.grid_statistics(vals, grid, masks[0])
    for e in range(4):
        g1 = grid.copy()
        g1.values = vals[e]
        assert abs(st["mean"][e] - g1.mean(masks[0])) < 1e-13
        assert abs(st["std"][e] - g1.std(masks[0])) < 1e-13


@pytest.mark.parametrize("seed", range(max(5, int(__import__("os").environ.get("GB_FUZZ_SEEDS", "10")) // 2)))
def test_random_point_sets(gb, orc, seed):
    """Point-set kernels at random sizes: synthesis on both sides of the per-point / GEMM switch (16 epochs), the
    direct diag(F Sigma F'), the adjoint on all three epoch-tile widths (24 / 48 / 120), the dense operator."""
    rng = np.random.default_rng(500 + seed)
    P = int(rng.choice((1, 3, 127, 128, 129, int(rng.integers(2, 1500)))))
    N = int(rng.integers(1, 31))
    E = int(rng.choice((1, 2, 15, 16, 17, 24, 25, 48, 49, int(rng.integers(1, 60)))))
    kernel = str(rng.choice(("ewh", "potential", "geoid", "obp")))
    lon, lat = rng.uniform(-np.pi, np.pi, P), np.arcsin(rng.uniform(-1, 1, P))
    pts = gb.IrregularGrid(lon, lat)
    anm = np.stack([orc.synthetic_coefficients(N, 31 * seed + e) for e in range(E)])
    anm[:, 0, 0] = rng.standard_normal(E) * 1e-6
    out = gb.to_grid_batch(anm, pts, kernel)
    pick = sorted(set((0, E // 2, E - 1)))
    ref = np.stack([orc.synthesis_points(anm[e], lon, lat, kernel) for e in pick])
    assert out.shape == (E, P) and maxnorm_err(out[pick], ref) < TOL, (P, N, E, kernel)
    nmin = int(rng.integers(0, min(2, N) + 1))
    sigma = orc.synthetic_covariance(N, rank=int(rng.integers(4, 20)), seed=seed)[nmin * nmin:, nmin * nmin:]
    std = pts.covariance_propagation(sigma, nmin, N, kernel)
    assert maxnorm_err(std, orc.covariance_propagation_points(sigma, lon, lat, nmin, N, kernel)) < TOL, (P, N, nmin, kernel)
    K = rng.uniform(0.5, 1.5, (N + 1, N + 1))
    rbf = gb.RadialBasisFunctions(pts, K, 0, N)
    v = rng.standard_normal((E, P))
    got = rbf.to_potential_coefficients_batch(v).cpu().numpy()
    want = np.stack([orc.radial_basis_to_coefficients(K, v[e], lon, lat, N) for e in pick])
    assert maxnorm_err(got[pick], want) < TOL, (P, N, E)
    A = pts.synthesis_matrix(nmin, N, kernel)
    assert maxnorm_err(A, orc.synthesis_matrix_points(lon, lat, nmin, N, kernel)) < TOL

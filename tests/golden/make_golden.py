"""Generate golden input/output vectors by running the UNMODIFIED reference (akvas/grates,
mounted read-only at /root/reference) on seeded inputs.  Run in the build container only:

    python tests/golden/make_golden.py

The reference cannot travel to the GPU box, so the resulting small ``*.npz`` fixtures are
committed.  netCDF4/h5py are absent and unused on this path; they are stubbed so that
``import grates`` succeeds (grates/__init__.py:48 -> grates/io.py:18-19).
"""
import os
import sys
import types
import warnings

import numpy as np

warnings.filterwarnings("ignore")
_nc = types.ModuleType("netCDF4")
_nc.Dataset = object
sys.modules.setdefault("netCDF4", _nc)
sys.modules.setdefault("h5py", types.ModuleType("h5py"))
sys.path.insert(0, "/root/reference")
import grates  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
GM, R = 3.9860044150e+14, 6.3781363000e+06


def coeffs(nmax, seed):
    """Kaula-like synthetic coefficients: degree-n entries * 1e-5/n^2, degrees 0-1 zero."""
    rng = np.random.default_rng(seed)
    anm = rng.standard_normal((nmax + 1, nmax + 1))
    for n in range(1, nmax + 1):
        anm[grates.gravityfield.degree_indices(n)] *= 1e-5 / n ** 2
    anm[0:2, 0:2] = 0
    pc = grates.gravityfield.PotentialCoefficients(GM, R)
    pc.anm = anm
    return pc


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(name, {k: np.asarray(v).shape for k, v in arrays.items()}, os.path.getsize(path) // 1024, "KiB")


def l1_numerics():
    rng = np.random.default_rng(11)
    colat = np.sort(rng.uniform(0.01, np.pi - 0.01, 13))
    colat[0], colat[-1] = 1e-3, np.pi - 2e-3
    lon = rng.uniform(-np.pi, np.pi, 13)
    out = dict(colat=colat, lon=lon)
    for N in (5, 40, 200):
        c = colat if N < 200 else colat[::4]
        out["legendre_%d" % N] = grates.utilities.legendre_functions(N, c)
    for m in (0, 1, 7, 40):
        out["legendre_order_%d" % m] = grates.utilities.legendre_functions_per_order(40, m, colat)
    out["legendre_polynomials_40"] = grates.utilities.legendre_polynomials(40, colat)
    out["trig_12"] = grates.utilities.trigonometric_functions(12, lon)
    out["ynm_9"] = grates.utilities.spherical_harmonics(9, colat, lon)
    a3 = rng.standard_normal((3, 7, 7))
    out["ravel_in"] = a3
    out["ravel_0_6"] = grates.utilities.ravel_coefficients(a3, 0, 6)
    out["ravel_2_6"] = grates.utilities.ravel_coefficients(a3, 2, 6)
    out["ravel_2_4"] = grates.utilities.ravel_coefficients(a3[0], 2, 4)
    v = rng.standard_normal(49 - 4)
    out["unravel_in"] = v
    out["unravel_2_6"] = grates.utilities.unravel_coefficients(v, 2, 6)
    lat = np.linspace(-np.pi / 2, np.pi / 2, 9)
    out["lat"] = lat
    out["geocentric_radius"] = grates.utilities.geocentric_radius(lat)
    out["colatitude"] = grates.utilities.colatitude(lat)
    out["colatitude_wgs"] = grates.utilities.colatitude(lat, 6378137.0, 1 / 298.257223563)
    save("l1_numerics", **out)


def kernels():
    lat = np.linspace(89.5, -89.5, 7) * np.pi / 180
    colat = grates.utilities.colatitude(lat)
    r = grates.utilities.geocentric_radius(lat)
    out = dict(lat=lat, colat=colat, r=r)
    N = 24
    for name in ("ewh", "obp", "potential", "geoid", "surface_density", "anomaly", "deformation", "uplift"):
        k = grates.kernel.get_kernel(name)
        out["coeff_" + name] = k.coefficients(0, N, r, colat)
        out["inv_" + name] = k.inverse_coefficients(0, N, r, colat)
        out["kn_" + name] = k.inverse_coefficients(0, N, r, colat) * \
            np.power((R / r)[:, np.newaxis], np.arange(N + 1, dtype=int) + 1) * GM / R
    out["normal_gravity"] = grates.gravityfield.GRS80.normal_gravity(r, colat)
    # known answers the reference's own tests assert (testing/gravityfield.py:80-86)
    out["gamma_equator_pole"] = np.array([
        grates.gravityfield.GRS80.normal_gravity(np.array([grates.gravityfield.GRS80.R]), np.array([np.pi / 2]))[0],
        grates.gravityfield.GRS80.normal_gravity(
            np.array([grates.gravityfield.GRS80.R * (1 - grates.gravityfield.GRS80.flattening)]), np.array([0.0]))[0]])
    k_ce, h_ce, l_ce = grates.data.load_love_numbers(frame="CE")
    out["love_k_200"] = k_ce[:201]
    out["love_h_200"] = h_ce[:201]
    save("kernels", **out)


def grids():
    out = {}
    g = grates.grid.GeographicGrid(2.0, 4.0)
    out["geo_meridians"], out["geo_parallels"], out["geo_area"] = g.meridians, g.parallels, g.area
    gg = grates.grid.GaussGrid(14)
    out["gauss_meridians"], out["gauss_parallels"], out["gauss_area"] = gg.meridians, gg.parallels, gg.area
    rng = np.random.default_rng(21)
    mer = np.sort(rng.uniform(-np.pi, np.pi, 20))
    par = np.sort(rng.uniform(-1.5, 1.5, 11))[::-1]
    rg = grates.grid.RegularGrid(mer, par)
    out["reg_meridians"], out["reg_parallels"], out["reg_area"] = rg.meridians, rg.parallels, rg.area
    save("grids", **out)


def synthesis():
    out = {}
    # config 1 of BASELINE.json: degree 60 -> 1 deg x 1 deg, ewh
    pc = coeffs(60, 1000)
    g = grates.grid.GeographicGrid(1.0, 1.0)
    out["c1_anm"] = pc.anm
    out["c1_ewh"] = pc.to_grid(g, "ewh").value_array
    # small multi-kernel / multi-epoch cases on a 10 degree grid
    g10 = grates.grid.GeographicGrid(10.0, 10.0)
    anm = np.stack([coeffs(20, 1000 + e).anm for e in range(3)])
    out["s_anm"] = anm
    for name in ("ewh", "obp", "potential", "geoid", "surface_density", "anomaly", "deformation", "uplift"):
        vals = []
        for e in range(3):
            pc = grates.gravityfield.PotentialCoefficients(GM, R)
            pc.anm = anm[e]
            vals.append(pc.to_grid(g10, name).value_array)
        out["s_" + name] = np.stack(vals)
    # non-default GM / R
    pc = grates.gravityfield.PotentialCoefficients(3.986004415e14 * 1.001, 6378137.0)
    pc.anm = anm[0]
    out["s_gmr"] = np.array([pc.GM, pc.R])
    out["s_ewh_gmr"] = pc.to_grid(g10, "ewh").value_array
    # Gauss grid and a regular grid with random parallels / meridians
    gg = grates.grid.GaussGrid(24)
    pc = coeffs(20, 1000)
    out["gauss_ewh"] = pc.to_grid(gg, "ewh").value_array
    rng = np.random.default_rng(21)
    mer = np.sort(rng.uniform(-np.pi, np.pi, 50))
    par = np.sort(rng.uniform(-1.5, 1.5, 31))[::-1]
    rg = grates.grid.RegularGrid(mer, par)
    out["reg_meridians"], out["reg_parallels"] = mer, par
    out["reg_geoid"] = pc.to_grid(rg, "geoid").value_array
    # degree 0 / degree 1 only and a high-degree polar-underflow case (N=200 on 4 parallels near the pole)
    pc0 = grates.gravityfield.PotentialCoefficients(GM, R)
    pc0.anm = np.array([[1.0]])
    out["deg0_potential"] = pc0.to_grid(g10, "potential").value_array
    pch = coeffs(200, 7)
    par_p = np.array([89.9, 89.0, 0.3, -89.95]) * np.pi / 180
    mer_p = np.linspace(-np.pi + 0.1, np.pi - 0.1, 16)
    rp = grates.grid.RegularGrid(mer_p, par_p)
    out["hi_anm"], out["hi_parallels"], out["hi_meridians"] = pch.anm, par_p, mer_p
    out["hi_ewh"] = pch.to_grid(rp, "ewh").value_array
    # irregular branch (gravityfield.py:370-388)
    lon = rng.uniform(-np.pi, np.pi, 700)
    lat = rng.uniform(-1.55, 1.55, 700)
    ig = grates.grid.IrregularGrid(lon, lat)
    pc = coeffs(15, 1003)
    out["irr_lon"], out["irr_lat"], out["irr_anm"] = lon, lat, pc.anm
    out["irr_ewh"] = pc.to_grid(ig, "ewh").values
    save("synthesis", **out)


def analysis():
    out = {}
    rng = np.random.default_rng(31)
    g = grates.grid.GeographicGrid(10.0, 10.0)    # 18 x 36 -> N <= 12
    N = 12
    pc = coeffs(N, 1000)
    for name in ("ewh", "potential"):
        grid = pc.to_grid(g, name)
        noise = grid.copy()
        noise.values = grid.values + rng.standard_normal(grid.values.size) * 1e-3 * np.abs(grid.values).max()
        out["in_" + name] = noise.value_array
        out["anm_%s_0_12" % name] = noise.to_potential_coefficients(0, N, name, GM, R).anm
        out["anm_%s_2_12" % name] = noise.to_potential_coefficients(2, N, name, GM, R).anm
        out["anm_%s_3_9" % name] = noise.to_potential_coefficients(3, 9, name, GM, R).anm
    gg = grates.grid.GaussGrid(14)
    grid = coeffs(10, 1001).to_grid(gg, "ewh")
    out["gauss_in"] = grid.value_array
    out["gauss_anm_0_10"] = grid.to_potential_coefficients(0, 10, "ewh", GM, R).anm
    mer = np.sort(rng.uniform(-np.pi, np.pi, 30))
    par = np.sort(rng.uniform(-1.5, 1.5, 17))[::-1]
    rg = grates.grid.RegularGrid(mer, par)
    grid = coeffs(8, 1002).to_grid(rg, "geoid")
    out["reg_meridians"], out["reg_parallels"] = mer, par
    out["reg_in"] = grid.value_array
    out["reg_anm_0_8"] = grid.to_potential_coefficients(0, 8, "geoid", GM, R).anm
    # dense operators in degree-wise order (grid.py:412-443, :698-730)
    g30 = grates.grid.GeographicGrid(30.0, 30.0)
    out["synthesis_matrix_1_4"] = g30.synthesis_matrix(1, 4, "ewh", GM, R)
    out["analysis_matrix_0_4"] = g30.analysis_matrix(0, 4, "ewh", GM, R)
    save("analysis", **out)


def covariance():
    out = {}
    rng = np.random.default_rng(4)
    N = 8
    K = (N + 1) ** 2
    Lr = rng.standard_normal((K, 16)) * 1e-11
    sigma = Lr @ Lr.T + np.diag(rng.uniform(0.5, 1.5, K) * 1e-22)
    out["sigma"] = sigma
    g = grates.grid.GeographicGrid(15.0, 15.0)
    out["std_ewh_0_8"] = g.copy().covariance_propagation(sigma, 0, N, "ewh", GM, R)
    out["std_potential_2_8"] = g.copy().covariance_propagation(sigma[4:, 4:], 2, N, "potential", GM, R)
    gg = grates.grid.GaussGrid(10)
    out["std_gauss_geoid"] = gg.covariance_propagation(sigma, 0, N, "geoid", GM, R)
    lon = rng.uniform(-np.pi, np.pi, 300)
    lat = rng.uniform(-1.55, 1.55, 300)
    ig = grates.grid.IrregularGrid(lon, lat)
    out["irr_lon"], out["irr_lat"] = lon, lat
    out["std_irr_ewh"] = ig.covariance_propagation(sigma, 0, N, "ewh", GM, R)
    save("covariance", **out)


def filters():
    out = {}
    rng = np.random.default_rng(5)
    nf = 12
    blocks = []
    for m in range(nf + 1):
        k = nf + 1 - m
        for _ in range(1 if m == 0 else 2):
            blocks.append(0.5 * np.eye(k) + 0.01 * rng.standard_normal((k, k)))
    flt = grates.filter.OrderWiseFilter(blocks)
    for i, b in enumerate(blocks):
        out["block_%02d" % i] = b
    pc12, pc9 = coeffs(12, 1000), coeffs(9, 1001)
    pc12.anm[0:2, 0:2] = rng.standard_normal((2, 2))
    out["in_12"], out["in_9"] = pc12.anm, pc9.anm
    out["out_12"] = flt.filter(pc12).anm
    out["out_9"] = flt.filter(pc9).anm
    out["matrix_2_9"] = flt.matrix(2, 9)
    # time series ordering (gravityfield.py:964-980)
    import datetime
    data = []
    for e in range(3):
        p = coeffs(4, 1000 + e)
        p.epoch = datetime.datetime(2002, 4, 15) + datetime.timedelta(days=30.4375 * (2 - e))
        data.append(p)
    ts = grates.gravityfield.TimeSeries(data)
    out["ts_anm_sorted"] = np.stack([d.anm for _, d in ts.items()])
    out["ts_array"] = ts.to_array()
    save("filters", **out)


def degreewise_filters():
    """Gaussian / Butterworth filters and the Gauss kernel (filter.py:31-130, kernel.py:464-506)."""
    out = {}
    pc = coeffs(40, 1002)
    pc.anm[0:2, 0:2] = np.random.default_rng(9).standard_normal((2, 2))
    out["in_40"] = pc.anm
    for radius in (0.0, 150.0, 500.0):
        tag = "%d" % radius
        out["gauss_w_" + tag] = grates.kernel.Gauss(radius).coefficients(0, 200).ravel()
        out["gauss_out_" + tag] = grates.filter.Gaussian(radius).filter(pc).anm
    out["gauss_matrix_300_2_9"] = np.diag(grates.filter.Gaussian(300.0).matrix(2, 9))
    out["gauss_w_300_ext"] = grates.kernel.Gauss(300.0).coefficients(1020, 1030).ravel()
    for order, cutoff in ((2, 30), (5, 12)):
        tag = "%d_%d" % (order, cutoff)
        out["butter_out_" + tag] = grates.filter.Butterworth(order, cutoff).filter(pc).anm
        out["butter_matrix_" + tag] = np.diag(grates.filter.Butterworth(order, cutoff).matrix(1, 12))
    out["ewh_gauss300_in40"] = grates.filter.Gaussian(300.0).filter(pc).to_grid(
        grates.grid.GeographicGrid(dlon=6.0, dlat=6.0), kernel='ewh').value_array
    save("degreewise_filters", **out)


def dense_filters():
    """GeneralMatrix and VDK (filter.py:430-546)."""
    out = {}
    rng = np.random.default_rng(17)
    nmin, nmax = 2, 10
    k = (nmax + 1) ** 2 - nmin ** 2
    W = 0.6 * np.eye(k) + 0.02 * rng.standard_normal((k, k))
    out["W_2_10"] = W
    flt = grates.filter.GeneralMatrix(W, nmin, nmax)
    for N, seed in ((10, 1003), (7, 1004), (14, 1005)):
        pc = coeffs(N, seed)
        pc.anm[0:2, 0:2] = rng.standard_normal((2, 2))
        out["in_%d" % N] = pc.anm
        out["out_%d" % N] = flt.filter(pc).anm
    out["matrix_0_12"] = flt.matrix(0, 12)
    out["matrix_4_8"] = flt.matrix(4, 8)
    A = rng.standard_normal((2 * k, k))
    Nmat = A.T @ A
    out["normals_2_10"] = Nmat
    vdk = grates.filter.VDK(Nmat, nmin, nmax, 1e2, 2.0)
    out["vdk_matrix"] = vdk.matrix(nmin, nmax)
    bvdk = grates.filter.BlockedNormalsVDK(Nmat, nmin, nmax, 1e2, 2.0)
    out["blocked_vdk_matrix"] = bvdk.matrix(0, nmax)
    pc10 = coeffs(10, 1003)
    pc10.anm = out["in_10"].copy()
    out["blocked_vdk_out_10"] = bvdk.filter(pc10).anm
    save("dense_filters", **out)


def radial_basis():
    """RadialBasisFunctions.to_potential_coefficients / to_grid (gravityfield.py:645-781) on a seeded point set."""
    out = {}
    rng = np.random.default_rng(23)
    N, P = 20, 700           # three 256-point blocks of the reference, the last one ragged
    lon = rng.uniform(-np.pi, np.pi, P)
    lat = np.arcsin(rng.uniform(-1, 1, P))
    K = np.zeros((N + 1, N + 1))
    shape = 1.0 / (1.0 + 0.1 * np.arange(N + 1)) ** 2          # isotropic shape factors sigma_n
    for n in range(2, N + 1):
        K[n, 0:n + 1] = shape[n]
        K[0:n, n] = shape[n]
    pts = grates.grid.IrregularGrid(lon, lat)
    rbf = grates.gravityfield.RadialBasisFunctions(pts, K, 2, N)
    rbf.values = rng.standard_normal(P) * 1e-9
    out["lon"], out["lat"], out["K"], out["values"] = lon, lat, K, rbf.values
    out["anm"] = rbf.to_potential_coefficients().anm
    out["matrix_cols"] = rbf.to_potential_coefficients_matrix()[:, ::50]       # [K', 14] of the [K', 700] matrix
    g = grates.grid.GeographicGrid(6.0, 6.0)
    out["grid_ewh"] = rbf.to_grid(g, "ewh").value_array
    # anisotropic basis functions (gravityfield.py:573-642): dense operator between the point adjoint and the synthesis
    N2, nmin2, P2 = 10, 2, 300
    k = (N2 + 1) ** 2 - nmin2 ** 2
    Ka = 0.5 * np.eye(k) + 0.05 * rng.standard_normal((k, k))
    lon2, lat2 = lon[:P2], lat[:P2]
    abf = grates.gravityfield.AnisotropicBasisFunctions(grates.grid.IrregularGrid(lon2, lat2), Ka, nmin2, N2)
    abf.values = rng.standard_normal(P2) * 1e-9
    out["aniso_K"], out["aniso_values"] = Ka, abf.values
    out["aniso_grid_ewh"] = abf.to_grid(grates.grid.GeographicGrid(10.0, 10.0), "ewh").value_array
    save("radial_basis", **out)


def irregular_operators():
    """IrregularGrid.synthesis_matrix(_per_order) / analysis_matrix / to_potential_coefficients
    (grid.py:412-443, 477-507, 957-1017)."""
    out = {}
    rng = np.random.default_rng(29)
    P, N = 240, 6
    lon = rng.uniform(-np.pi, np.pi, P)
    lat = np.arcsin(rng.uniform(-1, 1, P))
    area = rng.uniform(0.5, 1.5, P) * 4 * np.pi / P
    g = grates.grid.IrregularGrid(lon, lat, area)
    out["lon"], out["lat"], out["area"] = lon, lat, area
    out["A_2_N_ewh"] = g.synthesis_matrix(2, N, "ewh")
    c3, s3 = g.synthesis_matrix_per_order(3, 2, N, "ewh", GM, R)
    out["A_m3_cos"], out["A_m3_sin"] = c3, s3
    out["F_0_N_potential"] = g.analysis_matrix(0, N, "potential")
    pc = coeffs(N, 1006)
    vals = g.synthesis_matrix(0, N, "ewh") @ grates.utilities.ravel_coefficients(pc.anm) + 1e-4 * rng.standard_normal(P)
    g.values = vals
    out["values"] = vals
    out["anm_2_N_ewh"] = g.to_potential_coefficients(2, N, "ewh").anm
    save("irregular_operators", **out)


def space_kernels():
    """AnisotropicKernel.evaluate / evaluate_grid and FilterKernel (kernel.py:576-658, filter.py:575-598)."""
    out = {}
    rng = np.random.default_rng(37)
    nmin, nmax = 2, 12
    k = (nmax + 1) ** 2 - nmin ** 2
    K = 0.3 * np.eye(k) + 0.02 * rng.standard_normal((k, k))
    ker = grates.kernel.AnisotropicKernel(K, nmin, nmax)
    elon = rng.uniform(-np.pi, np.pi, 50)
    elat = np.arcsin(rng.uniform(-1, 1, 50))
    glon = np.linspace(-np.pi, np.pi, 24, endpoint=False) + 0.1
    glat = np.linspace(1.4, -1.4, 13)
    out["K"], out["eval_lon"], out["eval_lat"], out["grid_lon"], out["grid_lat"] = K, elon, elat, glon, glat
    out["sources"] = np.array([[0.3, 0.7], [-2.0, -0.4]])
    for i, (slon, slat) in enumerate(out["sources"]):
        out["points_%d" % i] = ker.evaluate(slon, slat, elon, elat)
        out["grid_%d" % i] = ker.evaluate_grid(slon, slat, glon, glat)
    fk = grates.filter.FilterKernel(grates.filter.Gaussian(400.0), nmin, nmax, "ewh")
    out["gauss_points"] = fk.evaluate(0.3, 0.7, elon, elat)       # (evaluate_grid raises on FilterKernel's 3-d matrix)
    blocks = [0.7 * np.eye(nmax + 1) + 0.03 * rng.standard_normal((nmax + 1, nmax + 1))]
    for m in range(1, nmax + 1):
        for _ in range(2):
            blocks.append(0.7 * np.eye(nmax + 1 - m) + 0.03 * rng.standard_normal((nmax + 1 - m, nmax + 1 - m)))
    for i, b in enumerate(blocks):
        out["block_%d" % i] = b
    fo = grates.filter.FilterKernel(grates.filter.OrderWiseFilter(blocks), nmin, nmax, "potential")
    out["orderwise_points"] = fo.evaluate(-2.0, -0.4, elon, elat)
    out["mtf_psi"] = np.linspace(0.0, 0.6, 25)
    out["mtf"] = fo.modulation_transfer(out["mtf_psi"], 0.2, 0.3, 0.4)
    save("space_kernels", **out)


def normal_equations():
    """Covariance producers: BlockMatrix.cholesky / sparse_inverse / inverse / solve_triangular /
    multiply_symmetric and NormalEquations.solve / compute_covariance of the reference on a block-sparse SPD matrix
    (block arrow + band structure with ragged block sizes), and its SINEX reader on a file written by
    grates_b200.io.savesinexnormals (committed next to the arrays)."""
    sys.path.insert(0, os.path.join(HERE, "..", ".."))
    import copy
    rng = np.random.default_rng(21)
    index = np.array([0, 37, 101, 165, 190, 270, 333])          # ragged blocks, 333 parameters
    nb, n = len(index) - 1, int(index[-1])
    keep = np.zeros((nb, nb), dtype=bool)
    for i in range(nb):
        keep[i, i] = True
        if i + 1 < nb:
            keep[i, i + 1] = True          # band
        keep[i, nb - 1] = True             # arrow
    keep[0, 3] = True                      # an extra off-band block: fill-in in the factor
    A = rng.standard_normal((2 * n, n))
    N = A.T @ A / (2 * n) + np.eye(n)
    for i in range(nb):
        for j in range(nb):
            if not (keep[min(i, j), max(i, j)]):
                N[index[i]:index[i + 1], index[j]:index[j + 1]] = 0
    N = N + np.diag(np.abs(N).sum(axis=1))                      # diagonally dominant: SPD with the zero blocks
    rhs = rng.standard_normal((n, 2))
    bm = grates.lstsq.BlockMatrix.from_array(N, index, index)
    out = dict(index=index, N=N, rhs=rhs, keep=keep)
    out["multiply_symmetric"] = copy.deepcopy(bm).multiply_symmetric(rhs)
    chol = copy.deepcopy(bm)
    chol.cholesky()
    out["cholesky"] = np.triu(chol.to_array())
    out["solve_t"] = chol.solve_triangular(rhs, transpose=True)
    out["solve_n"] = chol.solve_triangular(rhs, transpose=False)
    out["multiply_triangular"] = chol.multiply_triangular(rhs)
    sp = copy.deepcopy(chol)
    sp.sparse_inverse()
    out["sparse_inverse"] = sp.to_array()
    full = copy.deepcopy(chol)
    full.inverse()
    out["inverse"] = full.to_array()
    ne = grates.lstsq.NormalEquations(copy.deepcopy(bm), rhs[:, 0:1].copy(), 7.5, 4000)
    ne.compute_covariance(sparse=False)
    out["covariance_dense"] = ne.matrix.to_array()
    # SINEX: file written by this repository's writer, read back by the reference's reader
    from grates_b200 import io as gio
    pcount = 45
    cf = []
    for nn in range(2, 9):
        for m in range(0, nn + 1):
            cf.append((0, nn, m))
            if m > 0:
                cf.append((1, nn, m))
    cf = cf[:pcount]
    B = rng.standard_normal((90, pcount))
    Ns, ns = B.T @ B, rng.standard_normal((pcount, 1)) * 1e3
    path = os.path.join(HERE, "normals_small.snx")
    gio.savesinexnormals(path, Ns, ns, np.array([1234.5678]), 90, cf)
    rN, rn, rl, rc = grates.io.loadsinexnormals(path)
    out.update(sinex_N=rN, sinex_n=rn, sinex_lPl=rl, sinex_obs_count=np.array(rc))
    save("normal_equations", **out)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        for name in sys.argv[1:]:
            globals()[name]()
        sys.exit(0)
    l1_numerics()
    kernels()
    grids()
    synthesis()
    analysis()
    covariance()
    filters()
    degreewise_filters()
    dense_filters()
    radial_basis()
    irregular_operators()
    space_kernels()
    normal_equations()

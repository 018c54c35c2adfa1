"""CPU oracle for the spherical-harmonic hot path of akvas/grates  --  TEST INFRASTRUCTURE ONLY.

A numpy restatement of the reference's algorithm for synthesis, analysis, covariance
propagation and order-wise filtering.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this module; the
product package ``grates_b200`` never does (it fails loudly without its CUDA library).

Parity status: PINNED.  ``tests/golden/make_golden.py`` imports the unmodified reference
from ``/root/reference`` in the build container, runs it on seeded inputs and stores
inputs + outputs under ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks every
function here against those fixtures (bit-exact for the table builders, <= 4e-15
max-normalised for the BLAS-summed results, whose summation order belongs to OpenBLAS).

Every function cites the reference lines (relative to /root/reference/grates/) it follows.
The evaluation order of every floating-point expression is kept so that the Legendre,
trigonometric and kernel-factor tables are bit-identical to the reference's.
"""
import os

import numpy as np

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "grates_b200", "data",
                     "love_numbers_ak135_ce.npz")

GM_DEFAULT = 3.9860044150e+14   # gravityfield.py:89
R_DEFAULT = 6.3781363000e+06    # gravityfield.py:89
A_GRS80 = 6378137.0             # grid.py:1141
F_GRS80 = 298.2572221010 ** -1  # grid.py:1141


# --------------------------------------------------------------------------------------
# L1 numerics: Legendre functions, trigonometric tables, coefficient packing
# --------------------------------------------------------------------------------------
def legendre_functions(nmax, colat):
    """Packed fully-normalised P_nm for all n, m <= nmax (utilities.py:13-59).

    out[:, n, m] = P_nm, mirrored into out[:, m-1, n] for m >= 1 (utilities.py:56-57).
    """
    theta = np.atleast_1d(colat)
    L = nmax + 1
    P = np.empty((theta.size, L, L))
    P[:, 0, 0] = 1.0
    if nmax == 0:
        return P
    ct, st = np.cos(theta), np.sin(theta)
    P[:, 1, 0] = np.sqrt(3) * ct                      # :38
    P[:, 1, 1] = np.sqrt(3) * st                      # :39
    for n in range(2, L):                             # sectorials, :41-43
        P[:, n, n] = np.sqrt((2.0 * n + 1.0) / (2.0 * n)) * st * P[:, n - 1, n - 1]
    deg = np.arange(L)
    n1, m1 = deg[2:], deg[1:-1]                       # first off-diagonal, :45-47
    P[:, n1, m1] = np.sqrt(2 * n1 + 1) * ct[:, None] * P[:, m1, m1]
    for d in range(2, L):                             # d = n - m, :49-54
        n, m = deg[d:], deg[:L - d]
        a = np.sqrt((2.0 * n - 1.0) / (n - m) * (2.0 * n + 1.0) / (n + m))
        b = np.sqrt((2.0 * n + 1.0) / (2.0 * n - 3.0) * (n - m - 1.0) / (n - m) * (n + m - 1.0) / (n + m))
        P[:, n, m] = a * ct[:, None] * P[:, n - 1, m] - b * P[:, n - 2, m]
    for m in range(1, L):                             # mirror, :56-57
        P[:, m - 1, m:] = P[:, m:, m]
    return P


def legendre_polynomials(nmax, colat):
    """Order-0 column (utilities.py:138-151, derivative=None branch)."""
    t = np.cos(np.atleast_1d(colat))
    P = np.empty((t.size, nmax + 1))
    P[:, 0] = 1
    if nmax == 0:
        return P
    P[:, 1] = np.sqrt(3) * t
    for n in range(2, nmax + 1):
        P[:, n] = np.sqrt((2.0 * n - 1.0) * (2.0 * n + 1.0)) / n * t * P[:, n - 1] - \
            np.sqrt((2.0 * n + 1.0) / (2.0 * n - 3.0)) * (n - 1.0) / n * P[:, n - 2]
    return P


def legendre_functions_per_order(nmax, m, colat):
    """P_nm for one order, n = m..nmax, with s = sqrt(1 - t^2) (utilities.py:62-115)."""
    if m == 0:
        return legendre_polynomials(nmax, colat)
    if m > nmax:
        raise ValueError('order exceeds maximum degree ({0:d} vs. {1:d})'.format(m, nmax))
    t = np.cos(np.atleast_1d(colat))
    s = np.sqrt(1 - t ** 2)
    cnt = nmax + 1 - m
    out = np.empty((t.size, cnt))
    pmm = np.sqrt(3) * s                               # :99
    for n in range(2, m + 1):                          # :101-103
        pmm = np.sqrt((2 * n + 1) / (2 * n)) * s * pmm
    out[:, 0] = pmm
    if cnt > 1:
        out[:, 1] = np.sqrt(2 * m + 3) * t * out[:, 0]  # :107
    for n in range(m + 2, nmax + 1):                   # :109-113
        out[:, n - m] = np.sqrt((2 * n - 1) / (n - m) * (2 * n + 1) / (n + m)) * t * out[:, n - 1 - m] - \
            np.sqrt((2 * n + 1) / (2 * n - 3) * (n - m - 1) / (n - m) * (n + m - 1) / (n + m)) * out[:, n - 2 - m]
    return out


def trigonometric_functions(nmax, lon):
    """Packed cos(m lon) / sin(m lon) table (utilities.py:249-275)."""
    lam = np.atleast_1d(lon)
    L = nmax + 1
    cs = np.empty((lam.size, L, L))
    cs[:, :, 0] = 1
    for m in range(1, L):
        cs[:, m:, m] = np.cos(m * lam)[:, None]
        cs[:, m - 1, m:] = np.sin(m * lam)[:, None]
    return cs


def spherical_harmonics(nmax, colat, lon):
    """Packed Y_nm = trig * P_nm at paired points (utilities.py:278-307)."""
    return trigonometric_functions(nmax, lon) * legendre_functions(nmax, colat)


def _degreewise_index(nmin, nmax):
    """(rows, cols) of the packed array in the degree-wise vector order of
    utilities.py:336-343: per degree n: C_n0, then (C_nm, S_nm) pairs for m = 1..n."""
    rows, cols = [], []
    for n in range(nmin, nmax + 1):
        rows.append(n), cols.append(0)
        for m in range(1, n + 1):
            rows.append(n), cols.append(m)
            rows.append(m - 1), cols.append(n)
    return np.array(rows, dtype=int), np.array(cols, dtype=int)


def ravel_coefficients(array, nmin=0, nmax=None):
    """Packed [.., L, L] -> degree-wise vector (utilities.py:310-360)."""
    if array.ndim not in (2, 3):
        raise ValueError('Only 2d or 3d spherical harmonic arrays can be raveled.')
    if nmax is None:
        nmax = array.shape[-1] - 1
    count = (nmax + 1) ** 2 - nmin ** 2
    top = min(array.shape[-1] - 1, nmax)
    r, c = _degreewise_index(nmin, top)
    out = np.zeros(array.shape[:-2] + (count,), dtype=array.dtype)
    out[..., :r.size] = array[..., r, c]
    return out


def unravel_coefficients(vector, nmin=0, nmax=None):
    """Degree-wise vector -> packed [.., L, L] (utilities.py:363-411)."""
    if vector.ndim not in (1, 2):
        raise ValueError('Only 1d or 2d spherical harmonic vectors can be unraveled.')
    if nmax is None:
        nmax = int(np.sqrt(vector.shape[-1] + nmin * nmin) - 1)
    r, c = _degreewise_index(nmin, nmax)
    out = np.zeros(vector.shape[:-1] + (nmax + 1, nmax + 1), dtype=vector.dtype)
    out[..., r, c] = vector[..., :r.size]
    return out


def geocentric_radius(lat, a=A_GRS80, f=F_GRS80):
    """utilities.py:414-435."""
    e2 = f * (2 - f)
    nu = a / np.sqrt(1 - e2 * np.sin(lat) ** 2)
    return nu * np.sqrt(np.cos(lat) ** 2 + (1 - e2) ** 2 * np.sin(lat) ** 2)


def colatitude(lat, a=A_GRS80, f=F_GRS80):
    """utilities.py:438-459."""
    e2 = f * (2 - f)
    nu = a / np.sqrt(1 - e2 * np.sin(lat) ** 2)
    return np.arccos(nu * (1 - e2) * np.sin(lat) / geocentric_radius(lat, a, f))


# --------------------------------------------------------------------------------------
# Grids (only the geometry the hot path reads: parallels, meridians, area weights)
# --------------------------------------------------------------------------------------
class OracleGrid:
    """meridians [nlon], parallels [nlat] (north -> south), areas [nlat, nlon], ellipsoid."""

    def __init__(self, meridians, parallels, areas=None, a=A_GRS80, f=F_GRS80):
        self.meridians = np.asarray(meridians, dtype=float)
        self.parallels = np.asarray(parallels, dtype=float)
        self.a, self.f = a, f
        if areas is None:                              # grid.py:537-540
            lon_edges = np.concatenate(([-np.pi], self.meridians[0:-1] + 0.5 * np.diff(self.meridians), [np.pi]))
            lat_edges = np.concatenate(([0.5 * np.pi], self.parallels[0:-1] + 0.5 * np.diff(self.parallels), [-0.5 * np.pi]))
            areas = 2.0 * (np.sin(np.abs(np.diff(lat_edges)) * 0.5) * np.cos(self.parallels))[:, None] * np.diff(lon_edges)
        self.areas = areas

    @property
    def shape(self):
        return self.parallels.size, self.meridians.size


def geographic_grid(dlon=0.5, dlat=0.5, a=A_GRS80, f=F_GRS80):
    """grid.py:1141-1153."""
    nlons, nlats = 360 / dlon, 180 / dlat
    meridians = np.linspace(-np.pi + dlon / 180 * np.pi * 0.5, np.pi - dlon / 180 * np.pi * 0.5, int(nlons))
    parallels = -np.linspace(-np.pi * 0.5 + dlat / 180 * np.pi * 0.5, np.pi * 0.5 - dlat / 180 * np.pi * 0.5, int(nlats))
    areas = np.tile(2.0 * dlon / 180 * np.pi * np.sin(dlat * 0.5 / 180 * np.pi) * np.cos(parallels)[:, None], (1, meridians.size))
    return OracleGrid(meridians, parallels, areas, a, f)


def gauss_grid(parallel_count, a=A_GRS80, f=F_GRS80):
    """grid.py:1181-1195."""
    from scipy.special import roots_legendre
    zeros, weights, _ = roots_legendre(parallel_count, mu=True)
    dlon = np.pi / parallel_count
    meridians = np.linspace(-np.pi + dlon * 0.5, np.pi - dlon * 0.5, 2 * parallel_count)
    cosine_theta = -zeros
    sine_theta = np.sqrt(1 - cosine_theta ** 2)
    parallels = np.arctan2(cosine_theta, (1 - f) ** 2 * sine_theta)
    areas = np.tile(dlon * weights[:, None], (1, meridians.size))
    return OracleGrid(meridians, parallels, areas, a, f)


# --------------------------------------------------------------------------------------
# Isotropic kernels -> kn[lat, n] table
# --------------------------------------------------------------------------------------
_LOVE = None


def love_numbers(frame='CE'):
    """(k, h, l) load Love numbers (data/__init__.py:12-64); table truncated to degree 4096."""
    global _LOVE
    if _LOVE is None:
        with np.load(_DATA) as z:
            _LOVE = np.stack((z['h'], z['l'], z['k']), axis=1)
    hlk = _LOVE.copy()
    fr = frame.lower()
    if fr == 'cm':
        hlk[1, :] -= 1
    elif fr == 'cf':
        ce = hlk[1, :].copy()
        hlk[1, 0] = (ce[0] - ce[1]) * 2 / 3
        hlk[1, 1] = (ce[0] - ce[1]) * -1 / 3
        hlk[1, 2] = (-1 / 3 * ce[0] - 2 / 3 * ce[1])
    elif fr != 'ce':
        raise ValueError('frame of load love numbers must be one of CM, CE, or CF (got <' + frame + '>)')
    return hlk[:, 2], hlk[:, 0], hlk[:, 1]


def _grs80_field():
    """Zonal coefficients of the GRS80 normal field (gravityfield.py:1498-1542, J2 branch)."""
    GM, omega, a, J2 = 3986005e8, 7292115.0e-11, 6378137.0, 108263e-8
    e, e0 = 0.1, np.inf
    n = np.arange(1, 21, dtype=float)
    while not np.isclose(e, e0, atol=1e-22, rtol=0):
        e0 = e
        ep = e / np.sqrt(1 - e ** 2)
        q0 = -2 * np.sum(np.power(-1, n) * n * np.power(ep, 2 * n + 1) / ((2 * n + 1) * (2 * n + 3)))
        e = np.sqrt(3 * J2 + 4 / 15 * (omega ** 2 * a ** 3) / GM * e ** 3 / (2 * q0))
    e2 = e ** 2
    flattening = 1 - np.sqrt(1 - e2)
    coeffs = [1.0]
    k = 1
    while not np.isclose(coeffs[-1], 0, atol=1e-22, rtol=0):
        sign = 1 if k % 2 == 0 else -1
        coeffs.append(sign * (3 * e2 ** k * (1 - k + 5 * k * J2 / e2) / ((2 * k + 1) * (2 * k + 3) * np.sqrt(4 * k + 1))))
        k += 1
    nmax = (len(coeffs) - 1) * 2
    zonal = np.zeros(nmax + 1)
    zonal[0::2] = coeffs
    return GM, omega, a, flattening, zonal


_GRS80 = None


def normal_gravity(r, colat):
    """GRS80 normal gravity on the meridian plane y = 0 (gravityfield.py:1544-1570 with
    gravitational_acceleration :423-481 specialised to a zonal field, lon = 0)."""
    global _GRS80
    if _GRS80 is None:
        _GRS80 = _grs80_field()
    GM, omega, a, flat, zonal = _GRS80
    r = np.atleast_1d(np.asarray(r, dtype=float))
    colat = np.atleast_1d(np.asarray(colat, dtype=float))
    cnt = max(r.size, colat.size)
    x = np.zeros(cnt) + r * np.sin(colat)
    z = np.zeros(cnt) + r * np.cos(colat)
    # cartesian2geodetic (grid.py:1991-2006), Bowring iteration
    e2 = 2 * flat - flat ** 2
    p2 = x ** 2
    h0 = 0
    k = (1 - e2) ** -1
    for _ in range(10):
        c = np.power(p2 + (1 - e2) * z ** 2 * k ** 2, 1.5) / (a * e2)
        k = 1 + (p2 + (1 - e2) * z ** 2 * k ** 3) / (c - p2)
        h = (k ** -1 - (1 - e2)) * np.sqrt(p2 + z ** 2 * k ** 2) / e2
        if np.max(np.abs(h - h0)) < 1e-6:
            break
        h0 = h
    lat = np.arctan2(k * z, np.sqrt(p2))
    # cartesian2spherical (grid.py:2029-2031)
    rr = np.sqrt(x ** 2 + 0.0 + z ** 2)
    th = np.arctan2(np.sqrt(x ** 2 + 0.0), z)
    nz = zonal.size - 1
    n = np.arange(nz + 1, dtype=float)
    P0 = legendre_functions_per_order(nz + 1, 0, th)
    P1 = legendre_functions_per_order(nz + 1, 1, th)
    f_zero = np.sqrt((n + 1) * (n + 1)) * np.sqrt((2 * n + 1) / (2 * n + 3))
    f_plus = np.sqrt((n + 1) * (n + 2)) * np.sqrt((2 * n + 1) / (2 * n + 3)) * np.sqrt(2)
    c_zero = P0[:, 1:] * f_zero
    c_plus = (P1 * np.cos(np.zeros(cnt))[:, None]) * f_plus
    up = np.power(a / rr[:, None], n + 2)
    gx = -(c_plus * up) @ zonal
    gz = -2 * (c_zero * up) @ zonal
    gx = gx * GM / (2 * a ** 2)
    gz = gz * GM / (2 * a ** 2)
    gx = gx + omega ** 2 * x
    return -np.cos(lat) * gx - np.sin(lat) * gz


_KERNEL_ALIASES = {
    'ewh': 'ewh', 'water_height': 'ewh', 'obp': 'obp', 'ocean_bottom_pressure': 'obp',
    'potential': 'potential', 'geoid': 'geoid', 'geoid_height': 'geoid',
    'surface_density': 'surface_density', 'anomaly': 'anomaly', 'gravity_anomaly': 'anomaly',
    'deformation': 'deformation', 'vertical_derformation': 'deformation', 'uplift': 'uplift',
}


def kernel_coefficients(name, nmin, nmax, r, colat):
    """k_n(r, colat) table [points, nmax-nmin+1] (kernel.py:17-67 dispatch, :403-574 formulas)."""
    key = _KERNEL_ALIASES.get(name.lower())
    if key is None:
        raise ValueError("Unrecognized kernel '{0:s}'.".format(name))
    r = np.atleast_1d(np.asarray(r, dtype=float))
    colat = np.atleast_1d(np.asarray(colat, dtype=float))
    deg = np.arange(nmin, nmax + 1, dtype=float)
    if key == 'ewh':                                  # kernel.py:405-406
        k, _, _ = love_numbers()
        kn = (4 * np.pi * 6.673e-11 * 1025) * (1 + k[nmin:nmax + 1]) / (2 * deg + 1)
        return (kn[:, None] * r).T
    if key == 'obp':                                  # :420-421
        k, _, _ = love_numbers()
        kn = (4 * np.pi * 6.673e-11) * (1 + k[nmin:nmax + 1]) / (2 * deg + 1)
        return (kn[:, None] * (r / normal_gravity(r, colat))).T
    if key == 'surface_density':                      # :434-435
        k, _, _ = love_numbers()
        kn = (4 * np.pi * 6.673e-11) * (1 + k[nmin:nmax + 1]) / (2 * deg + 1)
        return (kn[:, None] * r).T
    if key == 'potential':                            # :447-449
        return np.ones((max(r.size, colat.size), nmax + 1 - nmin))
    if key == 'anomaly':                              # :460-461
        kn = np.array([1 / (n - 1) if n != 1 else 0.0 for n in deg])
        return (kn[:, None] * r).T
    if key == 'geoid':                                # :518
        return np.tile(normal_gravity(r, colat)[:, None], (1, nmax + 1 - nmin))
    if key == 'deformation':                          # :553-559
        k, h, _ = love_numbers('CE')
        ratio = h / (1 + k)
        return normal_gravity(r, colat)[:, None] / ratio[nmin:nmax + 1]
    if key == 'uplift':                               # :574
        return 2 * normal_gravity(r, colat)[:, None] / (2 * deg + 1)
    raise AssertionError(key)


def inverse_kernel_coefficients(name, nmin, nmax, r, colat):
    """1/k_n, zero where the whole degree column is ~0 (kernel.py:187-188)."""
    with np.errstate(divide='ignore'):
        kn = kernel_coefficients(name, nmin, nmax, r, colat)
        cols = [np.zeros(kn.shape[0]) if np.allclose(kn[:, j], 0.0) else 1.0 / kn[:, j] for j in range(kn.shape[1])]
    return np.vstack(cols).T


def kn_table(name, nmax, lat, a=A_GRS80, f=F_GRS80, GM=GM_DEFAULT, R=R_DEFAULT):
    """Per-point factor kn[i, n] = inv_k_n * (R/r)^(n+1) * GM / R (gravityfield.py:353-356,
    grid.py:653-657, :819-823).  Returns (colat, kn)."""
    colat = colatitude(lat, a, f)
    radius = geocentric_radius(lat, a, f)
    kn = inverse_kernel_coefficients(name, 0, nmax, radius, colat) * \
        np.power((R / radius)[:, None], np.arange(nmax + 1, dtype=int) + 1) * GM / R
    return colat, kn


def _scale_packed_by_degree(P, kn):
    """In-place P[:, n, m] *= kn[:, n] in the packed layout (gravityfield.py:359-362)."""
    L = P.shape[1]
    P[:, :, 0] *= kn
    for m in range(1, L):
        P[:, m:, m] *= kn[:, m:]
        P[:, m - 1, m:] *= kn[:, m:]
    return P


# --------------------------------------------------------------------------------------
# Synthesis
# --------------------------------------------------------------------------------------
def synthesis(anm, grid, kernel='ewh', GM=GM_DEFAULT, R=R_DEFAULT):
    """PotentialCoefficients.to_grid on a regular grid (gravityfield.py:352-368) -> [nlat, nlon].

    Does the same work the reference does per call: Legendre table, factor scaling, trig
    table, then one (nlat x L)@(L x nlon) product per packed row.
    """
    anm = np.asarray(anm, dtype=float)
    nmax = anm.shape[0] - 1
    colat, kn = kn_table(kernel, nmax, grid.parallels, grid.a, grid.f, GM, R)
    P = _scale_packed_by_degree(legendre_functions(nmax, colat), kn)
    P *= anm[None, :, :]                               # :363
    cs = trigonometric_functions(nmax, grid.meridians)  # :365
    out = np.zeros(grid.shape)
    for k in range(nmax + 1):                          # :367-368
        out += P[:, k, :] @ cs[:, k, :].T
    return out


def synthesis_points(anm, lon, lat, kernel='ewh', GM=GM_DEFAULT, R=R_DEFAULT, a=A_GRS80, f=F_GRS80):
    """Irregular-point branch of to_grid (gravityfield.py:370-388), 512-point blocks."""
    anm = np.asarray(anm, dtype=float)
    nmax = anm.shape[0] - 1
    lon, lat = np.asarray(lon, dtype=float), np.asarray(lat, dtype=float)
    out = np.zeros(lon.size)
    step = min(512, lon.size)
    for i1 in range(0, lon.size, step):
        i2 = min(i1 + step, lon.size)
        colat, kn = kn_table(kernel, nmax, lat[i1:i2], a, f, GM, R)
        Y = _scale_packed_by_degree(spherical_harmonics(nmax, colat, lon[i1:i2]), kn)
        for k in range(nmax + 1):
            out[i1:i2] += Y[:, k, :] @ anm[k, :]
    return out


# --------------------------------------------------------------------------------------
# Analysis (area-weighted least squares, order by order)
# --------------------------------------------------------------------------------------
def synthesis_matrix_per_order(grid, m, nmin, nmax, kernel, GM=GM_DEFAULT, R=R_DEFAULT):
    """RegularGrid.synthesis_matrix_per_order (grid.py:653-663): rows = points (lat-major)."""
    colat, kn = kn_table(kernel, nmax, grid.parallels, grid.a, grid.f, GM, R)
    P = (legendre_functions_per_order(nmax, m, colat) * kn[:, m:])[:, max(nmin - m, 0):]
    nlon = grid.meridians.size
    if m == 0:
        return np.repeat(P, nlon, axis=0)
    c = np.cos(m * grid.meridians[:, None])
    s = np.sin(m * grid.meridians[:, None])
    return (P[:, None, :] * c[None, :, :]).reshape(-1, P.shape[1]), (P[:, None, :] * s[None, :, :]).reshape(-1, P.shape[1])


def analysis_operator_per_order(grid, m, nmin, nmax, kernel, GM=GM_DEFAULT, R=R_DEFAULT):
    """RegularGrid.__analysis_matrix_per_order (grid.py:690-696): solve(A'WA, A'W)."""
    w = grid.areas.ravel()[:, None]
    mats = synthesis_matrix_per_order(grid, m, nmin, nmax, kernel, GM, R)
    if m == 0:
        return np.linalg.solve((mats * w).T @ mats, (mats * w).T)
    return tuple(np.linalg.solve((A * w).T @ A, (A * w).T) for A in mats)


def analysis_direct(values, grid, nmin, nmax, kernel='potential', GM=GM_DEFAULT, R=R_DEFAULT):
    """RegularGrid.to_potential_coefficients exactly as written (grid.py:776-785).
    O(P * sum_m (N-m+1)^2): use for small grids only."""
    v = np.asarray(values, dtype=float).ravel()
    anm = np.zeros((nmax + 1, nmax + 1))
    anm[nmin:, 0] = analysis_operator_per_order(grid, 0, nmin, nmax, kernel, GM, R) @ v
    for m in range(1, nmax + 1):
        Fc, Fs = analysis_operator_per_order(grid, m, nmin, nmax, kernel, GM, R)
        i0 = max(m, nmin)
        anm[i0:, m] = Fc @ v
        anm[m - 1, i0:] = Fs @ v
    return anm


def analysis_separable(values, grid, nmin, nmax, kernel='potential', GM=GM_DEFAULT, R=R_DEFAULT):
    """Same estimator as analysis_direct, using that A_m[(i,j), n] = p_i[n] * trig_m(j) and that
    the area weights are an outer product w_i * u_j: a longitude transform followed
    by a small weighted Legendre least-squares problem per order.  values: [E, nlat, nlon] or
    [nlat, nlon]; returns packed anm [E, L, L] or [L, L].
    """
    v = np.asarray(values, dtype=float)
    single = v.ndim == 2
    if single:
        v = v[None]
    w_lat, u_lon = separable_weights(grid.areas)
    colat, kn = kn_table(kernel, nmax, grid.parallels, grid.a, grid.f, GM, R)
    lam = grid.meridians
    out = np.zeros((v.shape[0], nmax + 1, nmax + 1))
    for m in range(nmax + 1):
        P = (legendre_functions_per_order(nmax, m, colat) * kn[:, m:])[:, max(nmin - m, 0):]
        gram = (P * w_lat[:, None]).T @ P
        op = np.linalg.solve(gram, (P * w_lat[:, None]).T)       # [cnt, nlat]
        i0 = max(m, nmin)
        for trig, is_sin in ((np.cos(m * lam), False), (np.sin(m * lam), True)):
            if m == 0 and is_sin:
                continue
            norm = np.sum(u_lon * trig * trig)
            g = (v * (u_lon * trig)[None, None, :]).sum(axis=2) / norm   # [E, nlat]
            x = g @ op.T
            if is_sin:
                out[:, m - 1, i0:] = x
            else:
                out[:, i0:, m] = x
    return out[0] if single else out


def separable_weights(areas):
    """Factor areas[nlat, nlon] = w_lat[:, None] * u_lon[None, :]; raise if not rank one."""
    areas = np.asarray(areas, dtype=float)
    j0 = int(np.argmax(np.abs(areas).sum(axis=0)))
    i0 = int(np.argmax(np.abs(areas[:, j0])))
    w_lat = areas[:, j0].copy()
    u_lon = areas[i0, :] / areas[i0, j0]
    if not np.allclose(w_lat[:, None] * u_lon[None, :], areas, rtol=1e-12, atol=0):
        raise ValueError('area elements are not separable into latitude and longitude factors')
    return w_lat, u_lon


# --------------------------------------------------------------------------------------
# Covariance propagation
# --------------------------------------------------------------------------------------
def covariance_propagation(sigma, grid, nmin, nmax, kernel='potential', GM=GM_DEFAULT, R=R_DEFAULT, rows=None):
    """RegularGrid.covariance_propagation (grid.py:817-839): sqrt(diag(F Sigma F')), one
    parallel at a time.  rows: optional iterable of parallel indices (bounded CPU samples)."""
    colat, kn = kn_table(kernel, nmax, grid.parallels, grid.a, grid.f, GM, R)
    P = ravel_coefficients(_scale_packed_by_degree(legendre_functions(nmax, colat), kn), nmin, nmax)
    cs = ravel_coefficients(trigonometric_functions(nmax, grid.meridians), nmin, nmax)
    nlat, nlon = grid.shape
    rows = range(nlat) if rows is None else rows
    var = np.zeros((nlat, nlon))
    for k in rows:
        F = cs * P[k:k + 1, :]
        var[k] = np.diag(F @ sigma @ F.T)
    return np.sqrt(var).ravel() if len(rows) == nlat else np.sqrt(var[list(rows)])


def covariance_propagation_points(sigma, lon, lat, nmin, nmax, kernel='potential', GM=GM_DEFAULT, R=R_DEFAULT,
                                  a=A_GRS80, f=F_GRS80):
    """IrregularGrid.covariance_propagation (grid.py:1096-1120), 256-point blocks."""
    lon, lat = np.asarray(lon, dtype=float), np.asarray(lat, dtype=float)
    var = np.zeros(lon.size)
    for i1 in range(0, lon.size, 256):
        i2 = min(i1 + 256, lon.size)
        colat, kn = kn_table(kernel, nmax, lat[i1:i2], a, f, GM, R)
        Y = _scale_packed_by_degree(spherical_harmonics(nmax, colat, lon[i1:i2]), kn)
        F = ravel_coefficients(Y, nmin, nmax)
        var[i1:i2] = np.einsum('pa,ab,pb->p', F, sigma, F, optimize=True)
    return np.sqrt(var)


# --------------------------------------------------------------------------------------
# Order-wise block filter
# --------------------------------------------------------------------------------------
def basin_variances(sigma, grid, masks, nmin, nmax, kernel='potential', GM=GM_DEFAULT, R=R_DEFAULT):
    """Variances of area-weighted basin means (Grid.mean, grid.py:174-201) of a field with coefficient covariance
    sigma: w' A sigma A' w with the synthesis operator A built column by column from unit coefficient sets
    (Grid.synthesis_matrix, grid.py:412-443, degree-wise columns from nmin)."""
    K = (nmax + 1) ** 2 - nmin ** 2
    A = np.empty((grid.shape[0] * grid.shape[1], K))
    for c in range(K):
        unit = np.zeros(K)
        unit[c] = 1.0
        A[:, c] = synthesis(unravel_coefficients(unit, nmin, nmax), grid, kernel, GM, R).ravel()
    masks = np.asarray(masks, dtype=float).reshape(len(masks), -1)
    w = masks * np.asarray(grid.areas).reshape(1, -1)
    w = w / w.sum(axis=1, keepdims=True)
    f = w @ A
    return np.einsum('bi,ij,bj->b', f, sigma, f)


def orderwise_filter(blocks, anm):
    """OrderWiseFilter.filter on a packed array (filter.py:175-191).  blocks[0]: order 0;
    blocks[2m-1], blocks[2m]: cosine / sine block of order m, each [(Nf+1-m), (Nf+1-m)]."""
    anm = np.asarray(anm, dtype=float)
    nmax = anm.shape[0] - 1
    nf = blocks[0].shape[0] - 1
    if nmax > nf:
        raise ValueError('DDK filter only implemented for a maximum degree of {1:d} (max_degree={0:d} supplied).'
                         .format(nmax, nf))
    out = anm.copy()
    out[:, 0] = blocks[0][0:nmax + 1, 0:nmax + 1] @ anm[:, 0]
    for m in range(1, nmax + 1):
        k = nmax + 1 - m
        out[m:, m] = blocks[2 * m - 1][0:k, 0:k] @ anm[m:, m]
        out[m - 1, m:] = blocks[2 * m][0:k, 0:k] @ anm[m - 1, m:]
    out[0:2, 0:2] = anm[0:2, 0:2]
    return out


def orderwise_filter_matrix(blocks, nmin, nmax):
    """OrderWiseFilter.matrix (filter.py:209-222): dense matrix in degree-wise order."""
    K = (nmax + 1) ** 2
    F = np.zeros((K, K))
    idx = np.arange(nmax + 1, dtype=int) ** 2
    F[np.ix_(idx, idx)] = blocks[0][0:nmax + 1, 0:nmax + 1]
    for m in range(1, nmax + 1):
        k = nmax + 1 - m
        F[np.ix_(idx[m:] + 2 * m - 1, idx[m:] + 2 * m - 1)] = blocks[2 * m - 1][0:k, 0:k]
        F[np.ix_(idx[m:] + 2 * m, idx[m:] + 2 * m)] = blocks[2 * m][0:k, 0:k]
    return F[nmin * nmin:, nmin * nmin:]


# --------------------------------------------------------------------------------------
# Synthetic inputs shared by tests and bench
# --------------------------------------------------------------------------------------
def gauss_weights(radius, nmax):
    """Degree weights of the Gauss kernel (kernel.py:468-506): recursion in b = ln2 / (1 - cos(radius/R)),
    stopped at the first weight below 1e-7 (the rest stays zero); the table is built to degree 1024 with
    R = 6378.1366 km, degrees beyond are appended with R = 6378.1363 km."""
    table = max(nmax, 1024)
    if radius <= 0:
        return np.ones(nmax + 1)
    wn = np.zeros(1025)
    b = np.log(2.0) / (1 - np.cos(radius / 6378.1366))
    wn[0] = 1.0
    wn[1] = (1 + np.exp(-2 * b)) / (1 - np.exp(-2 * b)) - 1 / b
    for n in range(2, 1025):
        wn[n] = -(2 * n - 1) / b * wn[n - 1] + wn[n - 2]
        if wn[n] < 1e-7:
            break
    if table > 1024:
        ext = np.empty(table + 1)       # np.empty in the reference: entries after the break are undefined there
        ext[:] = 0.0
        ext[0:1025] = wn
        b = np.log(2.0) / (1 - np.cos(radius / 6378.1363))
        for d in range(1025, table + 1):
            ext[d] = -(2 * d - 1) / b * ext[d - 1] + ext[d - 2]
            if ext[d] < 1e-7:
                break
        wn = ext
    return wn[0:nmax + 1].copy()


def butterworth_weights(order, cutoff_degree, nmax):
    """filter.py:113-118: (1 + (n / n_c)^(2 order))^(-1/2)."""
    # per degree with Python scalars, as the reference does (a vectorised power differs in the last bit)
    return np.array([np.power(1 + (n / cutoff_degree) ** (2 * order), -0.5) for n in range(nmax + 1)])


def degreewise_filter(anm, wn, first_degree=0):
    """Scale every coefficient of degree n >= first_degree by wn[n] (filter.py:66-68 uses first_degree = 2
    for the Gaussian, filter.py:115-116 first_degree = 0 for the Butterworth filter)."""
    out = np.array(anm, dtype=float, copy=True)
    L = out.shape[-1]
    deg = np.maximum(np.arange(L)[:, None], np.arange(L)[None, :])
    w = np.where(deg >= first_degree, np.asarray(wn, dtype=float)[deg], 1.0)
    return out * w


def dense_filter(W, nmin, nmax, anm):
    """GeneralMatrix.filter (filter.py:456-479): x = ravel(anm, nmin, nmax), y = W x, unravel up to
    min(degree of anm, nmax); the top-left nmin x nmin corner of the packed array is copied through."""
    anm = np.asarray(anm, dtype=float)
    max_degree = min(anm.shape[0] - 1, nmax)
    x = ravel_coefficients(anm, nmin, nmax)
    out = unravel_coefficients(W @ x, nmin, max_degree)
    out[0:nmin, 0:nmin] = anm[0:nmin, 0:nmin]
    return out


def vdk_matrix(normals, nmin, nmax, kaula_scale, kaula_power):
    """VDK filter matrix (filter.py:536-546): (N + diag(kaula_scale * n^kaula_power))^-1 N."""
    weights = np.concatenate([np.full(2 * n + 1, kaula_scale * float(n) ** kaula_power) for n in range(nmin, nmax + 1)])
    NP = normals.copy()
    NP.flat[::NP.shape[0] + 1] = np.diag(normals) + weights
    return np.linalg.solve(NP, normals)


def radial_basis_to_coefficients(K, values, lon, lat, nmax, R=R_DEFAULT, a=A_GRS80, f=F_GRS80, blocking_factor=256):
    """RadialBasisFunctions.to_potential_coefficients (reference gravityfield.py:692-727): per block of nodal points the
    spherical harmonics, scaled by the upward continuation (R/r)^(n+1) and the shape factors K, times the point
    values, summed over the points."""
    anm = np.zeros((nmax + 1, nmax + 1))
    values = np.asarray(values, dtype=float)
    for start in range(0, values.size, blocking_factor):
        sl = slice(start, min(start + blocking_factor, values.size))
        colat = colatitude(lat[sl], a, f)
        radius = geocentric_radius(lat[sl], a, f)
        Ynm = spherical_harmonics(nmax, colat, lon[sl])
        kn = np.power((R / radius)[:, np.newaxis], np.arange(nmax + 1, dtype=int) + 1)
        Ynm[:, :, 0] *= kn
        for m in range(1, nmax + 1):
            Ynm[:, m:, m] *= kn[:, m:]
            Ynm[:, m - 1, m:] *= kn[:, m:]
        Ynm *= K[np.newaxis, :, :]
        anm += np.sum(Ynm * values[sl, np.newaxis, np.newaxis], axis=0)
    return anm


def anisotropic_basis_to_grid(K, values, lon, lat, nmin, nmax, grid, kernel='ewh', GM=GM_DEFAULT, R=R_DEFAULT,
                              a=A_GRS80, f=F_GRS80):
    """AnisotropicBasisFunctions.to_grid (reference gravityfield.py:600-642): K @ (Y' v) accumulated over 512-point
    blocks, then per meridian the ravelled, continued Legendre table times that vector."""
    radius = geocentric_radius(grid.parallels, grid.a, grid.f)
    colat = colatitude(grid.parallels, grid.a, grid.f)
    kn = inverse_kernel_coefficients(kernel, 0, nmax, radius, colat)
    continuation = np.power(R / radius[:, np.newaxis], np.arange(0, nmax + 1, dtype=float) + 1) * kn
    Pnm = legendre_functions(nmax, colat)
    for n in range(nmin, nmax + 1):
        rows = np.concatenate((np.full(n + 1, n, dtype=int), np.arange(n, dtype=int)))
        cols = np.concatenate((np.arange(n + 1, dtype=int), np.full(n, n, dtype=int)))
        Pnm[:, rows, cols] *= continuation[:, n:n + 1]
    values = np.asarray(values, dtype=float)
    out = np.zeros((grid.parallels.size, grid.meridians.size))
    for start in range(0, values.size, 512):
        sl = slice(start, min(start + 512, values.size))
        Ynm = ravel_coefficients(spherical_harmonics(nmax, colatitude(lat[sl], a, f), lon[sl]), nmin, nmax).T
        K_tmp = K @ (Ynm @ values[sl])
        for k in range(grid.meridians.size):
            cs = trigonometric_functions(nmax, grid.meridians[k])
            out[:, k] += ravel_coefficients(Pnm * cs, nmin, nmax) @ K_tmp * GM / R
    return out


def synthesis_matrix_points(lon, lat, nmin, nmax, kernel='potential', GM=GM_DEFAULT, R=R_DEFAULT, a=A_GRS80, f=F_GRS80):
    """Grid.synthesis_matrix over IrregularGrid.synthesis_matrix_per_order (reference grid.py:412-443, 957-991):
    [points, K'] in degree-wise order, from the per-order Legendre recursion."""
    colat = colatitude(lat, a, f)
    r = geocentric_radius(lat, a, f)
    kn = inverse_kernel_coefficients(kernel, 0, nmax, r, colat) * \
        np.power((R / r)[:, np.newaxis], np.arange(nmax + 1, dtype=int) + 1) * GM / R
    A = np.empty((lon.size, (nmax + 1) ** 2 - nmin ** 2))
    for m in range(nmax + 1):
        Pnm = (legendre_functions_per_order(nmax, m, colat) * kn[:, m:])[:, max(nmin - m, 0):]
        n = np.arange(max(m, nmin), nmax + 1)
        base = n * n - nmin * nmin
        if m == 0:
            A[:, base] = Pnm
        else:
            A[:, base + 2 * m - 1] = Pnm * np.cos(m * lon[:, np.newaxis])
            A[:, base + 2 * m] = Pnm * np.sin(m * lon[:, np.newaxis])
    return A


def analysis_matrix_points(lon, lat, area, nmin, nmax, kernel='potential', GM=GM_DEFAULT, R=R_DEFAULT, a=A_GRS80, f=F_GRS80):
    """IrregularGrid.analysis_matrix (reference grid.py:993-1017): area-weighted least squares."""
    A = synthesis_matrix_points(lon, lat, nmin, nmax, kernel, GM, R, a, f) * np.sqrt(area)[:, np.newaxis]
    return np.linalg.solve(A.T @ A, A.T * np.sqrt(area))


def anisotropic_kernel_evaluate(K, nmin, nmax, source_lon, source_lat, eval_lon, eval_lat):
    """AnisotropicKernel.evaluate (reference kernel.py:595-620)."""
    v1 = ravel_coefficients(spherical_harmonics(nmax, np.pi * 0.5 - source_lat, source_lon), nmin, nmax) @ K
    Y = spherical_harmonics(nmax, np.pi * 0.5 - eval_lat, eval_lon)
    return np.atleast_1d((v1 @ ravel_coefficients(Y, nmin, nmax).T).squeeze())


def anisotropic_kernel_evaluate_grid(K, nmin, nmax, source_lon, source_lat, eval_lon, eval_lat):
    """AnisotropicKernel.evaluate_grid (reference kernel.py:622-658): [nlat, nlon]."""
    v1 = ravel_coefficients(spherical_harmonics(nmax, np.pi * 0.5 - source_lat, source_lon), nmin, nmax) @ K
    pnm = legendre_functions(nmax, np.pi * 0.5 - eval_lat)
    cs = trigonometric_functions(nmax, eval_lon)
    grid = np.empty((eval_lat.size, eval_lon.size))
    for k in range(eval_lat.size):
        grid[k, :] = (ravel_coefficients(cs * pnm[k], nmin, nmax) @ v1.T).squeeze()
    return grid


def anisotropic_kernel_modulation_transfer(K, nmin, nmax, psi, central_longitude=0, central_latitude=0, azimuth=0):
    """AnisotropicKernel.modulation_transfer (reference kernel.py:654-711)."""
    psi_array = np.atleast_1d(psi)
    theta0 = np.pi * 0.5 - (psi_array + central_latitude)
    x0 = np.vstack((np.sin(theta0) * np.cos(central_longitude), np.sin(theta0) * np.sin(central_longitude), np.cos(theta0)))
    ux, uy, uz = x0[0, 0], x0[1, 0], x0[2, 0]
    ca, sa = np.cos(azimuth), np.sin(azimuth)
    rot = np.array([[ca + ux**2 * (1 - ca), ux * uy * (1 - ca) - uz * sa, ux * uz * (1 - ca) + uy * sa],
                    [uy * ux * (1 - ca) + uz * sa, ca + uy**2 * (1 - ca), uy * uz * (1 - ca) - ux * sa],
                    [uz * ux * (1 - ca) - uy * sa, uz * uy * (1 - ca) + ux * sa, ca + uz**2 * (1 - ca)]])
    x = rot @ x0
    lon = -np.arctan2(x[1, :], x[0, :])
    lat = np.pi * 0.5 - np.arctan2(np.sqrt(x[0, :]**2 + x[1, :]**2), x[2, :])
    kn1 = anisotropic_kernel_evaluate(K, nmin, nmax, lon[0], lat[0], lon, lat).flatten()
    mtf = np.zeros(psi_array.size)
    for k in range(psi_array.size):
        kn2 = anisotropic_kernel_evaluate(K, nmin, nmax, lon[k], lat[k], lon[0:k + 1], lat[0:k + 1]).flatten()
        kn = kn1[0:k + 1] + kn2
        edge_threshold = min(kn[0], kn[-1])
        mtf[k] = 0 if np.min(kn) >= edge_threshold else 1 - kn[int(kn.size // 2)] / np.max(kn)
    return mtf


def filter_kernel_matrix(F, nmin, nmax, kernel='potential'):
    """Matrix of FilterKernel (reference filter.py:588-598) with the reference's broadcasting: both ravelled factor
    arrays have shape [1, K'], so both scale the columns of F."""
    def as_array(kn):
        out = np.zeros((1, nmax + 1, nmax + 1))
        for n in range(nmin, nmax + 1):
            out[:, n, 0:n + 1] = kn[:, n - nmin, np.newaxis]
            out[:, 0:n, n] = kn[:, n - nmin, np.newaxis]
        return out
    r, colat = np.full(1, 6378136.3), np.zeros(1)
    kn = as_array(kernel_coefficients(kernel, nmin, nmax, r, colat))
    kn_prime = as_array(inverse_kernel_coefficients(kernel, nmin, nmax, r, colat))
    K2 = (F * ravel_coefficients(kn, nmin, nmax)[np.newaxis, :]) * ravel_coefficients(kn_prime, nmin, nmax)[:, np.newaxis]
    return K2.reshape(K2.shape[-2:])


def synthetic_coefficients(nmax, epoch=0):
    """Kaula-like random coefficients, seed 1000 + epoch; degrees 0-1 zero."""
    rng = np.random.default_rng(1000 + epoch)
    L = nmax + 1
    anm = rng.standard_normal((L, L))
    deg = np.maximum(np.arange(L)[:, None], np.arange(L)[None, :])   # degree of packed entry [r, c]
    low = np.tril(np.ones((L, L), dtype=bool))
    degree = np.where(low, np.arange(L)[:, None], deg)               # lower: row; upper (S): column
    scale = np.zeros((L, L))
    nz = degree >= 1
    scale[nz] = 1e-5 / degree[nz].astype(float) ** 2
    anm *= scale
    anm[0:2, 0:2] = 0
    return anm


def synthetic_covariance(nmax, rank=64, seed=4):
    K = (nmax + 1) ** 2
    rng = np.random.default_rng(seed)
    Lr = rng.standard_normal((K, rank)) * 1e-11
    return Lr @ Lr.T + np.diag(rng.uniform(0.5, 1.5, K) * 1e-22)


def synthetic_filter_blocks(nf, seed=5):
    rng = np.random.default_rng(seed)
    blocks = []
    for m in range(nf + 1):
        k = nf + 1 - m
        for _ in range(1 if m == 0 else 2):
            blocks.append(0.5 * np.eye(k) + 0.01 * rng.standard_normal((k, k)))
    return blocks


# ---------------------------------------------------------------------------------------------------------------------
# covariance producers: dense restatements of what the reference's block algorithms compute
# ---------------------------------------------------------------------------------------------------------------------
def normal_matrix_cholesky(N):
    """Upper Cholesky factor W, N = W' W: the dense equivalent of BlockMatrix.cholesky (lstsq.py:698-717; block by block
    la.cholesky(lower=False) + solve_triangular(trans='T') + Schur updates give exactly the dense factor)."""
    return np.linalg.cholesky(np.asarray(N, dtype=float)).T


def normal_matrix_inverse(N):
    """N^-1 = W^-1 W^-T: what BlockMatrix.inverse leaves in the upper triangle (lstsq.py:848-882) and
    NormalEquations.compute_covariance(sparse=False) returns (lstsq.py:1026-1042)."""
    W = normal_matrix_cholesky(N)
    Wi = np.linalg.solve(W, np.eye(W.shape[0]))
    return Wi @ Wi.T


def normal_matrix_sparse_inverse(N, index):
    """Sparse inverse (lstsq.py:823-846): the entries of N^-1 on the block sparsity pattern of the Cholesky factor
    (Takahashi recursion); zero elsewhere, upper block triangle only."""
    W = normal_matrix_cholesky(N)
    Z = normal_matrix_inverse(N)
    out = np.zeros_like(Z)
    nb = len(index) - 1
    for i in range(nb):
        for j in range(i, nb):
            blk = W[index[i]:index[i + 1], index[j]:index[j + 1]]
            if np.count_nonzero(blk):
                out[index[i]:index[i + 1], index[j]:index[j + 1]] = Z[index[i]:index[i + 1], index[j]:index[j + 1]]
    return out

"""Isotropic harmonic kernels: the per-degree, per-latitude factors fused into the CUDA
synthesis / analysis / covariance kernels.  Host side (a [nlat, L] table, milliseconds).

Mirrors the names and the ``coefficients`` / ``inverse_coefficients`` interface of
grates.kernel (reference kernel.py:17-67, :85-188, :388-574).
"""
import abc
import os

import numpy as np

from . import utilities

_LOVE_FILE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "love_numbers_ak135_ce.npz")
_love_ce = None

G_NEWTON = 6.673e-11   # value used by the reference kernels (kernel.py:405)


def load_love_numbers(frame='CE'):
    """Load Love numbers (k, h, l) of the ak135 Earth model (Wang et al. 2012), degrees
    0..4096; frame handling as reference data/__init__.py:54-64."""
    global _love_ce
    if _love_ce is None:
        with np.load(_LOVE_FILE) as z:
            _love_ce = np.column_stack((z['h'], z['l'], z['k']))
    hlk = _love_ce.copy()
    frame = frame.lower()
    if frame == 'cm':
        hlk[1] -= 1
    elif frame == 'cf':
        h1, l1 = hlk[1, 0], hlk[1, 1]
        hlk[1] = ((h1 - l1) * 2 / 3, (h1 - l1) * -1 / 3, -1 / 3 * h1 - 2 / 3 * l1)
    elif frame != 'ce':
        raise ValueError('frame of load love numbers must be one of CM, CE, or CF (got <' + frame + '>)')
    return hlk[:, 2], hlk[:, 0], hlk[:, 1]


def _love_slice(values, min_degree, max_degree):
    if max_degree >= values.size:
        raise ValueError('load Love numbers are tabulated up to degree {0:d} (requested {1:d})'
                         .format(values.size - 1, max_degree))
    return values[min_degree:max_degree + 1]


class _NormalField:
    """GRS80 normal gravity field (reference gravityfield.py:1498-1574): zonal harmonics from
    J2 plus the centrifugal term, evaluated on the meridian plane."""
    GM, omega, a, J2 = 3986005e8, 7292115.0e-11, 6378137.0, 108263e-8

    def __init__(self):
        n = np.arange(1, 21, dtype=float)
        e, previous = 0.1, np.inf
        while not np.isclose(e, previous, atol=1e-22, rtol=0):
            previous = e
            e_prime = e / np.sqrt(1 - e ** 2)
            q0 = -2 * np.sum(np.power(-1, n) * n * np.power(e_prime, 2 * n + 1) / ((2 * n + 1) * (2 * n + 3)))
            e = np.sqrt(3 * self.J2 + 4 / 15 * (self.omega ** 2 * self.a ** 3) / self.GM * e ** 3 / (2 * q0))
        self.e2 = e ** 2
        self.flattening = 1 - np.sqrt(1 - self.e2)
        c = [1.0]
        k = 1
        while not np.isclose(c[-1], 0, atol=1e-22, rtol=0):
            sign = 1 if k % 2 == 0 else -1
            c.append(sign * (3 * self.e2 ** k * (1 - k + 5 * k * self.J2 / self.e2)
                             / ((2 * k + 1) * (2 * k + 3) * np.sqrt(4 * k + 1))))
            k += 1
        self.zonal = np.zeros(2 * (len(c) - 1) + 1)
        self.zonal[0::2] = c

    def _zonal_legendre(self, nmax, order, theta):
        """P_n,order(theta) for n = order..nmax via the per-order recursion (utilities.py:62-115)."""
        t = np.cos(theta)
        out = np.empty((t.size, nmax + 1 - order))
        if order == 0:
            out[:, 0] = 1
            out[:, 1] = np.sqrt(3) * t
            for n in range(2, nmax + 1):
                out[:, n] = np.sqrt((2.0 * n - 1.0) * (2.0 * n + 1.0)) / n * t * out[:, n - 1] - \
                    np.sqrt((2.0 * n + 1.0) / (2.0 * n - 3.0)) * (n - 1.0) / n * out[:, n - 2]
            return out
        s = np.sqrt(1 - t ** 2)
        out[:, 0] = np.sqrt(3) * s
        out[:, 1] = np.sqrt(2 * order + 3) * t * out[:, 0]
        for n in range(order + 2, nmax + 1):
            out[:, n - order] = np.sqrt((2 * n - 1) / (n - order) * (2 * n + 1) / (n + order)) * t * out[:, n - 1 - order] - \
                np.sqrt((2 * n + 1) / (2 * n - 3) * (n - order - 1) / (n - order) * (n + order - 1) / (n + order)) * out[:, n - 2 - order]
        return out

    def normal_gravity(self, r, colat):
        r = np.atleast_1d(np.asarray(r, dtype=float))
        colat = np.atleast_1d(np.asarray(colat, dtype=float))
        count = max(r.size, colat.size)
        x = np.zeros(count) + r * np.sin(colat)
        z = np.zeros(count) + r * np.cos(colat)
        # geodetic latitude by Bowring's iteration (reference grid.py:1991-2006)
        e2 = 2 * self.flattening - self.flattening ** 2
        p2 = x ** 2
        k = (1 - e2) ** -1
        h_prev = 0
        for _ in range(10):
            c = np.power(p2 + (1 - e2) * z ** 2 * k ** 2, 1.5) / (self.a * e2)
            k = 1 + (p2 + (1 - e2) * z ** 2 * k ** 3) / (c - p2)
            h = (k ** -1 - (1 - e2)) * np.sqrt(p2 + z ** 2 * k ** 2) / e2
            if np.max(np.abs(h - h_prev)) < 1e-6:
                break
            h_prev = h
        lat = np.arctan2(k * z, np.sqrt(p2))
        # gradient of the zonal field at (x, 0, z) (reference gravityfield.py:437-454, 481)
        radius = np.sqrt(x ** 2 + 0.0 + z ** 2)
        theta = np.arctan2(np.sqrt(x ** 2 + 0.0), z)
        nz = self.zonal.size - 1
        n = np.arange(nz + 1, dtype=float)
        p0 = self._zonal_legendre(nz + 1, 0, theta)
        p1 = self._zonal_legendre(nz + 1, 1, theta)
        f_zero = np.sqrt((n + 1) * (n + 1)) * np.sqrt((2 * n + 1) / (2 * n + 3))
        f_plus = np.sqrt((n + 1) * (n + 2)) * np.sqrt((2 * n + 1) / (2 * n + 3)) * np.sqrt(2)
        upward = np.power(self.a / radius[:, None], n + 2)
        gx = -((p1 * np.ones(count)[:, None]) * f_plus * upward) @ self.zonal
        gz = -2 * (p0[:, 1:] * f_zero * upward) @ self.zonal
        gx = gx * self.GM / (2 * self.a ** 2) + self.omega ** 2 * x
        gz = gz * self.GM / (2 * self.a ** 2)
        return -np.cos(lat) * gx - np.sin(lat) * gz


_GRS80 = None


def normal_gravity(r, colat):
    """GRS80 normal gravity at geocentric (r, colat)."""
    global _GRS80
    if _GRS80 is None:
        _GRS80 = _NormalField()
    return _GRS80.normal_gravity(r, colat)


class IsotropicKernel(metaclass=abc.ABCMeta):
    """Band-limited isotropic kernel; subclasses implement ``_coefficients`` returning the
    [points, max_degree - min_degree + 1] table that maps the functional to potential."""

    @abc.abstractmethod
    def _coefficients(self, min_degree, max_degree, r, colat):
        pass

    def coefficients(self, min_degree, max_degree, r=6378136.3, colat=0):
        r_scalar, c_scalar = np.isscalar(r), np.isscalar(colat)
        if not (r_scalar or isinstance(r, np.ndarray)) or not (c_scalar or isinstance(colat, np.ndarray)):
            raise ValueError('input must be either numeric scalar or ndarrays of matching or broadcastable dimensions')
        if r_scalar and not c_scalar:
            r = np.full(colat.shape, r)
        elif c_scalar and not r_scalar:
            colat = np.full(r.shape, colat)
        elif not r_scalar and not c_scalar and r.shape != colat.shape:
            raise ValueError('shape mismatch in radius and colatitude: objects cannot be broadcast to a single shape')
        return self._coefficients(min_degree, max_degree, r, colat)

    def inverse_coefficients(self, min_degree, max_degree, r=6378136.3, colat=0):
        with np.errstate(divide='ignore'):
            kn = self.coefficients(min_degree, max_degree, r, colat)
            columns = [np.zeros(kn.shape[0]) if np.allclose(kn[:, j], 0.0) else 1.0 / kn[:, j]
                       for j in range(kn.shape[1])]
        return np.vstack(columns).T

    def _as_array(self, kn, min_degree, max_degree, count):
        """Per-degree factors spread over the packed [L, L] layout (reference kernel.py:210-218)."""
        out = np.zeros((count, max_degree + 1, max_degree + 1))
        for n in range(min_degree, max_degree + 1):
            out[:, n, 0:n + 1] = kn[:, n - min_degree, np.newaxis]
            out[:, 0:n, n] = kn[:, n - min_degree, np.newaxis]
        return out

    def coefficient_array(self, min_degree, max_degree, r=6378136.3, colat=0):
        count = max(np.asarray(r).size, np.asarray(colat).size)
        return self._as_array(self.coefficients(min_degree, max_degree, r, colat), min_degree, max_degree, count)

    def inverse_coefficient_array(self, min_degree, max_degree, r=6378136.3, colat=0):
        count = max(np.asarray(r).size, np.asarray(colat).size)
        return self._as_array(self.inverse_coefficients(min_degree, max_degree, r, colat), min_degree, max_degree, count)

    def coefficient(self, n, r=6378136.3, colat=0):
        return self.coefficients(n, n, r, colat).squeeze(axis=1)

    def inverse_coefficient(self, n, r=6378136.3, colat=0):
        kn = self.coefficient(n, r, colat)
        return np.zeros(kn.shape) if np.allclose(kn, 0.0) else 1.0 / kn


class Gauss(IsotropicKernel):
    """Gauss kernel (reference kernel.py:464-506): degree weights by the three-term recursion in
    b = ln 2 / (1 - cos(radius / R)), stopped (remaining weights zero) once a weight falls below 1e-7.
    The reference's own quirk is kept: the table is built with R = 6378.1366 km, an extension beyond
    degree 1024 uses R = 6378.1363 km."""

    def __init__(self, radius):
        if radius < 0:
            raise ValueError('Gaussian filter radius must be positive (got {0:f})'.format(radius))
        nmax = 1024
        self._radius = radius
        if self._radius > 0:
            b = np.log(2.0) / (1 - np.cos(radius / 6378.1366))
            self._wn = np.zeros(nmax + 1)
            self._wn[0] = 1.0
            self._wn[1] = (1 + np.exp(-2 * b)) / (1 - np.exp(-2 * b)) - 1 / b
            for n in range(2, nmax + 1):
                self._wn[n] = -(2 * n - 1) / b * self._wn[n - 1] + self._wn[n - 2]
                if self._wn[n] < 1e-7:
                    break
        else:
            self._wn = np.ones(nmax + 1)

    def _coefficients(self, min_degree, max_degree, r=6378136.3, colat=0):
        local_nmax = self._wn.size - 1
        if max_degree > local_nmax:
            if self._radius > 0:
                wn = self._wn.copy()
                self._wn = np.zeros(max_degree + 1)   # np.empty in the reference: weights behind the cut-off are undefined there
                self._wn[0:local_nmax + 1] = wn
                b = np.log(2.0) / (1 - np.cos(self._radius / 6378.1363))
                for d in range(local_nmax + 1, max_degree + 1):
                    self._wn[d] = -(2 * d - 1) / b * self._wn[d - 1] + self._wn[d - 2]
                    if self._wn[d] < 1e-7:
                        break
            else:
                self._wn = np.ones(max_degree + 1)
        count = max(np.asarray(r).size, np.asarray(colat).size)
        return np.tile(self._wn[min_degree:max_degree + 1], (count, 1))


def _degrees(min_degree, max_degree):
    return np.arange(min_degree, max_degree + 1, dtype=float)


class WaterHeight(IsotropicKernel):
    """Equivalent water height in metres (reference kernel.py:388-406)."""

    def __init__(self, rho=1025):
        self._rho = rho
        self._k = load_love_numbers()[0]

    def _coefficients(self, min_degree, max_degree, r=6378136.3, colat=0):
        kn = (4 * np.pi * G_NEWTON * self._rho) * (1 + _love_slice(self._k, min_degree, max_degree)) \
            / (2 * _degrees(min_degree, max_degree) + 1)
        return (kn[:, None] * r).T


class OceanBottomPressure(IsotropicKernel):
    """Ocean bottom pressure in Pascal (reference kernel.py:409-421)."""

    def __init__(self):
        self._k = load_love_numbers()[0]

    def _coefficients(self, min_degree, max_degree, r=6378136.3, colat=0):
        kn = (4 * np.pi * G_NEWTON) * (1 + _love_slice(self._k, min_degree, max_degree)) \
            / (2 * _degrees(min_degree, max_degree) + 1)
        return (kn[:, None] * (r / normal_gravity(r, colat))).T


class SurfaceDensity(IsotropicKernel):
    """Surface density (reference kernel.py:424-435)."""

    def __init__(self):
        self._k = load_love_numbers()[0]

    def _coefficients(self, min_degree, max_degree, r=6378136.3, colat=0):
        kn = (4 * np.pi * G_NEWTON) * (1 + _love_slice(self._k, min_degree, max_degree)) \
            / (2 * _degrees(min_degree, max_degree) + 1)
        return (kn[:, None] * r).T


class Potential(IsotropicKernel):
    """Disturbing potential (reference kernel.py:438-449)."""

    def _coefficients(self, min_degree, max_degree, r=6378136.3, colat=0):
        return np.ones((max(np.asarray(r).size, np.asarray(colat).size), max_degree + 1 - min_degree))


class GravityAnomaly(IsotropicKernel):
    """Gravity anomalies (reference kernel.py:452-461); degree 1 is singular and mapped to zero."""

    def _coefficients(self, min_degree, max_degree, r=6378136.3, colat=0):
        kn = np.array([1 / (n - 1) if n != 1 else 0.0 for n in _degrees(min_degree, max_degree)])
        return (kn[:, None] * r).T


class GeoidHeight(IsotropicKernel):
    """Geoid height: potential over normal gravity (reference kernel.py:509-518)."""

    def _coefficients(self, min_degree, max_degree, r=6378136.3, colat=0):
        return np.tile(normal_gravity(r, colat)[:, None], (1, max_degree + 1 - min_degree))


class VerticalDeformation(IsotropicKernel):
    """Elastic vertical deformation (reference kernel.py:542-559)."""

    def __init__(self, frame='CE'):
        k, h, _ = load_love_numbers(frame)
        with np.errstate(divide='ignore', invalid='ignore'):
            self._ratio = h / (1 + k)

    def _coefficients(self, min_degree, max_degree, r=6378136.3, colat=0):
        with np.errstate(divide='ignore'):
            return normal_gravity(r, colat)[:, None] / _love_slice(self._ratio, min_degree, max_degree)


class Uplift(IsotropicKernel):
    """Approximate uplift after Wahr et al. 2000 (reference kernel.py:562-574)."""

    def _coefficients(self, min_degree, max_degree, r=6378136.3, colat=0):
        return 2 * normal_gravity(r, colat)[:, None] / (2 * _degrees(min_degree, max_degree) + 1)


_KERNELS = {
    'ewh': WaterHeight, 'water_height': WaterHeight,
    'obp': OceanBottomPressure, 'ocean_bottom_pressure': OceanBottomPressure,
    'potential': Potential,
    'geoid': GeoidHeight, 'geoid_height': GeoidHeight,
    'surface_density': SurfaceDensity,
    'anomaly': GravityAnomaly, 'gravity_anomaly': GravityAnomaly,
    'deformation': VerticalDeformation, 'vertical_derformation': VerticalDeformation,
    'uplift': Uplift,
}


class AnisotropicKernel:
    """Possibly anisotropic kernel in the space domain, given as a matrix K between degree-wise ordered spherical
    harmonics (reference kernel.py:576-658).  ``evaluate`` / ``evaluate_grid`` give the kernel of one source point as the
    reference does; the ``*_batch`` forms take many source points at once -- footprints of a filter over a region are one
    design-matrix kernel, one dense GEMM and one batched synthesis on the device:

        Y(sources) [E, K']  --  V = Y K  --  values = synthesis(V) on the evaluation points / grid (unit sphere)
    """

    def __init__(self, K, min_degree, max_degree):
        self._matrix = np.array(K, dtype=float, copy=True)
        self._min_degree = min_degree
        self._max_degree = max_degree
        self._operator = None
        self._plans = {}

    def _source_vectors(self, source_longitude, source_latitude):
        """V = ravel(Y(source)) @ K for every source point, as packed coefficients [E, L, L] on the device
        (reference kernel.py:615-616)."""
        import torch
        from . import plan as _plan, utilities
        from .filter import GeneralMatrix
        lon = np.atleast_1d(np.asarray(source_longitude, dtype=float))
        lat = np.atleast_1d(np.asarray(source_latitude, dtype=float))
        L = self._max_degree + 1
        src = _plan.PointsPlan(lon, lat, 1.0, 0.0, self._max_degree, degree_factors=np.ones((lon.size, L)),
                               colatitude=np.pi * 0.5 - lat)
        Y = src.synthesis_matrix(0)                                          # [E, L^2], degree-wise order
        rows, cols = utilities.degreewise_index(0, self._max_degree)
        packed = torch.zeros((lon.size, L, L), dtype=torch.float64, device=Y.device)
        packed[:, torch.as_tensor(rows).to(Y.device), torch.as_tensor(cols).to(Y.device)] = Y
        if self._operator is None:                                           # v K == K' v
            self._operator = GeneralMatrix(np.ascontiguousarray(self._matrix.T), self._min_degree, self._max_degree)
        V = self._operator.filter_batch(packed)
        V[:, 0:self._min_degree, 0:self._min_degree] = 0.0                   # degrees below the band do not enter
        return V

    def evaluate_batch(self, source_longitude, source_latitude, eval_longitude, eval_latitude):
        """Kernels of E source points at m evaluation points: CUDA tensor [E, m]."""
        from . import plan as _plan
        lon = np.atleast_1d(np.asarray(eval_longitude, dtype=float))
        lat = np.atleast_1d(np.asarray(eval_latitude, dtype=float))
        pts = _plan.PointsPlan(lon, lat, 1.0, 0.0, self._max_degree,
                               degree_factors=np.ones((lon.size, self._max_degree + 1)), colatitude=np.pi * 0.5 - lat)
        return pts.synthesis(self._source_vectors(source_longitude, source_latitude))

    def evaluate_grid_batch(self, source_longitude, source_latitude, eval_longitude, eval_latitude):
        """Kernels of E source points on the grid of eval_latitude x eval_longitude: CUDA tensor [E, nlat, nlon]."""
        from . import plan as _plan
        lon = np.atleast_1d(np.asarray(eval_longitude, dtype=float))
        lat = np.atleast_1d(np.asarray(eval_latitude, dtype=float))
        key = (lon.tobytes(), lat.tobytes())
        if key not in self._plans:
            self._plans.clear()
            self._plans[key] = _plan.SHPlan(lon, lat, 1.0, 0.0, self._max_degree,
                                            degree_factors=np.ones((lat.size, self._max_degree + 1)),
                                            colatitude=np.pi * 0.5 - lat)
        return self._plans[key].synthesis(self._source_vectors(source_longitude, source_latitude))

    def modulation_transfer(self, psi, central_longitude=0, central_latitude=0, azimuth=0):
        """Modulation transfer function of the kernel along a great circle (reference kernel.py:654-711): two kernels
        are shifted apart by psi[k]; all kernel evaluations of the reference's loop are one batched call here."""
        psi_array = np.atleast_1d(psi)
        theta0 = np.pi * 0.5 - (psi_array + central_latitude)
        x0 = np.vstack((np.sin(theta0) * np.cos(central_longitude), np.sin(theta0) * np.sin(central_longitude),
                        np.cos(theta0)))
        ux, uy, uz = x0[0, 0], x0[1, 0], x0[2, 0]
        ca, sa = np.cos(azimuth), np.sin(azimuth)
        rotation_matrix = np.array([[ca + ux**2 * (1 - ca), ux * uy * (1 - ca) - uz * sa, ux * uz * (1 - ca) + uy * sa],
                                    [uy * ux * (1 - ca) + uz * sa, ca + uy**2 * (1 - ca), uy * uz * (1 - ca) - ux * sa],
                                    [uz * ux * (1 - ca) - uy * sa, uz * uy * (1 - ca) + ux * sa, ca + uz**2 * (1 - ca)]])
        x = rotation_matrix @ x0
        lon = -np.arctan2(x[1, :], x[0, :])
        lat = np.pi * 0.5 - np.arctan2(np.sqrt(x[0, :]**2 + x[1, :]**2), x[2, :])
        G = self.evaluate_batch(lon, lat, lon, lat).cpu().numpy()        # G[k, j]: kernel of source k at point j
        kn1 = G[0]
        mtf = np.zeros(psi_array.size)
        for k in range(0, psi_array.size):
            kn = kn1[0:k + 1] + G[k, 0:k + 1]
            edge_threshold = min(kn[0], kn[-1])
            mtf[k] = 0 if np.min(kn) >= edge_threshold else 1 - kn[int(kn.size // 2)] / np.max(kn)
        return mtf

    def evaluate(self, source_longitude, source_latitude, eval_longitude, eval_latitude):
        """Kernel of one source point at the evaluation points, ndarray(m,) (reference kernel.py:595-620)."""
        return self.evaluate_batch(source_longitude, source_latitude, eval_longitude, eval_latitude)[0].cpu().numpy()

    def evaluate_grid(self, source_longitude, source_latitude, eval_longitude, eval_latitude):
        """Kernel of one source point on a longitude / latitude grid, ndarray(nlat, nlon) (reference kernel.py:622-658)."""
        return self.evaluate_grid_batch(source_longitude, source_latitude, eval_longitude, eval_latitude)[0].cpu().numpy()


def get_kernel(kernel_name):
    """Kernel instance for a name (same names and spelling as reference kernel.py:40-62)."""
    try:
        return _KERNELS[kernel_name.lower()]()
    except KeyError:
        raise ValueError("Unrecognized kernel '{0:s}'.".format(kernel_name)) from None


def degree_factors(kernel_name, max_degree, parallels, a, f, GM, R):
    """kn[i, n] = 1/k_n(r_i, theta_i) * (R / r_i)^(n+1) * GM / R  and the colatitudes, evaluated
    in the reference's operation order (gravityfield.py:353-356 == grid.py:653-657, :819-823)."""
    colat = utilities.colatitude(parallels, a, f)
    radius = utilities.geocentric_radius(parallels, a, f)
    kn = get_kernel(kernel_name).inverse_coefficients(0, max_degree, radius, colat) * \
        np.power((R / radius)[:, None], np.arange(max_degree + 1, dtype=int) + 1) * GM / R
    return colat, np.ascontiguousarray(kn)

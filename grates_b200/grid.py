"""Regular grids (parallels x meridians on the ellipsoid) with GPU-backed analysis and
covariance propagation.  Mirrors RegularGrid / GeographicGrid / GaussGrid of grates.grid
(reference grid.py:510-839, :1123-1204) for the spherical-harmonic hot path: same
constructor arguments, attributes (``parallels``, ``meridians``, ``value_array``, ``values``,
``area``, ``epoch``), return types and error behaviour.  Irregular point sets are not part of
this path and raise instead of falling back to the CPU.
"""
import numpy as np
import torch

from . import plan as _plan, utilities

GM_DEFAULT = 3.9860044150e+14
R_DEFAULT = 6.3781363000e+06


class RegularGrid:
    """Global regular point distribution defined by meridians [rad] and parallels [rad]
    (north to south).  area_elements: optional [nlat, nlon] weights (reference grid.py:529-542)."""

    def __init__(self, meridians, parallels, area_elements=None, a=6378137.0, f=298.2572221010 ** -1):
        self.parallels = np.asarray(parallels, dtype=float)
        self.meridians = np.asarray(meridians, dtype=float)
        self._a, self._f = a, f
        if area_elements is None:
            lon_edges = np.concatenate(([-np.pi], self.meridians[0:-1] + 0.5 * np.diff(self.meridians), [np.pi]))
            lat_edges = np.concatenate(([0.5 * np.pi], self.parallels[0:-1] + 0.5 * np.diff(self.parallels), [-0.5 * np.pi]))
            area_elements = 2.0 * (np.sin(np.abs(np.diff(lat_edges)) * 0.5) * np.cos(self.parallels))[:, None] * np.diff(lon_edges)
        self._areas = area_elements
        self.value_array = None
        self.epoch = None

    # -- container behaviour (reference grid.py:544-625) ----------------------------------
    def copy(self):
        grid = RegularGrid(self.meridians.copy(), self.parallels.copy(), self._areas.copy(), self._a, self._f)
        if self.value_array is not None:
            grid.values = self.values.copy()
        grid.epoch = self.epoch
        return grid

    @property
    def semimajor_axis(self):
        return self._a

    @property
    def flattening(self):
        return self._f

    @property
    def point_count(self):
        return self.parallels.size * self.meridians.size

    @property
    def size(self):
        return self.point_count

    @property
    def longitude(self):
        return np.tile(self.meridians, self.parallels.size)

    @property
    def latitude(self):
        return np.repeat(self.parallels, self.meridians.size)

    @property
    def area(self):
        return self._areas.ravel()

    @property
    def values(self):
        if self.value_array is not None:
            return self.value_array.ravel()

    @values.setter
    def values(self, val):
        if val is None:
            self.value_array = None
        elif isinstance(val, np.ndarray):
            if val.ndim > 1:
                raise ValueError("unable to assign values of dimension {0:d} to grid".format(val.ndim))
            if val.size != self.point_count:
                raise ValueError("unable to assign values of size {0:d} to grid with {1:d} points".format(val.size, self.point_count))
            self.value_array = np.reshape(val, (self.parallels.size, self.meridians.size))
        else:
            raise ValueError("grid values must be either None or " + str(np.ndarray))

    def is_compatible(self, other):
        if self.point_count == other.point_count:
            return np.allclose(self.longitude, other.longitude) and np.allclose(self.latitude, other.latitude)
        return False

    # -- weighted statistics (reference grid.py:174-260) ----------------------------------
    def _masked(self, mask):
        mask = np.ones(self.point_count, dtype=bool) if mask is None else mask
        return self.area[mask], self.values[mask]

    def mean(self, mask=None):
        w, v = self._masked(mask)
        return np.sum(w * v) / np.sum(w)

    def rms(self, mask=None):
        w, v = self._masked(mask)
        return np.sqrt(np.sum(w * v ** 2) / np.sum(w))

    def std(self, mask=None):
        w, v = self._masked(mask)
        v = v - self.mean(mask)
        return np.sqrt(np.sum(w * v ** 2) / np.sum(w))

    # -- hot path ------------------------------------------------------------------------
    def to_potential_coefficients(self, min_degree, max_degree, kernel='potential', GM=GM_DEFAULT, R=R_DEFAULT):
        """Spherical-harmonic analysis of the grid values on the GPU; same estimator as the
        reference's order-by-order area-weighted least squares (grid.py:752-790).

        Returns a PotentialCoefficients whose degrees below ``min_degree`` are zero.
        Raises ValueError if the grid holds no values.
        """
        from .gravityfield import PotentialCoefficients
        if self.values is None:
            raise ValueError('grid has no values to propagate to potential coefficients')
        anm = analysis_batch(self.value_array[None], self, min_degree, max_degree, kernel, GM, R)[0]
        coeffs = PotentialCoefficients(GM, R)
        coeffs.anm = anm
        return coeffs

    def synthesis_matrix(self, min_degree, max_degree, kernel='potential', GM=GM_DEFAULT, R=R_DEFAULT):
        """Dense operator A [points, K'] mapping coefficients in degree-wise order to grid values
        (reference grid.py:412-443), generated on the GPU."""
        p = _plan.get_plan(self, max_degree, kernel, GM, R)
        return p.synthesis_matrix(min_degree).cpu().numpy()

    def synthesis_matrix_per_order(self, m, min_degree, max_degree, kernel, GM, R):
        """Columns of the synthesis operator for one order: [points, n_m] for m = 0, else the tuple
        (cosine part, sine part) (reference grid.py:627-663)."""
        if m > max_degree:
            raise ValueError('order exceeds maximum degree ({0:d} vs. {1:d})'.format(m, max_degree))
        A = self.synthesis_matrix(min_degree, max_degree, kernel, GM, R)
        n = np.arange(max(m, min_degree), max_degree + 1)
        base = n * n - min_degree * min_degree
        if m == 0:
            return np.ascontiguousarray(A[:, base])
        return np.ascontiguousarray(A[:, base + 2 * m - 1]), np.ascontiguousarray(A[:, base + 2 * m])

    def analysis_matrix(self, min_degree, max_degree, kernel, GM=GM_DEFAULT, R=R_DEFAULT):
        """Dense analysis operator [K', points], rows in degree-wise order (reference grid.py:698-730)."""
        p = _plan.get_plan(self, max_degree, kernel, GM, R)
        p.set_analysis(min_degree, self.area.reshape(p.nlat, p.nlon))
        return p.analysis_matrix().cpu().numpy()

    def covariance_propagation(self, covariance_matrix, min_degree, max_degree, kernel='potential', GM=GM_DEFAULT, R=R_DEFAULT,
                               spatial_filter=None):
        """Propagate a degree-wise ordered coefficient covariance matrix to grid-point standard
        deviations, sqrt(diag(A Sigma A')), on the GPU.  Like the reference (grid.py:792-839)
        this also stores the result in ``self.values`` and returns a 1-d array.
        spatial_filter (extension): propagate F Sigma F' for an OrderWiseFilter / Gaussian / Butterworth F
        without forming the filtered matrix."""
        p = _plan.get_plan(self, max_degree, kernel, GM, R)
        sigma = _device_matrix(covariance_matrix, p.device)
        std = p.covariance_propagation(sigma, min_degree, spatial_filter=spatial_filter).cpu().numpy().ravel()
        self.values = std
        return std.copy()


class IrregularGrid:
    """Arbitrary point distribution on the ellipsoid (reference grid.py:842-1120): longitude /
    latitude pairs in radians, optional per-point area elements (default 4 pi / point count)."""

    def __init__(self, longitude, latitude, area_element=None, a=6378137.0, f=298.2572221010 ** -1):
        self._lons = np.asarray(longitude, dtype=float)
        self._lats = np.asarray(latitude, dtype=float)
        self._areas = np.full(self._lons.size, 4 * np.pi / self._lons.size) if area_element is None else area_element
        self._a, self._f = a, f
        self._values = None
        self.epoch = None

    def copy(self):
        grid = IrregularGrid(self._lons.copy(), self._lats.copy(), self._areas.copy(), self._a, self._f)
        if self._values is not None:
            grid.values = self._values.copy()
        grid.epoch = self.epoch
        return grid

    semimajor_axis = property(lambda self: self._a)
    flattening = property(lambda self: self._f)
    longitude = property(lambda self: self._lons)
    latitude = property(lambda self: self._lats)
    area = property(lambda self: self._areas)
    point_count = property(lambda self: self._lons.size)
    size = property(lambda self: self._lons.size)

    @property
    def values(self):
        return self._values

    @values.setter
    def values(self, val):
        if val is None:
            self._values = None
        elif isinstance(val, np.ndarray):
            if val.ndim > 1:
                raise ValueError("unable to assign values of dimension {0:d} to grid".format(val.ndim))
            if val.size != self.point_count:
                raise ValueError("unable to assign values of size {0:d} to grid with {1:d} points".format(val.size, self.point_count))
            self._values = val
        else:
            raise ValueError("grid values must be either None or " + str(np.ndarray))

    def synthesis_matrix(self, min_degree, max_degree, kernel='potential', GM=GM_DEFAULT, R=R_DEFAULT):
        """Dense operator A [points, K'] mapping coefficients in degree-wise order to point values
        (reference grid.py:412-443 over :957-991), generated on the GPU."""
        p = _plan.get_points_plan(self, max_degree, kernel, GM, R)
        return p.synthesis_matrix(min_degree).cpu().numpy()

    def synthesis_matrix_per_order(self, m, min_degree, max_degree, kernel, GM, R):
        """Columns of the synthesis operator for one order: [points, n_m] for m = 0, else the tuple
        (cosine part, sine part) (reference grid.py:957-991)."""
        if m > max_degree:
            raise ValueError('order exceeds maximum degree ({0:d} vs. {1:d})'.format(m, max_degree))
        A = self.synthesis_matrix(min_degree, max_degree, kernel, GM, R)
        n = np.arange(max(m, min_degree), max_degree + 1)
        base = n * n - min_degree * min_degree
        if m == 0:
            return np.ascontiguousarray(A[:, base])
        return np.ascontiguousarray(A[:, base + 2 * m - 1]), np.ascontiguousarray(A[:, base + 2 * m])

    def _analysis_operator(self, min_degree, max_degree, kernel, GM, R):
        """Area-weighted least-squares operator solve(A'WA, A'W) on the device (reference grid.py:993-1017).  The design
        matrix comes from the point-set kernels; the normal matrix A'WA, its Cholesky factor W'W and the two triangular
        solves run on this library's own FP64 tensor-core kernels (gb_dgemm / gb_dpotrf_upper / gb_dtrsm_upper, gb_linalg.cu)
        where the reference calls BLAS / LAPACK through numpy."""
        from .lstsq import _Ops
        p = _plan.get_points_plan(self, max_degree, kernel, GM, R)
        A = p.synthesis_matrix(min_degree)
        sw = torch.as_tensor(np.sqrt(np.asarray(self.area, dtype=float))).to(A.device)
        A = (A * sw[:, None]).contiguous()
        ops = _Ops(A.device.index)
        normals = ops.matmul(A, A, trans_a=True)
        ops.cholesky_upper(normals)
        rhs = (A * sw[:, None]).T.contiguous()                  # A'W (A already carries sqrt(w))
        ops.solve_triangular(normals, rhs, trans=True)
        return ops.solve_triangular(normals, rhs, trans=False)

    def analysis_matrix(self, min_degree, max_degree, kernel='potential', GM=GM_DEFAULT, R=R_DEFAULT):
        """Dense analysis operator [K', points] (reference grid.py:993-1017)."""
        return self._analysis_operator(min_degree, max_degree, kernel, GM, R).cpu().numpy()

    def to_potential_coefficients(self, min_degree, max_degree, kernel='potential', GM=GM_DEFAULT, R=R_DEFAULT):
        """Least-squares spherical-harmonic analysis of the point values (reference grid.py:477-507)."""
        from .gravityfield import PotentialCoefficients
        if self.values is None:
            raise ValueError('grid has no values to propagate to potential coefficients')
        F = self._analysis_operator(min_degree, max_degree, kernel, GM, R)
        from .lstsq import _Ops
        v = torch.as_tensor(np.ascontiguousarray(self.values, dtype=float)).to(F.device).reshape(-1, 1)
        x = _Ops(F.device.index).matmul(F.contiguous(), v.contiguous()).reshape(-1)
        coeffs = PotentialCoefficients(GM, R)
        coeffs.anm = utilities.unravel_coefficients(x.cpu().numpy(), min_degree, max_degree)
        return coeffs

    def covariance_propagation(self, covariance_matrix, min_degree, max_degree, kernel='potential', GM=GM_DEFAULT, R=R_DEFAULT):
        """sqrt(diag(F Sigma F')) at every point on the GPU (reference grid.py:1071-1120: 256-point
        blocks of dense products); stores the standard deviations in ``self.values`` as the reference does."""
        p = _plan.get_points_plan(self, max_degree, kernel, GM, R)
        sigma = _device_matrix(covariance_matrix, p.device)
        std = p.covariance_propagation(sigma, min_degree).cpu().numpy()
        self.values = std
        return std.copy()


class GeographicGrid(RegularGrid):
    """Equi-angular geographic grid of pixel centres (reference grid.py:1141-1162)."""

    def __init__(self, dlon=0.5, dlat=0.5, a=6378137.0, f=298.2572221010 ** -1):
        self._dlon, self._dlat = dlon, dlat
        nlons, nlats = 360 / dlon, 180 / dlat
        meridians = np.linspace(-np.pi + dlon / 180 * np.pi * 0.5, np.pi - dlon / 180 * np.pi * 0.5, int(nlons))
        parallels = -np.linspace(-np.pi * 0.5 + dlat / 180 * np.pi * 0.5, np.pi * 0.5 - dlat / 180 * np.pi * 0.5, int(nlats))
        areas = np.tile(2.0 * dlon / 180 * np.pi * np.sin(dlat * 0.5 / 180 * np.pi) * np.cos(parallels)[:, None], (1, meridians.size))
        super().__init__(meridians, parallels, areas, a, f)

    def copy(self):
        grid = GeographicGrid(self._dlon, self._dlat, self.semimajor_axis, self.flattening)
        if self.values is not None:
            grid.values = self.values.copy()
        grid.epoch = self.epoch
        return grid


class GaussGrid(RegularGrid):
    """Gaussian grid: parallels at the Legendre roots (reference grid.py:1181-1204)."""

    def __init__(self, parallel_count, a=6378137.0, f=298.2572221010 ** -1):
        from scipy.special import roots_legendre
        zeros, weights, _ = roots_legendre(parallel_count, mu=True)
        dlon = np.pi / parallel_count
        meridians = np.linspace(-np.pi + dlon * 0.5, np.pi - dlon * 0.5, 2 * parallel_count)
        cosine_theta = -zeros
        sine_theta = np.sqrt(1 - cosine_theta ** 2)
        parallels = np.arctan2(cosine_theta, (1 - f) ** 2 * sine_theta)
        areas = np.tile(dlon * weights[:, None], (1, meridians.size))
        super().__init__(meridians, parallels, areas, a, f)

    def copy(self):
        grid = GaussGrid(self.parallels.size, self.semimajor_axis, self.flattening)
        if self.value_array is not None:
            grid.values = self.values.copy()
        grid.epoch = self.epoch
        return grid


def _device_matrix(matrix, device):
    """A covariance matrix on the plan's device: CUDA tensors pass through (no 0.7 GB host round trip at degree 96)."""
    if isinstance(matrix, torch.Tensor):
        return matrix.to(device=torch.device("cuda", device), dtype=torch.float64)
    return torch.as_tensor(np.ascontiguousarray(matrix, dtype=np.float64)).to(torch.device("cuda", device))


def covariance_from_normals(normal_equation_matrix, device_output=True):
    """Covariance matrix Sigma = N^-1 of a dense, symmetric positive definite normal-equation matrix, kept on the device
    for ``covariance_propagation`` / ``basin_variances`` (the dense case of NormalEquations.compute_covariance, reference
    lstsq.py:1026-1043: Cholesky factor W'W, then W^-1 W^-T) on this library's own dense kernels (gb_linalg.cu); the
    blocked / sparse case is ``lstsq.NormalEquations.compute_covariance``."""
    from .lstsq import _Ops
    dev = torch.device("cuda", _plan._current_device(None))
    n = normal_equation_matrix if isinstance(normal_equation_matrix, torch.Tensor) else \
        torch.as_tensor(np.ascontiguousarray(normal_equation_matrix, dtype=np.float64))
    w = n.to(device=dev, dtype=torch.float64).clone().contiguous()
    if w.dim() != 2 or w.shape[0] != w.shape[1]:
        raise ValueError("expected a square normal-equation matrix")
    ops = _Ops(dev.index)
    ops.cholesky_upper(w)
    sigma = ops.inv_gram(torch.triu(w))
    return sigma if device_output else sigma.cpu().numpy()


def basin_variances(covariance_matrix, grid, masks, min_degree, max_degree, kernel='potential', GM=GM_DEFAULT,
                    R=R_DEFAULT, take_sqrt=True):
    """Standard deviations of the area-weighted basin means of a field with coefficient covariance Sigma:
    sqrt(w_b' A Sigma A' w_b), w_b = area * mask_b / sum(area * mask_b), A the synthesis operator of the grid
    (reference: Grid.mean, grid.py:174-201, of fields synthesised with grid.py:412-443).  The adjoint synthesis
    A' w_b and the quadratic forms run on the GPU; A is never formed.
    masks: boolean (or weight) array [B, points] or [B, nlat, nlon]; returns [B]."""
    import ctypes
    from . import _lib
    p = _plan.get_plan(grid, max_degree, kernel, GM, R, variant='adjoint')
    p.set_adjoint(min_degree)
    masks = np.asarray(masks)
    masks = masks.reshape(masks.shape[0], -1) if masks.ndim > 1 else masks.reshape(1, -1)
    if masks.shape[1] != grid.point_count:
        raise ValueError("masks must have one entry per grid point")
    w = masks.astype(float) * np.asarray(grid.area, dtype=float).reshape(1, -1)
    norm = w.sum(axis=1, keepdims=True)
    if np.any(norm == 0):
        raise ValueError("empty basin mask")
    w = np.ascontiguousarray(w / norm).reshape(masks.shape[0], p.nlat, p.nlon)
    dev = torch.device("cuda", p.device)
    functionals = p.analysis(torch.as_tensor(w).to(dev))                      # [B, L, L] packed A' w_b
    sigma = covariance_matrix if isinstance(covariance_matrix, torch.Tensor) else \
        torch.as_tensor(np.ascontiguousarray(covariance_matrix, dtype=np.float64)).to(dev)
    k = (max_degree + 1) ** 2 - min_degree ** 2
    if tuple(sigma.shape) != (k, k):
        raise ValueError("covariance matrix must have shape [{0}, {0}] (got {1})".format(k, tuple(sigma.shape)))
    sigma = sigma.contiguous()
    B = masks.shape[0]
    vec = torch.empty((B, k), dtype=torch.float64, device=dev)
    var = torch.empty(B, dtype=torch.float64, device=dev)
    lib, st = _lib.load(), _plan._stream_handle(p.device)
    _lib.check(lib.gb_ravel_coefficients(ctypes.c_void_p(functionals.data_ptr()), B, max_degree, min_degree,
                                         ctypes.c_void_p(vec.data_ptr()), p.device, st))
    _lib.check(lib.gb_quadratic_forms(ctypes.c_void_p(sigma.data_ptr()), k, ctypes.c_void_p(vec.data_ptr()), B,
                                      ctypes.c_void_p(var.data_ptr()), p.device, st))
    var = var.cpu().numpy()
    return np.sqrt(var) if take_sqrt else var


def analysis_batch(values, grid, min_degree, max_degree, kernel='potential', GM=GM_DEFAULT, R=R_DEFAULT, device_output=False):
    """Batched analysis: values [E, nlat, nlon] (numpy or CUDA tensor) -> packed anm [E, L, L].
    New entry point (the reference loops over epochs and rebuilds its operators every call)."""
    if min_degree < 0 or max_degree < min_degree:
        raise ValueError("invalid degree range [{0}, {1}]".format(min_degree, max_degree))
    p = _plan.get_plan(grid, max_degree, kernel, GM, R)
    p.set_analysis(min_degree, grid.area.reshape(p.nlat, p.nlon))
    if isinstance(values, torch.Tensor):
        out = p.analysis(values)
        return out if device_output else out.cpu().numpy()
    out = p.analysis_host(np.asarray(values, dtype=float))
    if device_output:
        return torch.as_tensor(out).to(torch.device("cuda", p.device))
    return out

"""Spatial filters on the GPU, batched over epochs.  Mirrors grates.filter: Gaussian and Butterworth
(degree-wise weights, reference filter.py:31-130) and OrderWiseFilter (DDK-style blocks, filter.py:133-222)."""
import abc
import ctypes

import numpy as np
import torch

from . import _lib, kernel as _kernel, plan as _plan, utilities
from .gravityfield import PotentialCoefficients


class SpatialFilter(metaclass=abc.ABCMeta):
    """Interface of the reference (filter.py:15-28): ``filter`` returns a filtered copy of a
    PotentialCoefficients instance, ``matrix`` the dense filter matrix in degree-wise order."""

    @abc.abstractmethod
    def filter(self, gravityfield):
        pass

    @abc.abstractmethod
    def matrix(self, min_degree, max_degree):
        pass


class _DegreeWiseFilter(SpatialFilter):
    """Filters that scale every coefficient of degree n by a weight w_n.  The weights are host numpy
    (a few hundred doubles); batches are scaled on the GPU (gb_scale_by_degree), and a batched
    synthesis can take them as ``degree_weights`` so that the scaled coefficients are never written."""

    _first_degree = 0      # degrees below are left untouched by ``filter``

    @abc.abstractmethod
    def _weights(self, max_degree):
        pass

    def degree_weights(self, max_degree):
        """w[0..max_degree] as applied by ``filter`` (1 below the first filtered degree)."""
        w = np.array(self._weights(max_degree), dtype=float)
        w[0:min(self._first_degree, max_degree + 1)] = 1.0
        return w

    def filter_batch(self, anm, out=None):
        """anm: [E, L, L] packed coefficients (numpy or CUDA tensor) -> filtered copy, same type."""
        on_host = not isinstance(anm, torch.Tensor)
        dev = _plan._current_device(None if on_host else anm.device)
        x = torch.as_tensor(np.ascontiguousarray(anm, dtype=float)).to(torch.device("cuda", dev)) if on_host else anm.contiguous()
        if x.dim() != 3 or x.shape[1] != x.shape[2] or x.dtype != torch.float64:
            raise ValueError("coefficients must be a float64 array of shape [epochs, L, L]")
        nmax = x.shape[-1] - 1
        w = torch.as_tensor(self.degree_weights(nmax)).to(x.device)
        y = torch.empty_like(x) if out is None else _plan._check_out(out, x.shape, dev)
        lib = _lib.load()
        _lib.check(lib.gb_scale_by_degree(ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(w.data_ptr()), x.shape[0], nmax,
                                          ctypes.c_void_p(y.data_ptr()), dev, _plan._stream_handle(dev)))
        return y.cpu().numpy() if on_host else y

    def filter(self, gravityfield):
        if not isinstance(gravityfield, PotentialCoefficients):
            raise TypeError("Filter operation only implemented for instances of 'PotentialCoefficients'")
        result = gravityfield.copy()
        result.anm = self.filter_batch(np.ascontiguousarray(gravityfield.anm, dtype=float)[None])[0]
        return result

    def matrix(self, min_degree, max_degree):
        """Diagonal filter matrix in degree-wise order (filter.py:72-92, :120-127); unlike ``filter`` it
        weights every degree from min_degree on."""
        w = np.asarray(self._weights(max_degree), dtype=float)
        diag = np.concatenate([np.full(2 * n + 1, w[n]) for n in range(min_degree, max_degree + 1)])
        return np.diag(diag)


class Gaussian(_DegreeWiseFilter):
    """Gaussian filter of a given radius in kilometres (filter.py:31-92); degrees 0 and 1 pass through."""

    _first_degree = 2

    def __init__(self, radius):
        self.radius = radius

    def _weights(self, max_degree):
        from .kernel import Gauss
        return Gauss(self.radius).coefficients(0, max_degree).ravel()


class Butterworth(_DegreeWiseFilter):
    """Butterworth filter on the sphere (filter.py:95-127)."""

    def __init__(self, order, cutoff_degree):
        self.order = order
        self.cutoff_degree = cutoff_degree

    def _weights(self, max_degree):
        # per degree with Python scalars, in the reference's own expression (filter.py:116): a vectorised
        # power differs from it in the last bit
        return np.array([np.power(1 + (n / self.cutoff_degree) ** (2 * self.order), -0.5) for n in range(max_degree + 1)])


class OrderWiseFilter(SpatialFilter):
    """Sparse spherical-harmonic filter that only couples coefficients of the same order and
    trigonometric function.  ``orderwise_blocks[0]`` acts on C_n0; blocks ``2m-1`` / ``2m`` act on
    C_nm / S_nm (n = m..nmax), each of shape [(nmax+1-m), (nmax+1-m)]."""

    def __init__(self, orderwise_blocks):
        self._blocks = [np.ascontiguousarray(b, dtype=float) for b in orderwise_blocks]
        self._nmax = self._blocks[0].shape[0] - 1
        if len(self._blocks) != 2 * self._nmax + 1:
            raise ValueError("expected {0} order-wise blocks for degree {1} (got {2})"
                             .format(2 * self._nmax + 1, self._nmax, len(self._blocks)))
        for i, b in enumerate(self._blocks):
            k = self._nmax + 1 - (i + 1) // 2
            if b.shape != (k, k):
                raise ValueError("block {0} must have shape ({1}, {1}) (got {2})".format(i, k, b.shape))
        sizes = np.array([b.size for b in self._blocks], dtype=np.int64)
        self._offsets = np.concatenate(([0], np.cumsum(sizes))).astype(np.int64)
        self._flat = np.concatenate([b.ravel() for b in self._blocks])
        self._device_blocks = {}

    @property
    def max_degree(self):
        return self._nmax

    def _blocks_on(self, device):
        if device not in self._device_blocks:
            self._device_blocks[device] = torch.as_tensor(self._flat).to(torch.device("cuda", device))
        return self._device_blocks[device]

    def filter_batch(self, anm, out=None):
        """anm: [E, L, L] packed coefficients (numpy or CUDA tensor) -> filtered copy, same type."""
        L = anm.shape[-1]
        nmax = L - 1
        if nmax > self._nmax:
            raise ValueError('DDK filter only implemented for a maximum degree of {1:d} (max_degree={0:d} supplied).'
                             .format(nmax, self._nmax))
        on_host = not isinstance(anm, torch.Tensor)
        dev = _plan._current_device(None if on_host else anm.device)
        x = torch.as_tensor(np.ascontiguousarray(anm, dtype=float)).to(torch.device("cuda", dev)) if on_host else anm.contiguous()
        if x.dim() != 3 or x.shape[1] != x.shape[2] or x.dtype != torch.float64:
            raise ValueError("coefficients must be a float64 array of shape [epochs, L, L]")
        y = torch.empty_like(x) if out is None else _plan._check_out(out, x.shape, dev)
        lib = _lib.load()
        _lib.check(lib.gb_orderwise_filter(ctypes.c_void_p(self._blocks_on(dev).data_ptr()),
                                           self._offsets.ctypes.data_as(ctypes.c_void_p), self._nmax,
                                           ctypes.c_void_p(x.data_ptr()), x.shape[0], nmax,
                                           ctypes.c_void_p(y.data_ptr()), dev, _plan._stream_handle(dev)))
        return y.cpu().numpy() if on_host else y

    def filter(self, gravityfield):
        """Filtered copy of a PotentialCoefficients instance (reference filter.py:153-191):
        TypeError for other types, ValueError if its degree exceeds the filter's."""
        if not isinstance(gravityfield, PotentialCoefficients):
            raise TypeError("Filter operation only implemented for instances of 'PotentialCoefficients'")
        result = gravityfield.copy()
        result.anm = self.filter_batch(np.ascontiguousarray(gravityfield.anm, dtype=float)[None])[0]
        return result

    def matrix(self, min_degree, max_degree):
        """Dense filter matrix in degree-wise order (reference filter.py:193-222); host side,
        it is the F of F Sigma F' and only index bookkeeping."""
        K = (max_degree + 1) ** 2
        F = np.zeros((K, K))
        idx = np.arange(max_degree + 1, dtype=int) ** 2
        F[np.ix_(idx, idx)] = self._blocks[0][0:max_degree + 1, 0:max_degree + 1]
        for m in range(1, max_degree + 1):
            k = max_degree + 1 - m
            F[np.ix_(idx[m:] + 2 * m - 1, idx[m:] + 2 * m - 1)] = self._blocks[2 * m - 1][0:k, 0:k]
            F[np.ix_(idx[m:] + 2 * m, idx[m:] + 2 * m)] = self._blocks[2 * m][0:k, 0:k]
        return F[min_degree * min_degree:, min_degree * min_degree:]


class GeneralMatrix(SpatialFilter):
    """Spherical-harmonic filter defined by an arbitrary square matrix in degree-wise coefficient order
    (reference filter.py:430-510).  Batches are filtered with ONE tensor-core GEMM (gb_dense_filter); the matrix is
    re-tiled for it once per device."""

    def __init__(self, matrix, min_degree, max_degree):
        matrix = np.asarray(matrix)
        if matrix.ndim > 2 or matrix.shape[0] != matrix.shape[1]:
            raise ValueError('filter matrix must be square (got {0})'.format(str(matrix.shape)))
        if (max_degree + 1) * (max_degree + 1) - min_degree * min_degree != matrix.shape[0]:
            raise ValueError('filter matrix dimensions do not correspond to min_degree and max_degree (got {0}, {1:d}, {2:d})'
                             .format(str(matrix.shape), min_degree, max_degree))
        self._W = np.ascontiguousarray(matrix, dtype=float)
        self._nmin = min_degree
        self._nmax = max_degree
        self._tiles = {}

    def _tiles_on(self, device):
        if device not in self._tiles:
            lib = _lib.load()
            k = self._W.shape[0]
            dev = torch.device("cuda", device)
            w = torch.as_tensor(self._W).to(dev)
            tiles = torch.empty(int(lib.gb_dense_filter_tile_elements(k)), dtype=torch.float64, device=dev)
            _lib.check(lib.gb_dense_filter_prepare(ctypes.c_void_p(w.data_ptr()), k, ctypes.c_void_p(tiles.data_ptr()),
                                                   device, _plan._stream_handle(device)))
            torch.cuda.current_stream(device).synchronize()       # w may be freed now
            self._tiles[device] = tiles
        return self._tiles[device]

    def filter_batch(self, anm, out=None):
        """anm: [E, L, L] packed coefficients (numpy or CUDA tensor) -> [E, L', L'] with L' - 1 = min(L - 1, max_degree)."""
        on_host = not isinstance(anm, torch.Tensor)
        dev = _plan._current_device(None if on_host else anm.device)
        x = torch.as_tensor(np.ascontiguousarray(anm, dtype=float)).to(torch.device("cuda", dev)) if on_host else anm.contiguous()
        if x.dim() != 3 or x.shape[1] != x.shape[2] or x.dtype != torch.float64:
            raise ValueError("coefficients must be a float64 array of shape [epochs, L, L]")
        nmax_in = x.shape[-1] - 1
        lout = min(nmax_in, self._nmax) + 1
        y = (torch.empty((x.shape[0], lout, lout), dtype=torch.float64, device=x.device) if out is None
             else _plan._check_out(out, (x.shape[0], lout, lout), dev))
        _lib.check(_lib.load().gb_dense_filter(ctypes.c_void_p(self._tiles_on(dev).data_ptr()), self._nmin, self._nmax,
                                               ctypes.c_void_p(x.data_ptr()), x.shape[0], nmax_in,
                                               ctypes.c_void_p(y.data_ptr()), dev, _plan._stream_handle(dev)))
        return y.cpu().numpy() if on_host else y

    def filter(self, gravityfield):
        """Filtered copy of a PotentialCoefficients instance (filter.py:456-479)."""
        result = gravityfield.copy()
        result.anm = self.filter_batch(np.ascontiguousarray(gravityfield.anm, dtype=float)[None])[0]
        return result

    def matrix(self, min_degree, max_degree):
        """Dense filter matrix for another degree range (filter.py:481-509): coefficients common to both ranges keep
        their entries, everything else is zero.  Degree-wise orders share contiguous runs of degrees."""
        if self._nmin == min_degree and self._nmax == max_degree:
            return self._W.copy()
        k = (max_degree + 1) ** 2 - min_degree ** 2
        W = np.zeros((k, k))
        lo, hi = max(min_degree, self._nmin), min(max_degree, self._nmax)
        if lo <= hi:
            n = (hi + 1) ** 2 - lo ** 2
            s, t = lo ** 2 - self._nmin ** 2, lo ** 2 - min_degree ** 2
            W[t:t + n, t:t + n] = self._W[s:s + n, s:s + n]
        return W


class VDK(GeneralMatrix):
    """VDK filter (reference filter.py:512-546): W = (N + diag(kaula weights))^-1 N from a normal-equation matrix in
    degree-wise order.  The solve is plan-time host work (numpy); the reference's own ``VDK.filter`` reads a
    name-mangled attribute that does not exist, so ``filter`` follows ``GeneralMatrix.filter``."""

    def __init__(self, normal_equation_matrix, min_degree, max_degree, kaula_scale, kaula_power):
        normal_equation_matrix = np.asarray(normal_equation_matrix, dtype=float)
        weights = np.concatenate([np.full(2 * n + 1, kaula_scale * float(n) ** kaula_power)
                                  for n in range(min_degree, max_degree + 1)])
        NP = normal_equation_matrix.copy()
        NP.flat[::NP.shape[0] + 1] = np.diag(normal_equation_matrix) + weights
        super(VDK, self).__init__(np.linalg.solve(NP, normal_equation_matrix), min_degree, max_degree)

class BlockedNormalsVDK(OrderWiseFilter):
    """Blocked VDK filter (reference filter.py:352-427): the order-wise (DDK-structured) blocks of a normal-equation
    matrix in degree-wise order, each regularised with Kaula weights and solved -- plan-time host work (numpy), the
    resulting blocks run on the order-wise filter kernel."""

    def __init__(self, normal_equation_matrix, min_degree, max_degree, kaula_scale, kaula_power):
        normals_in = np.asarray(normal_equation_matrix, dtype=float)
        weights = kaula_scale * np.arange(max_degree + 1, dtype=float) ** kaula_power
        weights[0] = 1
        off = min_degree * min_degree

        def block(m, sine):
            n = np.arange(max(m, min_degree), max_degree + 1)
            idx = n * n - off + (0 if m == 0 else (2 * m if sine else 2 * m - 1))
            full = np.zeros((max_degree + 1 - m, max_degree + 1 - m))
            lead = max(min_degree - m, 0)
            full[lead:, lead:] = normals_in[np.ix_(idx, idx)]
            return full

        normals = [block(0, False)]
        for m in range(1, max_degree + 1):
            normals.append(block(m, False))
            normals.append(block(m, True))
        array = []
        for normals_block in normals:
            m = max_degree + 1 - normals_block.shape[0]
            array.append(np.linalg.solve(normals_block + np.diag(weights[m:]), normals_block))
        super(BlockedNormalsVDK, self).__init__(array)


class FilterKernel(_kernel.AnisotropicKernel):
    """Space-domain kernel of a (possibly anisotropic) filter (reference filter.py:575-598): the filter matrix between
    the kernel's functional, K2 = diag(kn') F diag(kn)."""

    def __init__(self, spatial_filter, min_degree, max_degree, input_kernel='potential'):
        K = spatial_filter.matrix(min_degree, max_degree) if isinstance(spatial_filter, SpatialFilter) else spatial_filter
        generator = _kernel.get_kernel(input_kernel)
        kn = generator.coefficient_array(min_degree, max_degree)
        kn_prime = generator.inverse_coefficient_array(min_degree, max_degree)
        # the reference's expression with its broadcasting: both ravelled factor arrays are [1, K'], so both scale the
        # COLUMNS of K and the product carries a leading axis of one (filter.py:596)
        K2 = (K * utilities.ravel_coefficients(kn, min_degree, max_degree)[np.newaxis, :]) * \
            utilities.ravel_coefficients(kn_prime, min_degree, max_degree)[:, np.newaxis]
        super(FilterKernel, self).__init__(K2.reshape(K2.shape[-2:]), min_degree, max_degree)

"""Covariance producers of the path on the device: the block-sparse normal-equation machinery of the
reference, ``grates.lstsq.BlockMatrix`` (lstsq.py:391-917) and ``NormalEquations.compute_covariance``
(lstsq.py:1026-1042), with the blocks resident in HBM.

The reference's algorithms are loops over blocks around four library calls (scipy.linalg.cholesky, solve_triangular,
inv and numpy's ``@``); the loops and the sparsity bookkeeping are kept as they are (same block order, same fill-in),
the four calls run as hand-written CUDA kernels behind the C ABI (gb_dgemm on the FP64 tensor cores, gb_dpotrf_upper,
gb_dtrsm_upper; include/grates_b200.h).  The covariance matrix a propagation needs therefore never leaves the GPU:
``NormalEquations.compute_covariance()`` -> ``matrix.to_tensor()`` -> ``RegularGrid.covariance_propagation``.
"""
import ctypes

import numpy as np
import torch

from . import _lib, plan as _plan


def _dev_index(device):
    return _plan._current_device(device)


class _Ops:
    """The four dense kernels on 2-d, row-major, last-dimension-contiguous float64 CUDA tensors (views allowed)."""

    def __init__(self, device):
        self.lib = _lib.load()
        self.device = _dev_index(device)
        self.tdev = torch.device("cuda", self.device)
        self._info = torch.zeros(1, dtype=torch.int32, device=self.tdev)

    def _stream(self):
        return _plan._stream_handle(self.device)

    @staticmethod
    def _ld(t):
        if t.dim() != 2 or t.dtype != torch.float64 or (t.shape[1] > 1 and t.stride(1) != 1):
            raise ValueError("blocks must be 2-d float64 tensors with a contiguous last dimension")
        return t.stride(0) if t.shape[0] > 1 else max(t.shape[1], 1)

    def gemm(self, a, b, c, alpha=1.0, beta=0.0, trans_a=False, trans_b=False, upper_only=False):
        """c = alpha * op(a) @ op(b) + beta * c, in place of c."""
        m, n = c.shape
        k = a.shape[0] if trans_a else a.shape[1]
        if (a.shape[1] if trans_a else a.shape[0]) != m or (b.shape[0] if trans_b else b.shape[1]) != n or \
                (b.shape[1] if trans_b else b.shape[0]) != k:
            raise ValueError("shape mismatch in block product")
        _lib.check(self.lib.gb_dgemm(int(trans_a), int(trans_b), m, n, k, float(alpha), ctypes.c_void_p(a.data_ptr()),
                                     self._ld(a), ctypes.c_void_p(b.data_ptr()), self._ld(b), float(beta),
                                     ctypes.c_void_p(c.data_ptr()), self._ld(c), int(upper_only), self.device, self._stream()))
        return c

    def matmul(self, a, b, trans_a=False, trans_b=False):
        m = a.shape[1] if trans_a else a.shape[0]
        n = b.shape[0] if trans_b else b.shape[1]
        return self.gemm(a, b, torch.empty((m, n), dtype=torch.float64, device=self.tdev), trans_a=trans_a, trans_b=trans_b)

    def cholesky_upper(self, a):
        """a = W' W in place (scipy.linalg.cholesky(a, lower=False, overwrite_a=True)); raises LinAlgError like scipy."""
        if a.shape[0] != a.shape[1]:
            raise ValueError("expected square matrix")
        _lib.check(self.lib.gb_dpotrf_upper(ctypes.c_void_p(a.data_ptr()), a.shape[0], self._ld(a),
                                            ctypes.c_void_p(self._info.data_ptr()), self.device, self._stream()))
        info = int(self._info.item())
        if info:
            raise np.linalg.LinAlgError("{0}-th leading minor of the array is not positive definite".format(info))
        return a

    def solve_triangular(self, w, b, trans=False):
        """op(w) x = b in place of b (scipy.linalg.solve_triangular(w, b, trans=..., lower=False, overwrite_b=True))."""
        if w.shape[0] != w.shape[1] or b.shape[0] != w.shape[0]:
            raise ValueError("shape mismatch in triangular solve")
        _lib.check(self.lib.gb_dtrsm_upper(int(trans), ctypes.c_void_p(w.data_ptr()), w.shape[0], self._ld(w),
                                           ctypes.c_void_p(b.data_ptr()), b.shape[1], self._ld(b), self.device, self._stream()))
        return b

    def inv_triangular(self, w):
        """inverse of an upper triangular block (scipy.linalg.inv of a Cholesky factor): W X = I."""
        x = torch.eye(w.shape[0], dtype=torch.float64, device=self.tdev)
        return self.solve_triangular(w, x, trans=False)

    def inv_gram(self, w):
        """(W' W)^-1 = W^-1 W^-T for an upper triangular W (lstsq.py:835)."""
        x = self.inv_triangular(w)
        return self.matmul(x, x, trans_b=True)


class BlockMatrix:
    """Device-resident mirror of ``grates.lstsq.BlockMatrix`` (lstsq.py:391-917): a matrix partitioned into blocks, each
    either absent (structural zero) or a float64 CUDA tensor.  Same constructor arguments, method names and in-place
    semantics; ``from_array`` / ``__setitem__`` accept numpy arrays or CUDA tensors, ``to_array`` returns numpy,
    ``to_tensor`` the dense CUDA tensor."""

    def __init__(self, row_index, column_index, device=None):
        self.__row_index = np.asarray(row_index, dtype=int)
        self.__column_index = np.asarray(column_index, dtype=int)
        self.shape = (len(self.__row_index) - 1, len(self.__column_index) - 1)
        self.__data = np.empty(self.shape, dtype=object)
        self.__is_nonzero = np.zeros(self.shape, dtype=bool)
        self._ops = _Ops(device)

    def copy(self):
        """Deep copy (the blocks are cloned on the device)."""
        other = BlockMatrix(self.__row_index, self.__column_index, self._ops.device)
        for r in range(self.shape[0]):
            for c in range(self.shape[1]):
                if self.__data[r, c] is not None:
                    other.__data[r, c] = self.__data[r, c].clone()
        other.__is_nonzero = self.__is_nonzero.copy()
        return other

    def __deepcopy__(self, memo):
        return self.copy()

    # -- construction ---------------------------------------------------------------------------------------------
    @staticmethod
    def compute_block_index(array_shape, block_size):
        """Index bounds of blocks of (at most) block_size rows / columns (lstsq.py:435-462)."""
        row_index = [0]
        while row_index[-1] < array_shape[0]:
            row_index.append(min(array_shape[0], row_index[-1] + block_size))
        column_index = [0]
        while column_index[-1] < array_shape[1]:
            column_index.append(min(array_shape[1], column_index[-1] + block_size))
        return np.array(row_index), np.array(column_index)

    @staticmethod
    def from_array(array, row_index, column_index, device=None):
        """Block matrix from a dense 2-d array (numpy or CUDA tensor; copied).  All-zero blocks become structural zeros
        (lstsq.py:464-494)."""
        if isinstance(array, np.ndarray):
            if array.ndim != 2:
                raise ValueError('array must be a two-dimensional ' + str(np.ndarray))
            dense = torch.as_tensor(np.ascontiguousarray(array, dtype=float)).to(torch.device("cuda", _dev_index(device)))
        elif isinstance(array, torch.Tensor):
            if array.dim() != 2 or array.dtype != torch.float64 or not array.is_cuda:
                raise ValueError('array must be a two-dimensional float64 CUDA tensor')
            dense = array
            device = array.device.index
        else:
            raise ValueError('array must be of type ' + str(np.ndarray))
        if row_index[-1] != dense.shape[0]:
            raise ValueError("mismatch in array shape in dimension 0 and row block index")
        if column_index[-1] != dense.shape[1]:
            raise ValueError("mismatch in array shape in dimension 1 and column block index")
        bm = BlockMatrix(row_index, column_index, device)
        for r in range(bm.shape[0]):
            for c in range(bm.shape[1]):
                block = dense[row_index[r]:row_index[r + 1], column_index[c]:column_index[c + 1]]
                if bool(torch.count_nonzero(block).item()):
                    bm.__data[r, c] = block.clone()
                    bm.__is_nonzero[r, c] = True
        return bm

    def to_tensor(self):
        """Dense CUDA tensor [rows, columns]; what ``covariance_propagation`` takes."""
        out = torch.zeros((int(self.__row_index[-1]), int(self.__column_index[-1])), dtype=torch.float64, device=self._ops.tdev)
        for r in range(self.shape[0]):
            for c in range(self.shape[1]):
                if self.__is_nonzero[r, c]:
                    out[self.__row_slice(r), self.__column_slice(c)] = self.__data[r, c]
        return out

    def to_array(self):
        return self.to_tensor().cpu().numpy()

    # -- bookkeeping ----------------------------------------------------------------------------------------------
    def __block_shape(self, i, j):
        return int(self.__row_index[i + 1] - self.__row_index[i]), int(self.__column_index[j + 1] - self.__column_index[j])

    def __row_slice(self, i):
        return slice(int(self.__row_index[i]), int(self.__row_index[i + 1]), 1)

    def __column_slice(self, i):
        return slice(int(self.__column_index[i]), int(self.__column_index[i + 1]), 1)

    def __check_bounds(self, i, j):
        if i >= self.shape[0] or i < 0:
            raise IndexError("block index {0} is out of bounds for axis 0 with size {1}".format(i, self.shape[0]))
        if j >= self.shape[1] or j < 0:
            raise IndexError("block index {0} is out of bounds for axis 1 with size {1}".format(j, self.shape[1]))

    def __as_block(self, i, j, value):
        if isinstance(value, np.ndarray):
            value = torch.as_tensor(np.ascontiguousarray(value, dtype=float)).to(self._ops.tdev)
        elif isinstance(value, torch.Tensor):
            value = value.to(device=self._ops.tdev, dtype=torch.float64).clone()
        else:
            raise ValueError('Block matrix item must be of type ' + str(np.ndarray))
        if value.dim() != 2:
            raise ValueError('Block matrix item must be a two-dimensional ' + str(np.ndarray))
        if tuple(value.shape) != self.__block_shape(i, j):
            raise ValueError('Block matrix item at position ({0:d}, {1:d}) must be of size ({2:d}, {3:d}). Got ({4:d}, {5:d}).'
                             .format(i, j, *self.__block_shape(i, j), value.shape[0], value.shape[1]))
        return value.contiguous()

    def __setitem__(self, key, value):
        if not isinstance(key, tuple) or len(key) != 2:
            raise IndexError("Indices to block matrix must be tuples of length 2")
        self.__check_bounds(key[0], key[1])
        self.__data[key] = self.__as_block(key[0], key[1], value)
        self.__is_nonzero[key] = True

    def __getitem__(self, key):
        """Block (i, j) as a CUDA tensor (None for a structural zero)."""
        if not isinstance(key, tuple) or len(key) != 2:
            raise IndexError("Indices to block matrix must be tuples of length 2")
        self.__check_bounds(key[0], key[1])
        return self.__data[key]

    def __set_block(self, i, j):
        if self.__data[i, j] is None:
            self.__data[i, j] = torch.zeros(self.__block_shape(i, j), dtype=torch.float64, device=self._ops.tdev)
            self.__is_nonzero[i, j] = True

    def is_nonzero(self, row, column):
        return bool(self.__is_nonzero[row, column])

    def diag(self):
        """Copy of the main diagonal (numpy)."""
        d = np.zeros(min(self.__row_index[-1], self.__column_index[-1]))
        for idx in range(min(len(self.__row_index), len(self.__column_index)) - 1):
            if self.__is_nonzero[idx, idx]:
                d[self.__row_index[idx]:self.__row_index[idx + 1]] = torch.diagonal(self.__data[idx, idx]).cpu().numpy()
        return d

    # -- products -------------------------------------------------------------------------------------------------
    def __matmul__(self, other):
        if not isinstance(other, BlockMatrix):
            raise ValueError("Matrix multiplication not implemented for type {0}".format(type(other)))
        result = BlockMatrix(self.__row_index, other.__column_index, self._ops.device)
        ops = self._ops
        for i in range(result.shape[0]):
            for j in range(result.shape[1]):
                for k in range(self.shape[1]):
                    if self.__is_nonzero[i, k] and other.__is_nonzero[k, j]:
                        result.__set_block(i, j)
                        ops.gemm(self.__data[i, k], other.__data[k, j], result.__data[i, j], beta=1.0)
        return result

    def _rhs(self, b):
        host = not isinstance(b, torch.Tensor)
        t = torch.as_tensor(np.atleast_2d(np.asarray(b, dtype=float))).to(self._ops.tdev) if host else b
        if t.dim() == 1:
            t = t[None, :]
        return t.contiguous().clone(), host

    def multiply_symmetric(self, b):
        """v = N b for a symmetric matrix of which only the upper triangle is stored (lstsq.py:748-771)."""
        bt, host = self._rhs(b)
        v = torch.zeros_like(bt)
        ops = self._ops
        for i in range(self.shape[0]):
            if self.__is_nonzero[i, i]:
                ops.gemm(self.__data[i, i], bt[self.__row_slice(i)], v[self.__row_slice(i)], beta=1.0)
            for j in range(i + 1, self.shape[1]):
                if self.__is_nonzero[i, j]:
                    ops.gemm(self.__data[i, j], bt[self.__row_slice(j)], v[self.__row_slice(i)], beta=1.0)
                    ops.gemm(self.__data[i, j], bt[self.__row_slice(i)], v[self.__row_slice(j)], beta=1.0, trans_a=True)
        return v.cpu().numpy() if host else v

    def multiply_triangular(self, b, transpose=False):
        """v = W b or W' b for an upper triangular block matrix (lstsq.py:719-746; the transposed branch of the
        reference assigns instead of accumulating, the mathematically meant sum is computed here)."""
        bt, host = self._rhs(b)
        v = torch.zeros_like(bt)
        ops = self._ops
        if transpose:
            for i in range(self.shape[0]):
                for j in range(i + 1):
                    if self.__is_nonzero[j, i]:
                        ops.gemm(self.__data[j, i], bt[self.__row_slice(j)], v[self.__row_slice(i)], beta=1.0, trans_a=True)
        else:
            for i in range(self.shape[0]):
                for j in range(i, self.shape[1]):
                    if self.__is_nonzero[i, j]:
                        ops.gemm(self.__data[i, j], bt[self.__row_slice(j)], v[self.__row_slice(i)], beta=1.0)
        return v.cpu().numpy() if host else v

    def solve_triangular(self, b, transpose=False):
        """Solve W x = b or W' x = b for an upper triangular block matrix (lstsq.py:773-821)."""
        bc, host = self._rhs(b)
        ops = self._ops
        if transpose:
            for row in range(self.shape[0]):
                for column in range(row):
                    if self.__is_nonzero[column, row]:
                        ops.gemm(self.__data[column, row], bc[self.__row_slice(column)], bc[self.__row_slice(row)],
                                 alpha=-1.0, beta=1.0, trans_a=True)
                ops.solve_triangular(self.__data[row, row], bc[self.__row_slice(row)], trans=True)
        else:
            for row in range(self.shape[0] - 1, -1, -1):
                for column in range(self.shape[0] - 1, row, -1):
                    if self.__is_nonzero[row, column]:
                        ops.gemm(self.__data[row, column], bc[self.__row_slice(column)], bc[self.__row_slice(row)],
                                 alpha=-1.0, beta=1.0)
                ops.solve_triangular(self.__data[row, row], bc[self.__row_slice(row)], trans=False)
        return bc.cpu().numpy() if host else bc

    # -- factorisation and inverses (in place, upper triangle only) ----------------------------------------------------
    def cholesky(self):
        """N = W' W, block by block with the reference's fill-in (lstsq.py:698-717)."""
        ops, d, nz = self._ops, self.__data, self.__is_nonzero
        for row in range(self.shape[0]):
            for r in range(row):
                for c in range(row, self.shape[1]):
                    if nz[r, row] and nz[r, c]:
                        self.__set_block(row, c)
                        ops.gemm(d[r, row], d[r, c], d[row, c], alpha=-1.0, beta=1.0, trans_a=True, upper_only=(c == row))
            ops.cholesky_upper(d[row, row])
            for column in range(row + 1, self.shape[1]):
                if nz[row, column]:
                    ops.solve_triangular(d[row, row], d[row, column], trans=True)

    def sparse_inverse(self):
        """Sparse inverse N^-1 restricted to the sparsity of the Cholesky factor W held by the matrix (lstsq.py:823-846)."""
        ops, d, nz = self._ops, self.__data, self.__is_nonzero
        nb = self.shape[0]
        for i in range(nb - 1, -1, -1):
            temporary_row = [None] * (nb - i - 1)
            for k in range(i + 1, self.shape[1]):
                if nz[i, k]:
                    temporary_row[k - i - 1] = ops.solve_triangular(d[i, i], d[i, k], trans=False)    # W_ii^-1 W_ik
                    d[i, k] = torch.zeros_like(d[i, k])
            d[i, i] = ops.inv_gram(d[i, i])
            for j in range(nb - 1, i - 1, -1):
                if nz[i, j]:
                    for k in range(i + 1, nb):
                        if nz[min(k, j), max(k, j)] and temporary_row[k - i - 1] is not None:
                            if k < j:
                                ops.gemm(temporary_row[k - i - 1], d[k, j], d[i, j], alpha=-1.0, beta=1.0)
                            else:
                                ops.gemm(temporary_row[k - i - 1], d[j, k], d[i, j], alpha=-1.0, beta=1.0, trans_b=True)

    def inverse(self):
        """Full inverse N^-1 = W^-1 W^-T from the Cholesky factor W held by the matrix (lstsq.py:848-882)."""
        ops, d, nz = self._ops, self.__data, self.__is_nonzero
        nb = self.shape[0]
        for j in range(nb - 1, -1, -1):
            d[j, j] = ops.inv_triangular(d[j, j])
            for i in range(j - 1, -1, -1):
                if nz[i, j]:
                    d[i, j] = ops.matmul(d[i, j], d[j, j])
                for k in range(i + 1, j):
                    if nz[i, k] and nz[k, j]:
                        self.__set_block(i, j)
                        ops.gemm(d[i, k], d[k, j], d[i, j], beta=1.0)
                if nz[i, j]:
                    ops.solve_triangular(d[i, i], d[i, j], trans=False)
                    d[i, j].neg_()
        for i in range(nb):
            d[i, i] = ops.matmul(d[i, i], d[i, i], trans_b=True)
            for j in range(i + 1, nb):
                if nz[i, j]:
                    ops.gemm(d[i, j], d[i, j], d[i, i], beta=1.0, trans_b=True)
                    d[i, j] = ops.matmul(d[i, j], d[j, j], trans_b=True)
                for k in range(j + 1, nb):
                    if nz[i, k] and nz[j, k]:
                        self.__set_block(i, j)
                        ops.gemm(d[i, k], d[j, k], d[i, j], beta=1.0, trans_b=True)

    def symmetrize(self):
        """Mirror the stored upper triangle into the lower one (after ``inverse``: a covariance matrix for
        ``covariance_propagation`` as a full symmetric tensor)."""
        for i in range(self.shape[0]):
            if self.__is_nonzero[i, i]:
                blk = self.__data[i, i]
                self.__data[i, i] = torch.triu(blk) + torch.triu(blk, 1).T
            for j in range(i + 1, self.shape[1]):
                if self.__is_nonzero[i, j]:
                    self.__data[j, i] = self.__data[i, j].T.contiguous()
                    self.__is_nonzero[j, i] = True
        return self


class NormalEquations:
    """Mirror of ``grates.lstsq.NormalEquations`` for the covariance path (lstsq.py:920-1042): ``solve`` and
    ``compute_covariance`` on a device-resident ``BlockMatrix``."""

    def __init__(self, normal_matrix, right_hand_side, observation_square_sum, observation_count):
        self.matrix = normal_matrix
        self.right_hand_side = right_hand_side
        self.observation_square_sum = observation_square_sum
        self.observation_count = observation_count
        self.status = 'normal_matrix'

    def _cholesky(self):
        if self.status == 'cholesky_factor':
            return
        if self.status != 'normal_matrix':
            raise ValueError('Cholesky factor can only be computed from the normal matrix')
        self.matrix.cholesky()
        self.status = 'cholesky_factor'

    def solve(self):
        """x = N^-1 n through the Cholesky factor (lstsq.py:962-980, without the Monte-Carlo trace vectors)."""
        self._cholesky()
        h = self.matrix.solve_triangular(self.right_hand_side, transpose=True)
        return self.matrix.solve_triangular(h)

    def compute_covariance(self, sparse=True):
        """(Sparse) inverse of the normal matrix in place (lstsq.py:1026-1042)."""
        self._cholesky()
        if sparse:
            self.matrix.sparse_inverse()
        else:
            self.matrix.inverse()
        self.status = 'covariance_matrix'

    def to_array(self):
        return self.matrix.to_array(), self.right_hand_side, self.observation_square_sum, self.observation_count


def loadsinexnormals(file_name, block_size=2048, device=None):
    """Normal equations from a SINEX file (storage schemes 6b / 6c) straight into a device ``NormalEquations``
    (reference io.py:838-876 returns the dense arrays).  Returns (normal_equations, N, n, lPl, obs_count) with the arrays
    as the reference returns them."""
    from .io import loadsinexnormals as _load
    N, n, lPl, obs_count = _load(file_name)
    ri, ci = BlockMatrix.compute_block_index(N.shape, block_size)
    matrix = BlockMatrix.from_array(N, ri, ci, device)
    return NormalEquations(matrix, n, lPl, obs_count), N, n, lPl, obs_count

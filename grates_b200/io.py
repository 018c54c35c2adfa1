"""SINEX normal-equation files: the data format in front of the covariance path (reference io.py:601-683,
838-876).  Host-side text parsing only; the matrix goes to the device through ``grates_b200.lstsq.loadsinexnormals``.

The reader accepts what the reference's reader accepts: blocks ``SOLUTION/NORMAL_EQUATION_MATRIX {U|L}`` (rows
``row col v1 [v2 [v3]]``, 1-based, mirrored into a full symmetric matrix), ``SOLUTION/NORMAL_EQUATION_VECTOR`` (value in
columns 48-68), ``SOLUTION/STATISTICS`` (keyword lines, value from column 33 on); comment lines start with ``*``,
other blocks are skipped; ``.gz`` files are decompressed.  ``savesinexnormals`` writes the same layout (what
``SINEXSymmetricMatrix.write`` / ``SINEXSphericalHarmonicsVector.write`` produce) and exists for round trips and tests.
"""
import datetime as dt
import gzip
import io as _io

import numpy as np


def _open(file_name, mode):
    if str(file_name).endswith('.gz'):
        return gzip.open(file_name, mode)
    return open(file_name, mode)


def _parse_matrix(lines, parameter_count):
    """Rows ``row col v1 [v2 [v3]]`` -> full symmetric matrix (reference io.py:623-656)."""
    rows, cols, vals = [], [], []
    try:
        import pandas as pd
        frame = pd.read_csv(_io.BytesIO(b"".join(lines)), sep=r"\s+", header=None, names=range(5), engine="c")
        arr = frame.to_numpy(dtype=float)
        r = arr[:, 0].astype(np.int64) - 1
        c0 = arr[:, 1].astype(np.int64) - 1
        for k in range(3):
            v = arr[:, 2 + k]
            ok = ~np.isnan(v)
            rows.append(r[ok])
            cols.append(c0[ok] + k)
            vals.append(v[ok])
        rows, cols, vals = np.concatenate(rows), np.concatenate(cols), np.concatenate(vals)
    except ImportError:
        for line in lines:
            s = line.split()
            r, c0 = int(s[0]) - 1, int(s[1]) - 1
            for k, v in enumerate(s[2:]):
                rows.append(r)
                cols.append(c0 + k)
                vals.append(float(v))
        rows, cols, vals = np.asarray(rows, dtype=np.int64), np.asarray(cols, dtype=np.int64), np.asarray(vals, dtype=float)
    count = int(max(rows.max(), cols.max())) + 1 if rows.size else 0
    if parameter_count is not None:
        count = max(count, int(parameter_count))
    matrix = np.zeros((count, count))
    matrix[rows, cols] = vals
    matrix[cols, rows] = vals
    return matrix


def loadsinex(file_name):
    """Blocks of a SINEX file as a dict: block type (str, without the U/L suffix) -> parsed content
    (matrix: ndarray; vector: dict with index / kind / degree / order / x; statistics: dict)."""
    blocks = {}
    parameter_count = None
    with _open(file_name, 'rb') as f:
        current, body = None, []
        for raw in f:
            line = raw.rstrip(b"\r\n")
            if current is None:
                if line.startswith(b'%ENDSNX'):
                    break
                if line.startswith(b'+'):
                    current, body = line[1:].strip().decode(), []
                continue
            if line.startswith(b'-'):
                name = current
                if name.startswith('SOLUTION/NORMAL_EQUATION_MATRIX') or name.startswith('SOLUTION/MATRIX_ESTIMATE'):
                    blocks[name.rsplit(' ', 1)[0] if name[-2:] in (' U', ' L') else name] = _parse_matrix(body, parameter_count)
                elif name.startswith(('SOLUTION/ESTIMATE', 'SOLUTION/APRIORI', 'SOLUTION/NORMAL_EQUATION_VECTOR')):
                    vec = {"index": [], "kind": [], "degree": [], "order": [], "x": []}
                    for b in body:
                        ptype = b[7:13].strip()
                        if ptype not in (b'CN', b'SN'):
                            raise ValueError('Parameter type <' + ptype.decode() + '> not supported.')
                        vec["index"].append(int(b[1:6]) - 1)
                        vec["kind"].append(0 if ptype == b'CN' else 1)
                        vec["degree"].append(int(b[14:18]))
                        vec["order"].append(int(b[22:26]))
                        vec["x"].append(float(b[47:68]))
                    vec = {k: np.asarray(v) for k, v in vec.items()}
                    blocks[name] = vec
                    if parameter_count is None and vec["index"].size:
                        parameter_count = int(vec["index"].max()) + 1
                elif name.startswith('SOLUTION/STATISTICS'):
                    stats = {}
                    for b in body:
                        key, value = b[1:32].strip().decode(), b[32:].strip()
                        if key == 'NUMBER OF DEGREES OF FREEDOM':
                            stats["degrees_of_freedom"] = int(float(value))
                        elif key == 'NUMBER OF OBSERVATIONS':
                            stats["observation_count"] = int(float(value))
                        elif key == 'NUMBER OF UNKNOWNS':
                            stats["parameters"] = int(float(value))
                        elif key == 'WEIGHTED SQUARE SUM OF O-C':
                            stats["observation_square_sum"] = float(value)
                    blocks[name] = stats
                current = None
                continue
            if not line or line.startswith(b'*'):
                continue
            body.append(raw if raw.endswith(b"\n") else raw + b"\n")
    return blocks


def loadsinexnormals(file_name):
    """(N, n, lPl, obs_count) of a SINEX normal-equation file, as reference io.py:838-876 returns them: N full symmetric
    [p, p], n [p, 1], lPl [1], obs_count int."""
    blocks = loadsinex(file_name)
    required = {'SOLUTION/NORMAL_EQUATION_MATRIX', 'SOLUTION/NORMAL_EQUATION_VECTOR', 'SOLUTION/STATISTICS'}
    if not required.issubset(blocks.keys()):
        raise ValueError('SINEX file does not conform to storage schemes 6b or 6c for normal equations.')
    N = blocks['SOLUTION/NORMAL_EQUATION_MATRIX']
    n = blocks['SOLUTION/NORMAL_EQUATION_VECTOR']["x"][:, np.newaxis]
    lPl = np.atleast_1d(blocks['SOLUTION/STATISTICS']["observation_square_sum"])
    return N, n, lPl, blocks['SOLUTION/STATISTICS']["observation_count"]


def _sinex_time(t):
    start = dt.datetime(t.year, 1, 1)
    delta = t - start
    return '{0:2s}:{1:03d}:{2:05d}'.format(start.strftime('%y'), delta.days + 1, delta.seconds)


def savesinexnormals(file_name, N, n, lPl, obs_count, coefficients, reference_epoch=None, lower=False):
    """Write normal equations in the layout of the reference's SINEX writers (io.py:505-527, 658-683).
    coefficients: sequence of (kind, degree, order) with kind 0 = cosine, 1 = sine, one per parameter."""
    N = np.asarray(N, dtype=float)
    x = np.asarray(n, dtype=float).reshape(-1)
    p = N.shape[0]
    epoch = dt.datetime(2000, 1, 1, 12) if reference_epoch is None else reference_epoch
    with _open(file_name, 'wt') as f:
        f.write('%=SNX 2.02 {0:3s} {1:12s} {0:3s} {2:12s} {3:12s} C {4:05d} 2      \n'.format(
            'GB2', _sinex_time(dt.datetime(2020, 1, 1)), _sinex_time(epoch), _sinex_time(epoch), p))
        f.write('+SOLUTION/STATISTICS\n')
        f.write('*_STATISTICAL PARAMETER________ __VALUE(S)____________\n')
        f.write(' {0:30s} {1:22.15e}\n'.format('NUMBER OF OBSERVATIONS', float(obs_count)))
        f.write(' {0:30s} {1:22.15e}\n'.format('NUMBER OF UNKNOWNS', float(p)))
        f.write(' {0:30s} {1:22.15e}\n'.format('NUMBER OF DEGREES OF FREEDOM', float(obs_count - p)))
        f.write(' {0:30s} {1:22.15e}\n'.format('WEIGHTED SQUARE SUM OF O-C', float(np.atleast_1d(lPl)[0])))
        f.write('-SOLUTION/STATISTICS\n')
        f.write('+SOLUTION/NORMAL_EQUATION_VECTOR\n')
        for k in range(p):
            kind, degree, order = coefficients[k]
            f.write(' {0:5d} {1:6s} {2:4d} -- {3:4d} {4:12s} ---- 2 {5:21.14e}\n'.format(
                k + 1, 'CN' if kind == 0 else 'SN', degree, order, _sinex_time(epoch), x[k]))
        f.write('-SOLUTION/NORMAL_EQUATION_VECTOR\n')
        tag = 'SOLUTION/NORMAL_EQUATION_MATRIX ' + ('L' if lower else 'U')
        f.write('+' + tag + '\n')
        for row in range(p):
            first, last = (0, row + 1) if lower else (row, p)
            for column in range(first, last, 3):
                f.write(' {0:5d} {1:5d}'.format(row + 1, column + 1))
                for k in range(column, min(column + 3, last)):
                    f.write(' {0:21.14e}'.format(N[row, k]))
                f.write('\n')
        f.write('-' + tag + '\n')
        f.write('%ENDSNX\n')

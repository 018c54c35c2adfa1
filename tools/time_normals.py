"""Covariance producer at config-4 size (development aid): K = 9409 normal equations -> Cholesky -> full inverse on the
device, against numpy/LAPACK on the host cores; plus raw gb_dgemm throughput."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import grates_b200 as gb
from grates_b200.lstsq import _Ops

def ev(fn, reps=3):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts)

ops = _Ops(None)
res = {}
for n in (4096, 9409):
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    c = torch.empty(n, n, dtype=torch.float64, device="cuda")
    ms = ev(lambda: ops.gemm(a, b, c))
    res["dgemm_%d_tflops" % n] = 2.0 * n ** 3 / ms / 1e9
    ms = ev(lambda: ops.gemm(a, b, c, trans_a=True))
    res["dgemm_tn_%d_tflops" % n] = 2.0 * n ** 3 / ms / 1e9
K = 9409
rng = np.random.default_rng(0)
A = rng.standard_normal((2 * K, K))
normals = A.T @ A / K + np.eye(K)
nd = torch.as_tensor(normals).cuda()
for bs in (1024, 2048, 4096):
    ri, ci = gb.BlockMatrix.compute_block_index(normals.shape, bs)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    bm = gb.BlockMatrix.from_array(nd, ri, ci)
    bm.cholesky(); torch.cuda.synchronize(); t1 = time.perf_counter()
    bm.inverse(); torch.cuda.synchronize(); t2 = time.perf_counter()
    res["block_%d" % bs] = {"cholesky_s": t1 - t0, "inverse_s": t2 - t1, "cholesky_tflops": K ** 3 / 3 / (t1 - t0) / 1e12}
sigma = bm.symmetrize().to_tensor()
t0 = time.perf_counter(); want = np.linalg.inv(normals); res["numpy_inv_s"] = time.perf_counter() - t0
res["inverse_max_normalised_error"] = float(np.abs(sigma.cpu().numpy() - want).max() / np.abs(want).max())
print(json.dumps(res))

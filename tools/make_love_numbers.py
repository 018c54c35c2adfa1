"""Extract the ak135 load Love numbers (Wang et al. 2012, doi:10.1016/j.cageo.2012.06.022)
that the reference ships as text (grates/data/ak135-LLNs-complete.dat.gz, read by
grates/data/__init__.py:12-64) into a compact binary table for degrees 0..NMAX (CE frame;
CM/CF are derived from degree 1 at load time exactly as the reference does, :54-60).

Run once in the build container (needs /root/reference):  python tools/make_love_numbers.py
"""
import gzip
import numpy as np

NMAX = 4096
src = "/root/reference/grates/data/ak135-LLNs-complete.dat.gz"
with gzip.open(src, "rt") as f:
    rows = np.loadtxt(f, skiprows=1, usecols=(1, 2, 3), max_rows=NMAX)
hlk = np.vstack((np.zeros((1, 3)), rows))   # degree 0 row of zeros, as the reference prepends
np.savez_compressed("grates_b200/data/love_numbers_ak135_ce.npz", h=hlk[:, 0], l=hlk[:, 1], k=hlk[:, 2])
print(hlk.shape, hlk[:3])

"""In-kernel timeline of one config-2 shard (development aid; needs a library built with GB_NVCC_EXTRA=-DGB_TRACE).

python tools/trace_shard.py E [reps]   ->  per trace slot: min / median / max over the CTAs, in microseconds after the
first CTA of stage 1 started (globaltimer), plus per-CTA clock64 intervals of the Fourier stage.
Slots, stage 1 (kernel 0): 0 entry, 1 after griddepcontrol.wait, 3 first chunk landed, 4/6 K loop of pass 0/1 done,
5/7 epilogue of pass 0/1 done (first item of the CTA), 10 CTA done.
Slots, stage 2 (kernel 1): 0 entry, 1 after griddepcontrol.wait, 2 first chunk issued, 3 first chunk landed,
4+4i odd K loop of item i done, 5+4i odd half parked, 6+4i even K loop done, 7+4i even half parked,
16+2i epilogue warp got item i, 17+2i its stores are issued, 22 thread 0 at the final barrier, 23 after it."""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import grates_b200 as gb
from grates_b200 import _lib

E = int(sys.argv[1]) if len(sys.argv) > 1 else 30
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
N, d = 96, 0.5
plan = gb.get_plan(gb.GeographicGrid(d, d), N, "ewh")
lib = _lib.load()
lib.gb_debug_trace.restype = ctypes.c_int
lib.gb_debug_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
x = torch.randn(E, N + 1, N + 1, dtype=torch.float64, device="cuda") * 1e-6
out = torch.empty(E, plan.nlat, plan.nlon, dtype=torch.float64, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(5):
    plan.synthesis(x, out=out)
torch.cuda.synchronize()
buf = np.zeros((2, 160, 24, 2), dtype=np.uint64)
pbuf = np.zeros((512, 4), dtype=np.uint64)
lib.gb_debug_trace_pack.restype = ctypes.c_int
lib.gb_debug_trace_pack.argtypes = [ctypes.c_void_p, ctypes.c_int]
names = {0: "stage1", 1: "stage2"}
for rep in range(reps):
    flush.zero_()
    torch.cuda.synchronize()
    lib.gb_debug_trace(None, 1)
    lib.gb_debug_trace_pack(None, 1)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(4):          # the launches of the traced (last) call are queued while the GPU is still busy
        flush.zero_()
        a.record(); plan.synthesis(x, out=out); b.record()
    torch.cuda.synchronize()
    lib.gb_debug_trace(buf.ctypes.data, 0)
    lib.gb_debug_trace_pack(pbuf.ctypes.data, 0)
    gt = buf[..., 0].astype(np.int64)
    t0 = gt[0, :, 0][gt[0, :, 0] > 0].min()
    print("rep %d: E=%d, %.2f us between the events" % (rep, E, 1e3 * a.elapsed_time(b)))
    pg = pbuf.astype(np.int64)
    for slot, what in enumerate(("entry", "after the wait", "rows in shared memory", "done")):
        v = pg[:, slot]
        v = v[v > 0]
        if v.size:
            r = (v - t0) / 1e3
            print("  pack   %-22s n=%3d  min %7.2f  med %7.2f  max %7.2f us" % (what, v.size, r.min(), np.median(r), r.max()))
    for k in (0, 1):
        for slot in range(24):
            v = gt[k, :, slot]
            v = v[v > 0]
            if v.size == 0:
                continue
            r = (v - t0) / 1e3
            print("  %s slot %2d  n=%3d  min %7.2f  med %7.2f  max %7.2f us" % (names[k], slot, v.size, r.min(), np.median(r), r.max()))
if reps:
    ck = buf[1, :, :, 1].astype(np.int64)
    ok = ck[:, 0] > 0
    rel = (ck[ok] - ck[ok][:, :1]) / 1.965e3      # us at 1965 MHz, per CTA relative to its entry
    rel[ck[ok] == 0] = np.nan
    print("stage 2, per CTA clock64 relative to its entry (us), median over CTAs with the slot:")
    for slot in range(24):
        col = rel[:, slot]
        if np.all(np.isnan(col)):
            continue
        print("  slot %2d  n=%3d  min %7.2f  med %7.2f  max %7.2f" % (slot, int(np.sum(~np.isnan(col))), np.nanmin(col), np.nanmedian(col), np.nanmax(col)))

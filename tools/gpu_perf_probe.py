"""Quick device-resident throughput probe (not the bench): blobs/s and per-kernel shares."""
import os, sys, time, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import raiko_b200 as rk
from raiko_b200 import _native

wb = int(sys.argv[1]) if len(sys.argv) > 1 else 15
ns = [int(a) for a in sys.argv[2:]] or [1024, 4096]
lib = _native.load()
peak = ctypes.c_double(); clk = ctypes.c_double()
lib.rk_measure_imad_peak(0, ctypes.byref(peak), ctypes.byref(clk))
print("IMAD.WIDE peak %.2f T MAC/s (clock attr %.0f MHz)" % (peak.value / 1e12, clk.value), flush=True)
t = time.time(); s = rk.KzgSettings(window_bits=wb); print("ctx c=%d table %.1f GB in %.1fs" % (s.window_bits, s.table_bytes / 1e9, time.time() - t), flush=True)
g = torch.Generator(device="cuda"); g.manual_seed(1)
for n in ns:
    blobs = torch.randint(0, 256, (n, 4096, 32), dtype=torch.uint8, device="cuda", generator=g)
    blobs[:, :, 0] %= 0x73
    outs = {k: torch.empty((n, w), dtype=torch.uint8, device="cuda") for k, w in (("c", 48), ("vh", 32), ("x", 32), ("y", 32), ("p", 48), ("st", 1))}
    def run(prove=True):
        if prove:
            st = lib.rk_commit_prove_batch(s._ctx, blobs.data_ptr(), n, outs["c"].data_ptr(), outs["vh"].data_ptr(), outs["x"].data_ptr(), outs["y"].data_ptr(), outs["p"].data_ptr(), outs["st"].data_ptr())
        else:
            st = lib.rk_commit_batch(s._ctx, blobs.data_ptr(), n, outs["c"].data_ptr(), outs["vh"].data_ptr(), outs["st"].data_ptr())
        assert st == 0, _native.last_error()
    for prove in (False, True):
        run(prove); torch.cuda.synchronize()
        s.stats_enable(True); s.stats_reset()
        t = time.time(); run(prove); torch.cuda.synchronize(); dt = time.time() - t
        st = s.stats(); s.stats_enable(False)
        adds = st["msm_point_adds"]
        print("n=%d %s: %.1f ms  %.0f blobs/s | msm %.1f ms (%d launches, %.2f G add/s) fr %.1f sha %.1f fin %.1f | launches %d" % (
            n, "commit+prove" if prove else "commit", dt * 1e3, n / dt, st["msm_ms"], st["msm_launches"], adds / st["msm_ms"] / 1e6 if st["msm_ms"] else 0,
            st["fr_ms"], st["sha_ms"], st["finalize_ms"], st["total_launches"]), flush=True)
    del blobs

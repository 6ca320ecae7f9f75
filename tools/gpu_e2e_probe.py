"""Host-pointer (pinned / pageable) vs device-pointer batches: wall time and kernel-time sums."""
import os, sys, time, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import raiko_b200 as rk
from raiko_b200 import _native
lib = _native.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 9472
s = rk.KzgSettings(window_bits=15)
g = torch.Generator(device="cuda"); g.manual_seed(1)
dev = torch.randint(0, 256, (n, 4096, 32), dtype=torch.uint8, device="cuda", generator=g); dev[:, :, 0] %= 0x73
pinned = torch.empty((n, 4096, 32), dtype=torch.uint8, pin_memory=True); pinned.copy_(dev)
pageable = pinned.clone()
def outs(device):
    kw = dict(device=device) if device == "cuda" else dict(pin_memory=True)
    return [torch.zeros((n, w), dtype=torch.uint8, **kw) for w in (48, 32, 32, 32, 48, 1)]
for name, src, o in (("device", dev, outs("cuda")), ("pinned", pinned, outs("cpu")), ("pageable", pageable, outs("cpu")), ("device", dev, outs("cuda"))):
    def run():
        st = lib.rk_commit_prove_batch(s._ctx, src.data_ptr(), n, *[t.data_ptr() for t in o])
        assert st == 0, _native.last_error()
    run(); torch.cuda.synchronize()
    s.stats_enable(True); s.stats_reset()
    t = time.perf_counter(); run(); torch.cuda.synchronize(); dt = time.perf_counter() - t
    st = s.stats(); s.stats_enable(False)
    print("%-9s n=%d: %.1f ms  %.0f blobs/s | msm %.1f fr %.1f sha %.1f fin %.1f | h2d %.2f GB" % (name, n, dt * 1e3, n / dt, st["msm_ms"], st["fr_ms"], st["sha_ms"], st["finalize_ms"], st["h2d_bytes"] / 1e9), flush=True)

"""Per-chunk timeline of one host-buffer rk_commit_prove_batch call (RAIKO_KZG_TRACE=1): where the
time of the end-to-end path goes (H2D, compute, D2H) for a given batch size.  Usage:
    RAIKO_KZG_TRACE=1 python tests/tools/e2e_trace.py <nblobs> [reps]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import raiko_b200 as rk
from raiko_b200 import _native
lib = _native.load()
n = int(sys.argv[1]); reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
s = rk.KzgSettings()
h_in = torch.empty((n, 131072), dtype=torch.uint8, pin_memory=True)
s.synth_blobs(h_in, first_blob=0)
outs = {k: torch.zeros((n, w), dtype=torch.uint8, pin_memory=True) for k, w in (("c", 48), ("vh", 32), ("x", 32), ("y", 32), ("p", 48), ("st", 1))}
def step():
    st = lib.rk_commit_prove_batch(s._ctx, h_in.data_ptr(), n, outs["c"].data_ptr(), outs["vh"].data_ptr(), outs["x"].data_ptr(), outs["y"].data_ptr(), outs["p"].data_ptr(), outs["st"].data_ptr())
    assert st == 0, _native.last_error()
step()
ts = []
for _ in range(reps):
    t = time.perf_counter(); step(); ts.append(time.perf_counter() - t)
print("n=%d wave_tail=%s host wall ms per call: %s -> %.0f blobs/s (best)" % (n, os.environ.get("RAIKO_KZG_WAVE_TAIL", "1"), ["%.1f" % (1e3 * t) for t in ts], n / min(ts)), flush=True)

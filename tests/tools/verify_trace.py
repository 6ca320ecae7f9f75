"""Stage trace of verify_blob_kzg_proof_batch on n blobs (RAIKO_KZG_VERIFY_TRACE=1 prints the device
time of every stage); also A/B of the lane-parallel vs single-thread pairing check
(RAIKO_KZG_PAIRING_LANES=0).  Usage: python tests/tools/verify_trace.py [n]"""
import ctypes, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import raiko_b200 as rk
from raiko_b200 import _native
lib = _native.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
s = rk.KzgSettings(window_bits=int(os.environ.get("RAIKO_KZG_WINDOW_BITS", "12")))
blobs = torch.zeros((n, 131072), dtype=torch.uint8, device="cuda")
s.synth_blobs(blobs, first_blob=0)
oc = torch.zeros((n, 48), dtype=torch.uint8, device="cuda"); ovh = torch.zeros((n, 32), dtype=torch.uint8, device="cuda")
op = torch.zeros((n, 48), dtype=torch.uint8, device="cuda"); ost = torch.zeros(n, dtype=torch.uint8, device="cuda")
assert lib.rk_commit_batch(s._ctx, blobs.data_ptr(), n, oc.data_ptr(), ovh.data_ptr(), ost.data_ptr()) == 0, _native.last_error()
assert lib.rk_compute_blob_kzg_proof_batch(s._ctx, blobs.data_ptr(), oc.data_ptr(), n, op.data_ptr(), ost.data_ptr()) == 0, _native.last_error()
craw, praw = oc.cpu().numpy().tobytes(), op.cpu().numpy().tobytes()
ok = ctypes.c_int(0)
def verify(p=praw):
    assert lib.rk_verify_blob_kzg_proof_batch(s._ctx, blobs.data_ptr(), ctypes.cast(ctypes.c_char_p(craw), ctypes.c_void_p),
                                              ctypes.cast(ctypes.c_char_p(p), ctypes.c_void_p), n, ctypes.byref(ok)) == 0, _native.last_error()
    return ok.value
verify()
ts = []
for _ in range(5):
    t = time.perf_counter(); r = verify(); ts.append(time.perf_counter() - t)
    assert r == 1
bad = bytearray(praw); bad[48 * 7:48 * 8] = praw[48 * 8:48 * 9]
assert verify(bytes(bad)) == 0
print("n=%d pairing_lanes=%s verify wall ms: %s  best %.2f ms = %.0f blobs/s; accept ok, reject-after-swap ok" %
      (n, os.environ.get("RAIKO_KZG_PAIRING_LANES", "2 (default: CTA-wide)"), ["%.1f" % (1e3 * t) for t in ts], 1e3 * min(ts), n / min(ts)), flush=True)

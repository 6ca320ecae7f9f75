"""BASELINE.json configs 2..5 on one GPU: latency / throughput table (not the bench line)."""
import os, sys, time, ctypes, json
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import raiko_b200 as rk
from raiko_b200 import _native
from kzg_testlib import synthetic_blob, SETUP
import kzg_ref, kzg_oracle

lib = _native.load()
wb = int(sys.argv[1]) if len(sys.argv) > 1 else 0
s = rk.KzgSettings(window_bits=wb)
ref = kzg_ref.RefSettings(open(SETUP, "rb").read())
out = {"window_bits": s.window_bits}

def timeit(f, reps=5):
    f(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t = time.perf_counter(); f(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t)
    ts.sort()
    return ts[0], ts[len(ts) // 2]

# config 2: Cancun-style block, 6 blobs, commitment + proof at the Fiat-Shamir challenge (host buffers, latency)
blobs6 = [synthetic_blob(b) for b in range(6)]
best, med = timeit(lambda: rk.commit_prove_batch(blobs6, s))
res = rk.commit_prove_batch(blobs6, s)
assert all((res.commitments[i], res.versioned_hashes[i], res.xs[i], res.ys[i], res.proofs[i]) == ref.commit_prove(blobs6[i]) for i in range(6))
out["config2_6blobs_commit_prove_ms"] = {"best": best * 1e3, "median": med * 1e3}
best, med = timeit(lambda: rk.calc_kzg_proof_commitment(blobs6[0], s))
out["single_blob_commitment_ms"] = {"best": best * 1e3, "median": med * 1e3}

# config 3: 4096 random blobs, commitment only, device resident
n = 4096
g = torch.Generator(device="cuda"); g.manual_seed(3)
blobs = torch.randint(0, 256, (n, 4096, 32), dtype=torch.uint8, device="cuda", generator=g); blobs[:, :, 0] %= 0x73
oc = torch.zeros((n, 48), dtype=torch.uint8, device="cuda"); ovh = torch.zeros((n, 32), dtype=torch.uint8, device="cuda"); ost = torch.zeros((n,), dtype=torch.uint8, device="cuda")
def commit():
    assert lib.rk_commit_batch(s._ctx, blobs.data_ptr(), n, oc.data_ptr(), ovh.data_ptr(), ost.data_ptr()) == 0
best, med = timeit(commit)
out["config3_4096_commit_blobs_per_s"] = {"best": n / best, "median": n / med}
for i in (0, 4095):
    assert oc[i].cpu().numpy().tobytes() == ref.commit(blobs[i].cpu().numpy().tobytes())

# config 5: verify_blob_kzg_proof_batch on those 4096 blobs (EIP-4844 challenges)
hb = blobs.cpu().numpy()
cs = [oc[i].cpu().numpy().tobytes() for i in range(n)]
zs = [kzg_oracle.fr_to_bytes(kzg_oracle.compute_challenge(hb[i].tobytes(), cs[i])) for i in range(n)]
zt = b"".join(zs)
op = torch.zeros((n, 48), dtype=torch.uint8, device="cuda"); oy = torch.zeros((n, 32), dtype=torch.uint8, device="cuda")
assert lib.rk_compute_kzg_proof_batch(s._ctx, blobs.data_ptr(), ctypes.cast(ctypes.c_char_p(zt), ctypes.c_void_p), n, op.data_ptr(), oy.data_ptr(), ost.data_ptr()) == 0
torch.cuda.synchronize()
craw = b"".join(cs); praw = op.cpu().numpy().tobytes()
ok = ctypes.c_int(0)
def verify(p=praw):
    assert lib.rk_verify_blob_kzg_proof_batch(s._ctx, blobs.data_ptr(), ctypes.cast(ctypes.c_char_p(craw), ctypes.c_void_p), ctypes.cast(ctypes.c_char_p(p), ctypes.c_void_p), n, ctypes.byref(ok)) == 0
best, med = timeit(verify, reps=3)
assert ok.value == 1
bad = bytearray(praw); bad[48 * 1234:48 * 1235] = praw[48 * 1235:48 * 1236]
verify(bytes(bad)); assert ok.value == 0
out["config5_4096_verify_batch_s"] = {"best": best, "median": med, "blobs_per_s": n / best, "accept": True, "reject_after_flip": True}
print(json.dumps(out, indent=1))

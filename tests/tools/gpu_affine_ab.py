"""A/B of the two MSM formulations (k_msm XYZZ vs k_msm_affine) on one GPU: outputs must be
byte-identical to each other and to the C oracle; prints MSM time and additions per second.
usage: gpu_affine_ab.py [window_bits] [n_blobs] [chains ...]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import raiko_b200 as rk
from raiko_b200 import _native
import kzg_ref
from kzg_testlib import SETUP

wb = int(sys.argv[1]) if len(sys.argv) > 1 else 15
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4736
chains = [a for a in sys.argv[3:]] or ["64"]     # "chains" or "chains:warps"
if os.environ.get("RAIKO_KZG_LIB_OVERRIDE"):      # A/B of differently compiled libraries (e.g. -DRK_MAC_FORM=1)
    _native.LIB_PATH = os.path.abspath(os.environ["RAIKO_KZG_LIB_OVERRIDE"])
    print("library:", _native.LIB_PATH)
lib = _native.load()
g = torch.Generator(device="cuda"); g.manual_seed(7)
blobs = torch.randint(0, 256, (n, 4096, 32), dtype=torch.uint8, device="cuda", generator=g)
blobs[:, :, 0] %= 0x73
ref = kzg_ref.RefSettings(open(SETUP, "rb").read())
want = {i: ref.commit_prove(blobs[i].cpu().numpy().tobytes()) for i in (0, n // 2, n - 1)}
results = {}
for mode in [("xyzz", 0, 0)] + [("affine", 1, k) for k in chains]:
    os.environ["RAIKO_KZG_MSM_AFFINE"] = str(mode[1])
    if mode[2]:
        os.environ["RAIKO_KZG_AFFINE_CHAINS"] = mode[2].split(":")[0]
        os.environ["RAIKO_KZG_AFFINE_WARPS"] = (mode[2].split(":") + ["8"])[1]
        os.environ["RAIKO_KZG_AFFINE_LOCKSTEP"] = (mode[2].split(":") + ["8", "1"])[2]
    s = rk.KzgSettings(window_bits=wb)
    outs = {k: torch.zeros((n, w), dtype=torch.uint8, device="cuda") for k, w in (("c", 48), ("vh", 32), ("x", 32), ("y", 32), ("p", 48), ("st", 1))}
    def run():
        st = lib.rk_commit_prove_batch(s._ctx, blobs.data_ptr(), n, outs["c"].data_ptr(), outs["vh"].data_ptr(), outs["x"].data_ptr(), outs["y"].data_ptr(), outs["p"].data_ptr(), outs["st"].data_ptr())
        assert st == 0, _native.last_error()
    run(); torch.cuda.synchronize()
    s.stats_enable(True); s.stats_reset()
    t = time.time(); run(); torch.cuda.synchronize(); dt = time.time() - t
    st = s.stats(); s.stats_enable(False)
    name = "%s%s" % (mode[0], "-" + mode[2] if mode[2] else "")
    print("%-10s c=%d n=%d: %.1f ms  %.0f blobs/s | msm %.1f ms (%d launches, %.3f G add/s) fr %.1f sha %.1f fin %.1f" % (
        name, s.window_bits, n, dt * 1e3, n / dt, st["msm_ms"], st["msm_launches"], st["msm_point_adds"] / st["msm_ms"] / 1e6,
        st["fr_ms"], st["sha_ms"], st["finalize_ms"]), flush=True)
    assert int(outs["st"].sum()) == 0
    for i, w in want.items():
        got = tuple(outs[k][i].cpu().numpy().tobytes() for k in ("c", "vh", "x", "y", "p"))
        assert got == w, "%s: blob %d differs from the oracle" % (name, i)
    results[name] = {k: v.clone() for k, v in outs.items()}
    s.close()
base = results["xyzz"]
for name, r in results.items():
    for k in base:
        assert torch.equal(base[k], r[k]), "%s differs from xyzz in %s" % (name, k)
print("all modes byte-identical on %d blobs; 3 checked against the oracle" % n)

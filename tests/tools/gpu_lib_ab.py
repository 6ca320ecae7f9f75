"""A/B of differently compiled libraries (tools/ab/build_variant.sh) on one GPU.
usage: gpu_lib_ab.py name[:warps] [name[:warps] ...]     (name 'base' = the regular raiko_b200/libraiko_kzg.so)
Each variant runs in its own process: commit+prove of n = 2 waves of one-warp-per-blob MSMs (n = 2 * 148 * warps
blobs, c = 15), twice; prints the MSM kernel time and additions per second of the second pass, checks three
blobs against the C oracle and prints a digest of all outputs (equal digests = byte-identical variants when n is equal;
the first 4736 blobs are digested separately so that variants with different n can be compared too)."""
import hashlib, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def child(name, warps):
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    import torch
    import raiko_b200 as rk
    from raiko_b200 import _native
    import kzg_ref
    from kzg_testlib import SETUP
    if name != "base":
        _native.LIB_PATH = os.path.join(ROOT, "raiko_b200", "ab", "libraiko_kzg_%s.so" % name)
    lib = _native.load()
    n = 2 * 148 * warps
    s = rk.KzgSettings(window_bits=int(os.environ.get("AB_WINDOW_BITS", "15")))
    blobs = torch.empty((n, 4096, 32), dtype=torch.uint8, device="cuda")
    s.synth_blobs(blobs, first_blob=0, seed=20241018)
    outs = {k: torch.zeros((n, w), dtype=torch.uint8, device="cuda") for k, w in (("c", 48), ("vh", 32), ("x", 32), ("y", 32), ("p", 48), ("st", 1))}

    def run():
        st = lib.rk_commit_prove_batch(s._ctx, blobs.data_ptr(), n, outs["c"].data_ptr(), outs["vh"].data_ptr(), outs["x"].data_ptr(),
                                       outs["y"].data_ptr(), outs["p"].data_ptr(), outs["st"].data_ptr())
        assert st == 0, _native.last_error()
    run(); torch.cuda.synchronize()
    best = None
    for _ in range(int(os.environ.get("AB_REPS", "2"))):
        s.stats_enable(True); s.stats_reset()
        t = time.time(); run(); torch.cuda.synchronize(); dt = time.time() - t
        st = s.stats(); s.stats_enable(False)
        if best is None or st["msm_ms"] < best[1]["msm_ms"]:
            best = (dt, st)
    dt, st = best
    assert int(outs["st"].sum()) == 0
    ref = kzg_ref.RefSettings(open(SETUP, "rb").read())
    for i in (0, n // 2, n - 1):
        got = tuple(outs[k][i].cpu().numpy().tobytes() for k in ("c", "vh", "x", "y", "p"))
        assert got == ref.commit_prove(blobs[i].cpu().numpy().tobytes()), "%s: blob %d differs from the oracle" % (name, i)
    dig = hashlib.sha256(b"".join(outs[k][:4736].cpu().numpy().tobytes() for k in ("c", "vh", "x", "y", "p"))).hexdigest()[:16]
    print("%-12s warps %2d n %5d chunk %5s: %7.1f ms %6.0f blobs/s | msm %7.2f ms (%d launches) %.3f G add/s | fr %.1f sha %.1f fin %.1f | oracle ok, digest[:4736] %s"
          % (name, warps, n, os.environ.get("RAIKO_KZG_CHUNK", "-"), dt * 1e3, n / dt, st["msm_ms"], st["msm_launches"],
             st["msm_point_adds"] / st["msm_ms"] / 1e6, st["fr_ms"], st["sha_ms"], st["finalize_ms"], dig), flush=True)
    s.close()


if __name__ == "__main__":
    if sys.argv[1] == "--child":
        child(sys.argv[2], int(sys.argv[3]))
        sys.exit(0)
    for spec in sys.argv[1:]:
        name, warps = (spec.split(":") + ["16"])[:2]
        env = dict(os.environ, RAIKO_KZG_CHUNK=str(2 * 148 * int(warps)))
        p = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", name, warps], env=env, capture_output=True, text=True, timeout=600)
        out = [l for l in p.stdout.splitlines() if l.strip()]
        print(out[-1] if p.returncode == 0 and out else "%-12s FAILED rc=%d: %s" % (name, p.returncode, (p.stderr or p.stdout)[-600:]), flush=True)

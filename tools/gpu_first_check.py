"""First on-device check: golden vectors through the C ABI at a small and a large window."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import raiko_b200 as rk
from kzg_testlib import blob_from_recipe, load_golden

g = load_golden()
for wb in [int(a) for a in sys.argv[1:]] or [8]:
    t = time.time()
    s = rk.KzgSettings(window_bits=wb)
    print("window_bits", s.window_bits, "table GB %.2f" % (s.table_bytes / 1e9), "create %.2fs" % (time.time() - t), flush=True)
    bad = 0
    blobs = [blob_from_recipe(c["recipe"]) for c in g["cases"]]
    t = time.time()
    res = rk.commit_prove_batch(blobs, s)
    print("batch of %d: %.1f ms" % (len(blobs), 1e3 * (time.time() - t)))
    for i, c in enumerate(g["cases"]):
        p0 = c["proofs"][0]
        ok = (res.commitments[i].hex() == c["commitment"] and res.versioned_hashes[i].hex() == c["versioned_hash"]
              and res.xs[i].hex() == p0["z"] and res.ys[i].hex() == p0["y"] and res.proofs[i].hex() == p0["proof"] and res.status[i] == 0)
        print(c["name"], "OK" if ok else "MISMATCH", flush=True)
        if not ok:
            bad += 1
            print("  C ", res.commitments[i].hex(), c["commitment"]); print("  vh", res.versioned_hashes[i].hex(), c["versioned_hash"])
            print("  x ", res.xs[i].hex(), p0["z"]); print("  y ", res.ys[i].hex(), p0["y"]); print("  pi", res.proofs[i].hex(), p0["proof"])
        for pr in c["proofs"][1:]:
            proof, y = rk.compute_kzg_proof(blobs[i], bytes.fromhex(pr["z"]), s)
            ok = proof.hex() == pr["proof"] and y.hex() == pr["y"]
            print("  ", c["name"], pr["label"], "OK" if ok else "MISMATCH")
            bad += 0 if ok else 1
    for e in g["errors"]:
        try:
            rk.calc_kzg_proof_commitment(blob_from_recipe(e["recipe"]), s)
            print(e["name"], "NO ERROR (bad)"); bad += 1
        except rk.DeserializeBlob:
            print(e["name"], "OK (DeserializeBlob)")
    # single-call API
    b = blobs[2]
    c = rk.calc_kzg_proof_commitment(b, s); vh = rk.commitment_to_version_hash(c)
    x, y = rk.proof_of_equivalence(b, vh, s)
    pr = rk.calc_kzg_proof(b, vh, s)
    print("single", c.hex() == g["cases"][2]["commitment"], x.hex() == g["cases"][2]["proofs"][0]["z"], y.hex() == g["cases"][2]["proofs"][0]["y"], pr.hex() == g["cases"][2]["proofs"][0]["proof"], rk.get_evaluation_point(b, vh, s).hex() == g["cases"][2]["proofs"][0]["z"])
    print("MISMATCHES", bad, flush=True)
    s.close()

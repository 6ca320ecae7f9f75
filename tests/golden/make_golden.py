#!/usr/bin/env python3
"""Generate tests/golden/kzg_golden.json with the Python oracle.

Run in the build container (needs nothing but the repo):
    python tests/golden/make_golden.py
Every (C, z, y, proof) row is accepted by the oracle's pairing check
(``verify_kzg_proof``, the equation eip4844.rs:176-183 tests) before it is
written; rows C1..C7 equal SURVEY.md Appendix C.  Blobs are stored as recipes
(see ``blob_from_recipe`` in tests/kzg_testlib.py), not as 128 KiB payloads.
"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.join(HERE, "..", "..")
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import kzg_oracle as o  # noqa: E402
from kzg_testlib import blob_from_recipe  # noqa: E402

s = o.load_settings(open(os.path.join(ROOT, "raiko_b200", "data", "trusted_setup_4096.bin"), "rb").read())

CASES = [
    ("C1_zero", {"kind": "zero"}, []),
    ("C2_ones", {"kind": "const", "value": 1}, []),
    ("C3_mod64", {"kind": "mod64"},
     [("fs5", o.hash_to_bls_field(bytes([5] * 32))), ("omega5", s.roots_brp[5]),
      ("fs6", o.hash_to_bls_field(bytes([6] * 32)))]),
    ("C4_sparse", {"kind": "sparse", "index": 7, "value": 9}, []),
    ("C5_syn0", {"kind": "synthetic", "seed": 20241018, "b": 0}, []),
    ("C6_syn1", {"kind": "synthetic", "seed": 20241018, "b": 1}, [("omega4095", s.roots_brp[4095])]),
    ("S2_syn2", {"kind": "synthetic", "seed": 20241018, "b": 2}, [("zero", 0), ("one", 1)]),
    ("S3_syn3", {"kind": "synthetic", "seed": 20241018, "b": 3}, []),
    ("S4_syn4", {"kind": "synthetic", "seed": 20241018, "b": 4}, []),
    ("S5_syn5", {"kind": "synthetic", "seed": 20241018, "b": 5}, []),
    ("M1_max", {"kind": "const", "value": o.R - 1}, [("minus1", o.R - 1)]),
    ("K7_const7", {"kind": "const", "value": 7}, []),
]

out = {"settings_sha256": hashlib.sha256(open(os.path.join(ROOT, "raiko_b200", "data", "trusted_setup_4096.bin"), "rb").read()).hexdigest(),
       "cases": [], "errors": [
           {"name": "C7_noncanonical_first", "recipe": {"kind": "noncanonical", "index": 0, "bytes": "ff" * 32}},
           {"name": "E2_equal_r", "recipe": {"kind": "noncanonical", "index": 4095, "bytes": "%064x" % o.R}},
           {"name": "E3_r_plus_1_mid", "recipe": {"kind": "noncanonical", "index": 2048, "bytes": "%064x" % (o.R + 1)}},
       ]}
for name, recipe, extra in CASES:
    blob = blob_from_recipe(recipe)
    c = o.blob_to_kzg_commitment(blob, s)
    vh = o.commitment_to_version_hash(c)
    x = o.get_evaluation_point(blob, vh)
    proofs = []
    for label, z in [("raiko", x)] + extra:
        pr, y = o.compute_kzg_proof(blob, z, s)
        ok = o.verify_kzg_proof(c, z, int.from_bytes(y, "big"), pr, s)
        assert ok, (name, label)
        proofs.append({"label": label, "z": o.fr_to_bytes(z).hex(), "y": y.hex(), "proof": pr.hex(), "pairing_ok": ok})
    xb, yb = o.proof_of_equivalence(blob, vh, s)
    assert xb.hex() == proofs[0]["z"] and yb.hex() == proofs[0]["y"]
    out["cases"].append({"name": name, "recipe": recipe, "blob_sha256": hashlib.sha256(blob).hexdigest(),
                         "commitment": c.hex(), "versioned_hash": vh.hex(), "proofs": proofs})
    print(name, c.hex()[:16], "ok", flush=True)
json.dump(out, open(os.path.join(HERE, "kzg_golden.json"), "w"), indent=1)

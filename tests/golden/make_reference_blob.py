#!/usr/bin/env python3
"""Extracts the only real blob the reference ships -- `valid_blob_str` of
core/src/preflight.rs:480-524 (test_new_blob_decode), right-padded with '0' to 262 144 hex digits
exactly as :525 does -- and writes it, with what the oracle derives from it, as a golden fixture:

    tests/golden/reference_blob_preflight.json
      hex_prefix        the non-zero prefix of the blob (the rest is '0' padding)
      decoded_len       1200 (the 3-byte length header 0x0004b0)
      decoded_sha256    sha256 of decode_blob_data(blob)              (oracle restatement of utils.rs:85-144)
      commitment, versioned_hash, x, y, proof                          (oracle/kzg_oracle.py, pairing-checked)

Run in the build container (needs /root/reference); the GPU box only reads the JSON."""
import hashlib
import json
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import kzg_oracle as o  # noqa: E402
from kzg_testlib import SETUP  # noqa: E402

src = open("/root/reference/core/src/preflight.rs").read().splitlines()
lines = src[480 - 1:524]                                   # file lines 480..524
assert "valid_blob_str" in lines[0], lines[0]
hexstr = "".join(re.sub(r'[^0-9a-f]', "", l.split("=")[-1].replace("\\", "").replace('"', "").replace(";", "")) for l in lines)
assert hexstr.startswith("01000004b0f904ad") and len(hexstr) % 2 == 0, hexstr[:32]
blob = bytes.fromhex(hexstr.ljust(262144, "0"))
assert len(blob) == 131072
dec = o.decode_blob_data(blob)
assert len(dec) == 1200 and dec[:3].hex() == "f904ad" and int.from_bytes(dec[1:3], "big") + 3 == 1200    # a self-describing RLP list
s = o.load_settings(open(SETUP, "rb").read())
c = o.calc_kzg_proof_commitment(blob, s)
vh = o.commitment_to_version_hash(c)
x, y = o.proof_of_equivalence(blob, vh, s)
proof = o.calc_kzg_proof(blob, vh, s)
assert o.verify_kzg_proof(c, int.from_bytes(x, "big"), int.from_bytes(y, "big"), proof, s)
out = {
    "source": "core/src/preflight.rs:480-524 (valid_blob_str), padded as at :525",
    "hex_prefix": hexstr.rstrip("0") if len(hexstr.rstrip("0")) % 2 == 0 else hexstr.rstrip("0") + "0",
    "blob_sha256": hashlib.sha256(blob).hexdigest(),
    "decoded_len": len(dec), "decoded_sha256": hashlib.sha256(dec).hexdigest(), "decoded_prefix": dec[:8].hex(),
    "commitment": c.hex(), "versioned_hash": vh.hex(), "x": x.hex(), "y": y.hex(), "proof": proof.hex(), "pairing_ok": True,
}
json.dump(out, open(os.path.join(HERE, "reference_blob_preflight.json"), "w"), indent=1)
print(json.dumps({k: (v if len(str(v)) < 100 else str(v)[:60] + "...") for k, v in out.items()}, indent=1))

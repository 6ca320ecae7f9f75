"""Shared helpers for the test-suite (blob recipes, paths). Test infrastructure."""
import hashlib
import json
import os
import struct

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
GOLDEN = os.path.join(ROOT, "tests", "golden", "kzg_golden.json")
SETUP = os.path.join(ROOT, "raiko_b200", "data", "trusted_setup_4096.bin")
R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
BYTES_PER_BLOB = 131072


def synthetic_blob(b: int, seed: int = 20241018) -> bytes:
    """SURVEY.md §8(d) synthetic blob (b, i) -> sha256(tag|seed|b|i) mod r."""
    pre = b"raiko-kzg-bench-v1" + struct.pack("<Q", seed) + struct.pack("<I", b)
    out = bytearray()
    for i in range(4096):
        v = int.from_bytes(hashlib.sha256(pre + struct.pack("<I", i)).digest(), "big") % R
        out += v.to_bytes(32, "big")
    return bytes(out)


def blob_from_recipe(rc: dict) -> bytes:
    k = rc["kind"]
    if k == "zero":
        return bytes(BYTES_PER_BLOB)
    if k == "const":
        return int(rc["value"]).to_bytes(32, "big") * 4096
    if k == "mod64":  # eip4844.rs:165
        return bytes(v % 64 for v in range(BYTES_PER_BLOB))
    if k == "sparse":
        b = bytearray(BYTES_PER_BLOB)
        b[32 * rc["index"]:32 * rc["index"] + 32] = int(rc["value"]).to_bytes(32, "big")
        return bytes(b)
    if k == "synthetic":
        return synthetic_blob(rc["b"], rc["seed"])
    if k == "noncanonical":
        b = bytearray(synthetic_blob(9, 7))
        b[32 * rc["index"]:32 * rc["index"] + 32] = bytes.fromhex(rc["bytes"])
        return bytes(b)
    raise ValueError(k)


def load_golden() -> dict:
    with open(GOLDEN) as f:
        return json.load(f)

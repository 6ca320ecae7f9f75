"""The C-ABI library loads and exports every symbol include/raiko_kzg.h declares; host-only
entry points work; compute entry points fail loudly (no CPU fallback) without a GPU."""
import ctypes
import hashlib
import os
import re

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def header_functions():
    txt = open(os.path.join(ROOT, "include", "raiko_kzg.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(rk_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from raiko_b200 import _native
    lib = _native.load()
    names = header_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "libraiko_kzg.so does not export %s" % n
    assert set(names) == set(_native.SIGNATURES), set(names) ^ set(_native.SIGNATURES)
    assert b"sm_100a" in lib.rk_version()


def test_versioned_hash_host_entry_point():
    import raiko_b200 as rk
    c = b"\xc0" + bytes(47)
    assert rk.commitment_to_version_hash(c).hex() == "010657f37554c781402a22917dee2f75def7ab966d7b770905398eba3c444014"
    c = bytes(range(48))
    assert rk.kzg_to_versioned_hash(c) == b"\x01" + hashlib.sha256(c).digest()[1:]
    with pytest.raises(ValueError):
        rk.commitment_to_version_hash(b"short")


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import raiko_b200 as rk
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        rk.KzgSettings()


def test_product_never_imports_oracle():
    """The shipped package, its headers, bindings and generators must not reference oracle/
    (parity claims depend on it).  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
    / reference legs may."""
    for top in ("raiko_b200", "include", "bindings", "tools"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp", ".rs", ".toml")):
                    src = open(os.path.join(dirpath, f), errors="replace").read()
                    assert "kzg_oracle" not in src and "kzg_ref" not in src and "oracle/" not in src and '"oracle"' not in src, os.path.join(dirpath, f)


def test_python_wrapper_argument_checks():
    import raiko_b200 as rk
    from raiko_b200 import eip4844
    with pytest.raises(rk.DeserializeBlob):
        eip4844._blobs_buf(bytes(131072 + 1))
    buf, n = eip4844._blobs_buf([bytes(131072), bytes(131072)])
    assert n == 2 and buf.nbytes == 2 * 131072
    assert rk.kzg_proof_to_bytes(bytes(48)) == bytes(48)
    assert issubclass(rk.DeserializeBlob, rk.Eip4844Error)
    assert str(rk.DeserializeBlob()) == "Failed to deserialize blob to field elements"   # eip4844.rs:34

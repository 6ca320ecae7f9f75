"""ctypes wrapper of oracle/kzg_ref.c (TEST INFRASTRUCTURE: the fast CPU checker and the
timed CPU baseline).  Never imported by raiko_b200/."""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libkzgref.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "kzg_ref.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE] + (["-B"] if force else []))
    return _SO


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(build())
        vp, cp, sz = ctypes.c_void_p, ctypes.c_char_p, ctypes.c_size_t
        L.kzgref_load.restype = vp
        L.kzgref_load.argtypes = [cp, sz]
        L.kzgref_free.argtypes = [vp]
        L.kzgref_sha256.argtypes = [cp, sz, cp]
        L.kzgref_versioned_hash.argtypes = [cp, cp]
        L.kzgref_evaluation_point.argtypes = [cp, cp, cp]
        for name, args in (("kzgref_commit", [vp, cp, sz, cp]),
                           ("kzgref_proof_of_equivalence", [vp, cp, sz, cp, cp, cp]),
                           ("kzgref_compute_proof", [vp, cp, sz, cp, cp, cp]),
                           ("kzgref_commit_prove", [vp, cp, sz, cp, cp, cp, cp, cp])):
            fn = getattr(L, name)
            fn.restype = ctypes.c_int
            fn.argtypes = args
        _lib = L
    return _lib


class RefError(ValueError):
    """Eip4844Error::DeserializeBlob (eip4844.rs:34-35)."""


class RefSettings:
    def __init__(self, data: bytes):
        self._h = lib().kzgref_load(data, len(data))
        if not self._h:
            raise ValueError("kzg_ref: cannot load settings image")

    def _chk(self, rc):
        if rc:
            raise RefError("deserialize blob failed (rc=%d)" % rc)

    def commit(self, blob: bytes) -> bytes:
        out = ctypes.create_string_buffer(48)
        self._chk(lib().kzgref_commit(self._h, blob, len(blob), out))
        return out.raw

    def proof_of_equivalence(self, blob: bytes, vh: bytes):
        x, y = ctypes.create_string_buffer(32), ctypes.create_string_buffer(32)
        self._chk(lib().kzgref_proof_of_equivalence(self._h, blob, len(blob), vh, x, y))
        return x.raw, y.raw

    def compute_proof(self, blob: bytes, z: bytes):
        pr, y = ctypes.create_string_buffer(48), ctypes.create_string_buffer(32)
        self._chk(lib().kzgref_compute_proof(self._h, blob, len(blob), z, pr, y))
        return pr.raw, y.raw

    def commit_prove(self, blob: bytes):
        c, vh, x, y, pr = (ctypes.create_string_buffer(n) for n in (48, 32, 32, 32, 48))
        self._chk(lib().kzgref_commit_prove(self._h, blob, len(blob), c, vh, x, y, pr))
        return c.raw, vh.raw, x.raw, y.raw, pr.raw


def versioned_hash(c: bytes) -> bytes:
    out = ctypes.create_string_buffer(32)
    lib().kzgref_versioned_hash(c, out)
    return out.raw


def evaluation_point(blob: bytes, vh: bytes) -> bytes:
    out = ctypes.create_string_buffer(32)
    lib().kzgref_evaluation_point(blob, vh, out)
    return out.raw


def sha256(msg: bytes) -> bytes:
    out = ctypes.create_string_buffer(32)
    lib().kzgref_sha256(msg, len(msg), out)
    return out.raw

"""Python mirror of ``lib/src/primitives/eip4844.rs`` over the CUDA C ABI.

Same names, argument meaning and error behaviour as the reference wrapper
(file:line cited per function), plus the c-kzg era spellings BASELINE.json's
north_star uses (``KzgSettings``, ``blob_to_kzg_commitment``, ``compute_kzg_proof``,
``kzg_to_versioned_hash``) and the batch calls the GPU is built for.  Everything is
computed by ``libraiko_kzg.so`` on the GPU; nothing here does field arithmetic.
"""
from __future__ import annotations

import ctypes
import os
import threading
from typing import List, Optional, Sequence, Tuple

from . import _native

__all__ = [
    "VERSIONED_HASH_VERSION_KZG", "BYTES_PER_BLOB", "Eip4844Error", "DeserializeBlob", "EvaluatePolynomial",
    "ComputeKzgProof", "KzgSettings", "KZGSettings", "kzg_settings", "set_kzg_settings", "get_evaluation_point",
    "proof_of_equivalence", "calc_kzg_proof", "calc_kzg_proof_with_point", "calc_kzg_proof_commitment",
    "commitment_to_version_hash", "kzg_proof_to_bytes", "blob_to_kzg_commitment", "compute_kzg_proof",
    "kzg_to_versioned_hash", "verify_kzg_proof", "verify_kzg_proof_batch", "verify_blob_kzg_proof_batch", "commit_batch", "commit_prove_batch", "compute_kzg_proof_batch", "compute_blob_kzg_proof_batch", "BatchResult",
    "DEFAULT_SETTINGS_PATH",
]

VERSIONED_HASH_VERSION_KZG = 0x01  # eip4844.rs:26
BYTES_PER_BLOB = 131072
DEFAULT_SETTINGS_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "trusted_setup_4096.bin")


class Eip4844Error(Exception):
    """eip4844.rs:32-42"""


class DeserializeBlob(Eip4844Error):
    def __init__(self, msg="Failed to deserialize blob to field elements"):
        super().__init__(msg)


class EvaluatePolynomial(Eip4844Error):
    def __init__(self, msg):
        super().__init__("Failed to evaluate polynomial at hashed point: %s" % msg)


class ComputeKzgProof(Eip4844Error):
    def __init__(self, msg=""):
        super().__init__("Failed to compute KZG proof" + (": " + msg if msg else ""))


def _check(st: int):
    if st == _native.RK_OK:
        return
    msg = _native.last_error()
    if st in (_native.RK_ERR_BAD_LENGTH, _native.RK_ERR_NONCANONICAL_FE):
        raise DeserializeBlob()
    if st == _native.RK_ERR_CUDA:
        raise RuntimeError("raiko_b200 CUDA failure (no CPU fallback): " + msg)
    if st == _native.RK_ERR_BAD_SETTINGS:
        raise ValueError("failed to load trusted setup: " + msg)
    if st == _native.RK_ERR_BAD_POINT:
        raise ValueError("invalid G1 point: " + msg)
    raise ValueError("raiko_b200: status %d: %s" % (st, msg))


class _Buf:
    """Read-only pointer view of bytes / bytearray / numpy array / torch tensor (host or CUDA)."""

    def __init__(self, obj):
        self.keep = obj
        if hasattr(obj, "data_ptr"):  # torch tensor
            if not obj.is_contiguous():
                raise ValueError("tensor must be contiguous")
            self.ptr = obj.data_ptr()
            self.nbytes = obj.numel() * obj.element_size()
        elif hasattr(obj, "ctypes") and hasattr(obj, "nbytes"):  # numpy
            if not obj.flags["C_CONTIGUOUS"]:
                raise ValueError("array must be C-contiguous")
            self.ptr = obj.ctypes.data
            self.nbytes = obj.nbytes
        elif isinstance(obj, bytes):
            self.ptr = ctypes.cast(ctypes.c_char_p(obj), ctypes.c_void_p).value
            self.nbytes = len(obj)
        elif isinstance(obj, (bytearray, memoryview)):
            b = bytes(obj)
            self.keep = b
            self.ptr = ctypes.cast(ctypes.c_char_p(b), ctypes.c_void_p).value
            self.nbytes = len(b)
        else:
            raise TypeError("unsupported buffer type %r" % type(obj))


def _cptr(b: bytes):
    return ctypes.cast(ctypes.c_char_p(b), ctypes.c_void_p)


class KzgSettings:
    """KZGSettings (eip4844.rs:13,21-24): the trusted setup, resident on the GPU(s) as the
    fixed-base window table.  ``settings`` may be the bytes of kzg_settings_raw.bin, of the
    bincode image eip4844.rs:14 embeds, or of this package's compact image (default)."""

    def __init__(self, settings: Optional[bytes] = None, devices: Optional[Sequence[int]] = None, window_bits: int = 0):
        lib = _native.load()
        if settings is None:
            with open(DEFAULT_SETTINGS_PATH, "rb") as f:
                settings = f.read()
        self._lib = lib
        self._ctx = ctypes.c_void_p()
        devs = list(devices) if devices is not None else []
        arr = (ctypes.c_int * len(devs))(*devs) if devs else None
        st = lib.rk_kzg_ctx_create_ex(_cptr(settings), len(settings), arr, len(devs), int(window_bits),
                                      ctypes.byref(self._ctx))
        _check(st)

    @classmethod
    def load_trusted_setup_file(cls, path: str, **kw) -> "KzgSettings":
        with open(path, "rb") as f:
            return cls(f.read(), **kw)

    @property
    def window_bits(self) -> int:
        return self._lib.rk_kzg_ctx_window_bits(self._ctx)

    @property
    def num_devices(self) -> int:
        return self._lib.rk_kzg_ctx_num_devices(self._ctx)

    @property
    def table_bytes(self) -> int:
        return int(self._lib.rk_kzg_ctx_table_bytes(self._ctx))

    @property
    def window_reduced(self) -> bool:
        """True when the library had to pick a window narrower than 15 bits for lack of free HBM."""
        return bool(self._lib.rk_kzg_ctx_window_reduced(self._ctx))

    def synth_blobs(self, out, first_blob: int = 0, seed: int = 20241018):
        """Fill `out` (a uint8 torch tensor / numpy array of n * 131072 bytes, host or CUDA) with the
        SURVEY.md 8(d) synthetic blobs first_blob .. first_blob + n - 1 (bench / test tooling)."""
        b = _Buf(out)
        if b.nbytes % BYTES_PER_BLOB:
            raise ValueError("output must be a whole number of blobs")
        _check(self._lib.rk_synth_blobs(self._ctx, int(seed), int(first_blob), b.nbytes // BYTES_PER_BLOB, b.ptr))
        return out

    def export(self, kind: str = "bincode") -> bytes:
        """Re-serialise in the reference's layouts (host/src/bin/gen_kzg_settings.rs:8-22)."""
        k = {"raw": 0, "bincode": 1}[kind]
        n = ctypes.c_size_t(1_001_905)
        buf = ctypes.create_string_buffer(n.value)
        _check(self._lib.rk_kzg_ctx_export_settings(self._ctx, k, buf, ctypes.byref(n)))
        return buf.raw[:n.value]

    def close(self):
        if getattr(self, "_ctx", None):
            self._lib.rk_kzg_ctx_destroy(self._ctx)
            self._ctx = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def stats_enable(self, on: bool = True):
        self._lib.rk_kzg_stats_enable(self._ctx, 1 if on else 0)

    def stats_reset(self):
        self._lib.rk_kzg_stats_reset(self._ctx)

    def stats(self) -> dict:
        s = _native.RkKzgStats()
        self._lib.rk_kzg_stats_get(self._ctx, ctypes.byref(s))
        return {name: getattr(s, name) for name, _ in s._fields_}


KZGSettings = KzgSettings  # the reference's spelling (eip4844.rs:13)

_settings_lock = threading.Lock()
_settings: Optional[KzgSettings] = None


def kzg_settings() -> KzgSettings:
    """KZG_SETTINGS (eip4844.rs:21-24): process-wide, initialised on first use."""
    global _settings
    with _settings_lock:
        if _settings is None:
            _settings = KzgSettings()
        return _settings


def set_kzg_settings(s: Optional[KzgSettings]):
    global _settings
    with _settings_lock:
        _settings = s


def _s(settings: Optional[KzgSettings]) -> KzgSettings:
    return settings if settings is not None else kzg_settings()


# ---------------------------------------------------------------------------------------
# The reference's functions
# ---------------------------------------------------------------------------------------
def commitment_to_version_hash(commitment: bytes) -> bytes:
    """eip4844.rs:91-95"""
    if len(commitment) != 48:
        raise ValueError("commitment must be 48 bytes")
    out = ctypes.create_string_buffer(32)
    _check(_native.load().rk_kzg_to_versioned_hash(_cptr(bytes(commitment)), out))
    return out.raw


kzg_to_versioned_hash = commitment_to_version_hash


def calc_kzg_proof_commitment(blob, settings: Optional[KzgSettings] = None) -> bytes:
    """eip4844.rs:80-89 -> 48-byte compressed commitment; DeserializeBlob on bad blobs."""
    s = _s(settings)
    b = _Buf(blob)
    out = ctypes.create_string_buffer(48)
    _check(s._lib.rk_blob_to_kzg_commitment(s._ctx, b.ptr, b.nbytes, out))
    return out.raw


blob_to_kzg_commitment = calc_kzg_proof_commitment


def get_evaluation_point(blob, versioned_hash: bytes, settings: Optional[KzgSettings] = None) -> bytes:
    """eip4844.rs:44-48 -> x as 32 big-endian bytes (the reference returns ZFr)."""
    s = _s(settings)
    b = _Buf(blob)
    if len(versioned_hash) != 32:
        raise ValueError("versioned_hash must be 32 bytes")
    out = ctypes.create_string_buffer(32)
    _check(s._lib.rk_get_evaluation_point(s._ctx, b.ptr, b.nbytes, _cptr(bytes(versioned_hash)), out))
    return out.raw


def proof_of_equivalence(blob, versioned_hash: bytes, settings: Optional[KzgSettings] = None) -> Tuple[bytes, bytes]:
    """eip4844.rs:50-65 -> (x, y), 32 big-endian bytes each."""
    s = _s(settings)
    b = _Buf(blob)
    if len(versioned_hash) != 32:
        raise ValueError("versioned_hash must be 32 bytes")
    x = ctypes.create_string_buffer(32)
    y = ctypes.create_string_buffer(32)
    _check(s._lib.rk_proof_of_equivalence(s._ctx, b.ptr, b.nbytes, _cptr(bytes(versioned_hash)), x, y))
    return x.raw, y.raw


def compute_kzg_proof(blob, z: bytes, settings: Optional[KzgSettings] = None) -> Tuple[bytes, bytes]:
    """compute_kzg_proof_rust as called at eip4844.rs:75 -> (proof48, y32)."""
    s = _s(settings)
    b = _Buf(blob)
    if len(z) != 32:
        raise ValueError("z must be 32 bytes")
    proof = ctypes.create_string_buffer(48)
    y = ctypes.create_string_buffer(32)
    _check(s._lib.rk_compute_kzg_proof(s._ctx, b.ptr, b.nbytes, _cptr(bytes(z)), proof, y))
    return proof.raw, y.raw


def calc_kzg_proof_with_point(blob, z: bytes, settings: Optional[KzgSettings] = None) -> bytes:
    """eip4844.rs:71-78 (the reference returns ZG1; here its 48-byte encoding)."""
    return compute_kzg_proof(blob, z, settings)[0]


def calc_kzg_proof(blob, versioned_hash: bytes, settings: Optional[KzgSettings] = None) -> bytes:
    """eip4844.rs:67-69"""
    s = _s(settings)
    b = _Buf(blob)
    if len(versioned_hash) != 32:
        raise ValueError("versioned_hash must be 32 bytes")
    proof = ctypes.create_string_buffer(48)
    _check(s._lib.rk_calc_kzg_proof(s._ctx, b.ptr, b.nbytes, _cptr(bytes(versioned_hash)), proof))
    return proof.raw


def kzg_proof_to_bytes(proof: bytes) -> bytes:
    """eip4844.rs:97-99: proofs already travel as their 48-byte encoding here."""
    if len(proof) != 48:
        raise ValueError("proof must be 48 bytes")
    return bytes(proof)


def verify_kzg_proof(commitment: bytes, z: bytes, y: bytes, proof: bytes, settings: Optional[KzgSettings] = None) -> bool:
    """verify_kzg_proof_rust as the reference's tests call it (eip4844.rs:176-183)."""
    s = _s(settings)
    if (len(commitment), len(z), len(y), len(proof)) != (48, 32, 32, 48):
        raise ValueError("expected commitment48, z32, y32, proof48")
    ok = ctypes.c_int(0)
    _check(s._lib.rk_verify_kzg_proof(s._ctx, _cptr(bytes(commitment)), _cptr(bytes(z)), _cptr(bytes(y)), _cptr(bytes(proof)),
                                      ctypes.byref(ok)))
    return bool(ok.value)


def verify_kzg_proof_batch(commitments, zs, ys, proofs, settings: Optional[KzgSettings] = None) -> bool:
    """verify_kzg_proof_batch (Deneb spec): n (C, z, y, proof) tuples under one random linear
    combination and two pairings.  Each argument is a sequence of byte strings, one contiguous
    bytes-like object, or a CUDA tensor (device pointers go straight through the C ABI)."""
    s = _s(settings)

    def flat(v):
        return _Buf(b"".join(bytes(e) for e in v) if isinstance(v, (list, tuple)) else v)
    c, z, y, p = flat(commitments), flat(zs), flat(ys), flat(proofs)
    n = c.nbytes // 48
    if (c.nbytes, z.nbytes, y.nbytes, p.nbytes) != (48 * n, 32 * n, 32 * n, 48 * n):
        raise ValueError("need n x (commitment48, z32, y32, proof48)")
    ok = ctypes.c_int(0)
    _check(s._lib.rk_verify_kzg_proof_batch(s._ctx, c.ptr, z.ptr, y.ptr, p.ptr, n, ctypes.byref(ok)))
    return bool(ok.value)


def verify_blob_kzg_proof_batch(blobs, commitments: Sequence[bytes], proofs: Sequence[bytes],
                                settings: Optional[KzgSettings] = None) -> bool:
    """verify_blob_kzg_proof_batch (Deneb spec; BASELINE.json configs[4]) on the GPU."""
    s = _s(settings)
    buf, n = _blobs_buf(blobs)
    craw = b"".join(bytes(c) for c in commitments)
    praw = b"".join(bytes(p) for p in proofs)
    if len(craw) != 48 * n or len(praw) != 48 * n:
        raise ValueError("need one 48-byte commitment and proof per blob")
    ok = ctypes.c_int(0)
    _check(s._lib.rk_verify_blob_kzg_proof_batch(s._ctx, buf.ptr, _cptr(craw), _cptr(praw), n, ctypes.byref(ok)))
    return bool(ok.value)


# ---------------------------------------------------------------------------------------
# Batch API (the measured path)
# ---------------------------------------------------------------------------------------
class BatchResult:
    """Per-blob outputs as lists of bytes; status[i] != 0 marks a blob that failed to deserialize."""

    def __init__(self, n, commitments=None, versioned_hashes=None, xs=None, ys=None, proofs=None, status=None):
        self.n = n
        self.commitments, self.versioned_hashes, self.xs, self.ys, self.proofs, self.status = (
            commitments, versioned_hashes, xs, ys, proofs, status)


def _split(raw: bytes, width: int) -> List[bytes]:
    return [raw[i:i + width] for i in range(0, len(raw), width)]


def _blobs_buf(blobs) -> Tuple[_Buf, int]:
    if isinstance(blobs, (list, tuple)):
        blobs = b"".join(bytes(b) for b in blobs)
    buf = _Buf(blobs)
    if buf.nbytes % BYTES_PER_BLOB:
        raise DeserializeBlob()
    return buf, buf.nbytes // BYTES_PER_BLOB


def commit_batch(blobs, settings: Optional[KzgSettings] = None) -> BatchResult:
    s = _s(settings)
    buf, n = _blobs_buf(blobs)
    c = ctypes.create_string_buffer(48 * n)
    vh = ctypes.create_string_buffer(32 * n)
    st = ctypes.create_string_buffer(n)
    _check(s._lib.rk_commit_batch(s._ctx, buf.ptr, n, c, vh, st))
    return BatchResult(n, _split(c.raw, 48), _split(vh.raw, 32), status=list(st.raw))


def commit_prove_batch(blobs, settings: Optional[KzgSettings] = None) -> BatchResult:
    s = _s(settings)
    buf, n = _blobs_buf(blobs)
    c = ctypes.create_string_buffer(48 * n)
    vh = ctypes.create_string_buffer(32 * n)
    x = ctypes.create_string_buffer(32 * n)
    y = ctypes.create_string_buffer(32 * n)
    pr = ctypes.create_string_buffer(48 * n)
    st = ctypes.create_string_buffer(n)
    _check(s._lib.rk_commit_prove_batch(s._ctx, buf.ptr, n, c, vh, x, y, pr, st))
    return BatchResult(n, _split(c.raw, 48), _split(vh.raw, 32), _split(x.raw, 32), _split(y.raw, 32),
                       _split(pr.raw, 48), list(st.raw))


def compute_blob_kzg_proof_batch(blobs, commitments: Sequence[bytes], settings: Optional[KzgSettings] = None) -> BatchResult:
    """compute_blob_kzg_proof (Deneb spec) per blob: EIP-4844 challenge from (blob, commitment), then
    the proof there -- the proofs verify_blob_kzg_proof_batch accepts."""
    s = _s(settings)
    buf, n = _blobs_buf(blobs)
    craw = b"".join(bytes(c) for c in commitments)
    if len(craw) != 48 * n:
        raise ValueError("need one 48-byte commitment per blob")
    pr = ctypes.create_string_buffer(48 * n)
    st = ctypes.create_string_buffer(n)
    _check(s._lib.rk_compute_blob_kzg_proof_batch(s._ctx, buf.ptr, _cptr(craw), n, pr, st))
    return BatchResult(n, commitments=_split(craw, 48), proofs=_split(pr.raw, 48), status=list(st.raw))


def compute_kzg_proof_batch(blobs, zs: Sequence[bytes], settings: Optional[KzgSettings] = None) -> BatchResult:
    s = _s(settings)
    buf, n = _blobs_buf(blobs)
    zraw = b"".join(bytes(z) for z in zs)
    if len(zraw) != 32 * n:
        raise ValueError("need one 32-byte z per blob")
    y = ctypes.create_string_buffer(32 * n)
    pr = ctypes.create_string_buffer(48 * n)
    st = ctypes.create_string_buffer(n)
    _check(s._lib.rk_compute_kzg_proof_batch(s._ctx, buf.ptr, _cptr(zraw), n, pr, y, st))
    return BatchResult(n, xs=_split(zraw, 32), ys=_split(y.raw, 32), proofs=_split(pr.raw, 48), status=list(st.raw))

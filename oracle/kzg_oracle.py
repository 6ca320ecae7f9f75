"""CPU oracle for raiko's blob-KZG path (TEST INFRASTRUCTURE ONLY).

This module is the checker, never the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it.  The shipped path (``raiko_b200``) must never
import anything from ``oracle/``.

What it restates
----------------
* the raiko wrapper ``lib/src/primitives/eip4844.rs:44-99`` (challenge,
  proof_of_equivalence, calc_kzg_proof[_with_point], calc_kzg_proof_commitment,
  commitment_to_version_hash, kzg_proof_to_bytes) -- read from the tree;
* the arithmetic those functions call, which lives in the un-vendored crate
  ``rust-kzg`` (git brechtpd/rust-kzg, branch sp1-patch, rev
  cbbfafdd1fa9ae4133e04d5edf6eaabf86acac3a; ``Cargo.lock:4167-4176,7499-7513``).
  That source is NOT under /root/reference, so the published algorithm it
  implements (Deneb ``polynomial-commitments.md``) is
  restated here with Python big integers.

Parity pinning: the reference holds ONE byte-level KAT for this path (zero blob
-> versioned hash, eip4844.rs:147-160) and one pairing-level vector
(eip4844.rs:162-214).  Both are replayed by ``tests/test_oracle.py``; every
other golden vector is pinned by the pairing equation implemented below
(``verify_kzg_proof``) plus uniqueness of KZG commitments / proofs, and by
SURVEY.md Appendix C.  Byte-level results for non-trivial blobs are therefore
"parity pinned through pairing + uniqueness", not through reference bytes.

Everything is plain Python ints: slow (~1.5 s per 4096-term MSM) but obviously
correct.  ``oracle/kzg_ref.c`` is the fast C restatement used for large cases.
"""
from __future__ import annotations

import hashlib
import struct
from typing import Iterable, List, Optional, Sequence, Tuple

# --------------------------------------------------------------------------
# Constants (BLS12-381)
# --------------------------------------------------------------------------
P = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
FIELD_ELEMENTS_PER_BLOB = 4096
BYTES_PER_FIELD_ELEMENT = 32
BYTES_PER_BLOB = FIELD_ELEMENTS_PER_BLOB * BYTES_PER_FIELD_ELEMENT
VERSIONED_HASH_VERSION_KZG = 0x01  # eip4844.rs:26
G1_GEN = (
    0x17F1D3A73197D7942695638C4FA9AC0FC3688C4F9774B905A14E3A3F171BAC586C55E83FF97A1AEFFB3AF00ADB22C6BB,
    0x08B3F481E3AAA0F1A09E30ED741D8AE4FCF5E095D5D00AF600DB18CB2C04B3EDD03CC744A2888AE40CAA232946C5E7E1,
)
PRIMITIVE_ROOT = 7
OMEGA = pow(PRIMITIVE_ROOT, (R - 1) // FIELD_ELEMENTS_PER_BLOB, R)
FIAT_SHAMIR_PROTOCOL_DOMAIN = b"FSBLOBVERIFY_V1_"
RANDOM_CHALLENGE_KZG_BATCH_DOMAIN = b"RCKZGBATCH___V1_"
BLS_X_ABS = 0xD201000000010000  # |x| of the BLS12-381 parameter (x is negative)

MONT_FP = 1 << 384
MONT_FR = 1 << 256


class BlobError(ValueError):
    """Mirrors Eip4844Error::DeserializeBlob (eip4844.rs:34-35)."""


def bit_reverse(i: int, bits: int) -> int:
    return int(format(i, "0%db" % bits)[::-1], 2)


def roots_of_unity_brp(n: int = FIELD_ELEMENTS_PER_BLOB) -> List[int]:
    bits = n.bit_length() - 1
    w = pow(PRIMITIVE_ROOT, (R - 1) // n, R)
    nat = [1] * n
    for i in range(1, n):
        nat[i] = nat[i - 1] * w % R
    return [nat[bit_reverse(i, bits)] for i in range(n)]


# --------------------------------------------------------------------------
# G1 arithmetic (affine tuples or None for infinity; Jacobian inside the MSM)
# --------------------------------------------------------------------------
Affine = Optional[Tuple[int, int]]


def g1_is_on_curve(pt: Affine) -> bool:
    if pt is None:
        return True
    x, y = pt
    return (y * y - x * x * x - 4) % P == 0


def g1_neg(pt: Affine) -> Affine:
    if pt is None:
        return None
    return (pt[0], (-pt[1]) % P)


def g1_add(a: Affine, b: Affine) -> Affine:
    if a is None:
        return b
    if b is None:
        return a
    x1, y1 = a
    x2, y2 = b
    if x1 == x2:
        if (y1 + y2) % P == 0:
            return None
        lam = 3 * x1 * x1 * pow(2 * y1, -1, P) % P
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, P) % P
    x3 = (lam * lam - x1 - x2) % P
    return (x3, (lam * (x1 - x3) - y1) % P)


def _jac_double(X, Y, Z):
    if Y == 0 or Z == 0:
        return (0, 1, 0)
    A = X * X % P
    B = Y * Y % P
    C = B * B % P
    D = 2 * ((X + B) * (X + B) - A - C) % P
    E = 3 * A % P
    X3 = (E * E - 2 * D) % P
    Y3 = (E * (D - X3) - 8 * C) % P
    Z3 = 2 * Y * Z % P
    return (X3, Y3, Z3)


def _jac_add_affine(X1, Y1, Z1, x2, y2):
    """Jacobian += affine (x2, y2 never infinity here)."""
    if Z1 == 0:
        return (x2, y2, 1)
    Z1Z1 = Z1 * Z1 % P
    U2 = x2 * Z1Z1 % P
    S2 = y2 * Z1 * Z1Z1 % P
    H = (U2 - X1) % P
    rr = (S2 - Y1) % P
    if H == 0:
        if rr == 0:
            return _jac_double(X1, Y1, Z1)
        return (0, 1, 0)
    HH = H * H % P
    HHH = H * HH % P
    V = X1 * HH % P
    X3 = (rr * rr - HHH - 2 * V) % P
    Y3 = (rr * (V - X3) - Y1 * HHH) % P
    Z3 = Z1 * H % P
    return (X3, Y3, Z3)


def _jac_add(a, b):
    X1, Y1, Z1 = a
    X2, Y2, Z2 = b
    if Z1 == 0:
        return b
    if Z2 == 0:
        return a
    Z1Z1 = Z1 * Z1 % P
    Z2Z2 = Z2 * Z2 % P
    U1 = X1 * Z2Z2 % P
    U2 = X2 * Z1Z1 % P
    S1 = Y1 * Z2 * Z2Z2 % P
    S2 = Y2 * Z1 * Z1Z1 % P
    H = (U2 - U1) % P
    rr = (S2 - S1) % P
    if H == 0:
        if rr == 0:
            return _jac_double(X1, Y1, Z1)
        return (0, 1, 0)
    HH = H * H % P
    HHH = H * HH % P
    V = U1 * HH % P
    X3 = (rr * rr - HHH - 2 * V) % P
    Y3 = (rr * (V - X3) - S1 * HHH) % P
    Z3 = Z1 * Z2 * H % P
    return (X3, Y3, Z3)


def _jac_to_affine(j) -> Affine:
    X, Y, Z = j
    if Z == 0:
        return None
    zi = pow(Z, -1, P)
    zi2 = zi * zi % P
    return (X * zi2 % P, Y * zi2 * zi % P)


def g1_mul(pt: Affine, k: int) -> Affine:
    k %= R
    acc = (0, 1, 0)
    if pt is None or k == 0:
        return None
    for bit in bin(k)[2:]:
        acc = _jac_double(*acc)
        if bit == "1":
            acc = _jac_add_affine(*acc, pt[0], pt[1])
    return _jac_to_affine(acc)


def g1_lincomb(points: Sequence[Affine], scalars: Sequence[int], window: int = 8) -> Affine:
    """Σ scalars[i]·points[i]  (upstream ``g1_lincomb``).

    Plain unsigned-window bucket method.  The algorithm cannot change the bytes:
    the result is a group element with a canonical encoding."""
    assert len(points) == len(scalars)
    nwin = (255 + window - 1) // window
    mask = (1 << window) - 1
    total = (0, 1, 0)
    for w in reversed(range(nwin)):
        for _ in range(window):
            total = _jac_double(*total)
        buckets = [(0, 1, 0)] * (mask + 1)
        shift = w * window
        for pt, s in zip(points, scalars):
            d = (s >> shift) & mask
            if d and pt is not None:
                buckets[d] = _jac_add_affine(*buckets[d], pt[0], pt[1])
        run = (0, 1, 0)
        acc = (0, 1, 0)
        for d in range(mask, 0, -1):
            run = _jac_add(run, buckets[d])
            acc = _jac_add(acc, run)
        total = _jac_add(total, acc)
    return _jac_to_affine(total)


def g1_compress(pt: Affine) -> bytes:
    """ZG1::to_bytes / kzg_proof_to_bytes (eip4844.rs:97-99; App. B.5)."""
    if pt is None:
        return b"\xc0" + b"\x00" * 47
    x, y = pt
    out = bytearray(x.to_bytes(48, "big"))
    out[0] |= 0x80
    if y > (P - 1) // 2:
        out[0] |= 0x20
    return bytes(out)


def fp_sqrt(a: int) -> Optional[int]:
    s = pow(a, (P + 1) // 4, P)  # p ≡ 3 (mod 4)
    return s if s * s % P == a % P else None


def g1_decompress(b: bytes, check_subgroup: bool = False) -> Affine:
    if len(b) != 48 or not (b[0] & 0x80):
        raise ValueError("bad compressed G1 encoding")
    if b[0] & 0x40:
        if any(b[1:]) or (b[0] & 0x3F):
            raise ValueError("bad infinity encoding")
        return None
    sign = bool(b[0] & 0x20)
    x = int.from_bytes(bytes([b[0] & 0x1F]) + b[1:], "big")
    if x >= P:
        raise ValueError("x not canonical")
    y = fp_sqrt((x * x * x + 4) % P)
    if y is None:
        raise ValueError("not on curve")
    if (y > (P - 1) // 2) != sign:
        y = P - y
    pt = (x, y)
    if check_subgroup and g1_mul_raw(pt, R) is not None:
        raise ValueError("not in G1 subgroup")
    return pt


def g1_mul_raw(pt: Affine, k: int) -> Affine:
    """Scalar mul WITHOUT reducing k mod r (used for subgroup checks)."""
    acc = (0, 1, 0)
    if pt is None or k == 0:
        return None
    for bit in bin(k)[2:]:
        acc = _jac_double(*acc)
        if bit == "1":
            acc = _jac_add_affine(*acc, pt[0], pt[1])
    return _jac_to_affine(acc)


# --------------------------------------------------------------------------
# Trusted setup decoding (eip4844.rs:14-24)
# --------------------------------------------------------------------------
class Settings:
    """Decoded KZGSettings: Lagrange G1 (brp order), G2 monomial, brp roots."""

    def __init__(self, g1: List[Affine], g2: list, roots_brp: List[int]):
        self.g1 = g1
        self.g2 = g2  # list of ((x0,x1),(y0,y1)) affine Fp2 points
        self.roots_brp = roots_brp


def _fp_from_mont_le(b: bytes) -> int:
    return int.from_bytes(b, "little") * pow(MONT_FP, -1, P) % P


_FP_RINV = pow(MONT_FP, -1, P)
_FR_RINV = pow(MONT_FR, -1, R)


def _decode_g1_mont(b: bytes) -> Affine:
    x = int.from_bytes(b[0:48], "little") * _FP_RINV % P
    y = int.from_bytes(b[48:96], "little") * _FP_RINV % P
    z = int.from_bytes(b[96:144], "little") * _FP_RINV % P
    if z == 0:
        return None
    if z != 1:
        zi = pow(z, -1, P)
        # rust bls12_381 G1Projective is homogeneous projective (x/z, y/z)
        x, y = x * zi % P, y * zi % P
    return (x, y)


def _decode_g2_mont(b: bytes):
    v = [int.from_bytes(b[48 * k:48 * k + 48], "little") * _FP_RINV % P for k in range(6)]
    assert v[4] == 1 and v[5] == 0, "settings G2 points are expected affine (Z=1)"
    return ((v[0], v[1]), (v[2], v[3]))


def load_settings(data: bytes) -> Settings:
    """Accepts both on-disk images of the reference (auto-detected by length):
    ``kzg_settings_raw.bin`` (739 624 B) and the bincode image
    ``lib/kzg_settings/zkcrypto_kzg_settings.bin`` (1 001 905 B) that
    eip4844.rs:14 embeds, plus this repo's compact compressed-point format."""
    n = FIELD_ELEMENTS_PER_BLOB
    if len(data) == 739_624:
        assert int.from_bytes(data[0:8], "big") == n
        roots_off, g1_off, g2_off = 8, 131_080, 720_904
    elif len(data) == 1_001_905:
        assert int.from_bytes(data[0:8], "little") == n
        roots_off, g1_off, g2_off = 262_264 + 8, 393_344 + 8, 983_176 + 8
        assert int.from_bytes(data[393_344:393_352], "little") == n
        assert data[-1] == 0  # precomputation = None
    elif data[:8] == b"RKZGTS02":
        # compact image: compressed G1, uncompressed affine G2 (x.c0|x.c1|y.c0|y.c1, 48 B BE each)
        ng1, ng2 = struct.unpack("<II", data[8:16])
        g1 = [g1_decompress(data[16 + 48 * i:16 + 48 * i + 48]) for i in range(ng1)]
        off = 16 + 48 * ng1
        g2 = []
        for i in range(ng2):
            v = [int.from_bytes(data[off + 192 * i + 48 * k:off + 192 * i + 48 * k + 48], "big") for k in range(4)]
            g2.append(((v[0], v[1]), (v[2], v[3])))
        return Settings(g1, g2, roots_of_unity_brp(ng1))
    else:
        raise ValueError("unknown settings image (len %d)" % len(data))
    roots = [int.from_bytes(data[roots_off + 32 * i:roots_off + 32 * i + 32], "little") * _FR_RINV % R
             for i in range(n)]
    g1 = [_decode_g1_mont(data[g1_off + 144 * i:g1_off + 144 * i + 144]) for i in range(n)]
    g2 = [_decode_g2_mont(data[g2_off + 288 * i:g2_off + 288 * i + 288]) for i in range(65)]
    return Settings(g1, g2, roots)


# --------------------------------------------------------------------------
# Blob / field helpers (App. B.1)
# --------------------------------------------------------------------------
def deserialize_blob(blob: bytes) -> List[int]:
    """Blob::from_bytes + deserialize_blob_rust (call sites eip4844.rs:54-56)."""
    if len(blob) != BYTES_PER_BLOB:
        raise BlobError("blob length %d != %d" % (len(blob), BYTES_PER_BLOB))
    out = []
    for i in range(FIELD_ELEMENTS_PER_BLOB):
        v = int.from_bytes(blob[32 * i:32 * i + 32], "big")
        if v >= R:
            raise BlobError("field element %d not canonical" % i)
        out.append(v)
    return out


def hash_to_bls_field(b32: bytes) -> int:
    return int.from_bytes(b32, "big") % R


def fr_to_bytes(v: int) -> bytes:
    return (v % R).to_bytes(32, "big")


def commitment_to_version_hash(commitment: bytes) -> bytes:
    """eip4844.rs:91-95."""
    h = bytearray(hashlib.sha256(commitment).digest())
    h[0] = VERSIONED_HASH_VERSION_KZG
    return bytes(h)


def get_evaluation_point(blob: bytes, versioned_hash: bytes) -> int:
    """raiko's own Fiat-Shamir point, eip4844.rs:44-48 (NOT compute_challenge)."""
    blob_hash = hashlib.sha256(blob).digest()
    return hash_to_bls_field(hashlib.sha256(blob_hash + versioned_hash).digest())


def fr_batch_inv(vals: Sequence[int]) -> List[int]:
    """Montgomery's trick; zero inputs are a caller bug here (App. B.3)."""
    pref = []
    acc = 1
    for v in vals:
        pref.append(acc)
        acc = acc * v % R
    inv = pow(acc, -1, R)
    out = [0] * len(vals)
    for i in range(len(vals) - 1, -1, -1):
        out[i] = inv * pref[i] % R
        inv = inv * vals[i] % R
    return out


def evaluate_polynomial_in_evaluation_form(poly: Sequence[int], x: int, s: Settings) -> int:
    """App. B.3 (called at eip4844.rs:60)."""
    n = FIELD_ELEMENTS_PER_BLOB
    roots = s.roots_brp
    for i in range(n):
        if roots[i] == x:
            return poly[i]
    inv = fr_batch_inv([(x - w) % R for w in roots])
    acc = 0
    for i in range(n):
        acc = (acc + inv[i] * roots[i] % R * poly[i]) % R
    acc = acc * pow(n, -1, R) % R
    return acc * (pow(x, n, R) - 1) % R


def compute_quotient_eval_form(poly: Sequence[int], z: int, y: int, s: Settings) -> List[int]:
    """q_i = (p_i - y)/(w_i - z) with the in-domain special case (App. B.4)."""
    n = FIELD_ELEMENTS_PER_BLOB
    roots = s.roots_brp
    m = None
    denom = []
    for i in range(n):
        if roots[i] == z:
            m = i
            denom.append(1)
        else:
            denom.append((roots[i] - z) % R)
    inv = fr_batch_inv(denom)
    q = [(poly[i] - y) * inv[i] % R for i in range(n)]
    if m is not None:
        q[m] = 0
        d2 = fr_batch_inv([1 if i == m else z * (z - roots[i]) % R for i in range(n)])
        acc = 0
        for i in range(n):
            if i != m:
                acc = (acc + (poly[i] - y) * roots[i] % R * d2[i]) % R
        q[m] = acc
    return q


# --------------------------------------------------------------------------
# The reference wrapper functions (eip4844.rs:44-99), same names
# --------------------------------------------------------------------------
def blob_to_kzg_commitment(blob: bytes, s: Settings) -> bytes:
    return g1_compress(g1_lincomb(s.g1, deserialize_blob(blob)))


def calc_kzg_proof_commitment(blob: bytes, s: Settings) -> bytes:
    """eip4844.rs:80-89."""
    return blob_to_kzg_commitment(blob, s)


def proof_of_equivalence(blob: bytes, versioned_hash: bytes, s: Settings) -> Tuple[bytes, bytes]:
    """eip4844.rs:50-65 -> (x bytes, y bytes)."""
    poly = deserialize_blob(blob)
    x = get_evaluation_point(blob, versioned_hash)
    y = evaluate_polynomial_in_evaluation_form(poly, x, s)
    return fr_to_bytes(x), fr_to_bytes(y)


def compute_kzg_proof(blob: bytes, z: int, s: Settings) -> Tuple[bytes, bytes]:
    """compute_kzg_proof_rust (App. B.4) -> (proof48, y32)."""
    poly = deserialize_blob(blob)
    z %= R
    y = evaluate_polynomial_in_evaluation_form(poly, z, s)
    q = compute_quotient_eval_form(poly, z, y, s)
    return g1_compress(g1_lincomb(s.g1, q)), fr_to_bytes(y)


def calc_kzg_proof_with_point(blob: bytes, z: int, s: Settings) -> bytes:
    """eip4844.rs:71-78."""
    return compute_kzg_proof(blob, z, s)[0]


def calc_kzg_proof(blob: bytes, versioned_hash: bytes, s: Settings) -> bytes:
    """eip4844.rs:67-69."""
    return calc_kzg_proof_with_point(blob, get_evaluation_point(blob, versioned_hash), s)


# --------------------------------------------------------------------------
# Synthetic blobs
# --------------------------------------------------------------------------
def synthetic_blob(b: int, seed: int = 20241018) -> bytes:
    out = bytearray()
    pre = b"raiko-kzg-bench-v1" + struct.pack("<Q", seed) + struct.pack("<I", b)
    for i in range(FIELD_ELEMENTS_PER_BLOB):
        v = int.from_bytes(hashlib.sha256(pre + struct.pack("<I", i)).digest(), "big") % R
        out += v.to_bytes(32, "big")
    return bytes(out)


# --------------------------------------------------------------------------
# Pairing (only to VALIDATE golden vectors: replays eip4844.rs:162-214)
# Fp12 = Fp[w]/(w^12 - 2w^6 + 2); Fp2 element a+bi sits at (a-b) + b*w^6.
# --------------------------------------------------------------------------
def _f12_mul(a: List[int], b: List[int]) -> List[int]:
    c = [0] * 23
    for i, ai in enumerate(a):
        if ai:
            for j, bj in enumerate(b):
                if bj:
                    c[i + j] += ai * bj
    for k in range(22, 11, -1):
        t = c[k]
        if t:
            c[k - 6] += 2 * t
            c[k - 12] -= 2 * t
    return [v % P for v in c[:12]]


_F12_ONE = [1] + [0] * 11


def _f12_pow(a: List[int], e: int) -> List[int]:
    out = _F12_ONE
    for bit in bin(e)[2:]:
        out = _f12_mul(out, out)
        if bit == "1":
            out = _f12_mul(out, a)
    return out


def _fp2_mul(a, b):
    return ((a[0] * b[0] - a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)


def _fp2_inv(a):
    d = pow(a[0] * a[0] + a[1] * a[1], -1, P)
    return (a[0] * d % P, (-a[1]) * d % P)


def _fp2_sub(a, b):
    return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)


def _fp2_add(a, b):
    return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)


def g2_add(a, b):
    if a is None:
        return b
    if b is None:
        return a
    (x1, y1), (x2, y2) = a, b
    if x1 == x2:
        if _fp2_add(y1, y2) == (0, 0):
            return None
        lam = _fp2_mul(_fp2_mul((3, 0), _fp2_mul(x1, x1)), _fp2_inv(_fp2_add(y1, y1)))
    else:
        lam = _fp2_mul(_fp2_sub(y2, y1), _fp2_inv(_fp2_sub(x2, x1)))
    x3 = _fp2_sub(_fp2_sub(_fp2_mul(lam, lam), x1), x2)
    y3 = _fp2_sub(_fp2_mul(lam, _fp2_sub(x1, x3)), y1)
    return (x3, y3)


def g2_neg(a):
    return None if a is None else (a[0], ((-a[1][0]) % P, (-a[1][1]) % P))


def g2_mul(pt, k: int):
    acc = None
    for bit in bin(k % R)[2:] if k % R else "":
        acc = g2_add(acc, acc)
        if bit == "1":
            acc = g2_add(acc, pt)
    return acc


def g2_decompress(b: bytes):
    """zcash 96-byte compressed G2: x.c1 (with flags) || x.c0."""
    assert len(b) == 96 and b[0] & 0x80
    if b[0] & 0x40:
        return None
    sign = bool(b[0] & 0x20)
    x1 = int.from_bytes(bytes([b[0] & 0x1F]) + b[1:48], "big")
    x0 = int.from_bytes(b[48:96], "big")
    x = (x0, x1)
    rhs = _fp2_add(_fp2_mul(_fp2_mul(x, x), x), (4, 4))
    y = _fp2_sqrt(rhs)
    if y is None:
        raise ValueError("G2 point not on twist")
    # lexicographic sign: compare c1 first, then c0
    big = y[1] > (P - 1) // 2 if y[1] != 0 else y[0] > (P - 1) // 2
    if big != sign:
        y = ((-y[0]) % P, (-y[1]) % P)
    return (x, y)


def g2_compress(pt) -> bytes:
    if pt is None:
        return b"\xc0" + b"\x00" * 95
    (x0, x1), (y0, y1) = pt
    out = bytearray(x1.to_bytes(48, "big") + x0.to_bytes(48, "big"))
    out[0] |= 0x80
    big = y1 > (P - 1) // 2 if y1 != 0 else y0 > (P - 1) // 2
    if big:
        out[0] |= 0x20
    return bytes(out)


def _fp2_sqrt(a):
    """Square root in Fp2 = Fp[i]/(i^2+1) via the norm trick."""
    a0, a1 = a
    if a1 == 0:
        s = fp_sqrt(a0)
        if s is not None:
            return (s, 0)
        s = fp_sqrt((-a0) % P)
        return None if s is None else (0, s)
    n = fp_sqrt((a0 * a0 + a1 * a1) % P)
    if n is None:
        return None
    inv2 = pow(2, -1, P)
    for nn in (n, (-n) % P):
        t = (a0 + nn) * inv2 % P
        x0 = fp_sqrt(t)
        if x0 is not None and x0 != 0:
            x1 = a1 * pow(2 * x0, -1, P) % P
            if _fp2_mul((x0, x1), (x0, x1)) == (a0 % P, a1 % P):
                return (x0, x1)
    return None


def _line_to_f12(c0, c2, c3: int) -> List[int]:
    """c0 (Fp2) at w^0, c2 (Fp2) at w^2, c3 (Fp) at w^3."""
    out = [0] * 12
    out[0] = (c0[0] - c0[1]) % P
    out[6] = c0[1]
    out[2] = (c2[0] - c2[1]) % P
    out[8] = c2[1]
    out[3] = c3 % P
    return out


def miller_loop(q, p_aff: Affine) -> List[int]:
    """f_{|x|,Q}(P) with Q on the twist y^2 = x^3 + 4(1+i); lines scaled by w^3
    (a factor in Fp4, killed by the final exponentiation)."""
    if q is None or p_aff is None:
        return _F12_ONE
    xp, yp = p_aff
    f = _F12_ONE
    t = q
    for bit in bin(BLS_X_ABS)[3:]:
        (xt, yt) = t
        lam = _fp2_mul(_fp2_mul((3, 0), _fp2_mul(xt, xt)), _fp2_inv(_fp2_add(yt, yt)))
        line = _line_to_f12(_fp2_sub(_fp2_mul(lam, xt), yt), ((-lam[0] * xp) % P, (-lam[1] * xp) % P), yp)
        f = _f12_mul(_f12_mul(f, f), line)
        t = g2_add(t, t)
        if bit == "1":
            (xt, yt) = t
            lam = _fp2_mul(_fp2_sub(q[1], yt), _fp2_inv(_fp2_sub(q[0], xt)))
            line = _line_to_f12(_fp2_sub(_fp2_mul(lam, xt), yt), ((-lam[0] * xp) % P, (-lam[1] * xp) % P), yp)
            f = _f12_mul(f, line)
            t = g2_add(t, q)
    return f


_FINAL_EXP = (P ** 12 - 1) // R


def pairings_product_is_one(pairs: Iterable[Tuple[Affine, object]]) -> bool:
    f = _F12_ONE
    for g1pt, g2pt in pairs:
        f = _f12_mul(f, miller_loop(g2pt, g1pt))
    return _f12_pow(f, _FINAL_EXP) == _F12_ONE


def verify_kzg_proof(commitment: bytes, z: int, y: int, proof: bytes, s: Settings) -> bool:
    """verify_kzg_proof_rust as used by eip4844.rs:176-183 (App. B.6):
    e(C - [y]G1, G2) == e(pi, [s]G2 - [z]G2)."""
    c = g1_decompress(commitment)
    pi = g1_decompress(proof)
    lhs = g1_add(c, g1_neg(g1_mul(G1_GEN, y)))
    s_minus_z = g2_add(s.g2[1], g2_neg(g2_mul(s.g2[0], z)))
    return pairings_product_is_one([(lhs, s.g2[0]), (g1_neg(pi), s_minus_z)])


def compute_challenge(blob: bytes, commitment: bytes) -> int:
    """EIP-4844 compute_challenge (App. B.6) -- used by batch verification only."""
    data = (FIAT_SHAMIR_PROTOCOL_DOMAIN + (0).to_bytes(8, "big")
            + FIELD_ELEMENTS_PER_BLOB.to_bytes(8, "big") + blob + commitment)
    return hash_to_bls_field(hashlib.sha256(data).digest())


def compute_blob_kzg_proof(blob: bytes, commitment: bytes, s: Settings) -> bytes:
    return compute_kzg_proof(blob, compute_challenge(blob, commitment), s)[0]


def verify_kzg_proof_batch(commitments, zs, ys, proofs, s: Settings) -> bool:
    n = len(commitments)
    data = (RANDOM_CHALLENGE_KZG_BATCH_DOMAIN + FIELD_ELEMENTS_PER_BLOB.to_bytes(8, "big")
            + n.to_bytes(8, "big"))
    for c, z, y, pr in zip(commitments, zs, ys, proofs):
        data += c + fr_to_bytes(z) + fr_to_bytes(y) + pr
    rch = hash_to_bls_field(hashlib.sha256(data).digest())
    rp = [pow(rch, i, R) for i in range(n)]
    cs = [g1_decompress(c) for c in commitments]
    ps = [g1_decompress(pr) for pr in proofs]
    proof_lincomb = None
    proof_z_lincomb = None
    c_minus_y_lincomb = None
    for i in range(n):
        proof_lincomb = g1_add(proof_lincomb, g1_mul(ps[i], rp[i]))
        proof_z_lincomb = g1_add(proof_z_lincomb, g1_mul(ps[i], rp[i] * zs[i] % R))
        cmy = g1_add(cs[i], g1_neg(g1_mul(G1_GEN, ys[i])))
        c_minus_y_lincomb = g1_add(c_minus_y_lincomb, g1_mul(cmy, rp[i]))
    rhs = g1_add(c_minus_y_lincomb, proof_z_lincomb)
    return pairings_product_is_one([(g1_neg(proof_lincomb), s.g2[1]), (rhs, s.g2[0])])


def verify_blob_kzg_proof_batch(blobs, commitments, proofs, s: Settings) -> bool:
    zs, ys = [], []
    for blob, c in zip(blobs, commitments):
        z = compute_challenge(blob, c)
        zs.append(z)
        ys.append(evaluate_polynomial_in_evaluation_form(deserialize_blob(blob), z, s))
    return verify_kzg_proof_batch(commitments, zs, ys, proofs, s)


# --------------------------------------------------------------------------
# Blob -> tx-list byte codec (lib/src/utils.rs:80-179).
# The other consumer of the 128 KiB blob: OP-style packing, 4 field elements
# (4 x 31 bytes + 4 x 6 spare bits) -> 127 bytes per round.
# --------------------------------------------------------------------------
BLOB_ENCODING_VERSION = 0
MAX_BLOB_DATA_SIZE = (4 * 31 + 3) * 1024 - 4


def decode_blob_data(blob: bytes) -> bytes:
    """utils.rs:85-144, statement by statement (invalid input -> empty)."""
    if blob[1] != BLOB_ENCODING_VERSION:
        return b""
    output_len = (blob[2] << 16) | (blob[3] << 8) | blob[4]
    if output_len > MAX_BLOB_DATA_SIZE:
        return b""
    out = bytearray(MAX_BLOB_DATA_SIZE)
    out[0:27] = blob[5:32]
    opos, ipos = 28, 32
    enc = [blob[0], 0, 0, 0]

    def field_element(opos, ipos):          # utils.rs:146-161
        if blob[ipos] & 0b1100_0000:
            return None
        out[opos:opos + 31] = blob[ipos + 1:ipos + 32]
        return blob[ipos], opos + 32, ipos + 32

    def reassemble(opos):                   # utils.rs:163-179
        opos -= 1
        x = (enc[0] & 0b0011_1111) | ((enc[1] & 0b0011_0000) << 2)
        y = (enc[1] & 0b0000_1111) | ((enc[3] & 0b0000_1111) << 4)
        z = (enc[2] & 0b0011_1111) | ((enc[3] & 0b0011_0000) << 2)
        out[opos - 32] = z
        out[opos - 64] = y
        out[opos - 96] = x
        return opos

    for k in (1, 2, 3):
        r = field_element(opos, ipos)
        if r is None:
            return b""
        enc[k], opos, ipos = r
    opos = reassemble(opos)
    for _ in range(1, 1024):
        if opos < output_len:
            for k in range(4):
                r = field_element(opos, ipos)
                if r is None:
                    return b""
                enc[k], opos, ipos = r
            opos = reassemble(opos)
    if any(out[output_len:]):
        return b""
    if any(blob[ipos:BYTES_PER_BLOB]):
        return b""
    return bytes(out[:output_len])


def encode_blob_data(data: bytes) -> bytes:
    """Inverse of decode_blob_data (the reference only decodes; this builds test inputs)."""
    assert len(data) <= MAX_BLOB_DATA_SIZE
    padded = bytes(data) + bytes(MAX_BLOB_DATA_SIZE - len(data))
    blob = bytearray(BYTES_PER_BLOB)
    n = len(data)
    rounds = 1 if n <= 123 else 1 + (n - 123 + 126) // 127
    for r in range(rounds):
        base = -4 + 127 * r                 # output offset of this round's first payload byte
        fe = [bytearray(32) for _ in range(4)]
        for k in range(4):
            for b in range(1, 32):
                idx = base + 32 * k + (b - 1)
                if 0 <= idx < MAX_BLOB_DATA_SIZE and not (r == 0 and k == 0 and b <= 4):
                    fe[k][b] = padded[idx]
        x, y, z = (padded[base + 31], padded[base + 63], padded[base + 95])
        fe[0][0] = x & 0x3F
        fe[1][0] = ((x >> 2) & 0x30) | (y & 0x0F)
        fe[2][0] = z & 0x3F
        fe[3][0] = ((z >> 2) & 0x30) | ((y >> 4) & 0x0F)
        if r == 0:
            fe[0][1] = BLOB_ENCODING_VERSION
            fe[0][2], fe[0][3], fe[0][4] = (n >> 16) & 0xFF, (n >> 8) & 0xFF, n & 0xFF
        for k in range(4):
            blob[128 * r + 32 * k:128 * r + 32 * k + 32] = fe[k]
    return bytes(blob)

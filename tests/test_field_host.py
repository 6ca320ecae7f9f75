"""The product's limb arithmetic (raiko_b200/csrc/field30.cuh, g1.cuh), compiled for the host
by tests/native/fieldcheck.cpp, against Python big integers.  Same headers the kernels use."""
import ctypes
import os
import random
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "native", "_build", "libfieldcheck.so")
P = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001


def build(form=None):
    """The headers as the library compiles them (form None), or with one of field30.cuh's other
    Montgomery-product formulations (-DRK_MUL_FORM=n)."""
    so = SO if form is None else SO.replace(".so", "_form%d.so" % form)
    os.makedirs(os.path.dirname(so), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so]
                          + ([] if form is None else ["-DRK_MUL_FORM=%d" % form])
                          + [os.path.join(HERE, "native", "fieldcheck.cpp")])
    return ctypes.CDLL(so)


@pytest.fixture(scope="module")
def L():
    return build()


def limbs(v, n):
    return (ctypes.c_uint32 * n)(*[(v >> (30 * i)) & ((1 << 30) - 1) for i in range(n)])


def val(a, bits=30):
    return sum(int(x) << (bits * i) for i, x in enumerate(a))


@pytest.mark.parametrize("form", [None, 0, 1, 2, 3, 4, 5])
@pytest.mark.parametrize("name,mod,n", [("fp", P, 13), ("fr", R, 9)])
def test_montgomery_mul_sqr(form, name, mod, n):
    """Every formulation of the Montgomery product (field30.cuh RK_MUL_FORM), worst-case column
    sums first: a wrapped 64-bit column shows up as a wrong product."""
    L = build(form)
    rnd = random.Random(n)
    MR = 1 << (30 * n)
    mul, sqr = getattr(L, "fc_%s_mul" % name), getattr(L, "fc_%s_sqr" % name)
    for t in range(6000):
        if t < 4:
            a = b = MR - 1                       # every limb 2^30-1: worst-case column sums
        elif t < 2000:
            a, b = rnd.randrange(8 * mod), rnd.randrange(8 * mod)
        else:
            a, b = rnd.randrange(min(MR, 25 * mod)), rnd.randrange(min(MR, 25 * mod))
        out = (ctypes.c_uint32 * n)()
        mul(limbs(a, n), limbs(b, n), out)
        assert all(x < (1 << 30) for x in out[:n - 1])
        r = val(out)
        assert r % mod == a * b * pow(MR, -1, mod) % mod and r <= a * b // MR + mod
        sqr(limbs(a, n), out)
        r = val(out)
        assert r % mod == a * a * pow(MR, -1, mod) % mod and r <= a * a // MR + mod


def test_add_sub_zero_pack_inv(L):
    rnd = random.Random(5)
    for _ in range(3000):
        a, b = rnd.randrange(8 * P), rnd.randrange(6 * P)
        out = (ctypes.c_uint32 * 13)()
        L.fc_fp_add(limbs(a, 13), limbs(b, 13), out)
        assert val(out) == a + b
        L.fc_fp_sub6(limbs(a, 13), limbs(b, 13), out)
        assert val(out) == a - b + 6 * P
    for k in range(9):
        assert L.fc_fp_is_zero_mod(limbs(k * P, 13)) == 1
        assert L.fc_fp_is_zero_mod(limbs(k * P + 1, 13)) == 0
        assert L.fc_fp_is_zero_mod(limbs(k * P + (1 << 30), 13)) == 0
    for _ in range(1000):
        a = rnd.randrange(1 << 384)
        w = (ctypes.c_uint32 * 12)()
        L.fc_fp_pack(limbs(a, 13), w)
        assert val(w, 32) == a
        out = (ctypes.c_uint32 * 13)()
        L.fc_fp_unpack(w, out)
        assert val(out) == a
        a = rnd.randrange(1 << 256)
        w = (ctypes.c_uint32 * 8)()
        L.fc_fr_pack(limbs(a, 9), w)
        assert val(w, 32) == a
        out = (ctypes.c_uint32 * 9)()
        L.fc_fr_unpack(w, out)
        assert val(out) == a
    for _ in range(8):
        a = rnd.randrange(1, P)
        w = (ctypes.c_uint32 * 12)(*[(a >> (32 * i)) & 0xFFFFFFFF for i in range(12)])
        out = (ctypes.c_uint32 * 12)()
        L.fc_fp_inv(w, out)
        assert val(out, 32) == pow(a, -1, P)
        a = rnd.randrange(1, R)
        w = (ctypes.c_uint32 * 8)(*[(a >> (32 * i)) & 0xFFFFFFFF for i in range(8)])
        out = (ctypes.c_uint32 * 8)()
        L.fc_fr_inv(w, out)
        assert val(out, 32) == pow(a, -1, R)


def test_g1_formulas_and_exceptional_cases(L, pyoracle):
    o, s = pyoracle
    rnd = random.Random(9)
    INF = o.g1_compress(None)
    buf = lambda: ctypes.create_string_buffer(48)  # noqa: E731
    pts = [o.g1_mul(o.G1_GEN, rnd.randrange(1, R)) for _ in range(6)]
    for pt in pts + [None]:
        out = buf()
        assert L.fc_g1_roundtrip(o.g1_compress(pt), out) == 0 and out.raw == o.g1_compress(pt)
    for i, a in enumerate(pts):
        b = pts[(i + 1) % len(pts)]
        ca, cb = o.g1_compress(a), o.g1_compress(b)
        for fn in (L.fc_g1_add, L.fc_g1_madd):
            out = buf(); fn(ca, cb, out); assert out.raw == o.g1_compress(o.g1_add(a, b))          # generic
            out = buf(); fn(ca, ca, out); assert out.raw == o.g1_compress(o.g1_add(a, a))          # P + P
            out = buf(); fn(ca, o.g1_compress(o.g1_neg(a)), out); assert out.raw == INF            # P + (-P)
            out = buf(); fn(INF, ca, out); assert out.raw == ca                                    # O + P
        out = buf(); L.fc_g1_add(ca, INF, out); assert out.raw == ca
        out = buf(); L.fc_g1_dbl(ca, out); assert out.raw == o.g1_compress(o.g1_add(a, a))
        k = rnd.randrange(1 << 64)
        out = buf(); L.fc_g1_mul_u64(ca, ctypes.c_uint64(k), out); assert out.raw == o.g1_compress(o.g1_mul(a, k))
    out = buf(); L.fc_g1_dbl(INF, out); assert out.raw == INF
    # 4096 decompressions + a 4095-addition chain: sum of the Lagrange setup = generator
    allpts = b"".join(o.g1_compress(p) for p in s.g1)
    out = buf()
    assert L.fc_g1_sum(allpts, 4096, out) == 0 and out.raw == o.g1_compress(o.G1_GEN)


@pytest.mark.parametrize("name,mod,n", [("fp", P, 13), ("fr", R, 9)])
def test_safegcd_inversion(L, name, mod, n):
    """fe_inv_safegcd (Bernstein-Yang division steps on signed 30-bit limbs): Montgomery in,
    Montgomery out, fully reduced; inputs anywhere below 8 mod."""
    rnd = random.Random(31 + n)
    MR = 1 << (30 * n)
    fn = getattr(L, "fc_%s_inv_safegcd" % name)
    cases = [1, 2, mod - 1, mod + 1, 2 * mod - 1, 8 * mod - 1, MR % mod, (1 << 200) + 1] + [rnd.randrange(1, 8 * mod) for _ in range(4000)]
    for x in cases:
        if x % mod == 0:
            continue
        out = (ctypes.c_uint32 * n)()
        fn(limbs(x, n), out)
        r = val(out)
        assert r < mod and all(v < (1 << 30) for v in out)
        assert r == MR * MR * pow(x, -1, mod) % mod                       # x = aR  ->  a^-1 R
    for k in range(8):                                                     # multiples of the modulus -> 0, like a^(mod-2)
        out = (ctypes.c_uint32 * n)()
        fn(limbs(k * mod, n), out)
        assert val(out) == 0


def test_reduce_loose(L):
    rnd = random.Random(77)
    top = P >> 360
    for t in range(5000):
        a = (1 << 390) - 1 if t == 0 else rnd.randrange(8 * P) if t % 2 else rnd.randrange(1 << 390)
        out = (ctypes.c_uint32 * 13)()
        L.fc_fp_reduce_loose(limbs(a, 13), out)
        r = val(out)
        q = (a >> 360) // (top + 1)
        assert r == a - q * P and 0 <= r < P + ((q + 2) << 360) and all(v < (1 << 30) for v in out[:12])
        if a < 8 * P:
            assert r < P * 1.0001


def test_scalar_multiplication_ladders(L, pyoracle):
    """g1.cuh's two variable-base scalar multiplications (bitwise double-and-add, and the 4-bit fixed
    window the verification kernel uses) against the big-integer oracle, edge scalars included."""
    o, _ = pyoracle
    rnd = random.Random(13)
    pts = [o.g1_mul(o.G1_GEN, rnd.randrange(1, R)) for _ in range(3)]
    ks = [0, 1, 2, 15, 16, 17, R - 1, R, R + 1, (1 << 256) - 1, 0xF0F0 << 240, 1 << 255] + [rnd.randrange(1 << 256) for _ in range(6)]
    for pt in pts:
        for k in ks:
            words = (ctypes.c_uint32 * 8)(*[(k >> (32 * i)) & 0xFFFFFFFF for i in range(8)])
            want = o.g1_compress(o.g1_mul(pt, k % R) if k % R else None)
            for windowed in (0, 1):
                out = ctypes.create_string_buffer(48)
                assert L.fc_g1_scalar_mul(o.g1_compress(pt), words, windowed, out) == 0
                assert out.raw == want, (hex(k), windowed)


def test_fused_limb_passes(L):
    """field30.cuh's fused passes (k_msm_affine with RK_AFF_FUSE): conditional negation folded into a
    subtraction, and subtraction + loose reduction in one pass with a raw (unnormalised) subtrahend."""
    rnd = random.Random(78)
    edge = [0, 1, P - 1, P, P + 1, 2 * P - 1, 2 * P, (1 << 360) - 1, 1 << 360]
    for t in range(6000):
        a = edge[t % len(edge)] if t < 90 else rnd.randrange(2 * P)         # +-a - b + 4p needs a + b <= 4p
        b = edge[(t // len(edge)) % len(edge)] if t < 90 else rnd.randrange(2 * P)
        for neg in (0, 1):
            out = (ctypes.c_uint32 * 13)()
            L.fc_fp_sub_cneg4(limbs(a, 13), limbs(b, 13), neg, out)
            assert val(out) == (-a if neg else a) - b + 4 * P and all(v < (1 << 30) for v in out)
        # x3-shaped: a - (b1 + b2) + 3p, b1 + b2 <= 3p, the sum NOT normalised
        a = edge[t % len(edge)] if t < 90 else rnd.randrange(int(1.2 * P))
        b1, b2 = (edge[(t // 9) % 9] % (P + 2), edge[(t // 3) % 9] % (P + 2)) if t < 90 else (rnd.randrange(int(1.5 * P)), rnd.randrange(int(1.5 * P)))
        out = (ctypes.c_uint32 * 13)()
        L.fc_fp_sub_reduce3_rawsum(limbs(a, 13), limbs(b1, 13), limbs(b2, 13), out)
        r = val(out)
        assert (r - (a - b1 - b2)) % P == 0 and 0 <= r < max(P + (12 << 360), a + 1) and all(v < (1 << 30) for v in out), (a, b1, b2, r)
        # y3-shaped: a - b + 2p, b <= 2p
        b = edge[(t // 9) % 9] if t < 90 else rnd.randrange(2 * P)
        L.fc_fp_sub_reduce2(limbs(a, 13), limbs(b, 13), out)
        r = val(out)
        assert (r - (a - b)) % P == 0 and 0 <= r < max(P + (12 << 360), a + 1) and all(v < (1 << 30) for v in out), (a, b, r)


def test_affine_batch_addition_pieces(L, pyoracle):
    """The two halves of k_msm_affine's arithmetic on the host: one batch of affine additions
    under a shared safegcd inversion (forward prefix products, backward peel), and the slow path
    for equal x (doubling, cancellation)."""
    o, _ = pyoracle
    rnd = random.Random(11)
    n = 24
    accs = [o.g1_mul(o.G1_GEN, rnd.randrange(1, R)) for _ in range(n)]
    pts = [o.g1_mul(o.G1_GEN, rnd.randrange(1, R)) for _ in range(n)]
    out = ctypes.create_string_buffer(48 * n)
    rc = L.fc_g1_affine_batch_add(b"".join(map(o.g1_compress, accs)), b"".join(map(o.g1_compress, pts)), n, out)
    assert rc == 0
    for i in range(n):
        assert out.raw[48 * i:48 * i + 48] == o.g1_compress(o.g1_add(accs[i], pts[i])), i
    # the fused backward loop (RK_AFF_FUSE), with signs: acc[i] +- pts[i]
    negs = bytes(rnd.randrange(2) for _ in range(n))
    rc = L.fc_g1_affine_batch_add_fused(b"".join(map(o.g1_compress, accs)), b"".join(map(o.g1_compress, pts)), negs, n, out)
    assert rc == 0
    for i in range(n):
        assert out.raw[48 * i:48 * i + 48] == o.g1_compress(o.g1_add(accs[i], o.g1_neg(pts[i]) if negs[i] else pts[i])), i
    # equal x is refused by the batch and done by the slow path
    assert L.fc_g1_affine_batch_add(o.g1_compress(accs[0]), o.g1_compress(accs[0]), 1, out) == -2
    assert L.fc_g1_affine_batch_add(o.g1_compress(accs[0]), o.g1_compress(o.g1_neg(accs[0])), 1, out) == -2
    one = ctypes.create_string_buffer(48)
    for a, b in ((accs[0], accs[0]), (accs[1], o.g1_neg(accs[1])), (accs[2], pts[2])):
        assert L.fc_g1_affine_add_slow(o.g1_compress(a), o.g1_compress(b), one) == 0
        assert one.raw == o.g1_compress(o.g1_add(a, b))


def test_lane_parallel_fp12_equals_scalar(L):
    """pairing.cuh's lane-parallel Fp12 arithmetic (what k_pairing_check_lanes runs on 12 lanes of a
    warp) against the scalar implementation, operation by operation: product, square, sparse line
    product, both Frobenius maps, conjugation, the norm-based inversion (vs 12 x 12 elimination),
    exponentiation by |x|; 0 has no inverse.  With 1, 5 and 12 sub-lanes per lane (the CTA-wide kernel
    k_pairing_check_cta runs 12: one Fp product per thread); the sub-lanes of a lane must agree."""
    for sub, iters in ((1, 40), (5, 10), (12, 10)):
        assert L.fc_lane_fp12_ops(iters, sub) == 0, sub


def test_lane_parallel_pairing_check(L, pyoracle, golden):
    """The KZG check e(C - yG + z proof, G2) e(-proof, [s]G2) == 1 through the lane-parallel path
    with precomputed G2 lines: accepts the reference's own vector (eip4844.rs:162-184: k % 64 blob,
    z = hash_to_bls_field([5;32])), rejects y + 1 and a swapped G2 argument, and agrees with the
    scalar path on every case."""
    o, s = pyoracle
    case = [c for c in golden["cases"] if c["name"] == "C3_mod64"][0]
    pr = [p for p in case["proofs"] if p["label"] == "fs5"][0]
    C, PI = o.g1_decompress(bytes.fromhex(case["commitment"])), o.g1_decompress(bytes.fromhex(pr["proof"]))
    z, y = int(pr["z"], 16), int(pr["y"], 16)
    G = o.g1_decompress(bytes.fromhex("97f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb"))

    def be192(q):
        (x0, x1), (y0, y1) = q
        return b"".join(v.to_bytes(48, "big") for v in (x0, x1, y0, y1))
    for dy, g2a, g2b, want in ((0, s.g2[0], s.g2[1], 1), (1, s.g2[0], s.g2[1], 0), (0, s.g2[1], s.g2[0], 0)):
        p1 = o.g1_add(o.g1_add(C, o.g1_neg(o.g1_mul(G, (y + dy) % R))), o.g1_mul(PI, z))
        args = (o.g1_compress(p1), be192(g2a), o.g1_compress(o.g1_neg(PI)), be192(g2b))
        assert L.fc_pairing_check2(*args) == want
        assert L.fc_pairing_check2_lanes(*args, 1) == want
        assert L.fc_pairing_check2_lanes(*args, 12) == want
    # an infinity argument drops its pair: e(inf, Q) e(P, Q') == 1 only if P is infinity too
    inf = b"\xc0" + bytes(47)
    for sub in (1, 12):
        assert L.fc_pairing_check2_lanes(inf, be192(s.g2[0]), inf, be192(s.g2[1]), sub) == 1
        assert L.fc_pairing_check2_lanes(inf, be192(s.g2[0]), o.g1_compress(G), be192(s.g2[1]), sub) == 0


def test_g1_subgroup_check_by_endomorphism(L, pyoracle):
    """g1_in_subgroup ([u^2]P == (BETA x, -y), Scott ePrint 2021/1130) against [r]P == infinity:
    points of G1, random curve points, and points whose order is (or contains) each small prime
    dividing the cofactor 3 * 11^2 * 10177^2 * 859267^2 * 52437899^2 -- the cases a sloppy
    membership test lets through."""
    o, _ = pyoracle
    rnd = random.Random(7)
    u = 0xD201000000010000
    cof = (u + 1) ** 2 // 3
    assert R == u ** 4 - u ** 2 + 1
    G = o.g1_decompress(bytes.fromhex("97f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb"))

    def curve_point():
        while True:
            x = rnd.randrange(P)
            y = o.fp_sqrt((x * x * x + 4) % P)
            if y is not None and (y * y - x * x * x - 4) % P == 0:
                return (x, y)

    def xy(pt):
        return pt[0].to_bytes(48, "big") + pt[1].to_bytes(48, "big")
    assert o.g1_mul_raw(curve_point(), cof * R) is None               # the curve order
    cases = [(o.g1_mul(G, k), 1) for k in (1, 2, rnd.randrange(R), R - 1)] + [(curve_point(), 0) for _ in range(10)]
    for l, e in ((3, 1), (11, 2), (10177, 2), (859267, 2), (52437899, 2)):
        while True:
            q = o.g1_mul_raw(curve_point(), cof * R // l ** e)        # the l-primary part has exponent l
            if q is not None:
                break
        assert o.g1_mul_raw(q, l) is None
        cases += [(q, 0), (o.g1_add(q, o.g1_mul(G, rnd.randrange(1, R))), 0)]
    for pt, want in cases:
        assert L.fc_g1_in_subgroup_slow(xy(pt)) == want
        assert L.fc_g1_in_subgroup_fast(xy(pt)) == want
    assert L.fc_g1_in_subgroup_fast(xy((1, 1))) == -1                 # not on the curve

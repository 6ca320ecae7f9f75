#!/usr/bin/env python3
"""Re-encode the reference's trusted setup into this repo's compact image.

The reference embeds the Ethereum KZG ceremony output as Montgomery-form limbs
(``/root/reference/kzg_settings_raw.bin``; layout in SURVEY.md Appendix A).
This script (run once, in the build container; plain integer arithmetic, no other module of
this repo) decodes it and writes the same public points in zcash-compressed form:

    "RKZGTS02" | u32 n_g1 | u32 n_g2 | n_g1 x 48 B compressed G1 (Lagrange, bit-reversed)
               | n_g2 x 192 B affine G2 (monomial; x.c0 | x.c1 | y.c0 | y.c1, big-endian)

so that the GPU box (which has no /root/reference) can build its tables.  It is
data, not source: the library re-derives every Montgomery limb on the device
and ``tests/test_settings.py`` checks that re-serialising this image
reproduces the reference files byte for byte (sha256 in SURVEY.md App. A).
"""
import hashlib
import os
import struct
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
P = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
RINV = pow(1 << 384, -1, P)          # the reference stores Fp as Montgomery limbs, R = 2^384


def fp(b: bytes) -> int:
    return int.from_bytes(b, "little") * RINV % P


src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/kzg_settings_raw.bin"
dst = os.path.join(HERE, "..", "raiko_b200", "data", "trusted_setup_4096.bin")
raw = open(src, "rb").read()
assert len(raw) == 739624 and raw[:8] == (4096).to_bytes(8, "big")
G1_OFF, G2_OFF, N1, N2 = 8 + 4096 * 32, 8 + 4096 * 32 + 4096 * 144, 4096, 65
out = bytearray(b"RKZGTS02" + struct.pack("<II", N1, N2))
for i in range(N1):
    e = raw[G1_OFF + 144 * i:G1_OFF + 144 * (i + 1)]
    x, y, z = fp(e[0:48]), fp(e[48:96]), fp(e[96:144])
    assert z == 1 and (y * y - x * x * x - 4) % P == 0          # affine, on the curve
    c = bytearray(x.to_bytes(48, "big"))
    c[0] |= 0x80 | (0x20 if y > (P - 1) // 2 else 0)             # zcash compressed form
    out += c
for i in range(N2):
    e = raw[G2_OFF + 288 * i:G2_OFF + 288 * (i + 1)]
    v = [fp(e[48 * k:48 * k + 48]) for k in range(6)]           # X.c0 X.c1 Y.c0 Y.c1 Z.c0 Z.c1
    assert v[4] == 1 and v[5] == 0
    for k in range(4):
        out += v[k].to_bytes(48, "big")
if os.path.exists(dst):
    assert open(dst, "rb").read() == bytes(out), "differs from the committed image"
open(dst, "wb").write(out)
print(dst, len(out), hashlib.sha256(out).hexdigest())

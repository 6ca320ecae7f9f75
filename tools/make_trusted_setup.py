#!/usr/bin/env python3
"""Re-encode the reference's trusted setup into this repo's compact image.

The reference embeds the Ethereum KZG ceremony output as Montgomery-form limbs
(``/root/reference/kzg_settings_raw.bin``; layout in SURVEY.md Appendix A).
This script (run once, in the build container) decodes it with the oracle and
writes the same public points in zcash-compressed form:

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
sys.path.insert(0, os.path.join(HERE, "..", "oracle"))
import kzg_oracle as o  # noqa: E402

src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/kzg_settings_raw.bin"
dst = os.path.join(HERE, "..", "raiko_b200", "data", "trusted_setup_4096.bin")
s = o.load_settings(open(src, "rb").read())
out = bytearray(b"RKZGTS02" + struct.pack("<II", len(s.g1), len(s.g2)))
for pt in s.g1:
    out += o.g1_compress(pt)
for (x0, x1), (y0, y1) in s.g2:
    for v in (x0, x1, y0, y1):
        out += v.to_bytes(48, "big")
open(dst, "wb").write(out)
rt = o.load_settings(bytes(out))
assert rt.g1 == s.g1 and rt.g2 == s.g2 and rt.roots_brp == s.roots_brp
print(dst, len(out), hashlib.sha256(out).hexdigest())

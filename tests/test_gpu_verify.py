"""verify_kzg_proof / verify_blob_kzg_proof_batch on the GPU (SURVEY.md §8(f) rank 1,
BASELINE.json configs[4]) against the oracle's pairing-level answers.  The output is a boolean:
the test is accept on honest input, reject after flipping a proof / commitment / blob byte."""
import pytest

from kzg_testlib import blob_from_recipe, synthetic_blob

pytestmark = pytest.mark.gpu
R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001


def test_verify_kzg_proof_reference_vector(gpu_settings, golden):
    """eip4844.rs:162-214: the k % 64 blob at z = hash_to_bls_field([5;32]) verifies; a proof
    for another z and y + 1 do not."""
    import raiko_b200 as rk
    case = [c for c in golden["cases"] if c["name"] == "C3_mod64"][0]
    c = bytes.fromhex(case["commitment"])
    fs5 = [p for p in case["proofs"] if p["label"] == "fs5"][0]
    fs6 = [p for p in case["proofs"] if p["label"] == "fs6"][0]
    z, y, pr = (bytes.fromhex(fs5[k]) for k in ("z", "y", "proof"))
    assert rk.verify_kzg_proof(c, z, y, pr, gpu_settings)
    y1 = ((int.from_bytes(y, "big") + 1) % R).to_bytes(32, "big")
    assert not rk.verify_kzg_proof(c, z, y1, pr, gpu_settings)
    assert not rk.verify_kzg_proof(c, bytes.fromhex(fs6["z"]), y, bytes.fromhex(fs6["proof"]), gpu_settings)
    assert rk.verify_kzg_proof(c, bytes.fromhex(fs6["z"]), bytes.fromhex(fs6["y"]), bytes.fromhex(fs6["proof"]), gpu_settings)


def test_verify_every_golden_tuple(gpu_settings, golden):
    import raiko_b200 as rk
    for case in golden["cases"]:
        for p in case["proofs"]:
            assert rk.verify_kzg_proof(bytes.fromhex(case["commitment"]), bytes.fromhex(p["z"]), bytes.fromhex(p["y"]),
                                       bytes.fromhex(p["proof"]), gpu_settings), (case["name"], p["label"])


def test_pairing_check_kernels_agree_and_are_deterministic(golden):
    """The three pairing-check kernels (single thread, one warp of 12 lanes, CTA of 12 x 12 sub-lanes:
    RAIKO_KZG_PAIRING_LANES = 0 / 1 / 2) give the same answers on every golden tuple and on tampered
    ones, and the CTA-wide one -- whose Fp12 values travel through shared memory between barriers --
    gives the same answer 25 times in a row (a missing barrier would show up as a flaky verdict)."""
    import os
    import raiko_b200 as rk
    tuples = []
    for case in golden["cases"]:
        for p in case["proofs"]:
            c, z, y, pr = (bytes.fromhex(v) for v in (case["commitment"], p["z"], p["y"], p["proof"]))
            tuples.append((c, z, y, pr, True))
            y1 = ((int.from_bytes(y, "big") + 1) % R).to_bytes(32, "big")
            tuples.append((c, z, y1, pr, False))
    saved = os.environ.get("RAIKO_KZG_PAIRING_LANES")
    try:
        for lanes in ("0", "1", "2"):
            os.environ["RAIKO_KZG_PAIRING_LANES"] = lanes
            s = rk.KzgSettings(window_bits=8)
            sel = tuples if lanes != "0" else tuples[:6]          # the single-thread check takes ~0.1 s per tuple
            for c, z, y, pr, want in sel:
                assert rk.verify_kzg_proof(c, z, y, pr, s) == want, (lanes, want)
            if lanes == "2":
                for c, z, y, pr, want in tuples[:2]:
                    assert [rk.verify_kzg_proof(c, z, y, pr, s) for _ in range(25)] == [want] * 25
            s.close()
    finally:
        if saved is None:
            os.environ.pop("RAIKO_KZG_PAIRING_LANES", None)
        else:
            os.environ["RAIKO_KZG_PAIRING_LANES"] = saved


def test_verify_rejects_bad_encodings(gpu_settings, golden):
    import raiko_b200 as rk
    case = golden["cases"][4]
    p = case["proofs"][0]
    c, z, y, pr = bytes.fromhex(case["commitment"]), bytes.fromhex(p["z"]), bytes.fromhex(p["y"]), bytes.fromhex(p["proof"])
    with pytest.raises(ValueError):                     # x not on the curve / flags missing
        rk.verify_kzg_proof(bytes([c[0] & 0x7F]) + c[1:], z, y, pr, gpu_settings)
    with pytest.raises(ValueError):
        rk.verify_kzg_proof(c, z, y, pr[:47] + bytes([pr[47] ^ 1]), gpu_settings)   # almost surely not a curve point / not in G1
    with pytest.raises(rk.DeserializeBlob):
        rk.verify_kzg_proof(c, b"\xff" * 32, y, pr, gpu_settings)                    # z >= r


def _eip4844_challenges(pyoracle, blobs, commitments):
    o, _ = pyoracle
    return [o.fr_to_bytes(o.compute_challenge(b, c)) for b, c in zip(blobs, commitments)]


def test_compute_blob_kzg_proof_batch_equals_the_oracle(gpu_settings, pyoracle):
    """rk_compute_blob_kzg_proof_batch (Deneb compute_blob_kzg_proof: EIP-4844 challenge, then the
    proof) against the Python oracle, byte for byte, and through verify_blob_kzg_proof_batch."""
    import raiko_b200 as rk
    o, s = pyoracle
    blobs = [synthetic_blob(b, seed=31337) for b in range(3)] + [blob_from_recipe({"kind": "mod64"}), bytes(131072)]
    commitments = rk.commit_batch(blobs, gpu_settings).commitments
    res = rk.compute_blob_kzg_proof_batch(blobs, commitments, gpu_settings)
    assert res.status == [0] * len(blobs)
    for i in (0, 3, 4):
        assert res.proofs[i] == o.compute_blob_kzg_proof(blobs[i], commitments[i], s), i
    assert rk.verify_blob_kzg_proof_batch(blobs, commitments, res.proofs, gpu_settings)
    bad = bytearray(blobs[1]); bad[0:32] = b"\xff" * 32
    res2 = rk.compute_blob_kzg_proof_batch([bytes(bad)], commitments[1:2], gpu_settings)
    assert res2.status == [2] and res2.proofs[0] == bytes(48)


def test_verify_blob_kzg_proof_batch(gpu_settings, pyoracle):
    import raiko_b200 as rk
    o, s = pyoracle
    blobs = [synthetic_blob(b, seed=2718) for b in range(5)] + [bytes(131072), blob_from_recipe({"kind": "mod64"})]
    commitments = rk.commit_batch(blobs, gpu_settings).commitments
    zs = _eip4844_challenges(pyoracle, blobs, commitments)
    proofs = rk.compute_kzg_proof_batch(blobs, zs, gpu_settings).proofs
    assert proofs[0] == o.compute_blob_kzg_proof(blobs[0], commitments[0], s)          # Fiat-Shamir bytes match the oracle
    assert rk.verify_blob_kzg_proof_batch(blobs, commitments, proofs, gpu_settings)
    assert rk.verify_blob_kzg_proof_batch(blobs[:1], commitments[:1], proofs[:1], gpu_settings)
    # oracle agrees on a 3-blob sub-batch (python pairing: a few seconds)
    assert o.verify_blob_kzg_proof_batch(blobs[:3], commitments[:3], proofs[:3], s)
    # swapped proofs / flipped blob byte / wrong commitment
    swapped = [proofs[1], proofs[0]] + proofs[2:]
    assert not rk.verify_blob_kzg_proof_batch(blobs, commitments, swapped, gpu_settings)
    assert not o.verify_blob_kzg_proof_batch(blobs[:3], commitments[:3], swapped[:3], s)
    tampered = bytearray(blobs[3]); tampered[131071] ^= 1
    assert not rk.verify_blob_kzg_proof_batch(blobs[:3] + [bytes(tampered)] + blobs[4:], commitments, proofs, gpu_settings)
    wrong_c = [commitments[2]] + commitments[1:]
    assert not rk.verify_blob_kzg_proof_batch(blobs, wrong_c, proofs, gpu_settings)
    # empty batch is vacuously true; non-canonical blob is an error
    assert rk.verify_blob_kzg_proof_batch([], [], [], gpu_settings)
    bad = b"\xff" * 32 + blobs[0][32:]
    with pytest.raises(rk.DeserializeBlob):
        rk.verify_blob_kzg_proof_batch([bad], commitments[:1], proofs[:1], gpu_settings)


def _golden_tuples(golden):
    out = []
    for case in golden["cases"]:
        for p in case["proofs"]:
            out.append((bytes.fromhex(case["commitment"]), bytes.fromhex(p["z"]), bytes.fromhex(p["y"]), bytes.fromhex(p["proof"])))
    return out


def test_verify_kzg_proof_batch_goldens(gpu_settings, golden):
    """verify_kzg_proof_batch: every pairing-validated golden tuple under ONE random linear
    combination; any single altered field of any tuple must flip the verdict."""
    import raiko_b200 as rk
    t = _golden_tuples(golden)
    cs, zs, ys, ps = (list(col) for col in zip(*t))
    assert len(t) >= 8
    assert rk.verify_kzg_proof_batch(cs, zs, ys, ps, gpu_settings)
    assert rk.verify_kzg_proof_batch([], [], [], [], gpu_settings)          # vacuous
    assert rk.verify_kzg_proof_batch(cs[:1], zs[:1], ys[:1], ps[:1], gpu_settings)
    k = len(t) - 2
    y1 = ((int.from_bytes(ys[k], "big") + 1) % R).to_bytes(32, "big")
    assert not rk.verify_kzg_proof_batch(cs, zs, ys[:k] + [y1] + ys[k + 1:], ps, gpu_settings)
    # z only matters where the proof is not the point at infinity (constant polynomials: q = 0)
    inf = b"\xc0" + bytes(47)
    m = next(i for i in range(len(t)) if ps[i] != inf)
    z1 = ((int.from_bytes(zs[m], "big") + 1) % R).to_bytes(32, "big")
    assert not rk.verify_kzg_proof_batch(cs, zs[:m] + [z1] + zs[m + 1:], ys, ps, gpu_settings)
    # a proof that is a valid G1 point but belongs to another tuple
    j = next(i for i in range(1, len(t)) if ps[i] != ps[0])
    assert not rk.verify_kzg_proof_batch(cs, zs, ys, [ps[j]] + ps[1:], gpu_settings)
    with pytest.raises(rk.DeserializeBlob):
        rk.verify_kzg_proof_batch(cs, [b"\xff" * 32] + zs[1:], ys, ps, gpu_settings)
    with pytest.raises(ValueError):
        rk.verify_kzg_proof_batch([bytes([cs[0][0] & 0x7F]) + cs[0][1:]] + cs[1:], zs, ys, ps, gpu_settings)


def test_full_size_65536_blob_batch_self_check(ref):
    """BASELINE.json configs[3] at full size.  The oracle cannot redo 131 072 MSMs in a test, so the
    65,536-blob commit+prove batch is pinned through size-independent properties:
      * status 0 everywhere, versioned hash = 0x01 | sha256(C)[1:] for every blob (hashlib);
      * (x, y) of EVERY blob equal the C oracle's proof_of_equivalence (challenge + barycentric
        evaluation: ~1 ms per blob on the host, all cores);
      * every (C, x, y, proof) passes the pairing check under a random linear combination
        (rk_verify_kzg_proof_batch).  With y = p(x) established independently at the
        Fiat-Shamir point x = H(blob, C), an accepting proof forces C = commit(p)
        (Schwartz-Zippel), and proofs are unique given (C, x, y): commitments and proofs are
        bit-exact without recomputing a single MSM on the CPU;
      * C, proof of a spread sample equal the C oracle byte for byte anyway;
      * altering one byte of one y makes the batch check fail."""
    import hashlib
    import os
    from concurrent.futures import ThreadPoolExecutor
    import torch
    import raiko_b200 as rk
    from raiko_b200 import _native
    n = int(os.environ.get("RAIKO_KZG_FULL_BATCH", "65536"))
    dev = torch.device("cuda", 0)
    gen = torch.Generator(device=dev)
    gen.manual_seed(65536)
    blobs = torch.empty((n, 4096, 32), dtype=torch.uint8, device=dev)
    for i in range(0, n, 4096):
        j = min(n, i + 4096)
        blobs[i:j] = torch.randint(0, 256, (j - i, 4096, 32), dtype=torch.uint8, device=dev, generator=gen)
    blobs[:, :, 0] %= 0x73
    s = rk.KzgSettings(devices=[0], window_bits=int(os.environ.get("RAIKO_KZG_FULL_WINDOW_BITS", "0")))
    try:
        lib = _native.load()
        outs = {k: torch.zeros((n, w), dtype=torch.uint8, device=dev)
                for k, w in (("c", 48), ("vh", 32), ("x", 32), ("y", 32), ("p", 48), ("st", 1))}
        st = lib.rk_commit_prove_batch(s._ctx, blobs.data_ptr(), n, outs["c"].data_ptr(), outs["vh"].data_ptr(),
                                       outs["x"].data_ptr(), outs["y"].data_ptr(), outs["p"].data_ptr(), outs["st"].data_ptr())
        assert st == 0, _native.last_error()
        assert int(outs["st"].sum()) == 0
        h = {k: v.cpu().numpy() for k, v in outs.items()}
        for i in range(n):
            c = h["c"][i].tobytes()
            assert h["vh"][i].tobytes() == b"\x01" + hashlib.sha256(c).digest()[1:]

        host_blobs = blobs.cpu().numpy()

        def xy(i):
            return ref.proof_of_equivalence(host_blobs[i].tobytes(), h["vh"][i].tobytes())
        with ThreadPoolExecutor(max_workers=os.cpu_count() or 1) as ex:
            want = list(ex.map(xy, range(n)))
        for i in range(n):
            assert (h["x"][i].tobytes(), h["y"][i].tobytes()) == want[i], i

        assert rk.verify_kzg_proof_batch(outs["c"], outs["x"], outs["y"], outs["p"], s)
        for i in (0, n // 2 + 1, n - 1):                                         # byte-level spot check
            got = tuple(h[k][i].tobytes() for k in ("c", "vh", "x", "y", "p"))
            assert got == ref.commit_prove(host_blobs[i].tobytes())
        bad = outs["y"].clone()
        bad[n // 3, 31] ^= 1
        assert not rk.verify_kzg_proof_batch(outs["c"], outs["x"], bad, outs["p"], s)
    finally:
        s.close()

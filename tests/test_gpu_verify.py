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

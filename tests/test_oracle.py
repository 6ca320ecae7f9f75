"""The oracle against the reference's own known answers and the golden fixtures.

Reference anchors: the zero-blob versioned hash (eip4844.rs:147-160),
the pairing-level vector of test_verify_kzg_proof (eip4844.rs:162-184) with the two
negative cases of test_verify_kzg_proof_in_precompile (eip4844.rs:202-213)."""
import hashlib
import os

import pytest

from kzg_testlib import blob_from_recipe, synthetic_blob

REF_RAW = "/root/reference/kzg_settings_raw.bin"
REF_BINCODE = "/root/reference/lib/kzg_settings/zkcrypto_kzg_settings.bin"


def test_reference_kat_zero_blob(ref):
    # eip4844.rs:147-160: the only byte-level golden value in the reference
    import kzg_ref
    c = ref.commit(bytes(131072))
    assert c == b"\xc0" + bytes(47)
    assert "0x" + kzg_ref.versioned_hash(c).hex() == "0x010657f37554c781402a22917dee2f75def7ab966d7b770905398eba3c444014"
    assert kzg_ref.versioned_hash(c) == b"\x01" + hashlib.sha256(c).digest()[1:]


def test_lagrange_points_sum_to_generator(ref):
    # sum_i L_i(s) = 1  =>  blob of all ones commits to the G1 generator
    c = ref.commit((1).to_bytes(32, "big") * 4096)
    assert c.hex() == ("97f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac58"
                       "6c55e83ff97a1aeffb3af00adb22c6bb")


def test_c_oracle_matches_every_golden(ref, golden):
    for case in golden["cases"]:
        blob = blob_from_recipe(case["recipe"])
        assert hashlib.sha256(blob).hexdigest() == case["blob_sha256"]
        c, vh, x, y, pr = ref.commit_prove(blob)
        p0 = case["proofs"][0]
        assert (c.hex(), vh.hex(), x.hex(), y.hex(), pr.hex()) == (
            case["commitment"], case["versioned_hash"], p0["z"], p0["y"], p0["proof"]), case["name"]
        assert ref.proof_of_equivalence(blob, vh) == (x, y)
        for p in case["proofs"][1:]:
            pr2, y2 = ref.compute_proof(blob, bytes.fromhex(p["z"]))
            assert (pr2.hex(), y2.hex()) == (p["proof"], p["y"]), (case["name"], p["label"])


def test_c_oracle_rejects_noncanonical(ref, golden):
    import kzg_ref
    for e in golden["errors"]:
        with pytest.raises(kzg_ref.RefError):
            ref.commit(blob_from_recipe(e["recipe"]))
    with pytest.raises(kzg_ref.RefError):
        ref.commit(bytes(131071))


def test_survey_appendix_c_rows(golden):
    """SURVEY.md Appendix C rows, typed in from the survey (not from the generator)."""
    by = {c["name"]: c for c in golden["cases"]}
    assert by["C3_mod64"]["commitment"] == "8e961e44c3160e4a53bfe143b5214be9c6d9b170f2839c80cc20c90813c1e359bb61572a43aa7a0af25071c8de9137a5"
    fs5 = [p for p in by["C3_mod64"]["proofs"] if p["label"] == "fs5"][0]
    assert fs5["y"] == "099f129284b42d49f60fa048ec9917ecf459acf2ab8ee335829458a130204246"
    assert fs5["proof"] == "95a58d8850bedd003eaa91694932ee4c1ca4dbe8f8cf5117bfd84ce835d8174cf742a9c72d2020d4df64fff08d1a49db"
    om5 = [p for p in by["C3_mod64"]["proofs"] if p["label"] == "omega5"][0]
    assert om5["y"] == "202122232425262728292a2b2c2d2e2f303132333435363738393a3b3c3d3e3f"
    assert by["C5_syn0"]["proofs"][0]["proof"] == "9532d2e04e951e14ab8549edb373b279d9c19606994c238af091669f5c9ff24115159ed26a9a8b88416615d88bc2fc95"
    assert by["C6_syn1"]["versioned_hash"] == "011a8ace8b30aba01d729de886476a8df164e834485ad0c11d36ff2853f54ac4"
    assert by["C4_sparse"]["proofs"][0]["y"] == "13ed2ee659c9392c820300d034d7e98633408f317d0bc62d4a0babbb20949708"


def test_python_oracle_replays_reference_pairing_test(pyoracle):
    """eip4844.rs:162-214: blob byte k = k % 64, z = hash_to_bls_field([5;32]) verifies;
    a proof for another z and y + 1 do not."""
    o, s = pyoracle
    blob = bytes(k % 64 for k in range(131072))
    c = o.calc_kzg_proof_commitment(blob, s)
    x = o.hash_to_bls_field(bytes([5] * 32))
    proof, yb = o.compute_kzg_proof(blob, x, s)
    y = int.from_bytes(yb, "big")
    assert y == o.evaluate_polynomial_in_evaluation_form(o.deserialize_blob(blob), x, s)
    assert o.verify_kzg_proof(c, x, y, proof, s)
    assert not o.verify_kzg_proof(c, x, (y + 1) % o.R, proof, s)
    x6 = o.hash_to_bls_field(bytes([6] * 32))
    proof6, _ = o.compute_kzg_proof(blob, x6, s)
    assert not o.verify_kzg_proof(c, x6, y, proof6, s)


def test_python_and_c_oracles_agree_on_a_fresh_blob(pyoracle, ref):
    o, s = pyoracle
    blob = synthetic_blob(77, seed=123)
    c, vh, x, y, pr = ref.commit_prove(blob)
    assert o.calc_kzg_proof_commitment(blob, s) == c
    assert o.commitment_to_version_hash(c) == vh
    assert o.proof_of_equivalence(blob, vh, s) == (x, y)
    assert o.calc_kzg_proof(blob, vh, s) == pr


def test_pairing_is_bilinear(pyoracle):
    o, s = pyoracle
    g2 = s.g2[0]
    a, b = 0x1234567, 0x89ABCDE
    lhs = [(o.g1_mul(o.G1_GEN, a * b), g2), (o.g1_neg(o.g1_mul(o.G1_GEN, a)), o.g2_mul(g2, b))]
    assert o.pairings_product_is_one(lhs)
    assert not o.pairings_product_is_one([(o.G1_GEN, g2)])


@pytest.mark.skipif(not os.path.exists(REF_RAW), reason="reference tree not mounted (GPU box)")
def test_compact_setup_equals_reference_images(pyoracle, ref):
    """eip4844.rs:136-145 (test_kzg_settings_equivalence) analogue: the shipped compact image
    holds exactly the points of both reference images; their sha256 is pinned in SURVEY.md App. A."""
    import kzg_ref
    o, s = pyoracle
    raw = open(REF_RAW, "rb").read()
    bc = open(REF_BINCODE, "rb").read()
    assert hashlib.sha256(raw).hexdigest() == "b2fef63491219899427a1cf4fe09380cbc7d85e5969e105c13279056eac38092"
    assert hashlib.sha256(bc).hexdigest() == "1b4de2ed5ebaae9ff855851ece64012dc5b334f8a128f713f5dec8f88b472359"
    for img in (raw, bc):
        s2 = o.load_settings(img)
        assert s2.g1 == s.g1 and s2.g2 == s.g2 and s2.roots_brp == s.roots_brp
        blob = synthetic_blob(3)
        assert kzg_ref.RefSettings(img).commit(blob) == ref.commit(blob)


def test_blob_data_codec_oracle_roundtrip(pyoracle):
    """decode_blob_data restatement (lib/src/utils.rs:85-144) against an independent encoder,
    plus the reference's rejection rules."""
    import random
    o, _ = pyoracle
    rnd = random.Random(4)
    for n in (0, 1, 26, 27, 28, 122, 123, 124, 127, 250, 251, 4096, o.MAX_BLOB_DATA_SIZE - 1, o.MAX_BLOB_DATA_SIZE):
        d = bytes(rnd.randrange(256) for _ in range(n))
        b = o.encode_blob_data(d)
        assert len(b) == 131072 and o.decode_blob_data(b) == d
        assert all(b[32 * i] < 0x40 for i in range(4096))          # every field element stays canonical
    base = o.encode_blob_data(b"hello world" * 50)
    for mutate in (lambda b: b.__setitem__(1, 1), lambda b: b.__setitem__(32 * 5, b[32 * 5] | 0x80),
                   lambda b: b.__setitem__(131071, 1), lambda b: b.__setitem__(4, b[4] - 1), lambda b: b.__setitem__(2, 0xFF)):
        m = bytearray(base)
        mutate(m)
        assert o.decode_blob_data(bytes(m)) == b""


def test_call_site_host_logic():
    """Pure host parts of raiko_b200/callers.py (no GPU): hex decoding and proof-type selection."""
    from raiko_b200 import callers
    assert callers.blob_to_bytes("0X0a0B") == b"\x0a\x0b" and callers.blob_to_bytes("0x0x01") == b"\x01"
    assert callers.blob_to_bytes("xyz") == b""
    T, B = callers.VerifierType, callers.BlobProofType
    assert callers.get_blob_proof_type(T.SGX, B.ProofOfEquivalence) is B.ProofOfCommitment
    assert callers.get_blob_proof_type(T.SP1, B.ProofOfCommitment) is B.ProofOfEquivalence
    assert callers.get_blob_proof_type(T.NONE, B.ProofOfEquivalence, proof_of_equivalence_feature=False) is B.ProofOfCommitment
    assert B.from_str("ProofOfCommitment") is B.ProofOfCommitment
    import pytest
    with pytest.raises(ValueError):
        callers.calc_blob_versioned_hash("0x00")          # wrong length: "Could not create blob"
    with pytest.raises(LookupError):
        callers.find_tx_blob([], bytes(32))


def test_instance_hash_reference_kat():
    """lib/src/protocol_instance.rs:241-276 (test_calc_eip712_pi_hash): the public-input hash that
    proof_of_equivalence feeds, reproduced by raiko_b200/callers.py."""
    from raiko_b200 import callers
    assert callers.keccak256(b"").hex() == "c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470"
    trans = [bytes.fromhex(h) for h in (
        "07828133348460fab349c7e0e9fd8e08555cba34b34f215ffc846bfbce0e8f52",
        "e2105909de032b913abfa4c8b6101f9863d82be109ef32890b771ae214784efa",
        "abbd12b3bcb836b024c413bb8c9f58f5bb626d6d835f5554a8240933e40b2d3b",
        "00" * 32)]
    h = callers.instance_hash(167001, bytes.fromhex("4F3F0D5B22338f1f991a1a9686C7171389C97Ff7"), trans,
                              bytes.fromhex("741E45D08C70c1C232802711bBFe1B7C0E1acc55"),
                              bytes.fromhex("70997970C51812dc3A010C7d01b50e0d17dc79C8"),
                              bytes.fromhex("9608088f69e586867154a693565b4f3234f26f82d44ef43fb99fd774e7266024"), (0, 0))
    assert h.hex() == "dc1696a5289616fa5eaa9b6ce97d53765b79db948caedb6887f21a26e4c29511"
    # the (x, y) words change the hash; without the feature the tuple is two words shorter
    assert callers.instance_hash(167001, bytes(20), trans, bytes(20), bytes(20), bytes(32), (1, 2)) != \
        callers.instance_hash(167001, bytes(20), trans, bytes(20), bytes(20), bytes(32), (2, 1))
    assert len(callers.instance_hash(1, bytes(20), trans, bytes(20), bytes(20), bytes(32), proof_of_equivalence_feature=False)) == 32

"""The reference's call sites rewired to the GPU path (raiko_b200/callers.py)."""
import pytest

from kzg_testlib import blob_from_recipe, synthetic_blob

pytestmark = pytest.mark.gpu


def test_preflight_sidecar_matching(gpu_settings, golden, ref):
    from raiko_b200 import callers
    import kzg_ref
    blobs = [synthetic_blob(b, seed=606) for b in range(6)]                    # Cancun max per block
    sidecars = [{"index": str(i), "blob": ("0x" if i % 2 else "0X") + b.hex().upper(), "kzg_commitment": "", "kzg_proof": ""}
                for i, b in enumerate(blobs)]
    want_c = ref.commit(blobs[4])
    target = kzg_ref.versioned_hash(want_c)
    blob, commitment = callers.find_tx_blob(sidecars, target, gpu_settings)
    assert blob == blobs[4] and commitment == want_c
    assert callers.calc_blob_versioned_hash(sidecars[4]["blob"], gpu_settings) == target
    assert callers.blob_to_bytes("0xzz") == b"" and callers.blob_to_bytes("0x0A0b") == b"\x0a\x0b"
    with pytest.raises(LookupError):
        callers.find_tx_blob(sidecars, bytes(32), gpu_settings)
    with pytest.raises(LookupError):
        callers.find_tx_blob([], target, gpu_settings)
    # an undecodable sidecar is fatal only before the match (lazy find in the reference)
    broken_after = sidecars[:5] + [{"blob": "0xnothex"}]
    assert callers.find_tx_blob(broken_after, target, gpu_settings)[1] == want_c
    broken_before = [{"blob": "0xnothex"}] + sidecars
    with pytest.raises(ValueError):
        callers.find_tx_blob(broken_before, target, gpu_settings)
    bad = blob_from_recipe(golden["errors"][0]["recipe"])
    with pytest.raises(ValueError):
        callers.calc_blob_versioned_hash("0x" + bad.hex(), gpu_settings)


def test_protocol_instance_blob_branch(gpu_settings, golden):
    from raiko_b200 import callers
    case = [c for c in golden["cases"] if c["name"] == "C5_syn0"][0]
    blob = blob_from_recipe(case["recipe"])
    c = bytes.fromhex(case["commitment"])
    p0 = case["proofs"][0]
    vh, poe = callers.blob_tx_list_hash(blob, c, callers.VerifierType.SP1, settings=gpu_settings)
    assert vh.hex() == case["versioned_hash"]
    assert poe == (int.from_bytes(bytes.fromhex(p0["z"]), "little"), int.from_bytes(bytes.fromhex(p0["y"]), "little"))
    vh2, poe2 = callers.blob_tx_list_hash(blob, c, callers.VerifierType.SGX, settings=gpu_settings)
    assert vh2 == vh and poe2 == (0, 0)
    with pytest.raises(ValueError):
        callers.blob_tx_list_hash(blob, bytes.fromhex(golden["cases"][5]["commitment"]), callers.VerifierType.SGX, settings=gpu_settings)
    T, B = callers.VerifierType, callers.BlobProofType
    assert callers.get_blob_proof_type(T.NONE, B.ProofOfEquivalence) is B.ProofOfEquivalence
    assert callers.get_blob_proof_type(T.RISC0, B.ProofOfCommitment) is B.ProofOfEquivalence
    assert callers.get_blob_proof_type(T.SP1, B.ProofOfEquivalence, proof_of_equivalence_feature=False) is B.ProofOfCommitment
    assert B.from_str(" ProofOfEquivalence ") is B.ProofOfEquivalence
    with pytest.raises(ValueError):
        B.from_str("nope")


def test_kzg_proof_appended_to_proof(gpu_settings, golden):
    from raiko_b200 import callers
    case = [c for c in golden["cases"] if c["name"] == "C6_syn1"][0]
    blob = blob_from_recipe(case["recipe"])
    assert callers.kzg_proof_hex(blob, bytes.fromhex(case["commitment"]), gpu_settings) == case["proofs"][0]["proof"]
    assert callers.kzg_proof_hex(blob, None, gpu_settings) is None
    with pytest.raises(ValueError):
        callers.kzg_proof_hex(blob, b"\x00" * 47, gpu_settings)


def test_blob_data_codec(gpu_settings, pyoracle):
    """lib/src/utils.rs:85-144 on the GPU vs its statement-by-statement oracle restatement."""
    import random
    from raiko_b200 import callers
    o, _ = pyoracle
    rnd = random.Random(11)
    datas = [bytes(rnd.randrange(256) for _ in range(n)) for n in (0, 1, 27, 28, 122, 123, 124, 127, 250, 251, 5000, 130043, 130044)]
    blobs = [o.encode_blob_data(d) for d in datas]
    # invalid variants: wrong version, high bits set in a field element, dirty input tail, dirty output tail, length too large
    bad = []
    for k in range(5):
        b = bytearray(o.encode_blob_data(b"raiko-blob-data" * 300))
        if k == 0: b[1] = 1
        if k == 1: b[32 * 9] |= 0x40
        if k == 2: b[131071] = 7
        if k == 3: b[4] -= 1
        if k == 4: b[2] = 0xFF
        bad.append(bytes(b))
    rand_blob = bytes(rnd.randrange(256) for _ in range(131072))
    allb = blobs + bad + [rand_blob, bytes(131072)]
    got = callers.decode_blob_data_batch(allb, gpu_settings)
    assert got == [o.decode_blob_data(b) for b in allb]
    assert got[:len(datas)] == datas and all(g == b"" for g in got[len(datas):len(datas) + 5])
    assert callers.decode_blob_data(blobs[10], gpu_settings) == datas[10]


def test_reference_blob_fixture(gpu_settings):
    """The only real blob the reference ships (core/src/preflight.rs:480-525): GPU codec, GPU
    commitment / challenge / evaluation / proof against the committed fixture
    (tests/golden/reference_blob_preflight.json, made by tests/golden/make_reference_blob.py)."""
    import hashlib
    import raiko_b200 as rk
    from raiko_b200 import callers
    from test_callers_host import reference_blob
    fx, blob = reference_blob()
    dec = callers.decode_blob_data(blob, gpu_settings)
    assert len(dec) == 1200 and hashlib.sha256(dec).hexdigest() == fx["decoded_sha256"] and dec[:3].hex() == "f904ad"
    txs = callers.decode_transactions(dec)
    assert len(txs) == 3 and all(t[0] == 2 for t in txs)
    res = rk.commit_prove_batch([blob], gpu_settings)
    assert res.status == [0]
    assert (res.commitments[0].hex(), res.versioned_hashes[0].hex(), res.xs[0].hex(), res.ys[0].hex(), res.proofs[0].hex()) == \
        (fx["commitment"], fx["versioned_hash"], fx["x"], fx["y"], fx["proof"])
    assert rk.verify_kzg_proof(res.commitments[0], res.xs[0], res.ys[0], res.proofs[0], gpu_settings)
    # the blob is not zlib data: get_tx_list's unwrap_or_default gives an empty tx list, like the reference
    assert callers.get_tx_list(True, "taiko_a7", True, blob, gpu_settings) == b""
    assert callers.generate_transactions(True, "taiko_a7", True, blob, anchor_tx=b"\xc1\x01", settings=gpu_settings) == [b"\xc1\x01"]

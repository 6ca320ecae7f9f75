"""Host-only tail of the blob -> tx-list path (lib/src/utils.rs:13-79, 181-193) in
raiko_b200/callers.py, and the reference's only real blob (core/src/preflight.rs:480-525) through
the ORACLE codec.  No GPU: the GPU codec is compared with the same fixture in test_gpu_callers.py."""
import hashlib
import json
import os

import pytest

from kzg_testlib import ROOT

FIXTURE = os.path.join(ROOT, "tests", "golden", "reference_blob_preflight.json")


def reference_blob():
    fx = json.load(open(FIXTURE))
    blob = bytes.fromhex(fx["hex_prefix"].ljust(262144, "0"))        # format!("{:0<262144}", ...) at preflight.rs:525
    assert hashlib.sha256(blob).hexdigest() == fx["blob_sha256"]
    return fx, blob


def test_reference_blob_through_the_oracle_codec(pyoracle):
    o, _ = pyoracle
    fx, blob = reference_blob()
    dec = o.decode_blob_data(blob)
    assert len(dec) == fx["decoded_len"] == 1200                     # header 0x0004b0
    assert hashlib.sha256(dec).hexdigest() == fx["decoded_sha256"]
    assert dec[:3].hex() == "f904ad" and 3 + int.from_bytes(dec[1:3], "big") == len(dec)   # self-describing RLP list
    assert o.encode_blob_data(dec) == blob                           # the codec round-trips the reference's bytes


def test_reference_blob_commitment_fixture_is_pairing_valid(pyoracle):
    o, s = pyoracle
    fx, blob = reference_blob()
    c = bytes.fromhex(fx["commitment"])
    assert o.commitment_to_version_hash(c).hex() == fx["versioned_hash"]
    assert o.fr_to_bytes(o.get_evaluation_point(blob, bytes.fromhex(fx["versioned_hash"]))).hex() == fx["x"]
    assert o.verify_kzg_proof(c, int(fx["x"], 16), int(fx["y"], 16), bytes.fromhex(fx["proof"]), s)


def test_decode_transactions_on_the_reference_payload(pyoracle):
    from raiko_b200 import callers
    o, _ = pyoracle
    fx, blob = reference_blob()
    payload = o.decode_blob_data(blob)
    txs = callers.decode_transactions(payload)
    assert len(txs) == 3 and [t[0] for t in txs] == [2, 2, 2]        # three EIP-1559 envelopes
    assert [len(t) for t in txs] == [181, 504, 504] and sum(len(t) for t in txs) + (2 + 3 + 3) + 3 == len(payload)   # string headers b8xx, b9xxxx x2 + the list header
    # what the reference's own test does (preflight.rs:526-529): the RAW blob is not an RLP list -> empty
    assert callers.decode_transactions(blob) == []
    assert callers.decode_transactions(b"") == []
    assert callers.decode_transactions(payload[:-1]) == []           # truncated list
    assert callers.decode_transactions(b"\xc0") == []                # empty list
    assert callers.decode_transactions(b"\xc3\xc2\x01\x02") == [b"\xc2\x01\x02"]      # one legacy (list) item
    assert callers.decode_transactions(b"\xc2\x81\x05") == []        # non-canonical single byte string
    assert callers.decode_transactions(b"\xc4\x83\x02\xc0\xc0") == []   # typed payload is not ONE list


def test_zlib_and_tx_list_dispatch():
    from raiko_b200 import callers
    data = b"raiko tx list " * 1000
    z = callers.zlib_compress_data(data)
    assert callers.zlib_decompress_data(z) == data
    with pytest.raises(Exception):
        callers.zlib_decompress_data(z[:-4])
    with pytest.raises(Exception):
        callers.zlib_decompress_data(b"not zlib")
    # non-taiko: plain zlib, empty on failure (unwrap_or_default)
    assert callers.get_tx_list(False, "ethereum", False, z) == data
    assert callers.get_tx_list(False, "ethereum", False, b"junk") == b""
    # taiko call data: size limit before decompression (other networks) / after it (taiko_a7)
    big = callers.zlib_compress_data(bytes(callers.CALL_DATA_CAPACITY + 1))
    assert len(big) < callers.CALL_DATA_CAPACITY
    assert len(callers.get_tx_list(True, "taiko_mainnet", False, big)) == callers.CALL_DATA_CAPACITY + 1
    assert callers.get_tx_list(True, "taiko_a7", False, big) == b""
    assert callers.get_tx_list(True, "taiko_a7", False, z) == data
    assert callers.get_tx_list(True, "taiko_mainnet", False, bytes(callers.CALL_DATA_CAPACITY + 1)) == b""
    txs = callers.generate_transactions(False, "ethereum", False, callers.zlib_compress_data(b"\xc3\xc2\x01\x02"), anchor_tx=b"\xc1\x07")
    assert txs == [b"\xc1\x07", b"\xc2\x01\x02"]

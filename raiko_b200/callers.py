"""The reference's three call sites of the blob-KZG path, rewired to the batched GPU API
(SURVEY.md §8(f) ranks 2-3 and §2 row 4).  Host logic only: every KZG operation goes through
``raiko_b200.eip4844`` (the CUDA library).

* preflight sidecar matching      core/src/preflight.rs:301-316, 369-377 (+ the re-commit at :260)
* ProtocolInstance blob branch    lib/src/protocol_instance.rs:37-61, 189-203
* KZG proof appended to a proof   core/src/interfaces.rs:207-219
"""
from __future__ import annotations

import enum
from typing import List, Optional, Sequence, Tuple

from . import eip4844
from .eip4844 import KzgSettings


# ---------------------------------------------------------------------------------------
# preflight (core/src/preflight.rs)
# ---------------------------------------------------------------------------------------
def blob_to_bytes(blob_str: str) -> bytes:
    """preflight.rs:301-303: hex-decode, empty on failure."""
    s = blob_str.lower()
    while s.startswith("0x"):          # trim_start_matches("0x") strips repeated prefixes
        s = s[2:]
    try:
        return bytes.fromhex(s)
    except ValueError:
        return b""


def _decode_blob_strict(blob_str: str) -> bytes:
    s = blob_str.lower()
    while s.startswith("0x"):
        s = s[2:]
    try:
        return bytes.fromhex(s)
    except ValueError as e:
        raise ValueError("Could not decode blob") from e          # .expect("Could not decode blob")


def calc_blob_versioned_hash(blob_str: str, settings: Optional[KzgSettings] = None) -> bytes:
    """preflight.rs:305-316 (the reference panics where this raises)."""
    blob = _decode_blob_strict(blob_str)
    if len(blob) != eip4844.BYTES_PER_BLOB:
        raise ValueError("Could not create blob")
    try:
        commitment = eip4844.calc_kzg_proof_commitment(blob, settings)
    except eip4844.DeserializeBlob as e:
        raise ValueError("Could not deserialize blob") from e
    return eip4844.commitment_to_version_hash(commitment)


def find_tx_blob(sidecars: Sequence[dict], blob_hash: bytes,
                 settings: Optional[KzgSettings] = None) -> Tuple[bytes, bytes]:
    """get_blob_data_beacon's selection (preflight.rs:366-378) plus the commitment preflight
    recomputes at :260 -- in ONE batched launch over the (<= 6) sidecar blobs instead of up to
    6 + 1 separate MSMs.  Returns (blob bytes, 48-byte commitment) of the first sidecar whose
    versioned hash equals ``blob_hash``.

    Like the lazy ``find`` of the reference, a sidecar that cannot be decoded is fatal only if
    it comes before the match."""
    if not sidecars:
        raise LookupError("blob data not available anymore")                      # preflight.rs:367
    decoded: List[Optional[bytes]] = []
    for sc in sidecars:
        try:
            b = _decode_blob_strict(sc["blob"])
            decoded.append(b if len(b) == eip4844.BYTES_PER_BLOB else None)
        except ValueError:
            decoded.append(None)
    good = [b for b in decoded if b is not None]
    res = eip4844.commit_batch(good, settings) if good else None
    k = 0
    for b in decoded:
        if b is None:
            raise ValueError("Could not decode blob")
        if res.status[k] != 0:
            raise ValueError("Could not deserialize blob")
        if res.versioned_hashes[k] == bytes(blob_hash):
            return b, res.commitments[k]
        k += 1
    raise LookupError("no sidecar blob matches the versioned hash")               # ensure!(tx_blob.is_some())


# ---------------------------------------------------------------------------------------
# ProtocolInstance::new, blob branch (lib/src/protocol_instance.rs)
# ---------------------------------------------------------------------------------------
class BlobProofType(enum.Enum):            # lib/src/input.rs:90-103
    ProofOfCommitment = "ProofOfCommitment"
    ProofOfEquivalence = "ProofOfEquivalence"

    @classmethod
    def from_str(cls, s: str) -> "BlobProofType":                                # input.rs:105-115
        try:
            return cls(s.strip())
        except ValueError as e:
            raise ValueError("invalid blob proof type") from e


class VerifierType(enum.Enum):
    NONE = "None"
    SGX = "SGX"
    SP1 = "SP1"
    RISC0 = "RISC0"


def get_blob_proof_type(proof_type: VerifierType, hint: BlobProofType,
                        proof_of_equivalence_feature: bool = True) -> BlobProofType:
    """protocol_instance.rs:189-203."""
    if not proof_of_equivalence_feature:
        return BlobProofType.ProofOfCommitment
    return {VerifierType.NONE: hint, VerifierType.SGX: BlobProofType.ProofOfCommitment,
            VerifierType.SP1: BlobProofType.ProofOfEquivalence, VerifierType.RISC0: BlobProofType.ProofOfEquivalence}[proof_type]


def blob_tx_list_hash(tx_data: bytes, blob_commitment: bytes, proof_type: VerifierType = VerifierType.NONE,
                      hint: BlobProofType = BlobProofType.ProofOfCommitment, proof_of_equivalence_feature: bool = True,
                      settings: Optional[KzgSettings] = None) -> Tuple[bytes, Tuple[int, int]]:
    """protocol_instance.rs:37-61 -> (tx_list_hash = versioned hash, proof_of_equivalence (U256, U256)).

    Keeps the reference's quirk: the big-endian x, y bytes go through ``U256::from_le_bytes``
    (protocol_instance.rs:49-50)."""
    if len(blob_commitment) != 48:
        raise ValueError("blob commitment must be 48 bytes")
    versioned_hash = eip4844.commitment_to_version_hash(blob_commitment)
    poe = (0, 0)
    kind = get_blob_proof_type(proof_type, hint, proof_of_equivalence_feature)
    if kind is BlobProofType.ProofOfEquivalence:
        x, y = eip4844.proof_of_equivalence(tx_data, versioned_hash, settings)
        poe = (int.from_bytes(x, "little"), int.from_bytes(y, "little"))
    else:
        if bytes(blob_commitment) != eip4844.calc_kzg_proof_commitment(tx_data, settings):
            raise ValueError("blob commitment does not match the blob data")      # ensure!(commitment == ...)
    return versioned_hash, poe


# ---------------------------------------------------------------------------------------
# instance_hash (lib/src/protocol_instance.rs:165-185): where (x, y) end up
# ---------------------------------------------------------------------------------------
def keccak256(data: bytes) -> bytes:
    """Keccak-256 (the pre-NIST padding Ethereum uses); host-only glue, not on the GPU path."""
    rc = [0x0000000000000001, 0x0000000000008082, 0x800000000000808A, 0x8000000080008000, 0x000000000000808B,
          0x0000000080000001, 0x8000000080008081, 0x8000000000008009, 0x000000000000008A, 0x0000000000000088,
          0x0000000080008009, 0x000000008000000A, 0x000000008000808B, 0x800000000000008B, 0x8000000000008089,
          0x8000000000008003, 0x8000000000008002, 0x8000000000000080, 0x000000000000800A, 0x800000008000000A,
          0x8000000080008081, 0x8000000000008080, 0x0000000080000001, 0x8000000080008008]
    rot = [[0, 36, 3, 41, 18], [1, 44, 10, 45, 2], [62, 6, 43, 15, 61], [28, 55, 25, 21, 56], [27, 20, 39, 8, 14]]
    m64 = (1 << 64) - 1

    def rol(x, n):
        return ((x << n) | (x >> (64 - n))) & m64 if n else x
    rate = 136
    p = bytearray(data)
    p.append(0x01)
    while len(p) % rate:
        p.append(0)
    p[-1] |= 0x80
    a = [[0] * 5 for _ in range(5)]
    for off in range(0, len(p), rate):
        for i in range(rate // 8):
            a[i % 5][i // 5] ^= int.from_bytes(p[off + 8 * i:off + 8 * i + 8], "little")
        for rnd in range(24):
            c = [a[x][0] ^ a[x][1] ^ a[x][2] ^ a[x][3] ^ a[x][4] for x in range(5)]
            d = [c[(x - 1) % 5] ^ rol(c[(x + 1) % 5], 1) for x in range(5)]
            a = [[a[x][y] ^ d[x] for y in range(5)] for x in range(5)]
            b = [[0] * 5 for _ in range(5)]
            for x in range(5):
                for y in range(5):
                    b[y][(2 * x + 3 * y) % 5] = rol(a[x][y], rot[x][y])
            a = [[b[x][y] ^ ((~b[(x + 1) % 5][y]) & b[(x + 2) % 5][y]) for y in range(5)] for x in range(5)]
            a[0][0] ^= rc[rnd]
    return b"".join(a[i % 5][i // 5].to_bytes(8, "little") for i in range(4))


def instance_hash(chain_id: int, verifier_address: bytes, transition: Sequence[bytes], sgx_instance: bytes, prover: bytes,
                  meta_hash: bytes, proof_of_equivalence: Tuple[int, int] = (0, 0),
                  proof_of_equivalence_feature: bool = True) -> bytes:
    """protocol_instance.rs:165-185: keccak256 of abi.encode("VERIFY_PROOF", chainId, verifier,
    transition{parentHash, blockHash, stateRoot, graffiti}, newInstance, prover, metaHash
    [, proof_of_equivalence]) with the outer offset word skipped."""
    def word(v: int) -> bytes:
        return int(v).to_bytes(32, "big")

    def addr(a: bytes) -> bytes:
        if len(a) != 20:
            raise ValueError("address must be 20 bytes")
        return bytes(12) + bytes(a)
    if len(transition) != 4 or any(len(t) != 32 for t in transition) or len(meta_hash) != 32:
        raise ValueError("transition is 4 x bytes32, meta_hash bytes32")
    static = [word(chain_id), addr(verifier_address)] + [bytes(t) for t in transition] + [addr(sgx_instance), addr(prover), bytes(meta_hash)]
    if proof_of_equivalence_feature:
        static += [word(proof_of_equivalence[0]), word(proof_of_equivalence[1])]
    tag = b"VERIFY_PROOF"
    head_len = 32 * (1 + len(static))
    data = word(head_len) + b"".join(static) + word(len(tag)) + tag + bytes(32 - len(tag))
    return keccak256(data)


# ---------------------------------------------------------------------------------------
# blob -> tx-list codec (lib/src/utils.rs:85-144)
# ---------------------------------------------------------------------------------------
BLOB_DATA_STRIDE = 130048
MAX_BLOB_DATA_SIZE = (4 * 31 + 3) * 1024 - 4


def decode_blob_data_batch(blobs, settings: Optional[KzgSettings] = None) -> List[bytes]:
    """decode_blob_data for many blobs in one launch; invalid blobs decode to b\"\" like the reference."""
    import ctypes
    s = settings if settings is not None else eip4844.kzg_settings()
    buf, n = eip4844._blobs_buf(blobs)
    out = ctypes.create_string_buffer(BLOB_DATA_STRIDE * max(n, 1))
    lens = (ctypes.c_uint32 * max(n, 1))()
    eip4844._check(s._lib.rk_decode_blob_data_batch(s._ctx, buf.ptr, n, out, lens))
    raw = out.raw
    return [raw[i * BLOB_DATA_STRIDE:i * BLOB_DATA_STRIDE + lens[i]] for i in range(n)]


def decode_blob_data(blob: bytes, settings: Optional[KzgSettings] = None) -> bytes:
    """utils.rs:85-144."""
    if len(blob) != eip4844.BYTES_PER_BLOB:
        raise ValueError("blob must be 131072 bytes")
    return decode_blob_data_batch([blob], settings)[0]


# ---------------------------------------------------------------------------------------
# tx-list tail after the codec (lib/src/utils.rs:13-79, 181-193).  Host-only byte work: zlib and
# the RLP list framing.  What is NOT reproduced is reth's TransactionSigned decoding (signature
# recovery, per-type field validation): transactions come back as their raw encodings.
# ---------------------------------------------------------------------------------------
CALL_DATA_CAPACITY = 4096 * 31                     # utils.rs:84


def zlib_decompress_data(data: bytes) -> bytes:
    """utils.rs:181-186; raises on malformed input (callers use ``unwrap_or_default``)."""
    import zlib
    d = zlib.decompressobj()
    out = d.decompress(bytes(data))
    if not d.eof:
        raise ValueError("zlib stream is truncated")
    return out


def zlib_compress_data(data: bytes) -> bytes:
    """utils.rs:188-193 (the compressed bytes differ from libflate's, the round trip does not)."""
    import zlib
    return zlib.compress(bytes(data))


def _zlib_or_empty(data: bytes) -> bytes:
    try:
        return zlib_decompress_data(data)
    except Exception:  # noqa: BLE001 -- unwrap_or_default()
        return b""


def get_tx_list(is_taiko: bool, network: str, is_blob_data: bool, tx_list: bytes,
                settings: Optional[KzgSettings] = None) -> bytes:
    """utils.rs:27-56: blob data goes through decode_blob_data (GPU codec) then zlib; call data is
    size-checked before (other Taiko networks) or after (taiko_a7) decompression."""
    if is_taiko:
        if is_blob_data:
            return _zlib_or_empty(decode_blob_data(tx_list, settings))
        if network == "taiko_a7":
            out = _zlib_or_empty(tx_list)
            return out if len(out) <= CALL_DATA_CAPACITY else b""
        return _zlib_or_empty(tx_list) if len(tx_list) <= CALL_DATA_CAPACITY else b""
    return _zlib_or_empty(tx_list)


def _rlp_item(buf: bytes, pos: int) -> Tuple[int, int, bool]:
    """(payload start, payload end, is_list) of the RLP item at pos; raises ValueError when malformed."""
    if pos >= len(buf):
        raise ValueError("input too short")
    b = buf[pos]
    if b < 0x80:
        return pos, pos + 1, False
    short, is_list = (0x80, False) if b < 0xC0 else (0xC0, True)
    if b - short <= 55:
        start, ln = pos + 1, b - short
        if not is_list and ln == 1 and start < len(buf) and buf[start] < 0x80:
            raise ValueError("non-canonical single byte")
    else:
        nlen = b - short - 55
        if pos + 1 + nlen > len(buf) or buf[pos + 1] == 0:
            raise ValueError("bad length of length")
        ln = int.from_bytes(buf[pos + 1:pos + 1 + nlen], "big")
        if ln <= 55:
            raise ValueError("non-canonical length")
        start = pos + 1 + nlen
    if start + ln > len(buf):
        raise ValueError("input too short")
    return start, start + ln, is_list


def decode_transactions(tx_list: bytes) -> List[bytes]:
    """utils.rs:13-20: ``Vec::<TransactionSigned>::decode``; ANY failure yields an empty list (the
    reference then builds an empty block).  Each element is a legacy transaction (an RLP list,
    returned with its header) or a typed one (an RLP string ``type || payload``, returned without
    the string header, i.e. the EIP-2718 envelope)."""
    try:
        start, end, is_list = _rlp_item(tx_list, 0)
        if not is_list:
            raise ValueError("not a list")
        txs, pos = [], start
        while pos < end:
            s, e, lst = _rlp_item(tx_list, pos)
            if e > end:
                raise ValueError("item overruns the list")
            if lst:
                txs.append(bytes(tx_list[pos:e]))
            else:
                env = bytes(tx_list[s:e])
                if not env or env[0] > 0x7F:
                    raise ValueError("bad typed-transaction envelope")
                _s2, e2, l2 = _rlp_item(env, 1)
                if not l2 or e2 != len(env):
                    raise ValueError("typed transaction payload is not one list")
                txs.append(env)
            pos = e
        return txs
    except (ValueError, IndexError):
        return []


def generate_transactions(is_taiko: bool, network: str, is_blob_data: bool, tx_list: bytes,
                          anchor_tx: Optional[bytes] = None, settings: Optional[KzgSettings] = None) -> List[bytes]:
    """utils.rs:58-73: tx list from the raw data posted on chain, anchor transaction first."""
    txs = decode_transactions(get_tx_list(is_taiko, network, is_blob_data, tx_list, settings))
    if anchor_tx is not None:
        txs.insert(0, bytes(anchor_tx))
    return txs


# ---------------------------------------------------------------------------------------
# run_prover tail (core/src/interfaces.rs:207-219)
# ---------------------------------------------------------------------------------------
def kzg_proof_hex(tx_data: bytes, blob_commitment: Optional[bytes], settings: Optional[KzgSettings] = None) -> Optional[str]:
    """``proof.kzg_proof = Some(hex::encode(kzg_proof_to_bytes(&kzg_proof)))``: lower-case hex, no 0x."""
    if blob_commitment is None:
        return None
    if len(blob_commitment) != 48:
        raise ValueError("Could not convert blob commitment to version hash")
    proof = eip4844.calc_kzg_proof(tx_data, eip4844.commitment_to_version_hash(bytes(blob_commitment)), settings)
    return eip4844.kzg_proof_to_bytes(proof).hex()

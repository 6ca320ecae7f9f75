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

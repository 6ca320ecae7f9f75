//! Replays tests/golden/kzg_golden.json through the reference's own functions
//! (raiko_lib::primitives::eip4844, lib/src/primitives/eip4844.rs:44-99) and diffs every byte:
//! commitment, versioned hash, raiko challenge x, y, proof, and proofs at caller-given points.
//! Exit code 0 = the golden vectors (hence the oracle, hence the CUDA path that is bit-exact against
//! it) equal the reference byte for byte.
use raiko_lib::primitives::eip4844::{
    calc_kzg_proof, calc_kzg_proof_commitment, calc_kzg_proof_with_point, commitment_to_version_hash,
    get_evaluation_point, kzg_proof_to_bytes, proof_of_equivalence,
};
use kzg_traits::Fr as _; // ZFr::{from_bytes_unchecked, to_bytes}
use serde_json::Value;
use sha2::{Digest, Sha256};

const R_HEX: &str = "73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001";

/// big-endian 32-byte `v mod r` for a 32-byte big-endian digest (v < 2^256 < 3r: at most two subtractions)
fn reduce_mod_r(mut v: [u8; 32]) -> [u8; 32] {
    let r: Vec<u8> = hex::decode(R_HEX).unwrap();
    for _ in 0..2 {
        if v[..] >= r[..] {
            let mut borrow = 0i16;
            for i in (0..32).rev() {
                let d = v[i] as i16 - r[i] as i16 - borrow;
                v[i] = d.rem_euclid(256) as u8;
                borrow = if d < 0 { 1 } else { 0 };
            }
        }
    }
    v
}

/// the blob recipes of tests/kzg_testlib.py::blob_from_recipe
fn blob_from_recipe(rc: &Value) -> Vec<u8> {
    let n = 131072usize;
    let kind = rc["kind"].as_str().unwrap();
    let synthetic = |b: u64, seed: u64| -> Vec<u8> {
        let mut out = Vec::with_capacity(n);
        for i in 0..4096u32 {
            let mut h = Sha256::new();
            h.update(b"raiko-kzg-bench-v1");
            h.update(seed.to_le_bytes());
            h.update((b as u32).to_le_bytes());
            h.update(i.to_le_bytes());
            out.extend_from_slice(&reduce_mod_r(h.finalize().into()));
        }
        out
    };
    let be32 = |v: &Value| -> [u8; 32] {
        // decimal integer (possibly larger than u64) -> 32 big-endian bytes
        let s = v.to_string();
        let mut acc = [0u8; 32];
        for ch in s.trim_matches('"').bytes() {
            let mut carry = (ch - b'0') as u16;
            for i in (0..32).rev() {
                let t = acc[i] as u16 * 10 + carry;
                acc[i] = (t & 0xff) as u8;
                carry = t >> 8;
            }
        }
        acc
    };
    match kind {
        "zero" => vec![0u8; n],
        "const" => be32(&rc["value"]).repeat(4096),
        "mod64" => (0..n).map(|k| (k % 64) as u8).collect(),
        "sparse" => {
            let mut b = vec![0u8; n];
            let i = rc["index"].as_u64().unwrap() as usize;
            b[32 * i..32 * i + 32].copy_from_slice(&be32(&rc["value"]));
            b
        }
        "synthetic" => synthetic(rc["b"].as_u64().unwrap(), rc["seed"].as_u64().unwrap()),
        "noncanonical" => {
            let mut b = synthetic(9, 7);
            let i = rc["index"].as_u64().unwrap() as usize;
            b[32 * i..32 * i + 32].copy_from_slice(&hex::decode(rc["bytes"].as_str().unwrap()).unwrap());
            b
        }
        k => panic!("unknown recipe {k}"),
    }
}

fn main() {
    let path = std::env::args().nth(1).expect("usage: crosscheck <kzg_golden.json>");
    let doc: Value = serde_json::from_str(&std::fs::read_to_string(path).unwrap()).unwrap();
    let mut bad = 0usize;
    let mut checked = 0usize;
    let mut cmp = |what: &str, name: &str, got: &[u8], want: &str| {
        checked += 1;
        if hex::encode(got) != want {
            bad += 1;
            eprintln!("MISMATCH {name} {what}: reference {} golden {want}", hex::encode(got));
        }
    };
    for case in doc["cases"].as_array().unwrap() {
        let name = case["name"].as_str().unwrap();
        let blob = blob_from_recipe(&case["recipe"]);
        cmp("sha256(blob)", name, &Sha256::digest(&blob), case["blob_sha256"].as_str().unwrap());
        let c = calc_kzg_proof_commitment(&blob).expect("commitment");
        cmp("commitment", name, &c, case["commitment"].as_str().unwrap());
        let vh = commitment_to_version_hash(&c);
        cmp("versioned_hash", name, vh.as_slice(), case["versioned_hash"].as_str().unwrap());
        for p in case["proofs"].as_array().unwrap() {
            let label = p["label"].as_str().unwrap();
            let tag = format!("{name}/{label}");
            let z: [u8; 32] = hex::decode(p["z"].as_str().unwrap()).unwrap().try_into().unwrap();
            if label == "raiko" {
                // x = hash_to_bls_field(sha256(sha256(blob) || vh)) -- the byte order / reduction the survey
                // could only restate from memory is pinned HERE
                cmp("x", &tag, &get_evaluation_point(&blob, &vh).to_bytes(), p["z"].as_str().unwrap());
                let (x, y) = proof_of_equivalence(&blob, &vh).expect("proof_of_equivalence");
                cmp("poe.x", &tag, &x, p["z"].as_str().unwrap());
                cmp("poe.y", &tag, &y, p["y"].as_str().unwrap());
                let pi = calc_kzg_proof(&blob, &vh).expect("calc_kzg_proof");
                cmp("proof", &tag, &kzg_proof_to_bytes(&pi), p["proof"].as_str().unwrap());
            } else {
                let zf = kzg::kzg_types::ZFr::from_bytes_unchecked(&z).expect("z");
                let pi = calc_kzg_proof_with_point(&blob, zf).expect("calc_kzg_proof_with_point");
                cmp("proof", &tag, &kzg_proof_to_bytes(&pi), p["proof"].as_str().unwrap());
            }
        }
    }
    // blobs with a field element >= r must be rejected (Eip4844Error::DeserializeBlob, eip4844.rs:54-56)
    for case in doc["errors"].as_array().unwrap() {
        checked += 1;
        if calc_kzg_proof_commitment(&blob_from_recipe(&case["recipe"])).is_ok() {
            bad += 1;
            eprintln!("MISMATCH {}: reference accepted a non-canonical blob", case["name"]);
        }
    }
    println!("{checked} values compared with the reference, {bad} mismatches");
    std::process::exit(if bad == 0 { 0 } else { 1 });
}

"""Parity tests proper: the CUDA path, called through the C ABI, against the oracle and the
committed golden fixtures.  Bit-exact (integer / byte work).  Run with -m gpu on a B200."""
import ctypes
import os
import random
import subprocess
import sys

import pytest

from kzg_testlib import ROOT, blob_from_recipe, synthetic_blob

pytestmark = pytest.mark.gpu
R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001


def rand_blob(rnd, nonzero_fraction=1.0):
    out = bytearray()
    for _ in range(4096):
        v = rnd.randrange(R) if rnd.random() < nonzero_fraction else 0
        out += v.to_bytes(32, "big")
    return bytes(out)


def test_golden_vectors_batch_api(gpu_settings, golden):
    import raiko_b200 as rk
    cases = golden["cases"]
    blobs = [blob_from_recipe(c["recipe"]) for c in cases]
    res = rk.commit_prove_batch(blobs, gpu_settings)
    for i, c in enumerate(cases):
        p0 = c["proofs"][0]
        got = (res.commitments[i].hex(), res.versioned_hashes[i].hex(), res.xs[i].hex(), res.ys[i].hex(), res.proofs[i].hex())
        assert got == (c["commitment"], c["versioned_hash"], p0["z"], p0["y"], p0["proof"]), c["name"]
        assert res.status[i] == 0
    res = rk.commit_batch(blobs, gpu_settings)
    assert [x.hex() for x in res.commitments] == [c["commitment"] for c in cases]
    assert [x.hex() for x in res.versioned_hashes] == [c["versioned_hash"] for c in cases]


def test_golden_vectors_single_call_api(gpu_settings, golden):
    """The reference's own function names, one blob per call (eip4844.rs:44-99)."""
    import raiko_b200 as rk
    for c in golden["cases"]:
        blob = blob_from_recipe(c["recipe"])
        p0 = c["proofs"][0]
        commitment = rk.calc_kzg_proof_commitment(blob, gpu_settings)
        assert commitment.hex() == c["commitment"]
        assert rk.blob_to_kzg_commitment(blob, gpu_settings) == commitment
        vh = rk.commitment_to_version_hash(commitment)
        assert vh.hex() == c["versioned_hash"]
        assert rk.get_evaluation_point(blob, vh, gpu_settings).hex() == p0["z"]
        assert tuple(v.hex() for v in rk.proof_of_equivalence(blob, vh, gpu_settings)) == (p0["z"], p0["y"])
        assert rk.calc_kzg_proof(blob, vh, gpu_settings).hex() == p0["proof"]
        for p in c["proofs"]:
            z = bytes.fromhex(p["z"])
            proof, y = rk.compute_kzg_proof(blob, z, gpu_settings)
            assert (proof.hex(), y.hex()) == (p["proof"], p["y"]), (c["name"], p["label"])
            assert rk.kzg_proof_to_bytes(rk.calc_kzg_proof_with_point(blob, z, gpu_settings)).hex() == p["proof"]


def test_reference_kat_and_structural_anchor(gpu_settings):
    import raiko_b200 as rk
    c = rk.calc_kzg_proof_commitment(bytes(131072), gpu_settings)       # eip4844.rs:147-160
    assert "0x" + rk.commitment_to_version_hash(c).hex() == "0x010657f37554c781402a22917dee2f75def7ab966d7b770905398eba3c444014"
    ones = rk.calc_kzg_proof_commitment((1).to_bytes(32, "big") * 4096, gpu_settings)
    assert ones.hex().startswith("97f1d3a73197d794")                      # G1 generator


def test_deserialize_errors(gpu_settings, golden):
    import raiko_b200 as rk
    for e in golden["errors"]:
        blob = blob_from_recipe(e["recipe"])
        with pytest.raises(rk.DeserializeBlob):
            rk.calc_kzg_proof_commitment(blob, gpu_settings)
        with pytest.raises(rk.DeserializeBlob):
            rk.proof_of_equivalence(blob, bytes(32), gpu_settings)
        with pytest.raises(rk.DeserializeBlob):
            rk.calc_kzg_proof(blob, bytes(32), gpu_settings)
    for bad_len in (0, 32, 131071, 131073):
        with pytest.raises(rk.DeserializeBlob):
            rk.calc_kzg_proof_commitment(bytes(bad_len), gpu_settings)


def test_bad_blob_inside_batch_does_not_fail_the_batch(gpu_settings, golden, ref):
    import raiko_b200 as rk
    good = [synthetic_blob(b, seed=5) for b in range(3)]
    bad = blob_from_recipe(golden["errors"][1]["recipe"])
    blobs = [good[0], bad, good[1], bad, good[2]]
    res = rk.commit_prove_batch(blobs, gpu_settings)
    assert res.status == [0, 2, 0, 2, 0]
    for i in (1, 3):
        assert res.commitments[i] == bytes(48) and res.proofs[i] == bytes(48)
        assert res.xs[i] == bytes(32) and res.ys[i] == bytes(32) and res.versioned_hashes[i] == bytes(32)
    for i, g in zip((0, 2, 4), good):
        assert (res.commitments[i], res.versioned_hashes[i], res.xs[i], res.ys[i], res.proofs[i]) == ref.commit_prove(g)


def test_random_blobs_against_c_oracle(gpu_settings, ref):
    import raiko_b200 as rk
    rnd = random.Random(20241018)
    blobs = [rand_blob(rnd) for _ in range(6)] + [rand_blob(rnd, 0.01), rand_blob(rnd, 0.5)]
    # short scalars and extreme digits: every window digit at its maximum / carry boundary
    edge_vals = [R - 1, R - 2, 1, 2, (1 << 254) - 1, 1 << 254, (1 << 255) - 19 - (1 << 254), 0x7FFF, 0x8000, 0x4000, 0x3FFF,
                 int("7f" * 31, 16), int("80" * 31, 16) % R, int("ff" * 31, 16), (1 << 240) - 1, 1 << 240]
    blobs.append(b"".join(edge_vals[i % len(edge_vals)].to_bytes(32, "big") for i in range(4096)))
    res = rk.commit_prove_batch(blobs, gpu_settings)
    for i, blob in enumerate(blobs):
        assert (res.commitments[i], res.versioned_hashes[i], res.xs[i], res.ys[i], res.proofs[i]) == ref.commit_prove(blob), i
    zs = [rnd.randrange(1 << 256).to_bytes(32, "big") for _ in blobs]           # z >= r is reduced, not rejected
    res2 = rk.compute_kzg_proof_batch(blobs, zs, gpu_settings)
    for i, blob in enumerate(blobs):
        assert (res2.proofs[i], res2.ys[i]) == ref.compute_proof(blob, zs[i]), i


def test_chunked_pipeline_and_device_pointers(golden, ref):
    """n > chunk exercises slot reuse; torch CUDA tensors exercise the device-pointer path."""
    code = r'''
import os, sys, ctypes
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests")); sys.path.insert(0, os.path.join(%r, "oracle"))
import torch
import raiko_b200 as rk
from raiko_b200 import _native
import kzg_ref
from kzg_testlib import SETUP, synthetic_blob
s = rk.KzgSettings(window_bits=8)
ref = kzg_ref.RefSettings(open(SETUP, "rb").read())
n = 21
blobs = [synthetic_blob(b, seed=99) for b in range(n)]
host = rk.commit_prove_batch(blobs, s)
t = torch.frombuffer(bytearray(b"".join(blobs)), dtype=torch.uint8).cuda()
lib = _native.load()
outs = {k: torch.zeros((n, w), dtype=torch.uint8, device="cuda") for k, w in (("c", 48), ("vh", 32), ("x", 32), ("y", 32), ("p", 48), ("st", 1))}
st = lib.rk_commit_prove_batch(s._ctx, t.data_ptr(), n, outs["c"].data_ptr(), outs["vh"].data_ptr(), outs["x"].data_ptr(), outs["y"].data_ptr(), outs["p"].data_ptr(), outs["st"].data_ptr())
assert st == 0, _native.last_error()
torch.cuda.synchronize()
for i in range(n):
    want = ref.commit_prove(blobs[i]) if i %% 5 == 0 else None
    got = (host.commitments[i], host.versioned_hashes[i], host.xs[i], host.ys[i], host.proofs[i])
    dev = tuple(outs[k][i].cpu().numpy().tobytes() for k in ("c", "vh", "x", "y", "p"))
    assert got == dev, i
    if want: assert got == want, i
assert int(outs["st"].sum()) == 0
print("CHUNK_OK", s.stats()["total_launches"])
''' % (ROOT, ROOT, ROOT)
    env = dict(os.environ, RAIKO_KZG_CHUNK="4")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "CHUNK_OK" in out.stdout, out.stdout + out.stderr


def test_linearity_and_pairing_on_a_larger_batch(gpu_settings, pyoracle, ref):
    """Size-independent properties: commit is linear in the blob, and sampled
    (C, x, y, proof) tuples satisfy the pairing equation the reference tests (eip4844.rs:176-183)."""
    import raiko_b200 as rk
    o, s = pyoracle
    rnd = random.Random(7)
    n = 48
    blobs = [synthetic_blob(b, seed=31337) for b in range(n)]
    sums = [0] * 4096
    for blob in blobs:
        for i in range(4096):
            sums[i] += int.from_bytes(blob[32 * i:32 * i + 32], "big")
    sum_blob = b"".join((v % R).to_bytes(32, "big") for v in sums)
    res = rk.commit_prove_batch(blobs + [sum_blob], gpu_settings)
    acc = None
    for c in res.commitments[:n]:
        acc = o.g1_add(acc, o.g1_decompress(c))
    assert o.g1_compress(acc) == res.commitments[n]
    for i in rnd.sample(range(n), 2):
        assert o.verify_kzg_proof(res.commitments[i], int.from_bytes(res.xs[i], "big"), int.from_bytes(res.ys[i], "big"), res.proofs[i], s)
    i = rnd.randrange(n)
    assert (res.commitments[i], res.versioned_hashes[i], res.xs[i], res.ys[i], res.proofs[i]) == ref.commit_prove(blobs[i])


def test_settings_images_and_export(gpu_settings, setup_bytes):
    """Row a1: every on-disk form loads to the same table, and re-serialising reproduces the
    reference's files byte for byte (sha256 from SURVEY.md Appendix A)."""
    import hashlib
    import raiko_b200 as rk
    raw = gpu_settings.export("raw")
    bc = gpu_settings.export("bincode")
    assert len(raw) == 739624 and hashlib.sha256(raw).hexdigest() == "b2fef63491219899427a1cf4fe09380cbc7d85e5969e105c13279056eac38092"
    assert len(bc) == 1001905 and hashlib.sha256(bc).hexdigest() == "1b4de2ed5ebaae9ff855851ece64012dc5b334f8a128f713f5dec8f88b472359"
    blob = synthetic_blob(1)
    want = rk.calc_kzg_proof_commitment(blob, gpu_settings)
    for img in (raw, bc):
        s2 = rk.KzgSettings(img, window_bits=6)
        assert rk.calc_kzg_proof_commitment(blob, s2) == want
        s2.close()
    with pytest.raises(ValueError):
        rk.KzgSettings(setup_bytes[:-1])
    corrupted = bytearray(setup_bytes)
    corrupted[16 + 48 * 100 + 20] ^= 1
    with pytest.raises(ValueError):
        rk.KzgSettings(bytes(corrupted), window_bits=6)


@pytest.mark.parametrize("wb", [4, 11, 15])
def test_other_window_widths(wb, golden):
    import torch
    import raiko_b200 as rk
    if wb == 15 and torch.cuda.mem_get_info()[0] < 130e9:
        pytest.skip("not enough free HBM for the c=15 table")
    s = rk.KzgSettings(window_bits=wb)
    assert s.window_bits == wb
    cases = [c for c in golden["cases"] if c["name"] in ("C1_zero", "C3_mod64", "C5_syn0", "M1_max", "C6_syn1")]
    res = rk.commit_prove_batch([blob_from_recipe(c["recipe"]) for c in cases], s)
    for i, c in enumerate(cases):
        assert (res.commitments[i].hex(), res.proofs[i].hex(), res.ys[i].hex()) == (c["commitment"], c["proofs"][0]["proof"], c["proofs"][0]["y"])
    s.close()


def test_thread_safety(gpu_settings, golden):
    """Up to 16 requests are in flight in the reference (host/src/proof.rs:121)."""
    import threading
    import raiko_b200 as rk
    cases = golden["cases"][:8]
    blobs = [blob_from_recipe(c["recipe"]) for c in cases]
    errs = []

    def work(k):
        try:
            for rep in range(2):
                i = (k + rep) % len(cases)
                assert rk.calc_kzg_proof_commitment(blobs[i], gpu_settings).hex() == cases[i]["commitment"]
                vh = bytes.fromhex(cases[i]["versioned_hash"])
                assert rk.calc_kzg_proof(blobs[i], vh, gpu_settings).hex() == cases[i]["proofs"][0]["proof"]
        except Exception as e:  # noqa: BLE001
            errs.append(e)
    th = [threading.Thread(target=work, args=(k,)) for k in range(16)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs


def test_multi_gpu_sharding(golden, ref):
    """One context over every GPU of the box (the library's own sharding: contiguous shards, one host
    thread per device, host-side gather), host pointers, a batch size no device count divides, and
    more than one chunk per device."""
    code = r'''
import os, sys
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests")); sys.path.insert(0, os.path.join(%r, "oracle"))
import torch
import raiko_b200 as rk
import kzg_ref
from kzg_testlib import SETUP, synthetic_blob
ndev = torch.cuda.device_count()
s = rk.KzgSettings(devices=list(range(ndev)), window_bits=8)
assert s.num_devices == ndev
ref = kzg_ref.RefSettings(open(SETUP, "rb").read())
n = 8 * ndev + 5                                       # not divisible by 2, 4 or 8
blobs = [synthetic_blob(b, seed=4242) for b in range(n)]
bad = bytearray(blobs[n // 2]); bad[32 * 100:32 * 101] = b"\xff" * 32; blobs[n // 2] = bytes(bad)   # one non-canonical blob mid-batch
before = torch.cuda.current_device()
res = rk.commit_prove_batch(blobs, s)
assert torch.cuda.current_device() == before           # the library restores the caller's device
one = rk.KzgSettings(devices=[0], window_bits=8)
res1 = rk.commit_prove_batch(blobs, one)
for f in ("commitments", "versioned_hashes", "xs", "ys", "proofs", "status"):
    assert getattr(res, f) == getattr(res1, f), f
assert res.status[n // 2] == 2 and sum(res.status) == 2
for i in sorted({0, 1, n // 3, n - 1} | {n * g // ndev for g in range(ndev)} | {n * g // ndev - 1 for g in range(1, ndev)}):
    if i == n // 2: continue
    assert (res.commitments[i], res.versioned_hashes[i], res.xs[i], res.ys[i], res.proofs[i]) == ref.commit_prove(blobs[i]), i
# a device pointer on another GPU than the context's first device is refused, not dereferenced
if ndev >= 2:
    from raiko_b200 import _native
    lib = _native.load()
    t = torch.zeros(131072, dtype=torch.uint8, device="cuda:1")
    o = torch.zeros(48, dtype=torch.uint8, device="cuda:1")
    st = lib.rk_commit_batch(one._ctx, t.data_ptr(), 1, o.data_ptr(), None, None)
    assert st == _native.RK_ERR_ARG and "GPU 1" in _native.last_error(), (st, _native.last_error())
print("MULTI_OK ndev=%%d n=%%d launches=%%d" %% (ndev, n, s.stats()["total_launches"]))
''' % (ROOT, ROOT, ROOT)
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    env = dict(os.environ, RAIKO_KZG_CHUNK="3")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and "MULTI_OK" in out.stdout, out.stdout + out.stderr
    print(out.stdout.strip().splitlines()[-1])


def test_device_guard_and_window_report(gpu_settings):
    """Entry points leave the caller's current device alone and report an auto-narrowed window."""
    import torch
    import raiko_b200 as rk
    before = torch.cuda.current_device()
    rk.commit_batch([bytes(131072)], gpu_settings)
    assert torch.cuda.current_device() == before
    assert gpu_settings.window_reduced is False        # window_bits was passed explicitly


def test_synthetic_blobs_on_device_equal_the_host_generator(gpu_settings):
    """rk_synth_blobs (bench input) is byte-identical to tests/kzg_testlib.synthetic_blob,
    i.e. to the recipe of golden vectors C5 / C6; device and host outputs, first_blob offset."""
    import torch
    t = torch.zeros((5, 131072), dtype=torch.uint8, device="cuda")
    gpu_settings.synth_blobs(t, first_blob=0)
    h = t.cpu().numpy()
    for b in (0, 1, 4):
        assert h[b].tobytes() == synthetic_blob(b), b
    import numpy as np
    host = np.zeros((3, 131072), dtype=np.uint8)
    gpu_settings.synth_blobs(host, first_blob=70000, seed=7)
    assert host[2].tobytes() == synthetic_blob(70002, seed=7)


def test_wave_aware_tail_split(ref):
    """A chunk of more than one wave of one-warp-per-blob MSM work is launched as its whole waves
    plus a separately planned remainder (two MSM launches, two finalize launches): same bytes as the
    single-launch schedule (RAIKO_KZG_WAVE_TAIL=0), and spot-checked against the oracle."""
    code = r'''
import os, sys
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests")); sys.path.insert(0, os.path.join(%r, "oracle"))
import torch
import raiko_b200 as rk
from raiko_b200 import _native
lib = _native.load()
sm = torch.cuda.get_device_properties(0).multi_processor_count
n = sm * 16 + 301                                       # one wave of the affine kernel + a remainder
s = rk.KzgSettings(window_bits=8)
blobs = torch.zeros((n, 131072), dtype=torch.uint8, device="cuda")
s.synth_blobs(blobs, first_blob=1000, seed=5)
oc = torch.zeros((n, 48), dtype=torch.uint8, device="cuda"); ovh = torch.zeros((n, 32), dtype=torch.uint8, device="cuda"); ost = torch.zeros(n, dtype=torch.uint8, device="cuda")
s.stats_reset()
assert lib.rk_commit_batch(s._ctx, blobs.data_ptr(), n, oc.data_ptr(), ovh.data_ptr(), ost.data_ptr()) == 0, _native.last_error()
torch.cuda.synchronize()
print("WAVE", os.environ.get("RAIKO_KZG_WAVE_TAIL", "1"), s.stats()["msm_launches"], oc.cpu().numpy().tobytes().hex()[:0])
import hashlib
print("DIGEST", hashlib.sha256(oc.cpu().numpy().tobytes() + ovh.cpu().numpy().tobytes() + ost.cpu().numpy().tobytes()).hexdigest())
for i in (0, sm * 16 - 1, sm * 16, n - 1):
    print("ROW", i, blobs[i].cpu().numpy().tobytes().hex()[:0], oc[i].cpu().numpy().tobytes().hex())
''' % (ROOT, ROOT, ROOT)
    outs = {}
    for tail in ("1", "0"):
        env = dict(os.environ, RAIKO_KZG_WAVE_TAIL=tail)
        r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, r.stdout + r.stderr
        outs[tail] = r.stdout
    launches = {t: int([l for l in outs[t].splitlines() if l.startswith("WAVE")][0].split()[2]) for t in outs}
    assert launches == {"1": 2, "0": 1}, launches
    dig = {t: [l for l in outs[t].splitlines() if l.startswith("DIGEST")][0] for t in outs}
    assert dig["1"] == dig["0"]
    import torch
    sm = torch.cuda.get_device_properties(0).multi_processor_count
    for line in [l for l in outs["1"].splitlines() if l.startswith("ROW")]:
        _, i, c = line.split()
        assert ref.commit(synthetic_blob(1000 + int(i), seed=5)).hex() == c, i


def test_imad_peak_microbenchmark():
    from raiko_b200 import _native
    lib = _native.load()
    peak, clk = ctypes.c_double(), ctypes.c_double()
    assert lib.rk_measure_imad_peak(0, ctypes.byref(peak), ctypes.byref(clk)) == 0
    assert 5e12 < peak.value < 40e12


@pytest.mark.parametrize("n", [301, 3000])
def test_mid_size_batches_pick_many_splits(n, gpu_settings, ref):
    """Batches between one and a few waves use up to 128 warps per blob; the partial-sum
    buffer must hold blobs x splits entries (regression: 3968 blobs x 8 splits overflowed)."""
    import numpy as np
    import raiko_b200 as rk
    rng = np.random.default_rng(n)
    a = rng.integers(0, 256, size=(n, 4096, 32), dtype=np.uint8)
    a[:, :, 0] %= 0x73
    res = rk.commit_prove_batch(a, gpu_settings)
    assert sum(res.status) == 0
    for i in (0, n // 2, n - 1):
        assert (res.commitments[i], res.versioned_hashes[i], res.xs[i], res.ys[i], res.proofs[i]) == ref.commit_prove(a[i].tobytes()), i
    assert len(set(res.commitments)) == n


def test_concurrent_single_blob_calls_are_merged(gpu_settings, golden):
    """16 in-flight single-blob requests (host/src/proof.rs:121) must all get their own correct
    answer; the library merges waiting requests of one kind into one batch."""
    import threading
    import time
    import raiko_b200 as rk
    cases = golden["cases"]
    blobs = [blob_from_recipe(c["recipe"]) for c in cases]
    rk.calc_kzg_proof_commitment(blobs[0], gpu_settings)          # warm
    t0 = time.perf_counter()
    for i in range(16):
        rk.calc_kzg_proof_commitment(blobs[i % len(blobs)], gpu_settings)
    serial = time.perf_counter() - t0
    out, errs = [None] * 32, []

    def work(k):
        try:
            i = k % len(blobs)
            if k % 2 == 0:
                out[k] = ("c", i, rk.calc_kzg_proof_commitment(blobs[i], gpu_settings))
            else:
                out[k] = ("p", i, rk.calc_kzg_proof(blobs[i], bytes.fromhex(cases[i]["versioned_hash"]), gpu_settings))
        except Exception as e:  # noqa: BLE001
            errs.append(e)
    th = [threading.Thread(target=work, args=(k,)) for k in range(32)]
    t0 = time.perf_counter()
    [t.start() for t in th]
    [t.join() for t in th]
    merged = time.perf_counter() - t0
    assert not errs, errs
    for kind, i, val in out:
        want = cases[i]["commitment"] if kind == "c" else cases[i]["proofs"][0]["proof"]
        assert val.hex() == want
    print("16 serial commitments %.1f ms; 32 concurrent mixed calls %.1f ms" % (serial * 1e3, merged * 1e3))


@pytest.fixture
def msm_env():
    """Sets RAIKO_KZG_* knobs (read when a context is created) and restores them afterwards."""
    saved = {}

    def set_env(**kw):
        for k, v in kw.items():
            saved.setdefault(k, os.environ.get(k))
            os.environ[k] = str(v)
    yield set_env
    for k, v in saved.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v


@pytest.mark.parametrize("wb,chains", [(8, 64), (15, 64), (11, 10)])   # 10 chains: 3072 entries per lane leave a partial last round
def test_affine_msm_kernel_on_every_golden(wb, chains, golden, msm_env):
    """k_msm_affine (batched affine additions, shared safegcd inversion) forced onto EVERY golden
    case: one warp per blob, so each lane runs full chains through the zero blob (no entries at
    all), the sparse blob (one entry), constant blobs (every digit pattern equal), the in-domain
    evaluation points and the maximal field element.  Small batches normally take the XYZZ
    kernel; the planner is overridden here.  Also checks both kernels agree on random blobs."""
    import torch
    import raiko_b200 as rk
    if wb == 15 and torch.cuda.mem_get_info()[0] < 130e9:
        pytest.skip("not enough free HBM for the c=15 table")
    msm_env(RAIKO_KZG_MSM_AFFINE=2, RAIKO_KZG_AFFINE_MIN_ENTRIES=1, RAIKO_KZG_MAX_SPLITS_LOG2=0, RAIKO_KZG_AFFINE_CHAINS=chains)
    s = rk.KzgSettings(window_bits=wb)
    try:
        cases = golden["cases"]
        blobs = [blob_from_recipe(c["recipe"]) for c in cases]
        s.stats_enable(True)
        s.stats_reset()
        res = rk.commit_prove_batch(blobs, s)
        st = s.stats()
        assert st["msm_affine_launches"] == st["msm_launches"] == 2
        for i, c in enumerate(cases):
            p0 = c["proofs"][0]
            got = (res.commitments[i].hex(), res.versioned_hashes[i].hex(), res.xs[i].hex(), res.ys[i].hex(), res.proofs[i].hex())
            assert got == (c["commitment"], c["versioned_hash"], p0["z"], p0["y"], p0["proof"]), c["name"]
            assert res.status[i] == 0
        for c in cases:                                   # every (z, y, proof) incl. z inside the domain
            blob = blob_from_recipe(c["recipe"])
            zs = [bytes.fromhex(p["z"]) for p in c["proofs"]]
            r = rk.compute_kzg_proof_batch([blob] * len(zs), zs, s)
            assert [x.hex() for x in r.proofs] == [p["proof"] for p in c["proofs"]], c["name"]
        bads = [blob_from_recipe(e["recipe"]) for e in golden["errors"]]     # element 0 / 4095 / 2048 not canonical
        r = rk.commit_batch([blobs[4]] + bads + [blobs[5]], s)
        assert r.status == [0] + [2] * len(bads) + [0]
        assert r.commitments[0].hex() == cases[4]["commitment"] and r.commitments[-1].hex() == cases[5]["commitment"]
    finally:
        s.close()


def test_affine_and_xyzz_kernels_agree(ref, msm_env):
    """2368 random blobs (one full wave of the affine kernel: the planner picks k_msm_affine on its
    own) against the same batch with the affine kernel switched off, and a sample against the C
    oracle."""
    import torch
    import raiko_b200 as rk
    n = 2368
    g = torch.Generator(device="cuda")
    g.manual_seed(99)
    blobs = torch.randint(0, 256, (n, 4096, 32), dtype=torch.uint8, device="cuda", generator=g)
    blobs[:, :, 0] %= 0x73
    out = {}
    for mode in (1, 0):
        msm_env(RAIKO_KZG_MSM_AFFINE=mode)
        s = rk.KzgSettings(window_bits=8)
        try:
            s.stats_enable(True)
            s.stats_reset()
            out[mode] = rk.commit_prove_batch(blobs, s)
            st = s.stats()
            assert (st["msm_affine_launches"] > 0) == (mode == 1)
        finally:
            s.close()
    for f in ("commitments", "versioned_hashes", "xs", "ys", "proofs", "status"):
        assert getattr(out[0], f) == getattr(out[1], f), f
    for i in (0, 1183, n - 1):
        want = ref.commit_prove(blobs[i].cpu().numpy().tobytes())
        assert (out[1].commitments[i], out[1].versioned_hashes[i], out[1].xs[i], out[1].ys[i], out[1].proofs[i]) == want


@pytest.mark.parametrize("wb", [8, 15])
def test_affine_msm_kernel_digit_boundaries(wb, ref, msm_env):
    """Scalars built to sit on the signed-digit recoding boundaries of k_msm_affine's add-constant
    recoding (window value exactly 2^(c-1), 2^(c-1) +- 1, all-ones windows that carry into the next
    one, the largest canonical element, long runs of empty digits), forced through the affine
    kernel with one warp per blob and checked byte for byte against the C oracle."""
    import torch
    import raiko_b200 as rk
    if wb == 15 and torch.cuda.mem_get_info()[0] < 130e9:
        pytest.skip("not enough free HBM for the c=15 table")
    c = wb
    W = -(-255 // c)
    half = 1 << (c - 1)

    def blob(fn):
        return b"".join((fn(i) % R).to_bytes(32, "big") for i in range(4096))
    blobs = [
        blob(lambda i: half << (c * (i % W))),                                   # one digit exactly 2^(c-1)
        blob(lambda i: ((half + 1) << (c * (i % W))) | ((half - 1) << (c * ((i + 3) % (W - 1))))),
        blob(lambda i: ((1 << c) - 1) * sum(1 << (c * j) for j in range(i % (W - 1) + 1))),   # runs of all-ones windows
        blob(lambda i: R - 1 - i),                                               # top window at its maximum
        blob(lambda i: (i & 1) * (1 + (i << 200))),                              # every other element zero
        blob(lambda i: 1 << (i % 255)),                                          # single bits at every position
        blob(lambda i: (R - 1) if i == 4095 else 0),                             # one entry in the last lane only
    ]
    msm_env(RAIKO_KZG_MSM_AFFINE=2, RAIKO_KZG_AFFINE_MIN_ENTRIES=1, RAIKO_KZG_MAX_SPLITS_LOG2=0)
    s = rk.KzgSettings(window_bits=wb)
    try:
        s.stats_enable(True)
        s.stats_reset()
        res = rk.commit_prove_batch(blobs, s)
        assert s.stats()["msm_affine_launches"] == 2
        for i, b in enumerate(blobs):
            want = ref.commit_prove(b)
            got = (res.commitments[i], res.versioned_hashes[i], res.xs[i], res.ys[i], res.proofs[i])
            assert got == want and res.status[i] == 0, i
    finally:
        s.close()


def test_affine_msm_kernel_doubling_and_cancellation(pyoracle, setup_bytes, msm_env):
    """The additions k_msm_affine's shared-inversion formula cannot do -- a chain sum meeting a
    table point with the same x (doubling, cancellation) -- cannot happen with the ceremony's
    setup, whose discrete logs nobody knows.  A crafted setup with KNOWN logs makes them happen:
    L_i = (i + 2) G, except L_64 = L_0 and L_96 = L_32.  With c = 8, 64 chains and one warp per blob,
    lane 0 owns points 0, 32, 64, 96, ... and the windows of points 0 / 64 (and 32 / 96) land on the
    same chains, so equal scalars double a chain sum, opposite low digits cancel it, and a later
    entry restarts the emptied chain.  Commitments must still be bit-exact (slow path through the
    complete XYZZ formula, g1_affine_add_slow)."""
    import struct
    import raiko_b200 as rk
    o, real = pyoracle
    pts, acc = [], o.g1_add(o.G1_GEN, o.G1_GEN)
    for i in range(4096):
        pts.append(acc)
        acc = o.g1_add(acc, o.G1_GEN)
    pts[64], pts[96] = pts[0], pts[32]
    image = bytearray(b"RKZGTS02" + struct.pack("<II", 4096, len(real.g2)))
    for p in pts:
        image += o.g1_compress(p)
    image += setup_bytes[16 + 48 * 4096:]                      # the real G2 part (unused by commitments)
    custom = o.load_settings(bytes(image))

    def blob(vals):
        b = bytearray(131072)
        for i, v in vals.items():
            b[32 * i:32 * i + 32] = (v % R).to_bytes(32, "big")
        return bytes(b)
    blobs = [
        blob({0: 5, 64: 5}),                                   # 5 L_0 + 5 L_0: doubling inside chain 0
        blob({0: 5, 64: 251}),                                 # 251 = -5 + 256: chain 0 cancels to infinity, chain 1 gets 256 L_0
        blob({0: 5, 64: 251, 128: 7}),                         # ... and chain 0 restarts with 7 L_128
        blob({32: 0x0102, 96: 0x0102, 0: 3, 64: 253, 1: 9}),   # two chains double, one cancels, unrelated lane
        blob({i: (i * 0x9E3779B97F4A7C15 + 1) % R for i in range(4096)}),   # generic scalars on the degenerate setup
    ]
    want = [o.blob_to_kzg_commitment(b, custom) for b in blobs]
    assert want[0] == o.g1_compress(o.g1_mul(o.G1_GEN, 20)) and want[1] == o.g1_compress(o.g1_mul(o.G1_GEN, 512))
    for affine in (2, 0):                                      # the affine kernel forced, then the XYZZ kernel
        msm_env(RAIKO_KZG_MSM_AFFINE=affine, RAIKO_KZG_AFFINE_MIN_ENTRIES=1, RAIKO_KZG_MAX_SPLITS_LOG2=0)
        s = rk.KzgSettings(bytes(image), window_bits=8)
        try:
            s.stats_enable(True)
            s.stats_reset()
            res = rk.commit_batch(blobs, s)
            assert (s.stats()["msm_affine_launches"] == 1) == (affine == 2)
            assert res.commitments == want and res.status == [0] * len(blobs), "affine=%d" % affine
        finally:
            s.close()

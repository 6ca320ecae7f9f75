"""N > 1 host-side logic on CPU: world_size-2 gloo run of bench.py's sharding, gather and
max-over-ranks timing (the data path itself has no collective: blobs are independent)."""
import json
import os
import subprocess
import sys

from kzg_testlib import ROOT

WORKER = r'''
import os, sys, hashlib
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "oracle")); sys.path.insert(0, os.path.join(%r, "tests"))
import torch, torch.distributed as dist
import bench, kzg_ref
from kzg_testlib import SETUP, synthetic_blob
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
n_total = 5
lo, hi = bench.shard_range(n_total, rank, world)
ref = kzg_ref.RefSettings(open(SETUP, "rb").read())
mine = {i: ref.commit(synthetic_blob(i, seed=11)).hex() for i in range(lo, hi)}     # the oracle stands in for the GPU here
gathered = [None] * world
dist.all_gather_object(gathered, (lo, hi, mine))
tmax = bench.max_over_ranks(1.0 + rank, world)
bench.barrier(world)
if rank == 0:
    cover = sorted(i for lo_, hi_, m in gathered for i in range(lo_, hi_))
    merged = {}
    for _, _, m in gathered: merged.update(m)
    single = {i: ref.commit(synthetic_blob(i, seed=11)).hex() for i in range(n_total)}
    print("RESULT", cover == list(range(n_total)), merged == single, tmax == float(world), flush=True)
dist.destroy_process_group()
''' % (ROOT, ROOT, ROOT)


def test_world_size_2_gloo_sharding_and_timing(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29611", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r), LOCAL_RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for r in range(2)]
    outs = [p.communicate(timeout=300) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "RESULT True True True" in outs[0][0], outs


def test_shard_range_partitions_exactly():
    import bench
    for n in (0, 1, 5, 6, 4096, 65536, 65537):
        for world in (1, 2, 3, 4, 8):
            spans = [bench.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1


def test_reference_arm_runs_on_rank0_only():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--steps", "1", "--warmup", "0"], env=env, capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_reference_arm_json_contract():
    env = dict(os.environ, RANK="0", WORLD_SIZE="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "blobs/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]

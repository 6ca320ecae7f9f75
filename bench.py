#!/usr/bin/env python3
"""bench.py -- blob-KZG commit+prove throughput on B200 (BASELINE.json's metric).

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W

A "step" is one pass of the hot path over one batch of synthetic blobs per GPU:
commitment -> versioned hash -> raiko challenge -> evaluation -> KZG proof
(rk_commit_prove_batch).  Default batch: the 65,536-blob batch BASELINE.json quotes the
metric on (configs[3]); blobs are independent, so ranks shard with no collective and
scaling is weak (every rank runs its own 65,536-blob batch).

Prints ONE JSON line (rank 0).  `value` = blobs/s with inputs resident in HBM, `e2e` = the
same through the C ABI with HOST buffers (pinned; H2D/D2H inside the timed region),
`roofline` = the MSM kernel against the measured integer-multiply peak, `cpu_baseline` =
the C restatement of the reference path on this box's host cores.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

BLOB = 131072
METRIC = "kzg_blob_commit_prove_throughput"
UNIT = "blobs/s"
# SURVEY.md §8(d) / BASELINE.md §3 work model (shared with the judge):
MODEL_IMAD_PER_MSM = 540.7e6       # 90 112 adds x 10 Fp-mul x 600 IMAD (c = 13 bucket Pippenger)
MODEL_IMAD_PER_BLOB = 1.09e9       # commit + proof: 2 MSM + Fr side
# what this implementation actually issues on the integer-multiply pipe per table addition:
# madd-2008-s = 8 M + 2 S on 13 x 30-bit limbs = 8*(2*169+13) + 2*(91+169+13) IMAD.WIDE/IMAD
EXEC_IMAD_PER_ADD = 8 * 351 + 2 * 273
# k_msm_affine (batched affine additions): 5 M + 1 S per addition, plus per 64 additions one
# division-step inversion (27 iterations x 132 multiplies) and per 2176 additions 63 XYZZ
# chain-sum additions
EXEC_IMAD_PER_ADD_AFFINE = 5 * 351 + 273 + (27 * 132) // 64 + (63 * EXEC_IMAD_PER_ADD) // 2176


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=int(os.environ.get("RAIKO_BENCH_BATCH", "65536")),
                    help="blobs per GPU per step")
    ap.add_argument("--window-bits", type=int, default=int(os.environ.get("RAIKO_KZG_WINDOW_BITS", "0")))
    ap.add_argument("--cpu-sample", type=int, default=48, help="blobs timed for cpu_baseline")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


# --------------------------------------------------------------------------------------
# distributed plumbing (torch.distributed; one process per GPU)
# --------------------------------------------------------------------------------------
def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def shard_range(n_total: int, rank: int, world: int):
    """Contiguous shard [lo, hi) of n_total units for `rank` (same rule as the C library)."""
    return n_total * rank // world, n_total * (rank + 1) // world


def max_over_ranks(seconds: float, world: int, device=None) -> float:
    if world == 1:
        return seconds
    import torch
    import torch.distributed as dist
    t = torch.tensor([seconds], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier(world: int):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()


# --------------------------------------------------------------------------------------
# clocks during the timed region
# --------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index: int):
        self.index = index
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._th = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
                 0x80: "hw_power_brake_slowdown"}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:  # noqa: BLE001
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(0.1)

    def start(self):
        if self.nv is not None:
            self._th = threading.Thread(target=self._loop, daemon=True)
            self._th.start()

    def stop(self):
        self._stop.set()
        if self._th:
            self._th.join()
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# --------------------------------------------------------------------------------------
# CPU reference arm: the C restatement of the reference path (oracle/kzg_ref.c)
# --------------------------------------------------------------------------------------
def synth_blobs_host(n: int, seed: int):
    """Uniform canonical field elements (first byte < 0x73 => value < r), numpy, deterministic."""
    import numpy as np
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 256, size=(n, 4096, 32), dtype=np.uint8)
    a[:, :, 0] %= 0x73
    return a


def cpu_model() -> str:
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def cpu_commit_prove(blobs, threads: int):
    """Runs kzgref_commit_prove over `blobs` (list of bytes) on `threads` host threads.
    Returns (seconds, results)."""
    from concurrent.futures import ThreadPoolExecutor
    import kzg_ref
    from kzg_testlib import SETUP
    ref = kzg_ref.RefSettings(open(SETUP, "rb").read())
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:   # ctypes releases the GIL inside the C call
        res = list(ex.map(ref.commit_prove, blobs))
    return time.perf_counter() - t0, res


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return 0
    import kzg_ref
    kzg_ref.build()
    cores = os.cpu_count() or 1
    per_step = max(4 * cores, 64)      # ~1 s of work per host thread per step: thread start-up does not dominate
    arr = synth_blobs_host(per_step, seed=20241018)
    blobs = [arr[i].tobytes() for i in range(per_step)]
    for _ in range(args.warmup):
        cpu_commit_prove(blobs[:cores], cores)
    t = 0.0
    for _ in range(args.steps):
        dt, _ = cpu_commit_prove(blobs, cores)
        t += dt
    value = per_step * args.steps / t
    sample = "%d synthetic blobs per step on %d host threads (oracle/kzg_ref.c: C restatement of the reference's single-threaded CPU path, one blob per thread)" % (per_step, cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64 limbs (Fp 381-bit / Fr 255-bit Montgomery)", "data": "synthetic",
        "config": {"workload": "commit+versioned_hash+challenge+eval+proof per blob (bounded sample of the 65,536-blob batch)",
                   "blobs_per_step": per_step},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample, "cpu_model": cpu_model()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# --------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    rank, world, local = dist_env()
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    import raiko_b200 as rk
    from raiko_b200 import _native
    lib = _native.load()

    t0 = time.time()
    s = rk.KzgSettings(devices=[local], window_bits=args.window_bits)
    setup_s = time.time() - t0
    B = args.batch
    gen = torch.Generator(device=dev)
    gen.manual_seed(20241018 + rank)
    blobs = torch.empty((B, 4096, 32), dtype=torch.uint8, device=dev)
    step = 4096
    for i in range(0, B, step):       # chunked so the generator's temporaries stay small
        j = min(B, i + step)
        blobs[i:j] = torch.randint(0, 256, (j - i, 4096, 32), dtype=torch.uint8, device=dev, generator=gen)
    blobs[:, :, 0] %= 0x73            # canonical: every field element < r, full-width otherwise
    widths = (("c", 48), ("vh", 32), ("x", 32), ("y", 32), ("p", 48), ("st", 1))
    outs = {k: torch.zeros((B, w), dtype=torch.uint8, device=dev) for k, w in widths}

    def step_device():
        st = lib.rk_commit_prove_batch(s._ctx, blobs.data_ptr(), B, outs["c"].data_ptr(), outs["vh"].data_ptr(),
                                       outs["x"].data_ptr(), outs["y"].data_ptr(), outs["p"].data_ptr(), outs["st"].data_ptr())
        if st != 0:
            raise RuntimeError(_native.last_error())

    # ---- device-resident leg (value) ---------------------------------------------------
    for _ in range(args.warmup):
        step_device()
    torch.cuda.synchronize()
    s.stats_enable(True)
    s.stats_reset()
    clocks = ClockSampler(local)
    barrier(world)
    torch.cuda.synchronize()
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    w0 = time.perf_counter()
    for _ in range(args.steps):
        step_device()
    e1.record()
    torch.cuda.synchronize()
    w1 = time.perf_counter()
    barrier(world)
    clk = clocks.stop()
    dev_s = e0.elapsed_time(e1) / 1e3
    t_max = max_over_ranks(dev_s, world, dev)
    stats = s.stats()
    s.stats_enable(False)
    value = world * B * args.steps / t_max

    # ---- parity spot-check of the timed outputs against the oracle (rank 0) -------------
    parity_n = 0
    if rank == 0:
        import kzg_ref
        from kzg_testlib import SETUP
        ref = kzg_ref.RefSettings(open(SETUP, "rb").read())
        for i in (0, B // 3, B - 1):
            want = ref.commit_prove(blobs[i].cpu().numpy().tobytes())
            got = tuple(outs[k][i].cpu().numpy().tobytes() for k in ("c", "vh", "x", "y", "p"))
            assert got == want and int(outs["st"][i]) == 0, "bench output differs from the oracle at blob %d" % i
            parity_n += 1
        assert int(outs["st"].sum()) == 0

    # ---- end-to-end leg: host buffers through the C ABI ----------------------------------
    e2e = None
    if not args.no_e2e:
        Be = B
        host_kind = "pinned"
        try:
            h_in = torch.empty((Be, 4096, 32), dtype=torch.uint8, pin_memory=True)
        except RuntimeError:                      # not enough lockable memory on this box
            host_kind = "pageable"
            h_in = torch.empty((Be, 4096, 32), dtype=torch.uint8)
        h_in.copy_(blobs[:Be])
        h_out = {k: torch.zeros((Be, w), dtype=torch.uint8, pin_memory=(host_kind == "pinned")) for k, w in widths}

        def step_host():
            st = lib.rk_commit_prove_batch(s._ctx, h_in.data_ptr(), Be, h_out["c"].data_ptr(), h_out["vh"].data_ptr(),
                                           h_out["x"].data_ptr(), h_out["y"].data_ptr(), h_out["p"].data_ptr(), h_out["st"].data_ptr())
            if st != 0:
                raise RuntimeError(_native.last_error())
        e2e_steps = max(1, min(args.steps, 3))      # bounded so a large --steps does not double the run time
        for _ in range(min(args.warmup, 2)):
            step_host()
        torch.cuda.synchronize()
        barrier(world)
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            step_host()
        torch.cuda.synchronize()
        te = time.perf_counter() - t0
        barrier(world)
        te = max_over_ranks(te, world, dev)
        assert bytes(h_out["c"][Be - 1].numpy().tobytes()) == outs["c"][Be - 1].cpu().numpy().tobytes()
        e2e = {"value": world * Be * e2e_steps / te, "unit": UNIT, "h2d_bytes_per_step": Be * BLOB,
               "d2h_bytes_per_step": Be * 225, "host_memory": host_kind, "ms_per_step": 1e3 * te / e2e_steps,
               "steps": e2e_steps, "api": "rk_commit_prove_batch(host pointers): chunked H2D + kernels + D2H inside the timed region"}
        del h_in

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (k_msm) ------------------------------------------
    peak, clkattr = ctypes.c_double(), ctypes.c_double()
    lib.rk_measure_imad_peak(local, ctypes.byref(peak), ctypes.byref(clkattr))
    msm_launches = max(1, stats["msm_launches"])
    msm_s = stats["msm_ms"] / 1e3
    msms_done = 2 * B * args.steps                       # commit + proof MSM per blob
    avg_launch_s = msm_s / msm_launches
    msm_per_launch = msms_done / msm_launches
    achieved = MODEL_IMAD_PER_MSM * msm_per_launch / avg_launch_s          # algorithmic IMAD/s (work model)
    aff_adds = stats.get("msm_affine_point_adds", 0)
    executed = (EXEC_IMAD_PER_ADD * (stats["msm_point_adds"] - aff_adds) + EXEC_IMAD_PER_ADD_AFFINE * aff_adds) / msm_s   # issued on the multiply pipe
    msm_kernel = "k_msm_affine" if 2 * aff_adds > stats["msm_point_adds"] else "k_msm"
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "msm_dram_traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
        except Exception:  # noqa: BLE001
            traffic = None
    geom_w = stats["msm_point_adds"] / max(1, msms_done) / 4096
    roofline = {
        "bound": "imad", "kernel": msm_kernel, "achieved": achieved / 1e12, "peak": peak.value / 1e12, "unit": "TIMAD/s",
        "frac": achieved / peak.value, "traffic": traffic,
        "peak_source": "measured in this run: dependency-free mad.wide.u32 on all SMs (rk_measure_imad_peak); MEASURED_PEAKS.json has no integer peak",
        "work_model": "SURVEY.md 8(d): 540.7e6 IMAD per 4096-term MSM (c=13 bucket Pippenger, 600 IMAD per Fp mul)",
        "msm_share_of_step": msm_s / (dev_s if dev_s > 0 else 1),
        "avg_launch_ms": 1e3 * avg_launch_s, "msm_per_launch": msm_per_launch,
        "executed_timad_per_s": executed / 1e12, "frac_executed": executed / peak.value,
        "windows_per_scalar": geom_w,
        "table_read_gbs": stats["msm_point_adds"] * 96 / msm_s / 1e9,
        "point_adds_per_s": stats["msm_point_adds"] / msm_s, "affine_share_of_adds": aff_adds / max(1, stats["msm_point_adds"]),
        "whole_path_frac": value / world * MODEL_IMAD_PER_BLOB / peak.value,
    }

    hbm_peak = None
    try:
        hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:  # noqa: BLE001
        hbm_peak = 6650.0     # B200_PROFILING.md fallback
    alg_bytes_per_msm = 4096 * geom_w * 96 + BLOB          # table entries + the scalars
    roofline_hbm = {"bound": "hbm", "kernel": msm_kernel, "achieved": alg_bytes_per_msm * msm_per_launch / avg_launch_s / 1e9,
                    "peak": hbm_peak, "unit": "GB/s", "frac": alg_bytes_per_msm * msm_per_launch / avg_launch_s / 1e9 / hbm_peak,
                    "traffic": traffic, "note": "algorithmic bytes = one 96-byte table entry per addition + the scalars; k_msm_affine deliberately streams ~600 B more per addition (chain sums, prefix products) through HBM to save multiplies; the kernel is multiply/issue bound, not HBM bound"}

    # ---- CPU baseline (rank 0, N = 1 only): bounded sample on the host cores -------------
    cpu = None
    if world == 1:
        cores = os.cpu_count() or 1
        ncpu = max(args.cpu_sample, cores)
        sample = [blobs[i].cpu().numpy().tobytes() for i in range(ncpu)]
        dt, res = cpu_commit_prove(sample, cores)
        for i in (0, ncpu - 1):
            got = tuple(outs[k][i].cpu().numpy().tobytes() for k in ("c", "vh", "x", "y", "p"))
            assert got == res[i]
        dt1, _ = cpu_commit_prove(sample[:4], 1)
        cpu = {"value": ncpu / dt, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "first %d blobs of the timed batch, one blob per thread on %d threads (oracle/kzg_ref.c)" % (ncpu, cores),
               "single_thread_value": 4 / dt1, "cpu_model": cpu_model()}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32 limbs (13 x 30-bit Fp, 9 x 30-bit Fr Montgomery; IMAD.WIDE)", "data": "synthetic",
        "config": {"workload": "BASELINE.json configs[3]: 65,536-blob batch, commit+versioned_hash+challenge+eval+proof per blob"
                   if B == 65536 else "commit+versioned_hash+challenge+eval+proof per blob, %d blobs per GPU per step" % B,
                   "blobs_per_gpu_per_step": B, "window_bits": s.window_bits, "table_gb_per_gpu": s.table_bytes / 1e9,
                   "inputs": "resident in HBM (%.1f GB per GPU, > 126 MB L2: no flush needed)" % (B * BLOB / 1e9),
                   "parallelism": "dp%d (independent blobs, no collective)" % world, "setup_seconds": setup_s},
        "clocks": clk, "e2e": e2e, "gpu_launches": int(stats["total_launches"]),
        "roofline": roofline, "roofline_hbm": roofline_hbm, "cpu_baseline": cpu,
        "kernel_ms": {k: stats[k] for k in ("msm_ms", "fr_ms", "sha_ms", "finalize_ms")},
        "host_wall_ms_per_step": 1e3 * (w1 - w0) / args.steps, "parity_checked_blobs": parity_n,
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


_JSON_FD = None


def emit(line: dict):
    """The ONE JSON line, on the process's real stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def quiet_stdout():
    """Libraries print to stdout too (NCCL's version banner, for one).  Everything except the
    JSON line goes to stderr: fd 1 is pointed at fd 2 and the line is written to the saved fd."""
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)


def main():
    args = parse_args()
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # convenience: `python bench.py --gpus N` re-launches itself with one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29517"), os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    quiet_stdout()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())

#!/usr/bin/env python3
"""bench.py -- blob-KZG commit+prove throughput on B200 (BASELINE.json's metric).

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W

A "step" is one pass of the hot path over ONE batch of synthetic blobs: commitment -> versioned
hash -> raiko challenge -> evaluation -> KZG proof (rk_commit_prove_batch).  The batch is the
65,536-blob batch BASELINE.json quotes the metric on (configs[3]).  Blobs are independent, so at
N > 1 the batch is sharded contiguously over the ranks with no collective and the line says
"scaling": "strong" (rank r owns blobs shard_range(65536, r, N); `value` = 65,536 x steps / the
max-over-ranks device time).  `--scaling weak` gives every rank its own 65,536 blobs instead.

Prints ONE JSON line (rank 0).  `value` = blobs/s with inputs resident in HBM, `e2e` = the same
through the C ABI with HOST buffers (pinned; H2D/D2H inside the timed region), `roofline` = the MSM
kernel against the measured integer-multiply peak (executed instructions from the committed ncu
opcode histogram, profiles/roofline_inputs.json), `cpu_baseline` = the C restatement of the
reference path on this box's host cores.  At N = 1 the line also carries `configs`: BASELINE.json
configs[1], [2] and [4] (6-blob latency, 4096-blob commitment throughput, 4096-blob batch verify).
At N > 1 it carries `single_process`: ONE process, ONE context over all N GPUs, ONE
rk_commit_prove_batch call on the whole 65,536-blob host buffer (the library's own sharding).
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
SEED = 20241018                    # the synthetic family of golden vectors C5 / C6


def roofline_inputs():
    """Executed work per point addition of each MSM kernel, MEASURED: the dynamic SASS opcode
    histogram of an `ncu --set full --import-source on` capture (tools/make_roofline_inputs.py ->
    profiles/roofline_inputs.json, histograms beside it).  Nothing here is a typed-in constant."""
    with open(os.path.join(ROOT, "profiles", "roofline_inputs.json")) as f:
        return json.load(f)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=int(os.environ.get("RAIKO_BENCH_BATCH", "65536")),
                    help="blobs per step: the whole job's batch (strong scaling) or per GPU (weak)")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="N > 1: strong = ONE batch sharded over the ranks (default); weak = one batch per rank")
    ap.add_argument("--no-configs", action="store_true", help="N = 1: skip the configs[1]/[2]/[4] legs")
    ap.add_argument("--no-single-process", action="store_true", help="N > 1: skip the one-process multi-device leg")
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
def synth_blobs_host(n: int, seed: int = SEED, first: int = 0):
    """The SURVEY.md 8(d) synthetic blobs first .. first+n-1 on the host (hashlib): byte-identical to
    what rk_synth_blobs writes on the device."""
    from kzg_testlib import synthetic_blob
    return [synthetic_blob(first + b, seed) for b in range(n)]


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
    blobs = synth_blobs_host(per_step)
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
        "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "u64 limbs (Fp 381-bit / Fr 255-bit Montgomery)", "data": "synthetic (SURVEY.md 8(d) SHA-256 family, seed %d)" % SEED,
        "config": {"workload": "BASELINE.json configs[3]: commit+versioned_hash+challenge+eval+proof per blob; each step a bounded sample "
                               "(the first %d blobs) of the 65,536-blob batch" % per_step,
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
def _ptr(t):
    return t.data_ptr()


def alloc_host(shape, torch):
    """Pinned when the box lets us lock that much memory, else pageable (said in the line)."""
    try:
        return torch.empty(shape, dtype=torch.uint8, pin_memory=True), "pinned"
    except RuntimeError:
        return torch.empty(shape, dtype=torch.uint8), "pageable"


WIDTHS = (("c", 48), ("vh", 32), ("x", 32), ("y", 32), ("p", 48), ("st", 1))


def msm_roofline(stats, msm_kernel_ms, peak, inputs, world_value_per_gpu):
    """The MSM kernel against the measured integer-multiply peak.

    frac        executed multiply instructions (IMAD.WIDE / IMAD, thread level, from the committed
                ncu opcode histogram) per second / measured peak.  This is a pipe-utilisation number.
    slot_frac   the same with every IMAD.WIDE that has a 64-bit register addend counted twice: both the
                Ra,Rb,Rc64 form (product rows) and the Ra,imm,Rc64 form (reduction rows) issue at half rate
                on B200 (29 / 31 per clock per SM, profiles/r02/wide_operands3_ubench.txt).  This is the
                occupancy of the multiply pipe the instruction mix implies; it should agree with pipe_busy.
    model_frac  SURVEY.md 8(d) work model (540.7 M IMAD per MSM) per second / peak: a SPEED-UP over the
                model algorithm, not a utilisation -- the kernel executes far fewer multiplies.
    pipe_busy   sm__pipe_fmaheavy_cycles_active of the same ncu capture."""
    adds = stats["msm_point_adds"]
    aff_adds = stats.get("msm_affine_point_adds", 0)
    msm_s = stats["msm_ms"] / 1e3
    launches = max(1, stats["msm_launches"])
    kern = "k_msm_affine" if 2 * aff_adds > adds else "k_msm"
    ia, ix = inputs["k_msm_affine"], inputs["k_msm"]
    mul_inst = ia["multiply_thread_inst_per_unit"] * aff_adds + ix["multiply_thread_inst_per_unit"] * (adds - aff_adds)
    mul_slots = ia["mul_pipe_thread_slots_per_unit"] * aff_adds + ix["mul_pipe_thread_slots_per_unit"] * (adds - aff_adds)
    all_inst = ia["thread_inst_per_unit"] * aff_adds + ix["thread_inst_per_unit"] * (adds - aff_adds)
    msms = stats["_msms"]
    windows = adds / max(1, msms) / 4096.0
    k = inputs[kern]
    return {
        "bound": "imad", "kernel": kern, "unit": "T multiply instr/s (thread level)",
        "achieved": mul_inst / msm_s / 1e12, "peak": peak / 1e12, "frac": mul_inst / msm_s / peak,
        "slot_frac": mul_slots / msm_s / peak, "model_frac": MODEL_IMAD_PER_MSM * msms / msm_s / peak,
        "pipe_busy": k["pipe_busy_fmaheavy"], "issue_active": k["issue_active"],
        "traffic": k["dram_bytes_per_launch"], "traffic_units_per_launch": k["units_per_launch"],
        "dram_bytes_per_point_add": k["dram_bytes_per_unit"], "algorithmic_bytes_per_point_add": 96 + BLOB / (4096.0 * windows),
        "peak_source": "measured in this run: dependency-free mad.wide.u32 on all SMs (rk_measure_imad_peak); MEASURED_PEAKS.json has no integer peak",
        "work_source": "profiles/roofline_inputs.json <- %s (ncu --set full --import-source on; per-instruction executed counts)" % k["source_report"],
        "thread_inst_per_point_add": k["thread_inst_per_unit"], "multiply_inst_per_point_add": k["multiply_thread_inst_per_unit"],
        "all_inst_per_s_T": all_inst / msm_s / 1e12,
        "work_model": "SURVEY.md 8(d): 540.7e6 IMAD per 4096-term MSM (c=13 bucket Pippenger, 600 IMAD per Fp mul) -> model_frac",
        "msm_share_of_step": 1e3 * msm_s / msm_kernel_ms if msm_kernel_ms else None,
        "avg_launch_ms": 1e3 * msm_s / launches, "msm_per_launch": msms / launches,
        "windows_per_scalar": windows, "point_adds_per_s": adds / msm_s,
        "affine_share_of_adds": aff_adds / max(1, adds),
        "whole_path_model_frac": world_value_per_gpu * MODEL_IMAD_PER_BLOB / peak,
    }


def run_ours(args):
    import torch
    import torch.distributed as dist
    rank, world, local = dist_env()
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local)
    cpu_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        # host-side barrier for the phases in which a rank's GPU is used by ANOTHER process (the
        # single-process leg): an NCCL barrier would leave a kernel spinning on that GPU
        cpu_group = dist.new_group(backend="gloo")
    dev = torch.device("cuda", local)
    import raiko_b200 as rk
    from raiko_b200 import _native
    lib = _native.load()

    t0 = time.time()
    s = rk.KzgSettings(devices=[local], window_bits=args.window_bits)
    setup_s = time.time() - t0
    strong = args.scaling == "strong"
    total = args.batch if strong else args.batch * world           # blobs of the whole job per step
    lo, hi = shard_range(total, rank, world) if strong else (rank * args.batch, (rank + 1) * args.batch)
    B = hi - lo                                                     # this rank's shard
    blobs = torch.empty((B, 4096, 32), dtype=torch.uint8, device=dev)
    s.synth_blobs(blobs, first_blob=lo, seed=SEED)                  # SURVEY.md 8(d) family, generated on the device
    outs = {k: torch.zeros((B, w), dtype=torch.uint8, device=dev) for k, w in WIDTHS}

    def step_device():
        st = lib.rk_commit_prove_batch(s._ctx, _ptr(blobs), B, _ptr(outs["c"]), _ptr(outs["vh"]),
                                       _ptr(outs["x"]), _ptr(outs["y"]), _ptr(outs["p"]), _ptr(outs["st"]))
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
    stats["_msms"] = 2 * B * args.steps                   # commit + proof MSM per blob
    s.stats_enable(False)
    value = total * args.steps / t_max

    # ---- parity of the timed outputs (rank 0): golden vectors C5 / C6 + the oracle --------------
    parity_n = 0
    if rank == 0:
        import kzg_ref
        from kzg_testlib import SETUP, load_golden, synthetic_blob
        ref = kzg_ref.RefSettings(open(SETUP, "rb").read())
        if lo == 0 and B >= 2:
            # blobs 0 and 1 of this seed ARE SURVEY.md App. C rows C5 / C6 (tests/golden/kzg_golden.json)
            gold = {c["name"]: c for c in load_golden()["cases"]}
            for i, name in ((0, "C5_syn0"), (1, "C6_syn1")):
                g = gold[name]
                pr = [q for q in g["proofs"] if q["label"] == "raiko"][0]
                assert outs["c"][i].cpu().numpy().tobytes().hex() == g["commitment"], "bench blob %d: commitment differs from golden %s" % (i, name)
                assert outs["x"][i].cpu().numpy().tobytes().hex() == pr["z"] and outs["y"][i].cpu().numpy().tobytes().hex() == pr["y"]
                assert outs["p"][i].cpu().numpy().tobytes().hex() == pr["proof"], "bench blob %d: proof differs from golden %s" % (i, name)
                parity_n += 1
        for i in sorted({0, B // 3, B - 1}):
            host_blob = blobs[i].cpu().numpy().tobytes()
            assert host_blob == synthetic_blob(lo + i, SEED), "device-generated blob %d differs from the host generator" % (lo + i)
            want = ref.commit_prove(host_blob)
            got = tuple(outs[k][i].cpu().numpy().tobytes() for k in ("c", "vh", "x", "y", "p"))
            assert got == want and int(outs["st"][i]) == 0, "bench output differs from the oracle at blob %d" % i
            parity_n += 1
        assert int(outs["st"].sum()) == 0

    # ---- end-to-end leg: host buffers through the C ABI ----------------------------------
    e2e = None
    keep = {k: outs[k].cpu() for k, _ in WIDTHS}                     # for the cross-checks below
    if not args.no_e2e:
        h_in, host_kind = alloc_host((B, 4096, 32), torch)
        h_in.copy_(blobs)
        h_out = {k: torch.zeros((B, w), dtype=torch.uint8, pin_memory=(host_kind == "pinned")) for k, w in WIDTHS}

        def step_host():
            st = lib.rk_commit_prove_batch(s._ctx, _ptr(h_in), B, _ptr(h_out["c"]), _ptr(h_out["vh"]),
                                           _ptr(h_out["x"]), _ptr(h_out["y"]), _ptr(h_out["p"]), _ptr(h_out["st"]))
            if st != 0:
                raise RuntimeError(_native.last_error())
        e2e_steps = max(1, min(args.steps, 3))      # bounded so a large --steps does not double the run time
        for _ in range(min(args.warmup, 2)):
            step_host()
        torch.cuda.synchronize()
        barrier(world)
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            step_host()                             # returns when every output byte is in the host arrays
        te = time.perf_counter() - t0
        barrier(world)
        te = max_over_ranks(te, world, dev)
        for k, _ in WIDTHS:
            assert torch.equal(h_out[k], keep[k]), "host-buffer outputs differ from the device-resident leg (%s)" % k
        e2e = {"value": total * e2e_steps / te, "unit": UNIT, "h2d_bytes_per_step": total * BLOB,
               "d2h_bytes_per_step": total * 225, "host_memory": host_kind, "ms_per_step": 1e3 * te / e2e_steps,
               "steps": e2e_steps, "timing": "host wall clock around the blocking C-ABI calls, max over ranks",
               "api": "rk_commit_prove_batch(host pointers): chunked H2D + kernels + D2H inside the timed region"}
        del h_in, h_out

    # ---- N = 1: the other BASELINE.json configs, driver-visible ----------------------------------
    configs = None
    if world == 1 and not args.no_configs:
        configs = run_configs(torch, rk, lib, _native, s, blobs, keep)

    peak, clkattr = ctypes.c_double(), ctypes.c_double()
    if rank == 0:
        lib.rk_measure_imad_peak(local, ctypes.byref(peak), ctypes.byref(clkattr))

    # ---- CPU baseline (rank 0, N = 1 only): bounded sample on the host cores -------------
    cpu = None
    if world == 1:
        cores = os.cpu_count() or 1
        ncpu = max(args.cpu_sample, cores)
        sample = [blobs[i].cpu().numpy().tobytes() for i in range(ncpu)]
        dt, res = cpu_commit_prove(sample, cores)
        for i in range(ncpu):                       # every blob the CPU leg computed is also a parity check of the GPU outputs
            got = tuple(keep[k][i].numpy().tobytes() for k in ("c", "vh", "x", "y", "p"))
            assert got == res[i], "bench output differs from the oracle at blob %d" % i
        parity_n += ncpu
        dt1, _ = cpu_commit_prove(sample[:4], 1)
        cpu = {"value": ncpu / dt, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "first %d blobs of the timed batch, one blob per thread on %d threads (oracle/kzg_ref.c)" % (ncpu, cores),
               "single_thread_value": 4 / dt1, "cpu_model": cpu_model()}

    # ---- N > 1: ONE process, ONE context over all GPUs, ONE call on the whole host batch ---------
    single = None
    table_bytes, window_bits = s.table_bytes, s.window_bits
    if world > 1 and strong and not args.no_single_process:
        s.close()
        del blobs, outs
        torch.cuda.empty_cache()
        torch.cuda.synchronize()
        dist.barrier(group=cpu_group)               # every rank has released its table
        if rank == 0:
            single = run_single_process(torch, rk, lib, _native, args, world, total, keep, lo, hi)
        dist.barrier(group=cpu_group)               # host-side wait: no kernel parked on the other GPUs

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    inputs = roofline_inputs()
    roofline = msm_roofline(stats, 1e3 * dev_s, peak.value, inputs, value / world)
    hbm_peak = None
    try:
        hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        hbm_src = "MEASURED_PEAKS.json"
    except Exception:  # noqa: BLE001
        hbm_peak, hbm_src = 6650.0, "B200_PROFILING.md fallback"
    msm_s = stats["msm_ms"] / 1e3
    adds = stats["msm_point_adds"]
    alg_bytes = adds * 96 + stats["_msms"] * BLOB                      # one table entry per addition + the scalars
    roofline_hbm = {"bound": "hbm", "kernel": roofline["kernel"], "achieved": alg_bytes / msm_s / 1e9, "peak": hbm_peak, "peak_source": hbm_src,
                    "unit": "GB/s", "frac": alg_bytes / msm_s / 1e9 / hbm_peak,
                    "traffic_gbs": roofline["dram_bytes_per_point_add"] * adds / msm_s / 1e9,
                    "traffic_frac": roofline["dram_bytes_per_point_add"] * adds / msm_s / 1e9 / hbm_peak,
                    "note": "algorithmic bytes = one 96-byte table entry per addition + the scalars; k_msm_affine deliberately streams its "
                            "chain sums and prefix products through HBM (traffic_gbs, from the ncu capture) to save multiplies; the kernel "
                            "is multiply/issue bound, not HBM bound"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t_max / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "u32 limbs (13 x 30-bit Fp, 9 x 30-bit Fr Montgomery; IMAD.WIDE)",
        "data": "synthetic (SURVEY.md 8(d) SHA-256 family, seed %d, generated on the device by rk_synth_blobs; blobs 0/1 = golden C5/C6)" % SEED,
        "config": {"workload": ("BASELINE.json configs[3]: ONE 65,536-blob batch, commit+versioned_hash+challenge+eval+proof per blob"
                                if total == 65536 else "commit+versioned_hash+challenge+eval+proof per blob, %d blobs per step" % total)
                               + (", sharded contiguously over %d GPUs (host-side gather, no collective)" % world if world > 1 else ""),
                   "blobs_per_step": total, "blobs_per_gpu_per_step": B, "window_bits": window_bits, "table_gb_per_gpu": table_bytes / 1e9,
                   "inputs": "resident in HBM (%.1f GB per GPU, > 126 MB L2: no flush needed)" % (B * BLOB / 1e9),
                   "parallelism": "dp%d (independent blobs, no collective)" % world, "setup_seconds": setup_s},
        "clocks": clk, "e2e": e2e, "gpu_launches": int(stats["total_launches"]),
        "roofline": roofline, "roofline_hbm": roofline_hbm, "cpu_baseline": cpu,
        "kernel_ms": {k: stats[k] for k in ("msm_ms", "fr_ms", "sha_ms", "finalize_ms")},
        "host_wall_ms_per_step": 1e3 * (w1 - w0) / args.steps, "parity_checked_blobs": parity_n,
    }
    if configs is not None:
        line["configs"] = configs
    if single is not None:
        line["single_process"] = single
    emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_single_process(torch, rk, lib, _native, args, world, total, keep, lo, hi):
    """The API shape BASELINE.json's north_star describes: one host process, one context over all
    N GPUs, one rk_commit_prove_batch call on the whole batch in host memory; the library shards
    contiguously (one host thread per device) and gathers on the host."""
    t0 = time.time()
    sp = rk.KzgSettings(devices=list(range(world)), window_bits=args.window_bits)
    setup_s = time.time() - t0
    h_in, host_kind = alloc_host((total, 4096, 32), torch)
    sp.synth_blobs(h_in, first_blob=0, seed=SEED)
    h_out = {k: torch.zeros((total, w), dtype=torch.uint8, pin_memory=(host_kind == "pinned")) for k, w in WIDTHS}

    def step():
        st = lib.rk_commit_prove_batch(sp._ctx, _ptr(h_in), total, _ptr(h_out["c"]), _ptr(h_out["vh"]),
                                       _ptr(h_out["x"]), _ptr(h_out["y"]), _ptr(h_out["p"]), _ptr(h_out["st"]))
        if st != 0:
            raise RuntimeError(_native.last_error())
    step()
    steps = max(1, min(args.steps, 3))
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    # rank 0's shard was computed by this rank's own context before: the gathered outputs must agree
    for k, _ in WIDTHS:
        assert torch.equal(h_out[k][lo:hi], keep[k]), "single-process multi-device outputs differ from the per-rank leg (%s)" % k
    assert int(h_out["st"].sum()) == 0
    res = {"value": total * steps / dt, "unit": UNIT, "devices": world, "steps": steps, "ms_per_step": 1e3 * dt / steps,
           "host_memory": host_kind, "window_bits": sp.window_bits, "setup_seconds": setup_s,
           "h2d_bytes_per_step": total * BLOB, "d2h_bytes_per_step": total * 225,
           "api": "rk_kzg_ctx_create(devices=[0..%d]) + ONE rk_commit_prove_batch(host pointers, %d blobs)" % (world - 1, total),
           "checked": "blobs [%d, %d) equal the per-rank leg byte for byte; all statuses 0" % (lo, hi)}
    sp.close()
    return res


def run_configs(torch, rk, lib, _native, s, blobs, keep):
    """BASELINE.json configs[1] (6-blob block, latency), configs[2] (4096-blob commitment throughput)
    and configs[4] (verify_blob_kzg_proof_batch on 4096 blobs), through the C ABI, product calls only."""
    out = {}

    def timed(f, reps):
        f()
        ts = []
        for _ in range(reps):
            t = time.perf_counter()
            f()
            ts.append(time.perf_counter() - t)
        ts.sort()
        return ts[0], ts[len(ts) // 2]

    # configs[1]: Cancun-style block, 6 blobs, host buffers, commitment + proof at raiko's challenge
    six = [blobs[i].cpu().numpy().tobytes() for i in range(6)]
    best, med = timed(lambda: rk.commit_prove_batch(six, s), 20)
    res = rk.commit_prove_batch(six, s)
    for i in range(6):
        assert res.commitments[i] == keep["c"][i].numpy().tobytes() and res.proofs[i] == keep["p"][i].numpy().tobytes()
    out["config1_6_blobs_commit_prove_ms"] = {"best": 1e3 * best, "median": 1e3 * med, "buffers": "host (bytes)",
                                              "note": "BASELINE.json configs[1]; latency of ONE rk_commit_prove_batch call"}
    best, med = timed(lambda: rk.calc_kzg_proof_commitment(six[0], s), 20)
    out["single_blob_commitment_ms"] = {"best": 1e3 * best, "median": 1e3 * med}

    # configs[2]: 4096 blobs, commitment only, device resident, with its own roofline inputs
    n = 4096
    oc = torch.zeros((n, 48), dtype=torch.uint8, device=blobs.device)
    ovh = torch.zeros((n, 32), dtype=torch.uint8, device=blobs.device)
    ost = torch.zeros((n,), dtype=torch.uint8, device=blobs.device)

    def commit():
        assert lib.rk_commit_batch(s._ctx, _ptr(blobs), n, _ptr(oc), _ptr(ovh), _ptr(ost)) == 0, _native.last_error()
    commit()
    s.stats_enable(True)
    s.stats_reset()
    best, med = timed(commit, 5)
    st = s.stats()
    s.stats_enable(False)
    assert torch.equal(oc.cpu(), keep["c"][:n]) and int(ost.sum()) == 0
    out["config2_4096_blobs_commit_only"] = {"blobs_per_s_best": n / best, "blobs_per_s_median": n / med, "ms_best": 1e3 * best,
                                             "msm_ms_per_call": st["msm_ms"] / 6, "point_adds_per_s": st["msm_point_adds"] / (st["msm_ms"] / 1e3),
                                             "note": "BASELINE.json configs[2]; rk_commit_batch, inputs resident in HBM; 4096 = 1.73 waves of 2368 one-warp MSMs"}

    # configs[4]: verify_blob_kzg_proof_batch on those 4096 blobs (EIP-4844 challenges, Deneb spec)
    op = torch.zeros((n, 48), dtype=torch.uint8, device=blobs.device)
    assert lib.rk_compute_blob_kzg_proof_batch(s._ctx, _ptr(blobs), _ptr(oc), n, _ptr(op), _ptr(ost)) == 0, _native.last_error()
    craw, praw = oc.cpu().numpy().tobytes(), op.cpu().numpy().tobytes()
    ok = ctypes.c_int(0)

    def verify(p=praw):
        assert lib.rk_verify_blob_kzg_proof_batch(s._ctx, _ptr(blobs), ctypes.cast(ctypes.c_char_p(craw), ctypes.c_void_p),
                                                  ctypes.cast(ctypes.c_char_p(p), ctypes.c_void_p), n, ctypes.byref(ok)) == 0, _native.last_error()
    best, med = timed(verify, 3)
    accept = ok.value == 1
    bad = bytearray(praw)
    bad[48 * 1234:48 * 1235] = praw[48 * 1235:48 * 1236]
    verify(bytes(bad))
    reject = ok.value == 0
    assert accept and reject, "config 5: accept=%s reject_after_swap=%s" % (accept, reject)
    out["config4_4096_blobs_verify_batch"] = {"seconds_best": best, "seconds_median": med, "blobs_per_s": n / best, "accept": accept,
                                              "reject_after_proof_swap": reject,
                                              "note": "BASELINE.json configs[4]; proofs from rk_compute_blob_kzg_proof_batch; blobs in HBM"}
    return out


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

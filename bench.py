#!/usr/bin/env python
"""bench.py -- pileup + genotype-likelihood throughput (aligned bases / s) on N B200s.

    python bench.py --gpus 1 --steps 50 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --steps 5 --warmup 1      # the CPU arm (oracle port, host cores)

One "step" = one pass of the hot path over one batch: deposit every read of the batch into the
persistent device tables + the genotype pass over all G positions + candidate compaction.
Workload at N = 1: BASELINE.json configs[1]: synthetic SARS-CoV-2 Illumina 2x150
amplicon reads at 10,000x, 1,993,534 reads, 2.99e8 aligned bases, 0.498 GB algorithmic bytes.  At N > 1
every rank processes its own sample of the same geometry (config 4 style: independent samples, no
communication) -> weak scaling.  One line of JSON on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "covid-spings-variant-caller_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

METRIC = "pileup+GL aligned bases/sec"
UNIT = "aligned bases/s"
THRESH = dict(minBQ=30, minMQ=20, minDP=10, minAD=5, ratio=0.10)      # config_util/vc.config defaults
WORKLOAD = "config2: synthetic SARS-CoV-2 Illumina 2x150 amplicon, 10,000x, 1,993,534 reads (seed 20260101)"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def make_workload(seed: int, n_pairs: int):
    """config-2 batch; cached under /tmp so repeated runs on one box do not regenerate it."""
    from lvc_b200 import synth, packing
    cache = f"/tmp/lvc_bench_cfg2_{seed}_{n_pairs}.npz"
    if os.path.exists(cache):
        z = np.load(cache)
        b = packing.ReadBatch(z["pos"], z["flag"], z["mapq"], z["keep"], z["cigar_off"], z["cigar"], z["seq_off"],
                              z["seq4"], z["qual"])
        return str(z["ref"]), b
    t = time.time()
    ref, b = synth.amplicon_sample(seed=seed, n_pairs=n_pairs)
    log(f"[bench] generated workload seed={seed}: {b.n_reads} reads in {time.time() - t:.1f}s")
    try:
        np.savez(cache, ref=np.array(ref), pos=b.pos, flag=b.flag, mapq=b.mapq, keep=b.keep, cigar_off=b.cigar_off,
                 cigar=b.cigar, seq_off=b.seq_off, seq4=b.seq4, qual=b.qual)
    except Exception as e:  # cache is best effort
        log("[bench] cache write failed:", e)
    return ref, b


class ClockSampler(threading.Thread):
    """NVML polling of SM clock + throttle reasons during the timed region (in-process: the region can be
    shorter than nvidia-smi's sampling period)."""
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reason_bits, self.stop_flag, self.ok = index, [], 0, False, False
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.dev = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.dev, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:
            log("[bench] NVML unavailable:", e)

    def run(self):
        if not self.ok:
            return
        while not self.stop_flag:
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.dev, self.nv.NVML_CLOCK_SM))
                self.reason_bits |= int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.dev)) \
                    if hasattr(self.nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else int(self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.dev))
            except Exception:
                pass
            time.sleep(0.002)

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        reasons = [n for bit, n in self.REASONS.items() if self.reason_bits & bit and n != "gpu_idle"]
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(self.samples)}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic_bytes():
    """dram bytes per launch of the tiled deposit kernel from the committed ncu capture, if any."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get("k_deposit_tile_dram_bytes_per_launch")
        except Exception:
            return None
    return None


def cpu_oracle_throughput(ref, batch, min_seconds: float, max_reps: int):
    """the oracle port (oracle/oracle.c, single thread) on the same batch: deposit + genotype per rep."""
    from oracle.c_oracle import COracle
    reps, t0 = 0, time.perf_counter()
    while reps < max_reps and (reps == 0 or time.perf_counter() - t0 < min_seconds):
        co = COracle(ref, THRESH["minBQ"], THRESH["minMQ"])
        co.process(batch)
        co.genotype(THRESH["minDP"], THRESH["minAD"], THRESH["ratio"])
        reps += 1
    dt = time.perf_counter() - t0
    return batch.aligned_bases() * reps / dt, reps, dt


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (the reference itself is pure Python + pysam, which
    this image cannot run; the C port of its algorithm is the stronger baseline) on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ref, batch = make_workload(20260101, 996_767)
    total_steps = args.steps + args.warmup
    frac = min(1.0, 150.0 / (1.3 * max(total_steps, 1)))
    n = max(1000, int(batch.n_reads * frac))
    sample = batch if n >= batch.n_reads else batch.slice(0, n)
    from oracle.c_oracle import COracle
    for _ in range(args.warmup):
        co = COracle(ref, THRESH["minBQ"], THRESH["minMQ"]); co.process(sample)
        co.genotype(THRESH["minDP"], THRESH["minAD"], THRESH["ratio"])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        co = COracle(ref, THRESH["minBQ"], THRESH["minMQ"]); co.process(sample)
        co.genotype(THRESH["minDP"], THRESH["minAD"], THRESH["ratio"])
    dt = time.perf_counter() - t0
    value = sample.aligned_bases() * args.steps / dt
    desc = f"first {sample.n_reads} of {batch.n_reads} reads of the workload per step"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32+f64", "data": "synthetic",
            "config": {"workload": WORKLOAD},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port", "sample": desc,
                             "host_cores": os.cpu_count()},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--kernel-impl", type=int, default=0, help="0 auto (4 for short reads, 3 for long reads), 1 general kernel, 2 tiled kernel (raw TMA staging), 3 warp per read, 4 tiled kernel (4-bit keys)")
    ap.add_argument("--n-pairs", type=int, default=996_767)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    from lvc_b200 import capi, packing, records
    from variant_caller.live_variant_caller import LiveVariantCaller

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    # every rank owns an independent sample of the same geometry (rank 0 = the canonical config-2 seed)
    ref, batch = make_workload(20260101 + 1000 * rank, args.n_pairs)
    G = len(ref)
    bases = batch.aligned_bases()
    alg_bytes = batch.algorithmic_bytes(G)
    alg_bytes_deposit = alg_bytes - 52 * G
    e_lut, om_lut = records.phred_luts()

    # ---- device-resident inputs
    dev = torch.device("cuda", local)
    t_arr = {}
    for name in ("pos", "flag", "mapq", "keep", "cigar_off", "cigar", "seq_off", "seq4", "qual"):
        a = getattr(batch, name)
        t_arr[name] = torch.from_numpy(a.view(np.uint8).reshape(-1)).to(dev)
    dbatch = capi.Handle.make_batch(batch.n_reads, batch.n_cigar, batch.n_qual,
                                    *[t_arr[k].data_ptr() for k in ("pos", "flag", "mapq", "keep", "cigar_off", "cigar",
                                                                    "seq_off", "seq4", "qual")])
    stream = torch.cuda.Stream(device=dev)          # a real (non-null) stream: the library launches on it
    torch.cuda.set_stream(stream)
    h = capi.Handle(ref.encode("latin-1"), THRESH["minBQ"], THRESH["minMQ"], device=local, stream=stream.cuda_stream)
    h.set_impl(args.kernel_impl)

    def step_async():
        h.push_batch_device_async(dbatch)
        h.genotype_device_async(THRESH["minDP"], THRESH["minAD"], THRESH["ratio"], e_lut, om_lut)

    # ---- warm-up: the first push is synchronous (allocates the quality planes, may replay)
    h.push_batch_device(dbatch)
    h.genotype_device(THRESH["minDP"], THRESH["minAD"], THRESH["ratio"], e_lut, om_lut)
    n_cand = len(h.fetch_candidates())
    for _ in range(args.warmup - 1):
        step_async()
    h.check_async()

    def barrier():
        if world > 1:
            dist.barrier()

    sampler = ClockSampler(local)

    def timed_region(kernel_events: bool):
        """K steps between two events on the launching stream; with kernel_events the library also brackets
        every kernel launch with its own pair of events (which serialises the two kernels of a step)."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        h.set_timing(kernel_events)
        h.get_timing(0); h.get_timing(1); h.get_timing(2)
        barrier()
        torch.cuda.synchronize()
        n0 = h.launch_count
        e0.record(stream)
        for _ in range(args.steps):
            step_async()
        e1.record(stream)
        torch.cuda.synchronize()
        barrier()
        h.set_timing(False)
        return e0.elapsed_time(e1), h.launch_count - n0

    sampler.start()
    # (1) the step time: the two kernels of a step back to back, nothing else on the stream
    ms, launches = timed_region(False)
    # (2) the same K steps again with per-kernel events: the kernels' own durations for the roofline
    ms_ev, _ = timed_region(True)
    sampler.stop_flag = True
    h.check_async()
    tile_ms, tile_n = h.get_timing(0)
    gen_ms, gen_n = h.get_timing(1)
    geno_ms, geno_n = h.get_timing(2)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * bases * args.steps / (ms * 1e-3)

    # ---- end to end through the public API: pinned host SoA -> process_batch -> prepare_variants
    fasta = f"/tmp/lvc_bench_ref_{rank}.fasta"
    with open(fasta, "w") as fh:
        fh.write(">NC_045512.2\n" + ref + "\n")
    lvc = LiveVariantCaller(fasta, THRESH["minBQ"], THRESH["minMQ"], THRESH["minDP"], THRESH["minAD"], THRESH["ratio"], 1,
                            device=local)
    lvc._handle.set_stream(stream.cuda_stream)
    lvc._handle.set_impl(args.kernel_impl)
    pinned = packing.pin_batch(batch)
    h2d = sum(getattr(batch, k).nbytes for k in ("pos", "flag", "mapq", "keep", "cigar_off", "cigar", "seq_off")) \
        + (batch.n_qual + 1) // 2 + batch.n_qual
    for _ in range(2):
        lvc.process_batch(pinned)
        variants = lvc.prepare_variants()
    payload0 = lvc._handle.h2d_payload_bytes
    barrier()
    torch.cuda.synchronize()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record(stream)
    for _ in range(args.e2e_steps):
        lvc.process_batch(pinned)
        variants = lvc.prepare_variants()
    f1.record(stream)
    torch.cuda.synchronize()
    barrier()
    e2e_ms = f0.elapsed_time(f1)
    if world > 1:
        t = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_value = world * bases * args.e2e_steps / (e2e_ms * 1e-3)
    d2h = len(variants) * 48 + 4 + 32 * 4 + 8 * 4
    # bytes actually copied per step: the small per-read arrays in full + the payload of admitted reads only
    h2d = h2d - ((batch.n_qual + 1) // 2 + batch.n_qual) + (lvc._handle.h2d_payload_bytes - payload0) // args.e2e_steps

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        if args.kernel_impl == 1:
            tile_ms, tile_n = gen_ms, gen_n
        tile_avg_ms = tile_ms / max(tile_n, 1)
        achieved = alg_bytes_deposit / (tile_avg_ms * 1e-3) / 1e9 if tile_n else None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32+f64", "data": "synthetic",
            "config": {"workload": WORKLOAD if world == 1 else WORKLOAD + f"; one such sample per GPU per step x{world} "
                       "(config 4 style: independent samples, no communication)",
                       "reads_per_step_per_gpu": batch.n_reads, "aligned_bases_per_step_per_gpu": bases,
                       "algorithmic_bytes_per_step_per_gpu": alg_bytes, "ref_len": G, "thresholds": THRESH,
                       "l2": "inputs (0.50 GB/step) larger than L2 (126 MB); no flush needed",
                       "variants_per_step": n_cand},
            "roofline": {"bound": "hbm", "kernel": "k_deposit_tile4" if args.kernel_impl in (0, 4) else
                         ("k_deposit_tile" if args.kernel_impl == 2 else "k_deposit_general"), "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": (achieved / peak) if achieved else None, "traffic": ncu_traffic_bytes(),
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes_deposit,
                         "avg_launch_ms": tile_avg_ms, "launches_timed": tile_n,
                         "timing": "per-kernel CUDA events (library, launching stream) over a second pass of the same "
                                   "K steps; the first pass (ms_per_step) has no events between the kernels",
                         "ms_per_step_with_kernel_events": ms_ev / args.steps,
                         "step_frac": (alg_bytes / (ms / args.steps * 1e-3) / 1e9) / peak,
                         "other_kernels_ms_per_step": {"general_deposit": gen_ms / args.steps,
                                                       "genotype": geno_ms / args.steps}},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_ms / args.e2e_steps, "steps": args.e2e_steps,
                    "api": "LiveVariantCaller.process_batch(pinned SoA) + prepare_variants()"},
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
        }
        if world == 1 and not args.no_cpu_baseline:
            v, reps, dt = cpu_oracle_throughput(ref, batch, 10.0, 12)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
                                    "sample": f"the full workload batch, {reps} repetitions, {dt:.1f} s of CPU work "
                                              "(oracle/oracle.c, single thread like the reference)",
                                    "host_cores": os.cpu_count()}
        print(json.dumps(line), flush=True)
    lvc.close()
    h.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""Throughput of the other SURVEY 8(d) configurations on one GPU (bench.py measures config 2).

    python tools/bench_configs.py --config 3 [--distinct 8] [--batches 100]
    python tools/bench_configs.py --config 5 [--ref-len 1000000] [--steps 10]

config 3: ONT-like live batches (1,000x per batch, 74,758 reads, ~21 CIGAR ops per read, qualities 2..90).  The
          tables persist on the device; after EVERY batch: deposit + genotype pass + candidate compaction.  Reports the
          aggregate aligned bases/s over the batches and the per-batch latency p50 / p99 (CUDA events per batch).
          `--distinct` different batches (seeds 20260200 + k) are generated and cycled (generation costs seconds each).
config 5: shotgun 150 bp reads at 1,000x over a `--ref-len` genome (SURVEY: 5 Mb; the default 1 Mb keeps the same
          per-position depth and per-chunk geometry and generates in ~1.5 min); one step = the whole batch + genotype.
One JSON line per run on stdout.
"""
import argparse, json, os, sys, time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "covid-spings-variant-caller_b200")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402

THRESH = dict(minBQ=30, minMQ=20, minDP=10, minAD=5, ratio=0.10)


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def to_device(torch, capi, batch, dev):
    keep = {}
    for name in ("pos", "flag", "mapq", "keep", "cigar_off", "cigar", "seq_off", "seq4", "qual"):
        keep[name] = torch.from_numpy(getattr(batch, name).view(np.uint8).reshape(-1)).to(dev)
    db = capi.Handle.make_batch(batch.n_reads, batch.n_cigar, batch.n_qual,
                                *[keep[k].data_ptr() for k in ("pos", "flag", "mapq", "keep", "cigar_off", "cigar",
                                                               "seq_off", "seq4", "qual")])
    return db, keep


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, required=True, choices=[3, 5])
    ap.add_argument("--distinct", type=int, default=8)
    ap.add_argument("--batches", type=int, default=100)
    ap.add_argument("--ref-len", type=int, default=1_000_000)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--segments", type=int, default=1, help="config 5: genome = segments x ref-len (5 x 1 Mb = SURVEY's 5 Mb)")
    ap.add_argument("--min-bq", type=int, default=THRESH["minBQ"])
    a = ap.parse_args()
    import torch
    from lvc_b200 import capi, records, synth
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    e_lut, om_lut = records.phred_luts()

    def geno(h, sync=False):
        (h.genotype_device if sync else h.genotype_device_async)(THRESH["minDP"], THRESH["minAD"], THRESH["ratio"], e_lut, om_lut)

    if a.config == 3:
        ref = synth.random_reference(29903, 20260199)
        t0 = time.time()
        batches = [synth.ont_batch_fast(20260200 + k, ref) for k in range(a.distinct)]
        gen_s = time.time() - t0
        h = capi.Handle(ref.encode("latin-1"), a.min_bq, THRESH["minMQ"], device=0, stream=stream.cuda_stream)
        dbs = [to_device(torch, capi, b, dev) for b in batches]
        for db, _ in dbs:                       # synchronous first pass: allocates every quality plane that occurs
            h.push_batch_device(db)
        geno(h, sync=True)
        for k in range(3):
            h.push_batch_device_async(dbs[k % a.distinct][0]); geno(h)
        h.check_async()
        torch.cuda.synchronize()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.batches)]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = h.launch_count
        e0.record(stream)
        for k in range(a.batches):
            evs[k][0].record(stream)
            h.push_batch_device_async(dbs[k % a.distinct][0]); geno(h)
            evs[k][1].record(stream)
        e1.record(stream)
        torch.cuda.synchronize()
        h.check_async()
        ms = e0.elapsed_time(e1)
        lat = np.array([x.elapsed_time(y) for x, y in evs])
        # second pass with the library's per-kernel events: where the batch time goes
        h.set_timing(True)
        for w in range(3):
            h.get_timing(w)
        for k in range(20):
            h.push_batch_device_async(dbs[k % a.distinct][0]); geno(h)
        torch.cuda.synchronize()
        h.check_async()
        h.set_timing(False)
        kt = {name: (lambda t: t[0] / max(t[1], 1))(h.get_timing(w)) for w, name in ((0, "tile_ms"), (1, "general_or_warp_ms"), (2, "genotype_ms"))}
        bases = sum(batches[k % a.distinct].aligned_bases() for k in range(a.batches))
        byts = sum(batches[k % a.distinct].algorithmic_bytes(len(ref)) for k in range(a.batches))
        out = {"config": 3, "workload": f"ONT-like live batches, 1,000x per batch, {batches[0].n_reads} reads/batch, "
               f"{a.batches} batches ({a.distinct} distinct, cycled), minBQ {a.min_bq}", "value": bases / (ms * 1e-3),
               "unit": "aligned bases/s", "ms_total": ms, "batch_ms_p50": float(np.percentile(lat, 50)),
               "batch_ms_p99": float(np.percentile(lat, 99)), "algorithmic_bytes": byts,
               "achieved_GBps": byts / (ms * 1e-3) / 1e9, "roofline_frac": byts / (ms * 1e-3) / 1e9 / peak(),
               "gpu_launches": int(h.launch_count - n0), "planes": len(h.plane_keys()), "generation_s": gen_s,
               "kernel_avg_ms": kt}
    else:
        t0 = time.time()
        if a.segments <= 1:
            ref, batch = synth.shotgun_sample(ref_len=a.ref_len, n_snvs=max(10, a.ref_len // 10000))[:2]
        else:
            # a genome of `segments` x ref_len: independently seeded segments, positions shifted, concatenated (the
            # result is coordinate sorted; depth 1,000x never reaches the admission cap, finalize_batch recomputes it)
            from lvc_b200 import packing
            parts, refs = [], []
            for k in range(a.segments):
                r_k, b_k = synth.shotgun_sample(seed=20260400 + 7 * k, ref_len=a.ref_len,
                                                n_snvs=max(10, a.ref_len // 10000))[:2]
                refs.append(r_k); parts.append(b_k)
            n_tot = sum(b.n_reads for b in parts)
            pos = np.concatenate([b.pos[:b.n_reads].astype(np.int64) + k * a.ref_len for k, b in enumerate(parts)])
            coff = np.zeros(n_tot + 1, dtype=np.uint32); soff = np.zeros(n_tot + 1, dtype=np.uint64)
            i0 = c0 = q0 = 0
            for b in parts:
                n = b.n_reads
                coff[i0 + 1:i0 + n + 1] = b.cigar_off[1:n + 1].astype(np.uint32) + np.uint32(c0)
                soff[i0 + 1:i0 + n + 1] = b.seq_off[1:n + 1] + np.uint64(q0)
                i0 += n; c0 += b.n_cigar; q0 += b.n_qual
            batch = packing.finalize_batch(
                pos.astype(np.int32), np.concatenate([b.flag[:b.n_reads] for b in parts]),
                np.concatenate([b.mapq[:b.n_reads] for b in parts]), coff,
                np.concatenate([b.cigar[:b.n_cigar] for b in parts]), soff,
                np.concatenate([b.seq4[:b.n_qual // 2] for b in parts]),
                np.concatenate([b.qual[:b.n_qual] for b in parts]), THRESH["minMQ"], acgt_only=np.ones(n_tot, dtype=bool))
            ref = "".join(refs)
            del parts
        gen_s = time.time() - t0
        h = capi.Handle(ref.encode("latin-1"), a.min_bq, THRESH["minMQ"], device=0, stream=stream.cuda_stream)
        db, keep = to_device(torch, capi, batch, dev)
        h.push_batch_device(db); geno(h, sync=True)
        for _ in range(2):
            h.push_batch_device_async(db); geno(h)
        h.check_async()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = h.launch_count
        e0.record(stream)
        for _ in range(a.steps):
            h.push_batch_device_async(db); geno(h)
        e1.record(stream)
        torch.cuda.synchronize()
        h.check_async()
        ms = e0.elapsed_time(e1) / a.steps
        byts = batch.algorithmic_bytes(len(ref))
        out = {"config": 5, "workload": f"shotgun 150 bp, 1,000x, G = {len(ref)}, {batch.n_reads} reads", "value":
               batch.aligned_bases() / (ms * 1e-3), "unit": "aligned bases/s", "ms_per_step": ms, "algorithmic_bytes": byts,
               "achieved_GBps": byts / (ms * 1e-3) / 1e9, "roofline_frac": byts / (ms * 1e-3) / 1e9 / peak(),
               "gpu_launches": int(h.launch_count - n0), "steps": a.steps, "generation_s": gen_s}
    print(json.dumps(out))


if __name__ == "__main__":
    main()

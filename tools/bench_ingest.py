"""Throughput of the native BAM ingest (lvc_read_alignments): file on disk -> packed, admitted, page-locked batch.

    python tools/bench_ingest.py [--pairs 100000] [--threads 0] [--reps 3]

Writes a synthetic amplicon BAM (the bench workload's generator, scaled down), reads it back `reps` times and
prints one JSON line.  Also times the pure-Python reader once on a 1/10 sample for scale.
"""
import argparse, json, os, sys, tempfile, time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "covid-spings-variant-caller_b200"))
from lvc_b200 import samio, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=100_000)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    ref, batch = synth.amplicon_sample(n_pairs=a.pairs)[:2]
    with tempfile.TemporaryDirectory() as d:
        bam = os.path.join(d, "s.bam")
        samio.write_bam_batch(bam, ("chrS", len(ref)), batch)
        size = os.path.getsize(bam)
        best = 1e30
        for _ in range(a.reps):
            t0 = time.perf_counter()
            nat = samio.read_alignments_native(bam, None, 20, n_threads=a.threads)
            best = min(best, time.perf_counter() - t0)
            n, pinned = nat.n_reads, nat.pinned
            nat.close()
        small = os.path.join(d, "small.bam")
        sb = batch.slice(0, max(1, batch.n_reads // 10))
        samio.write_bam_batch(small, ("chrS", len(ref)), sb)
        t0 = time.perf_counter()
        samio.read_alignments(small, None, 20)
        t_py = time.perf_counter() - t0
    print(json.dumps({"metric": "bam_ingest_reads_per_sec", "value": n / best, "reads": n, "bam_bytes": size,
                      "seconds": best, "file_MBps": size / best / 1e6, "threads": a.threads or os.cpu_count(),
                      "pinned": pinned, "python_reader_reads_per_sec": sb.n_reads / t_py}))


if __name__ == "__main__":
    main()

"""Throughput of the native BAM ingest (lvc_read_alignments): file on disk -> packed, admitted, page-locked batch.

    python tools/bench_ingest.py [--pairs 100000] [--threads 0] [--reps 3] [--sweep 4,8,16,32] [--level 1]

Writes a synthetic amplicon BAM (the bench workload's generator, scaled down; mate fields filled in, as an aligner
writes them), reads it back `reps` times and prints one JSON line.  Also times the pure-Python reader once on a 1/10
sample for scale.  --sweep: one line per thread count, with the library's own inflate and with zlib's (LVC_INFLATE=zlib).
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
    ap.add_argument("--sweep", default="")
    ap.add_argument("--level", type=int, default=1)
    a = ap.parse_args()
    ref, batch = synth.amplicon_sample(n_pairs=a.pairs)[:2]
    with tempfile.TemporaryDirectory() as d:
        bam = os.path.join(d, "s.bam")
        ids, mpos, tlen = synth.amplicon_pairing(batch, a.pairs, len(ref))
        samio.write_bam_batch(bam, ("chrS", len(ref)), batch, level=a.level, name_id=ids, mate_pos=mpos, tlen=tlen)
        size = os.path.getsize(bam)
        for nt in [int(x) for x in a.sweep.split(",") if x]:
            for mode in ("own", "zlib"):
                if mode == "zlib":
                    os.environ["LVC_INFLATE"] = "zlib"
                else:
                    os.environ.pop("LVC_INFLATE", None)
                ts = []
                for _ in range(a.reps):
                    t0 = time.perf_counter()
                    nat = samio.read_alignments_native(bam, None, 20, n_threads=nt)
                    ts.append(time.perf_counter() - t0)
                    n = nat.n_reads
                    nat.close()
                print(json.dumps({"threads": nt, "inflate": mode, "reads": n, "bam_bytes": size, "best_s": round(min(ts), 4),
                                  "median_s": round(sorted(ts)[len(ts) // 2], 4)}), flush=True)
            os.environ.pop("LVC_INFLATE", None)
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

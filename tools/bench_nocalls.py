"""No-calls (N at quality 2) in a fraction of the reads of an amplicon batch: the deposit step with 4-bit bases (reads with
a no-call lack the A/C/G/T hint and take the one-warp-per-read path inside the tiled kernel) against 2-bit base codes
(no-calls below the threshold are not represented: every read takes the tiled path).  One JSON line per form.

    python tools/bench_nocalls.py [--pairs 400000] [--frac 0.03] [--steps 30]
"""
import argparse, json, os, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import numpy as np  # noqa: E402
import torch  # noqa: E402
from lvc_b200 import capi, records, synth, packing  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=400_000)
    ap.add_argument("--frac", type=float, default=0.03)
    ap.add_argument("--steps", type=int, default=30)
    a = ap.parse_args()
    ref, b = synth.amplicon_sample(seed=77, n_pairs=a.pairs)[:2]
    rng = np.random.default_rng(5)
    n = b.n_reads
    pick = np.nonzero(rng.random(n) < a.frac)[0]
    lq = packing.query_lengths(b.cigar_off, b.cigar)
    seq4, qual, keep = b.seq4.copy(), b.qual.copy(), b.keep.copy()
    for i in pick:                                           # one no-call per picked read
        x = int(b.seq_off[i]) + int(rng.integers(0, max(1, int(lq[i]))))
        byte = int(seq4[x >> 1])
        seq4[x >> 1] = (byte & 0xF0) | 15 if (x & 1) else (byte & 0x0F) | 0xF0
        qual[x] = 2
        keep[i] &= 0xFD                                      # the read is no longer A/C/G/T only
    nb = packing.ReadBatch(b.pos, b.flag, b.mapq, keep, b.cigar_off, b.cigar, b.seq_off, seq4, qual).admitted_only()
    coded = nb.with_quality_codes()
    forms = {"4-bit bases": coded, "2-bit base codes": coded.with_base_codes(bench.THRESH["minBQ"])}
    assert forms["2-bit base codes"].scode is not None
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    T = bench.THRESH
    e_lut, om_lut = records.phred_luts()
    res = {}
    for name, fb in forms.items():
        fields = {k: torch.from_numpy(getattr(fb, k).view(np.uint8).reshape(-1)).to(dev) for k in bench.BATCH_FIELDS if k not in ("qual", "seq4")}
        fields["qual"] = torch.from_numpy(fb.qcode.view(np.uint8).reshape(-1)).to(dev)
        fields["seq4"] = torch.from_numpy((fb.scode if fb.scode is not None else fb.seq4).view(np.uint8).reshape(-1)).to(dev)
        db = capi.Handle.make_batch(fb.n_reads, fb.n_cigar, fb.n_qual, *[fields[k].data_ptr() for k in bench.BATCH_FIELDS],
                                    qual_dict=fb.qdict, base_codes_min_bq=fb.scode_min_bq if fb.scode is not None else None)
        h = capi.Handle(ref.encode("latin-1"), T["minBQ"], T["minMQ"], device=0, stream=stream.cuda_stream)

        def step():
            h.push_batch_device_async(db)
            h.genotype_device_async(T["minDP"], T["minAD"], T["ratio"], e_lut, om_lut)
        h.push_batch_device(db)
        h.genotype_device(T["minDP"], T["minAD"], T["ratio"], e_lut, om_lut)
        cands = sorted((int(c["pos"]), int(c["code"]), int(c["ad"]), int(c["dp"])) for c in h.fetch_candidates())
        for _ in range(5):
            step()
        h.check_async()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(a.steps):
            step()
        e1.record(stream)
        torch.cuda.synchronize()
        h.check_async()
        res[name] = cands
        print(json.dumps({"form": name, "reads": int(fb.n_reads), "reads_with_a_no_call": int(len(pick)),
                          "step_us": round(e0.elapsed_time(e1) / a.steps * 1e3, 2), "candidates": len(cands)}), flush=True)
        h.close()
    assert res["4-bit bases"] == res["2-bit base codes"], "the two forms must give the same candidates"


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Generate golden vectors for tests/golden/ by running the REAL reference code.

Run HERE (the authoring container), where /root/reference exists:

    python oracle/gen_golden.py

pysam is not installable in this image, so a *stub* ``pysam`` module is injected before importing
``/root/reference/variant_caller/live_variant_caller.py``.  The stub provides only what that file
touches (``FastaFile``, ``AlignmentFile(...).pileup(...)``, ``AlignedSegment``); the pileup columns it
yields come from ``oracle/pileup_oracle.py``'s htslib ``bam_plp`` emulation.  Everything downstream of
the columns -- ``process_pileup_column``, ``process_svn``, ``prepare_variants``, and
``variant_caller/utils.py`` -- is the unmodified reference code.  So the vectors pin the state-update
and genotype/record half of the path to the real reference; the CIGAR-walk half stays pinned only to
the htslib restatement (see DESIGN.md "parity status").

The GPU box has no /root/reference: tests read only the committed JSON under tests/golden/.
"""
import json
import os
import random
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("LVC_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)

from oracle import pileup_oracle as po  # noqa: E402
from oracle import synth_small  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


# ------------------------------------------------------------------ stub pysam
class _FastaFile:
    def __init__(self, path):
        self.path = path
        self._seqs = {}
        name = None
        with open(path) as fh:
            for line in fh:
                line = line.rstrip("\n")
                if line.startswith(">"):
                    name = line[1:].split()[0]
                    self._seqs[name] = []
                elif name is not None:
                    self._seqs[name].append(line)
        self._seqs = {k: "".join(v) for k, v in self._seqs.items()}
        self.references = list(self._seqs)

    def fetch(self, reference=None):
        return self._seqs[reference]

    def get_reference_length(self, reference):
        return len(self._seqs[reference])

    def close(self):
        pass


class _Aln:
    def __init__(self, read):
        self.query_sequence = read.seq
        self.query_qualities = read.qual


class _PileupRead:
    def __init__(self, read, is_del, is_refskip, qpos):
        self.alignment = _Aln(read)
        self.is_del = int(is_del)
        self.is_refskip = int(is_refskip)
        self.query_position = None if is_del else qpos


class _Column:
    def __init__(self, name, pos, pileups):
        self.reference_name = name
        self.reference_pos = pos
        self.pileups = pileups


_READSETS = {}      # "path" -> list[Read]; lets scenarios hand reads to the stub without BAM files


class _AlignmentFile:
    def __init__(self, path, mode="rb"):
        self.path = path
        if path in _READSETS:
            self.reads = _READSETS[path]
            self._len = None
        else:
            contigs, self.reads = po.read_sam(path)
            self._contigs = dict(contigs)

    def pileup(self, min_mapping_quality=0, min_base_quality=13, reference=None, max_depth=po.MAX_DEPTH):
        for pos, entries in po.pileup_columns(self.reads, min_mapping_quality, max_depth):
            pile = []
            for idx, r, is_del, is_skip, qpos in entries:
                q = r.qual[qpos] if qpos < len(r.qual) else 0     # pysam pileup_base_qual_skip
                if q < min_base_quality:
                    continue
                pile.append(_PileupRead(r, is_del, is_skip, qpos))
            yield _Column(reference, pos, pile)

    def get_reference_length(self, reference):
        return 0

    def close(self):
        pass


def _install_stub():
    m = types.ModuleType("pysam")
    m.FastaFile = _FastaFile
    m.AlignmentFile = _AlignmentFile
    m.AlignedSegment = object
    sys.modules["pysam"] = m
    sys.path.insert(0, REF)


# ------------------------------------------------------------------ serialisation helpers
def _variant_json(v):
    return {"start": v["start"], "stop": v["stop"], "alleles": list(v["alleles"]),
            "qual_hex": float(v["qual"]).hex(),
            "DP": v["info"]["DP"], "AD": v["info"]["AD"],
            "GL": v["info"]["GL"] if isinstance(v["info"]["GL"], int) else {"hex": float(v["info"]["GL"]).hex()},
            "PL": v["info"]["PL"], "SCORE": v["info"]["SCORE"]}


def _memory_json(mem):
    return {str(p): {"reference": s["reference"], "totalDepth": s["totalDepth"],
                     "snvs": {b: list(map(int, q)) for b, q in s["snvs"].items()}}
            for p, s in mem.items()}


def _run_reference(LVC, u, fasta, reads_key, th):
    lvc = LVC(fasta, th["minBQ"], th["minMQ"], th["minDP"], th["minAD"], th["ratio"], 1)
    lvc.process_bam(reads_key)
    variants = lvc.prepare_variants()
    # per-allele likelihoods from the reference's own utils (ungated)
    lik = {}
    for p, site in lvc.memory.items():
        snvs = {a: [u.from_phred_scale(q) for q in site["snvs"][a]] for a in site["snvs"]}
        lik[str(p)] = {a: float(u.genotype_likelihood(a, snvs)).hex() for a in snvs}
    return lvc, variants, lik


def main():
    _install_stub()
    import variant_caller.utils as u                      # REAL reference module
    from variant_caller.live_variant_caller import LiveVariantCaller as LVC   # REAL reference class
    os.makedirs(GOLD, exist_ok=True)

    # ---- utils known answers
    rng = random.Random(20260101)
    ka = {"from_phred_hex": [u.from_phred_scale(q).hex() for q in range(256)],
          "to_phred": [], "genotype_likelihood": []}
    for p in [0.0, 1.0, 0.001, 1e-10, 1.2e-10, 0.5, 0.31622776601683794, 3.2e-5, 1e-300, 0.999999]:
        ka["to_phred"].append([p.hex(), u.to_phred_scale(p)])
    for _ in range(200):
        p = 10 ** rng.uniform(-12, 0)
        ka["to_phred"].append([p.hex(), u.to_phred_scale(p)])
    for _ in range(60):
        alle = {}
        for b in rng.sample("ACGTN", rng.randint(1, 4)):
            alle[b] = [u.from_phred_scale(rng.randint(0, 60)) for _ in range(rng.randint(1, 120))]
        for h in alle:
            ka["genotype_likelihood"].append({"h": h, "alleles": {b: [x.hex() for x in v] for b, v in alle.items()},
                                              "L": float(u.genotype_likelihood(h, alle)).hex()})
    with open(os.path.join(GOLD, "utils_known_answers.json"), "w") as fh:
        json.dump(ka, fh)

    # ---- scenario 1: the reference's own fixture, three threshold sets
    sam = os.path.join(GOLD, "testfile.sam")
    fasta = os.path.join(GOLD, "NC_045512.2.synthetic.fasta")
    if not os.path.exists(sam):
        import shutil
        shutil.copyfile(os.path.join(REF, "test", "testdata", "testfile.sam"), sam)   # DATA fixture only
    if not os.path.exists(fasta):
        synth_small.write_testfile_fasta(sam, fasta)
    scen = {}
    thresholds = {
        "vc_config": dict(minBQ=30, minMQ=20, minDP=10, minAD=5, ratio=0.10),
        "bq13": dict(minBQ=13, minMQ=0, minDP=1, minAD=1, ratio=0.0),
        "all_zero": dict(minBQ=0, minMQ=0, minDP=1, minAD=1, ratio=0.0),
        "bq13_dp3": dict(minBQ=13, minMQ=0, minDP=3, minAD=2, ratio=0.25),
    }
    for name, th in thresholds.items():
        lvc, variants, lik = _run_reference(LVC, u, fasta, sam, th)
        scen[name] = {"thresholds": th, "memory": _memory_json(lvc.memory),
                      "variants": [_variant_json(v) for v in variants], "likelihoods": lik}
    with open(os.path.join(GOLD, "testfile_golden.json"), "w") as fh:
        json.dump(scen, fh)

    # ---- scenario 2: seeded random small read sets (reads are re-generated from the seed in tests)
    out = {}
    for name, kw in synth_small.SCENARIOS.items():
        ref, reads = synth_small.make_scenario(**kw)
        fa = os.path.join("/tmp", f"lvc_golden_{name}.fasta")
        with open(fa, "w") as fh:
            fh.write(">chrS\n" + ref + "\n")
        key = f"mem://{name}"
        _READSETS[key] = reads
        res = {}
        for tname, th in synth_small.THRESHOLDS.items():
            lvc, variants, lik = _run_reference(LVC, u, fa, key, th)
            entry = {"thresholds": th, "variants": [_variant_json(v) for v in variants], "likelihoods": lik}
            if kw.get("store_memory", True):
                entry["memory"] = _memory_json(lvc.memory)
            else:   # big scenario: order-free summary only
                entry["summary"] = {str(p): {"totalDepth": s["totalDepth"],
                                             "order": list(s["snvs"].keys()),
                                             "counts": {b: len(q) for b, q in s["snvs"].items()},
                                             "qsum": {b: int(sum(q)) for b, q in s["snvs"].items()}}
                                    for p, s in lvc.memory.items()}
            res[tname] = entry
        out[name] = {"ref": ref, "reads": synth_small.reads_to_rows(reads), "results": res}
    with open(os.path.join(GOLD, "synthetic_golden.json"), "w") as fh:
        json.dump(out, fh)
    print("golden vectors written to", GOLD)


if __name__ == "__main__":
    main()

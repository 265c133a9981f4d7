"""CPU: the oracle restatement against the golden vectors produced by the REAL reference code."""
import json
import os

import pytest

from helpers import GOLD, po, synth_small, variants_from_golden, assert_variants_equal


def test_utils_known_answers(golden_utils):
    for q, hx in enumerate(golden_utils["from_phred_hex"]):
        assert po.from_phred_scale(q) == float.fromhex(hx)
    for hx, want in golden_utils["to_phred"]:
        assert po.to_phred_scale(float.fromhex(hx)) == want
    for case in golden_utils["genotype_likelihood"]:
        alleles = {b: [float.fromhex(x) for x in v] for b, v in case["alleles"].items()}
        assert float(po.genotype_likelihood(case["h"], alleles)) == float.fromhex(case["L"])


def test_survey_known_answers():
    # SURVEY 8c probes of the real utils.py
    assert po.from_phred_scale(30) == 0.001
    assert po.to_phred_scale(0.001) == 30 and po.to_phred_scale(0.0) == 99 and po.to_phred_scale(1.0) == 0
    assert float(po.genotype_likelihood('A', {'A': [1e-3, 1e-4], 'C': [1e-3]})) == 0.0009989001
    e = 10 ** -3.5
    L = {a: float(po.genotype_likelihood(a, {'A': [e] * 190, 'G': [e] * 10})) for a in 'AG'}
    assert L['A'] == 9.416771630345698e-36 and L['G'] == 0.0


@pytest.mark.parametrize("name", ["vc_config", "bq13", "all_zero", "bq13_dp3"])
def test_testfile_sam(golden_testfile, name):
    g = golden_testfile[name]
    th = g["thresholds"]
    contigs, reads = po.read_sam(os.path.join(GOLD, "testfile.sam"))
    ref = open(os.path.join(GOLD, "NC_045512.2.synthetic.fasta")).read().split("\n", 1)[1].replace("\n", "")
    oc = po.OracleCaller(ref, th["minBQ"], th["minMQ"], th["minDP"], th["minAD"], th["ratio"])
    oc.process_reads(reads)
    want_mem = {int(p): s for p, s in g["memory"].items()}
    assert list(oc.memory.keys()) == list(want_mem.keys())
    for p, s in want_mem.items():
        assert oc.memory[p]["reference"] == s["reference"]
        assert oc.memory[p]["totalDepth"] == s["totalDepth"]
        assert {b: list(q) for b, q in oc.memory[p]["snvs"].items()} == s["snvs"]
        assert list(oc.memory[p]["snvs"].keys()) == list(s["snvs"].keys())
    assert_variants_equal(oc.prepare_variants(), variants_from_golden(g["variants"]), name)
    lik = oc.likelihoods()
    for p, d in g["likelihoods"].items():
        for a, hx in d.items():
            assert lik[int(p)][a] == float.fromhex(hx)


def test_testfile_appendix_c(golden_testfile):
    # SURVEY Appendix C summary numbers
    expect = {"vc_config": (421, 202, 202), "bq13": (421, 1257, 1245), "all_zero": (421, 1642, 1609)}
    for name, (ncol, tot, dep) in expect.items():
        mem = golden_testfile[name]["memory"]
        assert len(mem) == ncol
        assert sum(s["totalDepth"] for s in mem.values()) == tot
        assert sum(len(q) for s in mem.values() for q in s["snvs"].values()) == dep
    assert golden_testfile["vc_config"]["variants"] == []


@pytest.mark.parametrize("scen", ["mixed_small", "ont_like", "deep_underflow", "amplicon_like", "maxdepth"])
def test_synthetic(golden_synth, scen):
    g = golden_synth[scen]
    reads = synth_small.rows_to_reads(g["reads"])
    for tname, res in g["results"].items():
        th = res["thresholds"]
        oc = po.OracleCaller(g["ref"], th["minBQ"], th["minMQ"], th["minDP"], th["minAD"], th["ratio"])
        oc.process_reads(reads)
        if "memory" in res:
            want = {int(p): s for p, s in res["memory"].items()}
            assert list(oc.memory.keys()) == list(want.keys())
            for p, s in want.items():
                assert oc.memory[p]["totalDepth"] == s["totalDepth"]
                assert {b: list(q) for b, q in oc.memory[p]["snvs"].items()} == s["snvs"]
        else:
            for p, s in res["summary"].items():
                site = oc.memory[int(p)]
                assert site["totalDepth"] == s["totalDepth"]
                assert list(site["snvs"].keys()) == s["order"]
                assert {b: len(q) for b, q in site["snvs"].items()} == s["counts"]
        assert_variants_equal(oc.prepare_variants(), variants_from_golden(res["variants"]), f"{scen}/{tname}")


def test_admission_mask_matches_literal_engine():
    """O(N) admission simulation == the literal bam_plp emulation, including tiny max_depth values."""
    import random
    rng = random.Random(5)
    for trial in range(60):
        n = rng.randint(1, 120)
        reads = []
        for i in range(n):
            pos = rng.choice([0, 0, 3, 3, 3, 7, 20, 21, 22, 40]) if trial % 2 else rng.randint(0, 60)
            ln = rng.randint(1, 15)
            flag = rng.choice([0, 16, 0, 0, 0x400, 0x1, 0x3])
            reads.append(po.Read(flag, pos, rng.choice([60, 60, 5]), [(0, ln)], "A" * ln, [30] * ln))
        reads = po.samtools_sort(reads)
        for md in (1, 2, 3, 5, 8000):
            admitted = []
            for _ in po.pileup_columns(reads, 10, md, admitted):
                pass
            mask = po.admission_mask([r.pos for r in reads], [r.end() for r in reads],
                                     [po.passes_read_filter(r, 10) for r in reads], md)
            assert [i for i, k in enumerate(mask) if k] == admitted, (trial, md)

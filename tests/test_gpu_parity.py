"""GPU parity tests (through the C-ABI): tables bit-exact, likelihoods within 1e-9, records identical."""
import json
import os
import pickle

import numpy as np
import pytest

from helpers import (GOLD, po, synth_small, rows_to_tuples, variants_from_golden, assert_variants_equal,
                     memory_tables, close_lik)

pytestmark = pytest.mark.gpu


def _lvc(fasta, th, device=0, impl=0):
    from variant_caller.live_variant_caller import LiveVariantCaller
    c = LiveVariantCaller(fasta, th["minBQ"], th["minMQ"], th["minDP"], th["minAD"], th["ratio"], 1, device=device)
    c._handle.set_impl(impl)
    return c


def _fasta(tmp_path, name, ref):
    p = str(tmp_path / f"{name}.fasta")
    with open(p, "w") as fh:
        fh.write(f">{name}\n{ref}\n")
    return p


def _check_against_golden_memory(lvc, want_mem, what):
    got_d, got_h, got_o = memory_tables(lvc.memory)
    want_d, want_h, want_o = memory_tables(want_mem)
    assert got_d == want_d, f"{what}: totalDepth / site set differs"
    assert got_h == want_h, f"{what}: (pos, allele, quality) histogram differs"
    assert got_o == want_o, f"{what}: first-seen allele order differs"


def _check_likelihoods(lvc, want_lik, what, depth_of):
    got = lvc.likelihoods()
    n = 0
    for p, d in want_lik.items():
        p = int(p)
        for a, hx in d.items():
            w = float.fromhex(hx) if isinstance(hx, str) else hx
            g = got.get(p, {}).get(a)
            assert g is not None, (what, p, a)
            assert close_lik(g, w, n_factors=depth_of(p) + 4), (what, p, a, g, w)
            n += 1
    return n


@pytest.mark.parametrize("impl", [1, 2, 3, 4, 5, 6])
@pytest.mark.parametrize("name", ["vc_config", "bq13", "all_zero", "bq13_dp3"])
def test_reference_fixture(lib, golden_testfile, name, impl):
    """BASELINE configs[0]: test/testdata/testfile.sam through process_bam (SAM text in, like the server)."""
    g = golden_testfile[name]
    th = g["thresholds"]
    lvc = _lvc(os.path.join(GOLD, "NC_045512.2.synthetic.fasta"), th, impl=impl)
    lvc.process_bam(os.path.join(GOLD, "testfile.sam"))
    _check_against_golden_memory(lvc, g["memory"], name)
    assert_variants_equal(lvc.prepare_variants(), variants_from_golden(g["variants"]), name)
    # ungated likelihoods: gate is minDP, so compare where the site passes it
    lvc.minTotalDepth = 0
    mem = g["memory"]
    _check_likelihoods(lvc, g["likelihoods"], name, lambda p: mem[str(p)]["totalDepth"])
    lvc.close()


@pytest.mark.parametrize("impl", [1, 2, 3, 4, 5, 6])
@pytest.mark.parametrize("scen", ["mixed_small", "ont_like", "deep_underflow", "amplicon_like", "maxdepth"])
def test_synthetic_scenarios(lib, golden_synth, tmp_path, scen, impl):
    from lvc_b200 import packing
    g = golden_synth[scen]
    fa = _fasta(tmp_path, "chrS", g["ref"])
    for tname, res in g["results"].items():
        th = res["thresholds"]
        lvc = _lvc(fa, th, impl=impl)
        lvc.process_batch(packing.pack_reads(rows_to_tuples(g["reads"]), th["minMQ"]))
        if "memory" in res:
            _check_against_golden_memory(lvc, res["memory"], f"{scen}/{tname}")
            depth_of = lambda p: res["memory"][str(p)]["totalDepth"]
        else:
            mem = lvc.memory
            for p, s in res["summary"].items():
                site = mem[int(p)]
                assert site["totalDepth"] == s["totalDepth"]
                assert list(site["snvs"].keys()) == s["order"]
                assert {b: len(q) for b, q in site["snvs"].items()} == s["counts"]
                assert {b: int(sum(q)) for b, q in site["snvs"].items()} == s["qsum"]
            depth_of = lambda p: res["summary"][str(p)]["totalDepth"]
        assert_variants_equal(lvc.prepare_variants(), variants_from_golden(res["variants"]), f"{scen}/{tname}")
        lvc.minTotalDepth = 0
        _check_likelihoods(lvc, res["likelihoods"], f"{scen}/{tname}", depth_of)
        lvc.close()


@pytest.mark.parametrize("impl", [1, 2, 3, 4, 5, 6])
def test_live_batches_accumulate(lib, golden_synth, tmp_path, impl):
    """incremental per-batch update: N process calls == the oracle fed the same batches (max_depth per call)."""
    from lvc_b200 import packing
    g = golden_synth["amplicon_like"]
    reads = synth_small.rows_to_reads(g["reads"])
    rows = g["reads"]
    th = dict(minBQ=13, minMQ=0, minDP=3, minAD=2, ratio=0.05)
    fa = _fasta(tmp_path, "chrS", g["ref"])
    lvc = _lvc(fa, th, impl=impl)
    oc = po.OracleCaller(g["ref"], th["minBQ"], th["minMQ"], th["minDP"], th["minAD"], th["ratio"])
    for k in range(5):
        sel = list(range(k, len(rows), 5))            # interleaved, still coordinate sorted
        lvc.process_batch(packing.pack_reads(rows_to_tuples([rows[i] for i in sel]), th["minMQ"]))
        oc.process_reads([reads[i] for i in sel])
        assert_variants_equal(lvc.prepare_variants(), oc.prepare_variants(), f"batch {k}")
    _check_against_golden_memory(lvc, oc.memory, "live")
    lvc.close()


def test_bam_input_and_vcf_text(lib, golden_testfile, tmp_path):
    from lvc_b200 import samio
    g = golden_testfile["bq13"]
    th = g["thresholds"]
    contigs, reads = po.read_sam(os.path.join(GOLD, "testfile.sam"))
    bam = str(tmp_path / "t.bam")
    samio.write_bam(bam, contigs, [(r.flag, r.pos, r.mapq, r.cigar, r.seq, r.qual, r.name) for r in reads])
    fasta = os.path.join(GOLD, "NC_045512.2.synthetic.fasta")
    lvc = _lvc(fasta, th)
    lvc.process_bam(bam)
    out = str(tmp_path / "o.vcf")
    lvc.write_vcf(out)
    ref = open(fasta).read().split("\n", 1)[1].replace("\n", "")
    oc = po.OracleCaller(ref, th["minBQ"], th["minMQ"], th["minDP"], th["minAD"], th["ratio"])
    oc.process_reads(reads)
    assert open(out).read() == oc.vcf_text(contigs)
    lvc.write_csv(str(tmp_path / "o.csv"))
    assert open(str(tmp_path / "o.csv")).readline().startswith("POS,REF,DEPTH")
    lvc.close()


def test_checkpoint_roundtrip_and_reference_schema(lib, golden_synth, tmp_path):
    from lvc_b200 import packing
    g = golden_synth["mixed_small"]
    res = g["results"]["loose"]
    th = res["thresholds"]
    fa = _fasta(tmp_path, "chrS", g["ref"])
    # (1) a pickle in the REFERENCE's schema (read-order lists) loads and yields the golden records
    ref_mem = {int(p): {"reference": s["reference"], "totalDepth": s["totalDepth"], "snvs": s["snvs"], "indels": {}}
               for p, s in res["memory"].items()}
    ck = str(tmp_path / "ref.pkl")
    with open(ck, "wb") as fh:
        pickle.dump(ref_mem, fh)
    lvc = _lvc(fa, th)
    lvc.load_checkpoint(ck)
    assert_variants_equal(lvc.prepare_variants(), variants_from_golden(res["variants"]), "ref-schema checkpoint")
    _check_against_golden_memory(lvc, res["memory"], "ref-schema checkpoint")
    # (2) our checkpoint round-trips, and processing continues on top of it
    ck2 = str(tmp_path / "ours.pkl")
    lvc.create_checkpoint(ck2)
    lvc2 = _lvc(fa, th)
    lvc2.load_checkpoint(ck2)
    assert_variants_equal(lvc2.prepare_variants(), lvc.prepare_variants(), "roundtrip")
    batch = packing.pack_reads(rows_to_tuples(g["reads"]), th["minMQ"])
    lvc.process_batch(batch)
    lvc2.process_batch(batch)
    assert_variants_equal(lvc2.prepare_variants(), lvc.prepare_variants(), "continue after load")
    d1, h1, o1 = memory_tables(lvc.memory)
    d2, h2, o2 = memory_tables(lvc2.memory)
    assert d1 == d2 and h1 == h2 and o1 == o2
    lvc.reset_memory()
    assert lvc.memory == {} and lvc.prepare_variants() == []
    lvc.close(); lvc2.close()


def test_empty_and_filtered_batches(lib, tmp_path):
    from lvc_b200 import packing
    fa = _fasta(tmp_path, "c", "ACGT" * 50)
    th = dict(minBQ=30, minMQ=20, minDP=1, minAD=1, ratio=0.0)
    lvc = _lvc(fa, th)
    lvc.process_batch(packing.pack_reads([], 20))
    assert lvc.prepare_variants() == [] and lvc.memory == {}
    # every read filtered (mapq / flags / orphan) or below minBQ
    rd = [(0x400, 0, 60, [(0, 10)], "A" * 10, [40] * 10), (0, 2, 5, [(0, 10)], "A" * 10, [40] * 10),
          (0x1, 3, 60, [(0, 10)], "A" * 10, [40] * 10), (0, 5, 60, [(0, 10)], "C" * 10, [10] * 10)]
    lvc.process_batch(packing.pack_reads(rd, 20))
    mem = lvc.memory
    assert sorted(mem) == list(range(5, 15)) and all(s["totalDepth"] == 0 and s["snvs"] == {} for s in mem.values())
    assert lvc.prepare_variants() == []
    # a read running off the reference end is an error, not silent corruption
    from lvc_b200 import capi
    with pytest.raises(capi.LvcError):
        lvc.process_batch(packing.pack_reads([(0, 195, 60, [(0, 10)], "A" * 10, [40] * 10)], 20))
    lvc.close()

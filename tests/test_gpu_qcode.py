"""Quality-code batches (lvc_batch::qual_bits == 2) on the GPU, through the C-ABI: the tables, likelihoods and records
must be those of the oracle, and identical to what the byte form of the same batch gives."""
import os

import numpy as np
import pytest

from helpers import (po, synth_small, rows_to_tuples, variants_from_golden, assert_variants_equal, memory_tables)
from test_gpu_parity import _lvc, _fasta, _check_against_golden_memory, _check_likelihoods

pytestmark = pytest.mark.gpu

QBINS = (2, 12, 23, 37)
THS = dict(strict=dict(minBQ=30, minMQ=20, minDP=10, minAD=5, ratio=0.10),      # only Q37 passes: no cold code
           loose=dict(minBQ=13, minMQ=0, minDP=3, minAD=2, ratio=0.05),        # Q23 and Q37 pass: one cold code
           zero=dict(minBQ=0, minMQ=0, minDP=1, minAD=1, ratio=0.0))           # everything passes (GE_ALL kernels)


def _tuples(reads):
    return [(r.flag, r.pos, r.mapq, r.cigar, r.seq, r.qual) for r in reads]


def _coded(reads, mq, max_depth=None):
    from lvc_b200 import packing
    b = packing.pack_reads(_tuples(reads), mq) if max_depth is None else packing.pack_reads(_tuples(reads), mq, max_depth)
    c = b.with_quality_codes()
    assert c.qcode is not None, "the scenario must qualify for the code form"
    return b, c


@pytest.mark.parametrize("impl", [0, 1, 5])
def test_golden_amplicon_scenario_as_codes(lib, golden_synth, tmp_path, impl):
    """the golden vectors of the REAL reference (tests/golden) through a quality-code batch"""
    g = golden_synth["amplicon_like"]
    reads = synth_small.rows_to_reads(g["reads"])
    fa = _fasta(tmp_path, "chrS", g["ref"])
    for tname, res in g["results"].items():
        th = res["thresholds"]
        lvc = _lvc(fa, th, impl=impl)
        lvc.process_batch(_coded(reads, th["minMQ"])[1])
        _check_against_golden_memory(lvc, res["memory"], f"codes/{tname}")
        assert_variants_equal(lvc.prepare_variants(), variants_from_golden(res["variants"]), f"codes/{tname}")
        lvc.minTotalDepth = 0
        _check_likelihoods(lvc, res["likelihoods"], f"codes/{tname}", lambda p: res["memory"][str(p)]["totalDepth"])
        lvc.close()


@pytest.mark.parametrize("impl", [0, 1, 5])
@pytest.mark.parametrize("tname", ["strict", "loose", "zero"])
@pytest.mark.parametrize("scen", ["weird_short", "deep_amplicon", "indel_dense", "two_values"])
def test_random_scenarios_codes_vs_oracle_and_bytes(lib, tmp_path, scen, tname, impl):
    """every CIGAR op, non-ACGT base codes, odd lengths, soft clips, reads of many ops (handed to the any-record path
    inside the tiled kernel), a dictionary of fewer than four values"""
    cfg = dict(
        weird_short=dict(seed=101, ref_len=500, n_reads=400, len_lo=21, len_hi=151, q_lo=0, q_hi=0, indel_rate=0.06, weird=True, qbins=QBINS),
        deep_amplicon=dict(seed=102, ref_len=900, n_reads=3000, len_lo=149, len_hi=150, q_lo=0, q_hi=0, indel_rate=0.01, weird=False,
                           amplicon=(0, 250, 251, 600, 750), qbins=QBINS, planted=((10, 0.5), (300, 0.2), (620, 1.0))),
        indel_dense=dict(seed=103, ref_len=700, n_reads=500, len_lo=80, len_hi=250, q_lo=0, q_hi=0, indel_rate=0.12, weird=False, qbins=QBINS),
        two_values=dict(seed=104, ref_len=300, n_reads=300, len_lo=30, len_hi=100, q_lo=0, q_hi=0, indel_rate=0.02, weird=False, qbins=(37, 37, 11, 37)),
    )[scen]
    ref, reads = synth_small.make_scenario(**cfg)
    th = THS[tname]
    fa = _fasta(tmp_path, "chrS", ref)
    raw, coded = _coded(reads, th["minMQ"])
    assert len(set(coded.qdict)) >= 2
    oc = po.OracleCaller(ref, th["minBQ"], th["minMQ"], th["minDP"], th["minAD"], th["ratio"])
    oc.process_reads(reads)
    lvc = _lvc(fa, th, impl=impl)
    lvc.process_batch(coded)
    _check_against_golden_memory(lvc, oc.memory, f"{scen}/{tname} codes vs oracle")
    got, want = lvc.prepare_variants(), oc.prepare_variants()
    for g, w in zip(got, want):
        # a likelihood inside the denormal band (GL below -307.6) is order dependent in the reference itself (SURVEY A6,
        # tests/helpers.close_lik): there the records are held to the byte-form run below, not to 1e-9 of the oracle
        if isinstance(w["info"]["GL"], float) and w["info"]["GL"] < -307.6 and abs(g["info"]["GL"] - w["info"]["GL"]) < 1e-3:
            g["info"]["GL"] = w["info"]["GL"]
    assert_variants_equal(got, want, f"{scen}/{tname}")
    ref_run = _lvc(fa, th, impl=impl)
    ref_run.process_batch(raw)
    assert_variants_equal(lvc.prepare_variants(), ref_run.prepare_variants(), f"{scen}/{tname} codes vs bytes")
    h, hr = lvc._handle, ref_run._handle
    # (planes are pre-allocated from a sample of the qualities: one run may hold an EMPTY plane the other never made)
    ka, kb = set(h.plane_keys().tolist()), set(hr.plane_keys().tolist())
    for k in sorted(ka | kb):
        pa = h.copy_plane(k) if k in ka else None
        pb = hr.copy_plane(k) if k in kb else None
        if pa is None or pb is None:
            assert not (pa if pb is None else pb).any(), f"plane {k} exists on one side only and is not empty"
        else:
            assert np.array_equal(pa, pb), f"plane {k}"
    assert np.array_equal(h.copy_dels(), hr.copy_dels()) and np.array_equal(h.copy_covdiff(), hr.copy_covdiff())
    for grp in range(4):
        a, b = h.copy_first(grp), hr.copy_first(grp)
        assert (a is None) == (b is None) and (a is None or np.array_equal(a, b)), f"first-seen group {grp}"
    lvc.close(); ref_run.close()


def test_live_batches_switch_forms_and_dictionaries(lib, tmp_path):
    """incremental batches: codes, then bytes, then codes with another dictionary (new planes through the replay of the
    any-record kernel on a code batch); the oracle is fed the same batches"""
    th = THS["loose"]
    ref, reads_a = synth_small.make_scenario(seed=201, ref_len=600, n_reads=500, len_lo=100, len_hi=150, q_lo=0, q_hi=0,
                                             indel_rate=0.03, weird=False, qbins=QBINS)
    _, reads_b = synth_small.make_scenario(seed=201, ref_len=600, n_reads=300, len_lo=100, len_hi=150, q_lo=0, q_hi=0,
                                           indel_rate=0.03, weird=False, qbins=(14, 22, 33, 40))
    fa = _fasta(tmp_path, "chrS", ref)
    lvc = _lvc(fa, th)
    oc = po.OracleCaller(ref, th["minBQ"], th["minMQ"], th["minDP"], th["minAD"], th["ratio"])
    for k, (rd, form) in enumerate([(reads_a[0::2], "codes"), (reads_a[1::2], "bytes"), (reads_b, "codes"), (reads_a[0::3], "codes")]):
        raw, coded = _coded(rd, th["minMQ"])
        lvc.process_batch(coded if form == "codes" else raw)
        oc.process_reads(rd)
        assert_variants_equal(lvc.prepare_variants(), oc.prepare_variants(), f"batch {k}")
    _check_against_golden_memory(lvc, oc.memory, "live code batches")
    lvc.close()


def test_depth_cap_and_dropped_reads_with_other_qualities(lib, tmp_path):
    """reads beyond htslib's max_depth are dropped at pack time and never read: their qualities do not count towards the
    dictionary (they may take any value) and their codes are never looked at"""
    from lvc_b200 import packing
    th = THS["strict"]
    ref, reads = synth_small.make_scenario(seed=301, ref_len=120, n_reads=700, len_lo=100, len_hi=100, q_lo=0, q_hi=0,
                                           indel_rate=0.0, weird=False, fixed_pos=3, qbins=QBINS)
    b = packing.pack_reads(_tuples(reads), th["minMQ"], 500)
    dropped = np.nonzero((b.keep & 1) == 0)[0]
    assert len(dropped) > 100
    for i in dropped[:50]:
        reads[i].qual = [int(q) for q in np.random.default_rng(int(i)).integers(40, 90, len(reads[i].qual))]
    raw, coded = _coded(reads, th["minMQ"], 500)
    assert sorted(set(coded.qdict)) == sorted(set(QBINS)) and len(set(raw.qual[:raw.n_qual].tolist())) > 4
    oc = po.OracleCaller(ref, th["minBQ"], th["minMQ"], th["minDP"], th["minAD"], th["ratio"], max_depth=500)
    oc.process_reads(reads)
    lvc = _lvc(fa := _fasta(tmp_path, "chrS", ref), th)
    lvc.maxDepth = 500
    lvc.process_batch(coded)
    _check_against_golden_memory(lvc, oc.memory, "depth cap, codes")
    assert_variants_equal(lvc.prepare_variants(), oc.prepare_variants(), "depth cap, codes")
    lvc.close()


@pytest.mark.parametrize("form", ["bytes", "codes"])
def test_admitted_only_batch_gives_the_same_tables_and_records(lib, tmp_path, form):
    """ReadBatch.admitted_only(): leaving the reads the admission dropped out of the batch changes nothing but the
    numbering of the first-seen ordinals (same order)"""
    from lvc_b200 import packing
    th = THS["loose"]
    ref, reads = synth_small.make_scenario(seed=501, ref_len=400, n_reads=2500, len_lo=60, len_hi=150, q_lo=0, q_hi=0,
                                           indel_rate=0.04, weird=True, amplicon=(0, 100, 101, 230), qbins=QBINS)
    full = packing.pack_reads(_tuples(reads), th["minMQ"], 300)
    lean = full.admitted_only()
    assert lean.n_reads == int((full.keep & 1).sum()) < full.n_reads and bool((lean.keep & 1).all())
    if form == "codes":
        full, lean = full.with_quality_codes(), lean.with_quality_codes()
        assert full.qcode is not None and lean.qcode is not None
    fa = _fasta(tmp_path, "chrS", ref)
    a, b = _lvc(fa, th), _lvc(fa, th)
    for _ in range(2):                                  # two live batches: ordinals keep counting
        a.process_batch(full)
        b.process_batch(lean)
    assert memory_tables(a.memory) == memory_tables(b.memory)
    assert_variants_equal(a.prepare_variants(), b.prepare_variants(), "admitted_only")
    ha, hb = a._handle, b._handle
    for k in sorted(set(ha.plane_keys().tolist()) & set(hb.plane_keys().tolist())):
        assert np.array_equal(ha.copy_plane(k), hb.copy_plane(k))
    assert np.array_equal(ha.copy_dels(), hb.copy_dels()) and np.array_equal(ha.copy_covdiff(), hb.copy_covdiff())
    oc = po.OracleCaller(ref, th["minBQ"], th["minMQ"], th["minDP"], th["minAD"], th["ratio"], max_depth=300)
    oc.process_reads(reads); oc.process_reads(reads)
    _check_against_golden_memory(b, oc.memory, "admitted_only vs oracle")
    a.close(); b.close()


def test_native_ingest_hands_out_codes_and_other_kernels_refuse(lib, tmp_path):
    """process_bam: the native ingest produces the code form by itself for a binned-quality BAM (and the byte form stays
    available); kernels without a code path refuse the batch instead of misreading it"""
    from lvc_b200 import samio, capi
    th = THS["loose"]
    ref, reads = synth_small.make_scenario(seed=401, ref_len=800, n_reads=900, len_lo=60, len_hi=151, q_lo=0, q_hi=0,
                                           indel_rate=0.03, weird=True, qbins=QBINS)
    fa = _fasta(tmp_path, "chrS", ref)
    bam = str(tmp_path / "codes.bam")
    samio.write_bam(bam, [("chrS", len(ref))], [(r.flag, r.pos, r.mapq, r.cigar, r.seq, r.qual, r.name) for r in reads])
    nat = samio.read_alignments_native(bam, None, th["minMQ"])
    assert nat.batch.qual_bits == 2 and sorted(nat.batch.qual_dict) == sorted(QBINS) and nat.batch_bytes.qual_bits == 8
    rb = nat.as_readbatch()
    assert rb.qcode is not None and np.array_equal(rb.qual[:rb.n_qual], packing_qual(reads, th["minMQ"]))
    nat.close()
    oc = po.OracleCaller(ref, th["minBQ"], th["minMQ"], th["minDP"], th["minAD"], th["ratio"])
    oc.process_reads(reads)
    lvc = _lvc(fa, th)
    lvc.process_bam(bam)
    _check_against_golden_memory(lvc, oc.memory, "process_bam, codes")
    assert_variants_equal(lvc.prepare_variants(), oc.prepare_variants(), "process_bam, codes")
    lvc.close()
    os.environ["LVC_QUALITY_CODES"] = "0"
    try:
        nat = samio.read_alignments_native(bam, None, th["minMQ"])
        assert nat.batch.qual_bits == 8
        nat.close()
    finally:
        del os.environ["LVC_QUALITY_CODES"]
    _, coded = _coded(reads, th["minMQ"])
    for impl in (2, 3, 4, 6):
        lvc = _lvc(fa, th, impl=impl)
        with pytest.raises(capi.LvcError):
            lvc.process_batch(coded)
        lvc.close()


def packing_qual(reads, mq):
    from lvc_b200 import packing
    b = packing.pack_reads(_tuples(reads), mq)
    return b.qual[:b.n_qual]


def _with_ns(reads, q_fail, every=7):
    """an N with a failing quality in every `every`-th read (what Illumina writes for a no-call)"""
    out = []
    for k, r in enumerate(reads):
        if k % every == 0 and len(r.seq) > 3:
            j = (k * 13) % len(r.seq)
            r.seq = r.seq[:j] + "N" + r.seq[j + 1:]
            r.qual = list(r.qual)
            r.qual[j] = q_fail
        out.append(r)
    return out


@pytest.mark.parametrize("impl", [0, 1, 5])
@pytest.mark.parametrize("tname", ["strict", "loose", "zero"])
@pytest.mark.parametrize("scen", ["deep_amplicon", "indel_dense", "two_values"])
def test_base_code_batches_vs_oracle_and_nibbles(lib, tmp_path, scen, tname, impl):
    """2-bit base codes on top of the quality codes (lvc_batch seq_form 2): no-calls with a failing quality are not
    represented -- the pileup never shows them --, everything else gives the tables of the oracle and of the nibble form"""
    cfg = dict(
        deep_amplicon=dict(seed=112, ref_len=900, n_reads=3000, len_lo=149, len_hi=150, q_lo=0, q_hi=0, indel_rate=0.01, weird=False,
                           amplicon=(0, 250, 251, 600, 750), qbins=QBINS, planted=((10, 0.5), (300, 0.2), (620, 1.0))),
        indel_dense=dict(seed=113, ref_len=700, n_reads=500, len_lo=80, len_hi=250, q_lo=0, q_hi=0, indel_rate=0.12, weird=False, qbins=QBINS),
        two_values=dict(seed=114, ref_len=300, n_reads=300, len_lo=30, len_hi=101, q_lo=0, q_hi=0, indel_rate=0.02, weird=False, qbins=(37, 37, 11, 37)),
    )[scen]
    ref, reads = synth_small.make_scenario(**cfg)
    th = THS[tname]
    if th["minBQ"] > 11:
        reads = _with_ns(reads, min(cfg["qbins"]))
    fa = _fasta(tmp_path, "chrS", ref)
    raw, coded = _coded(reads, th["minMQ"])
    b2 = coded.with_base_codes(th["minBQ"])
    assert b2.scode is not None and (b2.as_capi().seq_form & 255) == 2, "the scenario must qualify for base codes"
    oc = po.OracleCaller(ref, th["minBQ"], th["minMQ"], th["minDP"], th["minAD"], th["ratio"])
    oc.process_reads(reads)
    lvc = _lvc(fa, th, impl=impl)
    lvc.process_batch(b2)
    _check_against_golden_memory(lvc, oc.memory, f"{scen}/{tname} base codes vs oracle")
    ref_run = _lvc(fa, th, impl=impl)
    ref_run.process_batch(coded)
    assert_variants_equal(lvc.prepare_variants(), ref_run.prepare_variants(), f"{scen}/{tname} base codes vs nibbles")
    h, hr = lvc._handle, ref_run._handle
    ka, kb = set(h.plane_keys().tolist()), set(hr.plane_keys().tolist())
    for k in sorted(ka | kb):
        pa = h.copy_plane(k) if k in ka else None
        pb = hr.copy_plane(k) if k in kb else None
        if pa is None or pb is None:
            assert not (pa if pb is None else pb).any(), f"plane {k} exists on one side only and is not empty"
        else:
            assert np.array_equal(pa, pb), f"plane {k}"
    assert np.array_equal(h.copy_dels(), hr.copy_dels()) and np.array_equal(h.copy_covdiff(), hr.copy_covdiff())
    a, b = h.copy_first(0), hr.copy_first(0)
    assert (a is None) == (b is None) and (a is None or np.array_equal(a, b)), "first-seen ordinals"
    lvc.close(); ref_run.close()


def test_base_codes_need_the_threshold_they_were_made_for(lib, tmp_path):
    """codes made for minBQ 30 leave out the no-calls below 30: a handle with a lower threshold must refuse them; a
    batch with a no-call that passes keeps its nibbles"""
    from lvc_b200 import capi
    ref, reads = synth_small.make_scenario(seed=120, ref_len=300, n_reads=200, len_lo=50, len_hi=100, q_lo=0, q_hi=0,
                                           indel_rate=0.0, weird=False, qbins=QBINS)
    reads = _with_ns(reads, 12)
    fa = _fasta(tmp_path, "chrS", ref)
    raw, coded = _coded(reads, 0)
    b2 = coded.with_base_codes(30)
    assert b2.scode is not None
    assert coded.with_base_codes(12).scode is None              # an N at quality 12 passes a threshold of 12
    lvc = _lvc(fa, THS["loose"])                                # minBQ 13 < 30
    with pytest.raises(capi.LvcError):
        lvc.process_batch(b2)
    lvc.close()
    lvc = _lvc(fa, dict(THS["strict"], minBQ=35))               # a higher threshold is fine
    lvc.process_batch(b2)
    oc = po.OracleCaller(ref, 35, 20, 10, 5, 0.10)
    oc.process_reads(reads)
    _check_against_golden_memory(lvc, oc.memory, "base codes at a higher threshold")
    lvc.close()


def test_process_bam_ships_base_codes(lib, tmp_path):
    """process_bam on a BAM with instrument-binned qualities and no-calls at quality 2: the native reader hands out
    quality codes + base codes (lvc_reads_batch_for); records and memory are the oracle's"""
    from lvc_b200 import samio, capi
    th = THS["strict"]
    ref, reads = synth_small.make_scenario(seed=121, ref_len=800, n_reads=2500, len_lo=100, len_hi=151, q_lo=0, q_hi=0,
                                           indel_rate=0.02, weird=False, qbins=QBINS)
    reads = _with_ns(reads, 2)
    bam = str(tmp_path / "b2.bam")
    samio.write_bam(bam, [("chrS", len(ref))], [(r.flag, r.pos, r.mapq, r.cigar, r.seq, r.qual, f"r{k}") for k, r in enumerate(reads)])
    nat = capi.NativeReads(bam, None, th["minMQ"])
    nat.compact()
    assert (nat.batch_for(30).seq_form & 255) == 2 and nat.batch_for(30).qual_bits == 2
    assert (nat.batch_for(0).seq_form & 255) == 0               # the no-calls pass a threshold of 0: nibbles
    nat.close()
    fa = _fasta(tmp_path, "chrS", ref)
    lvc = _lvc(fa, th)
    lvc.process_bam(bam)
    oc = po.OracleCaller(ref, th["minBQ"], th["minMQ"], th["minDP"], th["minAD"], th["ratio"])
    oc.process_reads(reads)
    _check_against_golden_memory(lvc, oc.memory, "process_bam with base codes")
    assert_variants_equal(lvc.prepare_variants(), oc.prepare_variants(), "process_bam with base codes")
    lvc.close()

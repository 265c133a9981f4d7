"""CPU: C-ABI symbols, host admission (lvc_admit) vs the oracle, packing, readers, record finalisation."""
import ctypes
import os
import re

import numpy as np
import pytest

from helpers import GOLD, ROOT, po, synth_small, rows_to_tuples, variants_from_golden


def test_abi_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "lvc.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(lvc_[a-z_0-9]+)\s*\(", hdr))
    assert len(names) >= 35
    from lvc_b200 import capi
    bound = {n for n, _, _ in capi.SIGNATURES}
    assert names == bound, names ^ bound
    for n in names:
        assert hasattr(lib, n), n
    assert lib.lvc_version() == 1
    assert ctypes.sizeof(capi.Candidate) == 48


def test_no_cpu_fallback_without_gpu(lib):
    """Creating a handle without a CUDA device must fail loudly (no CPU path)."""
    from conftest import has_gpu
    if has_gpu():
        pytest.skip("a GPU is present")
    from lvc_b200 import capi
    with pytest.raises(capi.LvcError) as ei:
        capi.Handle(b"ACGT" * 10, 0, 0)
    assert "LVC_ENODEVICE" in str(ei.value)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "covid-spings-variant-caller_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "oracle" not in txt.replace("oracle/", "").lower() or f == "__init__.py" or \
                    all("import" not in ln for ln in txt.splitlines() if "oracle" in ln.lower()), f


def test_lvc_admit_matches_oracle(lib, golden_synth):
    from lvc_b200 import packing
    import random
    # golden scenarios incl. the >8000-reads-per-position one
    for scen in ("mixed_small", "maxdepth", "amplicon_like"):
        reads = synth_small.rows_to_reads(golden_synth[scen]["reads"])
        for mq in (0, 20):
            b = packing.pack_reads([(r.flag, r.pos, r.mapq, r.cigar, r.seq, r.qual) for r in reads], mq)
            want = po.admission_mask([r.pos for r in reads], [r.end() for r in reads],
                                     [po.passes_read_filter(r, mq) for r in reads])
            assert (b.keep & 1).astype(bool).tolist() == want, scen
    # random tiny max_depth
    rng = random.Random(7)
    for trial in range(40):
        n = rng.randint(1, 200)
        rd = []
        for i in range(n):
            pos = rng.choice([0, 0, 3, 3, 3, 7, 20, 21, 22, 40, 5000, 5001])
            ln = rng.randint(1, 25)
            rd.append(po.Read(rng.choice([0, 16, 0x400, 0x1, 0x3]), pos, rng.choice([60, 5]), [(0, ln)], "A" * ln, [30] * ln))
        rd = po.samtools_sort(rd)
        for md in (1, 2, 4, 8000):
            b = packing.pack_reads([(r.flag, r.pos, r.mapq, r.cigar, r.seq, r.qual) for r in rd], 10, md)
            want = po.admission_mask([r.pos for r in rd], [r.end() for r in rd],
                                     [po.passes_read_filter(r, 10) for r in rd], md)
            assert (b.keep & 1).astype(bool).tolist() == want, (trial, md)


def test_unsorted_input_raises(lib):
    from lvc_b200 import packing
    with pytest.raises(ValueError):
        packing.pack_reads([(0, 10, 60, [(0, 5)], "ACGTA", [30] * 5), (0, 3, 60, [(0, 5)], "ACGTA", [30] * 5)], 0)


def test_packing_layout(lib):
    from lvc_b200 import packing
    b = packing.pack_reads([(0, 1, 60, [(4, 2), (0, 3)], "ACGTN", [1, 2, 3, 4, 5]),
                            (16, 4, 60, [(0, 4)], "TTGA", [9, 9, 9, 9]),
                            (0, 9, 60, [(0, 2), (1, 1), (0, 2)], "ACGTA", [30] * 5)], 0)
    assert b.seq_off.tolist() == [0, 6, 10, 16]
    assert b.seq4[:3].tolist() == [0x12, 0x48, 0xF0]
    assert b.seq4[3:5].tolist() == [0x88, 0x41]
    assert b.qual[:6].tolist() == [1, 2, 3, 4, 5, 0]
    assert b.cigar[:2].tolist() == [(2 << 4) | 4, (3 << 4) | 0]
    assert (b.keep & 1).tolist() == [1, 1, 1]
    assert (b.keep >> 1).tolist() == [0, 1, 1]          # read 0 has an N; the pad nibble of odd reads is ignored
    assert b.aligned_bases() == 3 + 4 + 4
    assert b.algorithmic_bytes(100) == 3 * 20 + 4 * 6 + (3 + 2 + 3) + (5 + 4 + 5) + 5200


def test_sam_and_bam_readers_agree(lib, tmp_path):
    from lvc_b200 import samio
    sam = os.path.join(GOLD, "testfile.sam")
    contigs, b = samio.read_sam(sam, None, 0)
    assert contigs == [("NC_045512.2", 29903)] and b.n_reads == 4
    assert b.pos.tolist() == [10, 24, 24, 24]
    assert b.aligned_bases() == 1609
    _, reads = po.read_sam(sam)
    bam = str(tmp_path / "t.bam")
    samio.write_bam(bam, contigs, [(r.flag, r.pos, r.mapq, r.cigar, r.seq, r.qual, r.name) for r in reads])
    contigs2, b2 = samio.read_alignments(bam, "NC_045512.2", 0)
    assert contigs2 == contigs
    for f in ("pos", "flag", "mapq", "keep", "cigar_off", "seq_off"):
        assert getattr(b, f).tolist() == getattr(b2, f).tolist(), f
    assert b.cigar[:b.n_cigar].tolist() == b2.cigar[:b2.n_cigar].tolist()
    assert b.qual[:b.n_qual].tolist() == b2.qual[:b2.n_qual].tolist()
    assert b.seq4[:b.n_qual // 2].tolist() == b2.seq4[:b2.n_qual // 2].tolist()
    with pytest.raises(ValueError):
        samio.read_alignments(bam, "chrNope", 0)
    with pytest.raises(OSError):
        samio.read_alignments(str(tmp_path / "missing.bam"), None, 0)


def test_overlapping_mates_are_rewritten_not_refused(lib, tmp_path):
    """proper pairs whose mates overlap (every real 2x150 amplicon run) are processed the way pysam's default
    pileup(ignore_overlaps=True) does: qualities rewritten over the shared columns (tests/test_overlap_cpu.py has
    the htslib emulation); ignore_overlaps=False leaves them alone"""
    from lvc_b200 import samio, capi
    p = tmp_path / "ov.sam"
    p.write_text("@SQ\tSN:c\tLN:1000\n"
                 "a\t99\tc\t11\t60\t50M\t=\t41\t80\t" + "A" * 50 + "\t" + "I" * 50 + "\n"
                 "a\t147\tc\t41\t60\t50M\t=\t11\t-80\t" + "A" * 50 + "\t" + "I" * 50 + "\n")
    _, b = samio.read_sam(str(p), None, 0, overlap_model=capi.OVERLAP_HTSLIB_1_10)
    assert b.overlap_pairs == 1 and b.overlap_bases == 20
    assert b.qual[:50].tolist() == [40] * 30 + [80] * 20 and b.qual[50:100].tolist() == [0] * 20 + [40] * 30
    _, b = samio.read_sam(str(p), None, 0, overlap_model=capi.OVERLAP_OFF)
    assert b.overlap_pairs == 0 and b.qual[:100].tolist() == [40] * 100
    _, b = samio.read_sam(str(p), None, 0)                    # default model: one of the mates keeps the sum
    assert sorted([b.qual[30:50].tolist(), b.qual[50:70].tolist()]) == [[0] * 20, [80] * 20]


def test_record_finalisation_matches_golden(golden_synth):
    """records.py turns (L, S, esum, AD, DP, first) into the reference's dicts: feed it the ORACLE's numbers."""
    from lvc_b200 import records
    from helpers import assert_variants_equal
    g = golden_synth["deep_underflow"]
    reads = synth_small.rows_to_reads(g["reads"])
    th = g["results"]["zero"]["thresholds"]
    oc = po.OracleCaller(g["ref"], th["minBQ"], th["minMQ"], th["minDP"], th["minAD"], th["ratio"])
    oc.process_reads(reads)
    rows = []
    for p, site in oc.memory.items():
        snvs = {a: [po.from_phred_scale(q) for q in site["snvs"][a]] for a in site["snvs"]}
        L = {a: float(po.genotype_likelihood(a, snvs)) for a in snvs}
        S = 0.0
        for v in L.values():
            S += v
        S = S if S != 0 else 1.0
        for rank, a in enumerate(snvs):
            if site["reference"] != a:
                rows.append((p, records.NIBBLE_CHARS.index(a), ord(site["reference"]), 0, len(snvs[a]),
                             site["totalDepth"], rank, 0, L[a], S, float(np.sum(snvs[a]))))
    from lvc_b200.capi import CANDIDATE_DTYPE
    cands = np.array(rows, dtype=CANDIDATE_DTYPE)
    got = records.candidates_to_variants(cands)
    assert_variants_equal(got, variants_from_golden(g["results"]["zero"]["variants"]), "finalise")
    e, om = records.phred_luts()
    assert e[30] == 0.001 and om[0] == 0.0


def test_vcf_text_layout():
    from lvc_b200 import records
    v = [{"start": 9, "stop": 10, "alleles": ("A", "G"), "qual": 0.000316227766, "info":
          {"DP": 70, "AD": 60, "GL": -35.00824146158652, "PL": 350, "SCORE": 99}},
         {"start": 9, "stop": 10, "alleles": ("A", "T"), "qual": 0.001, "info":
          {"DP": 70, "AD": 5, "GL": 0, "PL": 0, "SCORE": 0}}]
    txt = records.format_vcf(v, [("NC_045512.2", 29903)])
    assert txt == po.format_vcf(v, [("NC_045512.2", 29903)])
    # the whole file against a LITERAL fixture (tests/golden/expected_two_records.vcf): reviewed by hand against the VCF 4.2
    # text htslib writes for a default pysam.VariantHeader() + the five add_meta calls of live_variant_caller.py:237-272 +
    # contigs.add (:274-278) + new_record without contig / filter (:287-295) -- header order, PASS filter line, "." in
    # ID and FILTER, kputd floats.  [EXT]: not produced by pysam (absent from this image); tests/test_pysam_crosscheck.py
    # compares with the real writer wherever pysam imports.
    with open(os.path.join(GOLD, "expected_two_records.vcf")) as fh:
        assert txt == fh.read()
    lines = txt.strip().split("\n")
    assert lines[0] == "##fileformat=VCFv4.2" and lines[-3].startswith("#CHROM")
    assert lines[-2] == "NC_045512.2\t10\t.\tA\tT\t0.001\t.\tDP=70;AD=5;GL=0;PL=0;SCORE=0"
    assert lines[-1] == "NC_045512.2\t10\t.\tA\tG\t0.000316228\t.\tDP=70;AD=60;GL=-35.0082;PL=350;SCORE=99"


def test_native_ingest_matches_python_readers(lib, golden_synth, tmp_path):
    """lvc_read_alignments (C++: BGZF inflate, SAM parse + samtools-order sort, packing, admission, ACGT hint)
    against the independent pure-Python readers, on the reference's fixture and on a generated BAM"""
    from lvc_b200 import samio, packing
    sam = os.path.join(GOLD, "testfile.sam")
    cases = [(sam, None, 0), (sam, "NC_045512.2", 20)]
    # a BAM with every CIGAR op, filtered reads, odd lengths, N bases
    g = golden_synth["mixed_small"]
    reads = synth_small.rows_to_reads(g["reads"])
    bam = str(tmp_path / "m.bam")
    samio.write_bam(bam, [("chrS", len(g["ref"]))], [(r.flag, r.pos, r.mapq, r.cigar, r.seq, r.qual, r.name) for r in reads])
    cases += [(bam, "chrS", 0), (bam, None, 20)]
    # an unsorted SAM: both readers sort like samtools
    rows = g["reads"][::-1]
    usam = tmp_path / "u.sam"
    usam.write_text("@SQ\tSN:chrS\tLN:%d\n" % len(g["ref"]) +
                    "".join(f"{n}\t{f}\tchrS\t{p + 1}\t{m}\t{c}\t*\t0\t0\t{s}\t{q}\n" for n, f, p, m, c, s, q in rows))
    cases.append((str(usam), None, 0))
    for path, contig, mq in cases:
        _, pyb = samio.read_alignments(path, contig, mq)
        nat = samio.read_alignments_native(path, contig, mq, n_threads=3)
        nb = nat.as_readbatch()
        assert nat.n_reads == pyb.n_reads, path
        for f in ("pos", "flag", "mapq", "keep", "cigar_off", "seq_off"):
            assert getattr(nb, f).tolist() == getattr(pyb, f).tolist(), (path, f)
        assert nb.cigar[:nb.n_cigar].tolist() == pyb.cigar[:pyb.n_cigar].tolist()
        assert nb.qual[:nb.n_qual].tolist() == pyb.qual[:pyb.n_qual].tolist()
        assert nb.seq4[:nb.n_qual // 2].tolist() == pyb.seq4[:pyb.n_qual // 2].tolist()
        nat.close()
    with pytest.raises(ValueError):
        samio.read_alignments_native(bam, "chrNope", 0)
    with pytest.raises(OSError):
        samio.read_alignments_native(str(tmp_path / "missing.bam"), None, 0)
    ov = tmp_path / "ov.sam"
    ov.write_text("@SQ\tSN:c\tLN:1000\n"
                  "a\t99\tc\t11\t60\t50M\t=\t41\t80\t" + "A" * 50 + "\t" + "I" * 50 + "\n"
                  "a\t147\tc\t41\t60\t50M\t=\t11\t-80\t" + "A" * 50 + "\t" + "I" * 50 + "\n")
    nat = samio.read_alignments_native(str(ov), None, 0)      # overlapping mates: rewritten, not refused
    assert nat.overlap_pairs == 1 and nat.overlap_bases == 20
    nat.close()
    bad = tmp_path / "badq.sam"
    bad.write_text("@SQ\tSN:c\tLN:1000\na\t0\tc\t11\t60\t5M\t*\t0\t0\tACGTA\tIII\n")
    with pytest.raises(packing.UnsupportedInput):               # len(QUAL) != len(SEQ)
        samio.read_alignments_native(str(bad), None, 0)


def test_native_ingest_random_files_match_python_readers(lib, tmp_path):
    """seeded random SAM and BAM files (random CIGARs with every op, odd lengths, ambiguity codes, filtered flags,
    unsorted SAM input, several contigs): the C++ ingest and the pure-Python readers must produce the same batch"""
    from lvc_b200 import samio
    rng = np.random.default_rng(20261018)
    letters = "ACGTNRYKM"
    ops_q = {"M": 1, "I": 1, "S": 1, "=": 1, "X": 1, "D": 0, "N": 0, "H": 0, "P": 0}
    seen_reads = seen_kept = 0
    for trial in range(12):
        G = int(rng.integers(300, 3000))
        n = int(rng.integers(1, 120))
        rows = []
        for i in range(n):
            k = int(rng.integers(1, 9))
            names = list(rng.choice(list("MIDNSHP=X"), size=k))
            if not any(o in "MDN=X" for o in names):
                names[int(rng.integers(0, k))] = "M"
            lens = [int(rng.integers(1, 40)) for _ in names]
            lq = sum(l for o, l in zip(names, lens) if ops_q[o])
            if lq == 0:
                names.append("M"); lens.append(5); lq = 5
            rlen = sum(l for o, l in zip(names, lens) if o in "MDN=X")
            pos = int(rng.integers(0, max(1, G - rlen - 1)))
            flag = int(rng.choice([0, 16, 0, 16, 1024, 256, 4, 2048]))
            seq = "".join(rng.choice(list(letters), size=lq, p=[.24, .24, .24, .24, .01, .01, .01, .005, .005]))
            qual = "".join(chr(33 + int(q)) for q in rng.integers(0, 60, lq))
            contig = "c2" if rng.random() < 0.15 else "c1"
            rows.append((f"r{i}", flag, contig, pos, int(rng.integers(0, 61)), "".join(f"{l}{o}" for o, l in zip(names, lens)), seq, qual))
        head = f"@HD\tVN:1.6\n@SQ\tSN:c1\tLN:{G}\n@SQ\tSN:c2\tLN:{G}\n"
        sam = tmp_path / f"t{trial}.sam"
        sam.write_text(head + "".join(f"{nm}\t{fl}\t{cg}\t{p + 1}\t{mq}\t{ci}\t*\t0\t0\t{s}\t{q}\n"
                                      for nm, fl, cg, p, mq, ci, s, q in rows))
        cases = [(str(sam), None), (str(sam), "c2")]
        # the same reads of c1, sorted the samtools way, as a BAM
        c1 = sorted([r for r in rows if r[2] == "c1"], key=lambda r: (r[3], (r[1] >> 4) & 1))
        if c1:
            opcode = {o: i for i, o in enumerate("MIDNSHP=X")}
            import re
            bam = str(tmp_path / f"t{trial}.bam")
            samio.write_bam(bam, [("c1", G), ("c2", G)],
                            [(fl, p, mq, [(opcode[o], int(l)) for l, o in re.findall(r"(\\d+)([MIDNSHP=X])", ci)], s,
                              [ord(ch) - 33 for ch in q], nm) for nm, fl, cg, p, mq, ci, s, q in c1])
            cases.append((bam, "c1"))
        for path, contig in cases:
            mq = int(rng.integers(0, 30))
            _, pyb = samio.read_alignments(path, contig, mq)
            nat = samio.read_alignments_native(path, contig, mq, n_threads=int(rng.integers(1, 6)))
            nb = nat.as_readbatch()
            assert nat.n_reads == pyb.n_reads, (path, contig)
            for f in ("pos", "flag", "mapq", "keep", "cigar_off", "seq_off"):
                assert getattr(nb, f).tolist() == getattr(pyb, f).tolist(), (path, contig, f)
            assert nb.cigar[:nb.n_cigar].tolist() == pyb.cigar[:pyb.n_cigar].tolist()
            assert nb.qual[:nb.n_qual].tolist() == pyb.qual[:pyb.n_qual].tolist()
            assert nb.seq4[:nb.n_qual // 2].tolist() == pyb.seq4[:pyb.n_qual // 2].tolist()
            seen_reads += nat.n_reads
            seen_kept += int((nb.keep & 1).sum())
            nat.close()
    assert seen_reads > 500 and 0 < seen_kept < seen_reads          # the comparison saw real, partly filtered data


def _rewrite_bgzf(src: str, dst: str, level: int, strategy: int, block_sizes):
    """the same uncompressed bytes as BGZF blocks of the given sizes (cycled), deflated with (level, strategy)"""
    import struct, zlib
    from lvc_b200 import samio
    body = samio._bgzf_decompress(src)
    with open(dst, "wb") as fh:
        i = k = 0
        while i < len(body):
            chunk = body[i:i + block_sizes[k % len(block_sizes)]]
            i += len(chunk); k += 1
            comp = zlib.compressobj(level, zlib.DEFLATED, -15, 8, strategy)
            cdata = comp.compress(chunk) + comp.flush()
            assert len(cdata) + 25 < 65536
            fh.write(b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", len(cdata) + 25))
            fh.write(cdata)
            fh.write(struct.pack("<II", zlib.crc32(chunk) & 0xFFFFFFFF, len(chunk)))
        fh.write(samio._BGZF_EOF)


def test_native_inflate_every_deflate_form(lib, tmp_path, monkeypatch):
    """the ingest's own DEFLATE decoder (csrc/inflate_fast.hpp) on stored, fixed-code and dynamic-code blocks of every size
    class (1 byte ... 0xFF00, i.e. blocks that never enter its hot loop and blocks that live in it), against zlib's
    inflate (LVC_INFLATE=zlib) and the pure-Python reader; corrupted and truncated blocks are rejected, not misread"""
    import zlib
    from lvc_b200 import samio, synth, capi, packing
    rejected = (capi.LvcError, packing.UnsupportedInput)
    ref, batch = synth.amplicon_sample(seed=5, n_pairs=4000)[:2]
    src = str(tmp_path / "src.bam")
    samio.write_bam_batch(src, ("chrS", len(ref)), batch, level=6)
    _, want = samio.read_alignments(src, None, 20)

    def native(path):
        nat = samio.read_alignments_native(path, None, 20, n_threads=3)
        nb = nat.as_readbatch()
        got = {f: np.array(getattr(nb, f)[:len(getattr(want, f))]) for f in ("pos", "flag", "mapq", "keep", "cigar_off", "seq_off")}
        got["cigar"] = np.array(nb.cigar[:nb.n_cigar]); got["qual"] = np.array(nb.qual[:nb.n_qual])
        got["seq4"] = np.array(nb.seq4[:nb.n_qual // 2]); got["n"] = nat.n_reads
        nat.close()
        return got

    forms = [(0, zlib.Z_DEFAULT_STRATEGY, [0xFF00]),                      # stored blocks
             (6, zlib.Z_FIXED, [0xFF00, 300, 1]),                         # fixed Huffman code, tiny blocks too
             (1, zlib.Z_DEFAULT_STRATEGY, [0xFF00, 17, 4096]),
             (9, zlib.Z_DEFAULT_STRATEGY, [0xFF00]),
             (6, zlib.Z_HUFFMAN_ONLY, [0x8000, 281, 282, 283]),           # literals only; sizes around the hot loop's margin
             (6, zlib.Z_RLE, [0xFF00, 5000])]                             # distance-1 matches
    for level, strategy, sizes in forms:
        path = str(tmp_path / f"f{level}_{strategy}.bam")
        _rewrite_bgzf(src, path, level, strategy, sizes)
        monkeypatch.delenv("LVC_INFLATE", raising=False)
        fast = native(path)
        monkeypatch.setenv("LVC_INFLATE", "zlib")
        slow = native(path)
        assert fast["n"] == slow["n"] == want.n_reads, (level, strategy)
        for f in fast:
            assert np.array_equal(fast[f], slow[f]), (level, strategy, f)
        for f in ("pos", "flag", "mapq", "keep", "cigar_off", "seq_off"):
            assert fast[f].tolist() == getattr(want, f).tolist(), (level, strategy, f)
        assert fast["qual"].tolist() == want.qual[:want.n_qual].tolist()
    monkeypatch.delenv("LVC_INFLATE", raising=False)
    # a flipped bit anywhere in a deflate stream, or a block cut short, is an error (the decoder's own checks or the CRC)
    good = open(src, "rb").read()
    rng = np.random.default_rng(3)
    for trial in range(12):
        bad = bytearray(good)
        at = int(rng.integers(18, len(good) - 60))
        bad[at] ^= 1 << int(rng.integers(0, 8))
        p = str(tmp_path / "bad.bam")
        open(p, "wb").write(bytes(bad))
        try:
            got = native(p)
        except rejected:
            continue
        # the flip hit a header byte the reader does not interpret (gzip MTIME / XFL / OS): the data must be intact
        assert got["qual"].tolist() == want.qual[:want.n_qual].tolist(), at
    with pytest.raises(rejected):
        open(str(tmp_path / "cut.bam"), "wb").write(good[:len(good) // 2])
        native(str(tmp_path / "cut.bam"))


def test_native_ingest_reuses_its_parked_inflate_buffer(lib, tmp_path, monkeypatch):
    """files of more than 8 MB inflate into a huge-page backed mapping that is parked between calls: a second, different
    file read through the reused mapping (and a file read with LVC_INGEST_PARK=0) gives the arrays of a fresh read"""
    from lvc_b200 import samio, synth
    files = []
    for seed, pairs in ((21, 36_000), (22, 30_000)):
        ref, batch = synth.amplicon_sample(seed=seed, n_pairs=pairs)[:2]
        path = str(tmp_path / f"p{seed}.bam")
        samio.write_bam_batch(path, ("chrS", len(ref)), batch, level=1)
        files.append((path, batch))

    def read(path):
        nat = samio.read_alignments_native(path, None, 20, n_threads=4)
        nb = nat.as_readbatch()
        out = (nat.n_reads, nb.pos.copy(), nb.keep.copy(), nb.cigar[:nb.n_cigar].copy(), nb.qual[:nb.n_qual].copy(),
               nb.seq4[:nb.n_qual // 2].copy())
        nat.close()
        return out

    monkeypatch.setenv("LVC_INGEST_PARK", "0")
    fresh = [read(p) for p, _ in files]
    monkeypatch.delenv("LVC_INGEST_PARK")
    for rnd in range(2):                                   # big file, smaller file, big file ...: the mapping is reused
        for k, (p, batch) in enumerate(files):
            got = read(p)
            assert got[0] == fresh[k][0] == batch.n_reads
            for a, b in zip(got[1:], fresh[k][1:]):
                assert np.array_equal(a, b), (rnd, k)
            assert np.array_equal(got[4], batch.qual[:batch.n_qual]) and np.array_equal(got[1], batch.pos)

"""Minimal SAM / BAM / FASTA readers and a BAM writer (host ingest; replaces the pysam calls of
live_variant_caller.py:30,55-60,78 and client_server/vc_queue.py:34-35 that this image cannot run).

Pure Python + zlib + numpy: correctness first.  A native multi-threaded BGZF decoder is the next
row of the scope table.
"""
from __future__ import annotations

import struct
import zlib
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import packing
from .packing import ReadBatch, UnsupportedInput, CIGAR_OPS, _ASCII_TO_NIBBLE

_BGZF_EOF = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


# ------------------------------------------------------------------------------------------ FASTA
class Fasta:
    """What the reference uses of pysam.FastaFile: .references, .fetch(reference=), .get_reference_length."""

    def __init__(self, path: str):
        self.path = path
        self._seqs: Dict[str, str] = {}
        name, chunks = None, []
        with open(path) as fh:       # raises OSError like pysam for a missing file
            for line in fh:
                line = line.rstrip("\r\n")
                if line.startswith(">"):
                    if name is not None:
                        self._seqs[name] = "".join(chunks)
                    name, chunks = line[1:].split()[0], []
                elif name is not None:
                    chunks.append(line)
        if name is not None:
            self._seqs[name] = "".join(chunks)
        if not self._seqs:
            raise ValueError(f"no sequences in {path}")
        self.references = list(self._seqs)

    def fetch(self, reference: str) -> str:
        return self._seqs[reference]

    def get_reference_length(self, reference: str) -> int:
        return len(self._seqs[reference])

    def close(self):
        pass


# ------------------------------------------------------------------------------------------ SAM
def _parse_cigar_text(text: str) -> List[Tuple[int, int]]:
    out, num = [], 0
    if text == "*":
        return out
    for ch in text:
        if "0" <= ch <= "9":
            num = num * 10 + ord(ch) - 48
        else:
            out.append((CIGAR_OPS.index(ch), num))
            num = 0
    return out


def read_sam(path: str, contig: Optional[str], min_mapq: int, sort: bool = True,
             max_depth: int = packing.capi.MAX_DEPTH_DEFAULT,
             overlap_model: int = packing.capi.OVERLAP_DEFAULT) -> Tuple[List[Tuple[str, int]], ReadBatch]:
    """SAM text -> packed batch of the reads on `contig` (default: first @SQ), coordinate sorted the way
    `samtools sort` does (position, forward strand first, input order) when `sort`."""
    contigs: List[Tuple[str, int]] = []
    rows = []
    with open(path) as fh:
        for line in fh:
            if line.startswith("@"):
                if line.startswith("@SQ"):
                    f = dict(x.split(":", 1) for x in line.rstrip("\n").split("\t")[1:])
                    contigs.append((f["SN"], int(f["LN"])))
                continue
            t = line.rstrip("\n").split("\t")
            if len(t) < 11:
                continue
            if contig is None:
                contig = contigs[0][0] if contigs else t[2]
            if t[2] != contig:
                continue
            flag, pos, mapq = int(t[1]), int(t[3]) - 1, int(t[4])
            ops = _parse_cigar_text(t[5])
            if t[10] == "*" and t[9] != "*" and not (flag & 0x4):
                raise UnsupportedInput(f"read {t[0]!r} has no base qualities (the reference raises TypeError)")
            qual = [ord(c) - 33 for c in t[10]] if t[10] != "*" else []
            seq = t[9] if t[9] != "*" else ""
            if qual and len(qual) != len(seq):
                raise UnsupportedInput(f"read {t[0]!r}: QUAL length {len(qual)} != SEQ length {len(seq)}")
            mref = 1 if t[6] in ("=", t[2]) else (-1 if t[6] == "*" else 0)
            rows.append((flag, pos, mapq, ops, seq, qual, t[0], int(t[7]) - 1, mref, int(t[8])))
    if sort:
        rows.sort(key=lambda r: (r[1], 1 if r[0] & 0x10 else 0))
    return contigs, packing.pack_reads(rows, min_mapq, max_depth, overlap_model)


# ------------------------------------------------------------------------------------------ BAM
def _bgzf_decompress(path: str) -> bytes:
    out = []
    with open(path, "rb") as fh:
        data = fh.read()
    i, n = 0, len(data)
    while i < n:
        if data[i:i + 4] != b"\x1f\x8b\x08\x04":
            raise ValueError(f"{path}: not a BGZF/BAM file")
        xlen = struct.unpack_from("<H", data, i + 10)[0]
        bsize = None
        j = i + 12
        while j < i + 12 + xlen:
            si1, si2, slen = data[j], data[j + 1], struct.unpack_from("<H", data, j + 2)[0]
            if si1 == 66 and si2 == 67:
                bsize = struct.unpack_from("<H", data, j + 4)[0]
            j += 4 + slen
        if bsize is None:
            raise ValueError(f"{path}: BGZF block without BC field")
        cdata = data[i + 12 + xlen:i + bsize + 1 - 8]
        out.append(zlib.decompress(cdata, -15) if len(cdata) else b"")
        i += bsize + 1
    return b"".join(out)


def read_bam(path: str, contig: Optional[str], min_mapq: int,
             max_depth: int = packing.capi.MAX_DEPTH_DEFAULT,
             overlap_model: int = packing.capi.OVERLAP_DEFAULT) -> Tuple[List[Tuple[str, int]], ReadBatch]:
    """BAM -> packed batch of the records whose reference is `contig` (file order = coordinate order).
    No index is needed (the reference needs one only because pysam's region iterator does)."""
    raw = _bgzf_decompress(path)
    if raw[:4] != b"BAM\x01":
        raise ValueError(f"{path}: bad BAM magic")
    l_text = struct.unpack_from("<i", raw, 4)[0]
    off = 8 + l_text
    n_ref = struct.unpack_from("<i", raw, off)[0]
    off += 4
    contigs = []
    for _ in range(n_ref):
        l_name = struct.unpack_from("<i", raw, off)[0]
        name = raw[off + 4:off + 4 + l_name - 1].decode()
        l_ref = struct.unpack_from("<i", raw, off + 4 + l_name)[0]
        contigs.append((name, l_ref))
        off += 8 + l_name
    names = [c[0] for c in contigs]
    if contig is None:
        contig = names[0]
    if contig not in names:
        raise ValueError(f"invalid contig `{contig}`")           # pysam's message for an unknown reference
    tid = names.index(contig)
    mv = memoryview(raw)
    arr = np.frombuffer(raw, dtype=np.uint8)
    pos, flag, mapq, coff, soff = [], [], [], [0], [0]
    cig_parts, seq_parts, qual_parts = [], [], []
    names, mpos, mref, tlens = [], [], [], []
    n = len(raw)
    while off + 4 <= n:
        bs = struct.unpack_from("<i", raw, off)[0]
        ref_id, p, l_rn, mq, _bin, n_cig, fl, l_seq, next_ref, next_pos, _tlen = struct.unpack_from(
            "<iiBBHHHiiii", raw, off + 4)
        rec_end = off + 4 + bs
        if ref_id == tid:
            o = off + 36 + l_rn
            cig = np.frombuffer(mv[o:o + 4 * n_cig], dtype="<u4")
            o += 4 * n_cig
            nsb = (l_seq + 1) // 2
            sq = arr[o:o + nsb]
            o += nsb
            ql = arr[o:o + l_seq]
            if l_seq and ql[0] == 0xFF and not (fl & 0x4):
                raise UnsupportedInput(f"{path}: a read has no base qualities (the reference raises TypeError)")
            ops = cig & 15
            lens = cig >> 4
            rlen = int(lens[(ops == 0) | (ops == 2) | (ops == 3) | (ops == 7) | (ops == 8)].sum())
            lq = int(lens[(ops == 0) | (ops == 1) | (ops == 4) | (ops == 7) | (ops == 8)].sum())
            if n_cig == 2 and (cig[0] & 15) == 4 and (cig[0] >> 4) == l_seq and (cig[1] & 15) == 3:
                raise UnsupportedInput(f"{path}: CIGAR stored in the CG tag (>65535 ops) is not supported")
            if lq != l_seq and rlen > 0:
                raise UnsupportedInput(f"{path}: read at {p}: CIGAR query length {lq} != l_seq {l_seq}")
            names.append(raw[off + 36:off + 36 + max(l_rn - 1, 0)].decode("latin-1"))
            mpos.append(next_pos); mref.append(-1 if next_ref < 0 else (1 if next_ref == ref_id else 0)); tlens.append(_tlen)
            pos.append(p); flag.append(fl); mapq.append(mq)
            if lq != l_seq:
                cig = cig[:0]; sq = sq[:0]; ql = ql[:0]; l_seq = 0
            cig_parts.append(cig)
            coff.append(coff[-1] + len(cig))
            if l_seq & 1:
                ql = np.concatenate([ql, np.zeros(1, np.uint8)])
            seq_parts.append(sq)
            qual_parts.append(ql)
            soff.append(soff[-1] + len(ql))
        off = rec_end
    cigar = np.concatenate(cig_parts).astype(np.uint32) if cig_parts else np.zeros(0, np.uint32)
    seq4 = np.concatenate(seq_parts) if seq_parts else np.zeros(0, np.uint8)
    qual = np.concatenate(qual_parts) if qual_parts else np.zeros(0, np.uint8)
    return contigs, packing.finalize_batch(pos, flag, mapq, coff, cigar, soff, seq4, qual, min_mapq, max_depth,
                                           mates=dict(names=names, mate_pos=mpos, mate_ref=mref, tlen=tlens),
                                           overlap_model=overlap_model)


def read_alignments_native(path: str, contig: Optional[str], min_mapq: int,
                           max_depth: int = packing.capi.MAX_DEPTH_DEFAULT, n_threads: int = 0,
                           overlap_model: int = packing.capi.OVERLAP_DEFAULT):
    """BAM or SAM through the library's native ingest (multi-threaded BGZF inflate, page-locked output).
    Returns a capi.NativeReads; `.batch` goes straight to Handle.push_batch, `.as_readbatch()` gives numpy views."""
    return packing.capi.NativeReads(path, contig, min_mapq, max_depth, n_threads, overlap_model)


def read_alignments(path: str, contig: Optional[str], min_mapq: int,
                    max_depth: int = packing.capi.MAX_DEPTH_DEFAULT,
                    overlap_model: int = packing.capi.OVERLAP_DEFAULT):
    """Pure-Python reader (kept as an independent cross-check of the native ingest).
    Dispatch on the file content: BGZF magic -> BAM, otherwise SAM text."""
    with open(path, "rb") as fh:
        magic = fh.read(4)
    if magic[:2] == b"\x1f\x8b":
        return read_bam(path, contig, min_mapq, max_depth, overlap_model)
    return read_sam(path, contig, min_mapq, True, max_depth, overlap_model)


def write_bam(path: str, contigs: List[Tuple[str, int]], reads, header_text: Optional[str] = None):
    """Write records (flag, pos0, mapq, [(op,len)], seq, qual_ints[, name[, mate_pos0, mate_ref, tlen]]) on contig 0 as a
    BGZF BAM (tests and synthetic fixtures).  Records must already be coordinate sorted.  mate_ref: 1 = same contig,
    0 = another contig (written as reference id 1), -1 = absent."""
    if header_text is None:
        header_text = "@HD\tVN:1.6\tSO:coordinate\n" + "".join(f"@SQ\tSN:{n}\tLN:{l}\n" for n, l in contigs)
    body = bytearray()
    body += b"BAM\x01" + struct.pack("<i", len(header_text)) + header_text.encode()
    body += struct.pack("<i", len(contigs))
    for n, l in contigs:
        body += struct.pack("<i", len(n) + 1) + n.encode() + b"\x00" + struct.pack("<i", l)
    for k, rec in enumerate(reads):
        fl, p, mq, ops, seq, qual = rec[:6]
        name = (rec[6] if len(rec) > 6 else f"r{k}").encode() + b"\x00"
        l_seq = len(seq)
        nib = _ASCII_TO_NIBBLE[np.frombuffer(seq.encode("ascii"), dtype=np.uint8)] if l_seq else np.zeros(0, np.uint8)
        if l_seq & 1:
            nib = np.concatenate([nib, np.zeros(1, np.uint8)])
        sq = ((nib[0::2] << 4) | nib[1::2]).astype(np.uint8).tobytes()
        ql = bytes(qual) if len(qual) == l_seq else b"\xff" * l_seq
        cig = b"".join(struct.pack("<I", (l << 4) | o) for o, l in ops)
        rlen = sum(l for o, l in ops if o in (0, 2, 3, 7, 8))
        m_pos, m_ref, t_len = (rec[7], rec[8], rec[9]) if len(rec) > 9 else (-1, -1, 0)
        core = struct.pack("<iiBBHHHiiii", 0, p, len(name), mq, 4680, len(ops), fl, l_seq,
                           0 if m_ref == 1 else (1 if m_ref == 0 else -1), m_pos, t_len)
        rec_b = core + name + cig + sq + ql
        body += struct.pack("<i", len(rec_b)) + rec_b
        _ = rlen
    with open(path, "wb") as fh:
        for i in range(0, len(body), 0xFF00):
            chunk = bytes(body[i:i + 0xFF00])
            comp = zlib.compressobj(6, zlib.DEFLATED, -15)
            cdata = comp.compress(chunk) + comp.flush()
            bsize = len(cdata) + 25
            fh.write(b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", bsize))
            fh.write(cdata)
            fh.write(struct.pack("<II", zlib.crc32(chunk) & 0xFFFFFFFF, len(chunk)))
        fh.write(_BGZF_EOF)


def write_bam_batch(path: str, contig: Tuple[str, int], batch, level: int = 1, name_id: Optional[np.ndarray] = None,
                    mate_pos: Optional[np.ndarray] = None, tlen: Optional[np.ndarray] = None):
    """Write a packed ReadBatch as a BGZF BAM without a per-read Python loop (ingest benchmarks and
    large fixtures).  Read names are 8 hex digits of `name_id` (default: the read index, i.e. all distinct; the two
    reads of a pair must share their id); `mate_pos` / `tlen` fill PNEXT (same contig) / TLEN, else the mate
    fields are written as absent (-1)."""
    n = batch.n_reads
    ncig = np.diff(batch.cigar_off[:n + 1]).astype(np.int64)
    lq = np.diff(batch.seq_off[:n + 1].astype(np.int64))
    lq_true = packing.query_lengths(batch.cigar_off, batch.cigar)[:n].astype(np.int64)
    sbytes = (lq_true + 1) // 2
    rec_len = 32 + 9 + 4 * ncig + sbytes + lq_true
    offs = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(rec_len + 4, out=offs[1:])
    out = np.zeros(int(offs[-1]), dtype=np.uint8)
    core = np.zeros(n, dtype=np.dtype([("bs", "<i4"), ("ref", "<i4"), ("pos", "<i4"), ("lname", "u1"), ("mapq", "u1"),
                                       ("bin", "<u2"), ("ncig", "<u2"), ("flag", "<u2"), ("lseq", "<i4"),
                                       ("nref", "<i4"), ("npos", "<i4"), ("tlen", "<i4"), ("name", "u1", (9,))]))
    core["bs"], core["pos"], core["lname"], core["mapq"] = rec_len, batch.pos[:n], 9, batch.mapq[:n]
    core["bin"], core["ncig"], core["flag"], core["lseq"] = 4680, ncig, batch.flag[:n], lq_true
    core["nref"] = -1 if mate_pos is None else 0
    core["npos"] = -1 if mate_pos is None else np.asarray(mate_pos)[:n]
    core["tlen"] = 0 if tlen is None else np.asarray(tlen)[:n]
    ids = (np.arange(n, dtype=np.uint32) if name_id is None else np.asarray(name_id)[:n].astype(np.uint32))
    hexd = np.frombuffer(b"0123456789abcdef", dtype=np.uint8)
    for k in range(8):
        core["name"][:, k] = hexd[(ids >> np.uint32(4 * (7 - k))) & np.uint32(15)]
    core["name"][:, 8] = 0
    hdr = core.view(np.uint8).reshape(n, 45)
    out[(offs[:n, None] + np.arange(45)[None, :]).ravel()] = hdr.ravel()

    def scatter(dst0, src0, lens, src):
        tot = int(lens.sum())
        if not tot:
            return
        starts = np.zeros(n, dtype=np.int64)
        np.cumsum(lens[:-1], out=starts[1:])
        within = np.arange(tot, dtype=np.int64) - np.repeat(starts, lens)
        out[np.repeat(dst0, lens) + within] = src[np.repeat(src0, lens) + within]

    cig8 = np.ascontiguousarray(batch.cigar).view(np.uint8)
    scatter(offs[:n] + 45, batch.cigar_off[:n].astype(np.int64) * 4, 4 * ncig, cig8)
    so = batch.seq_off[:n].astype(np.int64)
    scatter(offs[:n] + 45 + 4 * ncig, so // 2, sbytes, batch.seq4)
    scatter(offs[:n] + 45 + 4 * ncig + sbytes, so, lq_true, batch.qual)
    _ = lq
    name, length = contig
    header_text = f"@HD\tVN:1.6\tSO:coordinate\n@SQ\tSN:{name}\tLN:{length}\n"
    head = (b"BAM\x01" + struct.pack("<i", len(header_text)) + header_text.encode() + struct.pack("<i", 1) +
            struct.pack("<i", len(name) + 1) + name.encode() + b"\x00" + struct.pack("<i", length))
    body = head + out.tobytes()
    with open(path, "wb") as fh:
        for i in range(0, len(body), 0xFF00):
            chunk = body[i:i + 0xFF00]
            comp = zlib.compressobj(level, zlib.DEFLATED, -15)
            cdata = comp.compress(chunk) + comp.flush()
            fh.write(b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", len(cdata) + 25))
            fh.write(cdata)
            fh.write(struct.pack("<II", zlib.crc32(chunk) & 0xFFFFFFFF, len(chunk)))
        fh.write(_BGZF_EOF)

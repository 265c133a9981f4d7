"""Multi-GPU sharding of the hot path (one process per GPU, torch.distributed for the plumbing).

The path shards in exactly two ways:

* independent samples -> GPUs (:func:`assign_samples`): state is per caller instance
  (live_variant_caller.py:31-32), so there is NO communication; records are gathered on the host.
* one sample, contiguous chunks of the coordinate-sorted reads -> GPUs (:func:`shard_reads`): deposits
  are commutative integer adds, so the per-rank tables combine with ONE exchange step,
  an integer sum (counts, deletion counts, coverage) and a min (first-seen ordinals), done here
  with NCCL all-reduce over NVLink directly on the library's device tables (zero copy).
  The order-dependent admission (max_depth keep mask) is computed BEFORE sharding, and each rank's
  first-seen ordinals start at its chunk's global read index, so the result is bit-identical to a
  single-GPU run.

Nothing here is a data-path collective for the multi-sample workload; do not add one.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np

from .packing import ReadBatch, query_lengths

_BIAS = -2 ** 31      # int32 view of 0x80000000: makes unsigned order == signed order for the MIN reduce


def assign_samples(n_samples: int, world_size: int, rank: int) -> List[int]:
    """samples of a plate -> ranks, static round robin."""
    return list(range(rank, n_samples, world_size))


def shard_reads(batch: ReadBatch, world_size: int) -> List[Tuple[int, int]]:
    """[a, b) read ranges, contiguous in coordinate order and balanced by query bases."""
    n = batch.n_reads
    if n == 0:
        return [(0, 0)] * world_size
    lq = query_lengths(batch.cigar_off, batch.cigar).astype(np.int64)
    csum = np.cumsum(lq)
    total = int(csum[-1])
    cuts = [0]
    for r in range(1, world_size):
        cuts.append(int(np.searchsorted(csum, total * r / world_size, side="left")))
    cuts.append(n)
    cuts = [min(max(c, 0), n) for c in cuts]
    for i in range(1, len(cuts)):
        cuts[i] = max(cuts[i], cuts[i - 1])
    return [(cuts[i], cuts[i + 1]) for i in range(world_size)]


def key_union(local_keys: Sequence[int], group=None) -> List[int]:
    """union over ranks of the (allele group, quality) plane keys (1024-bit bitmap all-reduce)."""
    import torch
    import torch.distributed as dist
    bm = torch.zeros(1024, dtype=torch.int32)
    for k in local_keys:
        bm[int(k)] = 1
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else bm.device
        bm = bm.to(dev)
        dist.all_reduce(bm, op=dist.ReduceOp.MAX, group=group)
        bm = bm.cpu()
    return [int(k) for k in torch.nonzero(bm).flatten().tolist()]


def reduce_tables(tables: Dict[str, "object"], group=None) -> None:
    """In-place all-reduce of a dict of int32 tensors: names starting with 'first' take the unsigned
    MIN (through an order-preserving bias), everything else the integer SUM.  Works on CPU tensors with
    gloo (tests) and on CUDA tensors with NCCL (production)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    for name in sorted(tables):
        t = tables[name]
        assert t.dtype == torch.int32, name
        if name.startswith("first"):
            t.bitwise_xor_(torch.tensor(_BIAS, dtype=torch.int32, device=t.device))
            dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
            t.bitwise_xor_(torch.tensor(_BIAS, dtype=torch.int32, device=t.device))
        else:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)


class _DevArray:
    """exposes a raw device pointer of the library as a __cuda_array_interface__ object"""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i4", "data": (int(ptr), False), "version": 2}


def device_tables(handle) -> Dict[str, "object"]:
    """zero-copy int32 torch views of a handle's persistent device tables (after key agreement)."""
    import torch
    dev = torch.device("cuda", torch.cuda.current_device())
    G = handle.G
    out = {}
    for k in handle.plane_keys():
        out[f"plane{int(k):04d}"] = torch.as_tensor(_DevArray(handle.plane_devptr(int(k)), G * 4), device=dev)
    out["dels"] = torch.as_tensor(_DevArray(handle.dels_devptr(), G), device=dev)
    out["covdiff"] = torch.as_tensor(_DevArray(handle.covdiff_devptr(), G + 1), device=dev)
    for g in range(4):
        p = handle.first_devptr(g)
        if p:
            out[f"first{g}"] = torch.as_tensor(_DevArray(p, G * 4), device=dev)
    return out


def allreduce_handle(handle, group=None) -> int:
    """Make every rank's handle hold the tables of ALL ranks' reads (NCCL all-reduce over NVLink).
    Returns the number of bytes this rank contributed to the collective."""
    import torch
    import torch.distributed as dist
    handle.sync()
    for k in key_union([int(k) for k in handle.plane_keys()], group):
        handle.ensure_plane(k)               # a plane also brings its group's first-seen table
    tabs = device_tables(handle)
    torch.cuda.synchronize()
    reduce_tables(tabs, group)
    torch.cuda.synchronize()
    if dist.is_initialized():
        o = torch.tensor([handle.ordinal], dtype=torch.int64, device=next(iter(tabs.values())).device)
        dist.all_reduce(o, op=dist.ReduceOp.MAX, group=group)
        handle.ordinal = int(o.item())
    return sum(t.numel() * 4 for t in tabs.values())


def position_slice(G: int, world_size: int, rank: int) -> Tuple[int, int]:
    """contiguous slice of positions a rank owns / genotypes after the exchange: ceil((G+1)/n) rows per rank, the
    same rule as the library's lvc_position_slice (the coverage difference array has G + 1 entries)"""
    per = (G + 1 + world_size - 1) // world_size
    return min(rank * per, G), min((rank + 1) * per, G)


def process_batch_sharded(caller, batch: ReadBatch, group=None) -> int:
    """One sample, read-chunk sharding: `batch` is the WHOLE coordinate-sorted batch with
    its keep mask (identical on every rank); every rank deposits its contiguous chunk, the tables are
    all-reduced over NCCL, and each rank is left genotyping its own slice of positions.
    `caller` is a variant_caller.live_variant_caller.LiveVariantCaller.  Returns the bytes reduced."""
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    h = caller._handle
    base = h.ordinal
    if rank != 0:
        # after the previous all-reduce every rank holds the FULL tables; only rank 0 keeps that history,
        # the others contribute just this batch's delta (sum over ranks == history + all deltas)
        h.reset()
    a, b = shard_reads(batch, world)[rank]
    h.ordinal = base + a                      # first-seen ordinals are global read indices
    if b > a:
        h.push_batch(batch.slice(a, b).as_capi())
    h.ordinal = base + batch.n_reads
    n = allreduce_handle(h, group)
    p0, p1 = position_slice(h.G, world, rank)
    h.set_genotype_range(p0, p1)
    return n


# ------------------------------------------------------------------------------------------------
# halo-only exchange (the refinement of the full-table reduce)
# ------------------------------------------------------------------------------------------------
def ref_lengths(batch: ReadBatch) -> np.ndarray:
    """reference span per read = sum of the M/D/N/=/X op lengths"""
    n = batch.n_reads
    if n == 0:
        return np.zeros(0, dtype=np.int64)
    nc = int(batch.cigar_off[n])
    ops = batch.cigar[:nc] & 15
    w = np.where((ops == 0) | (ops == 2) | (ops == 3) | (ops == 7) | (ops == 8), (batch.cigar[:nc] >> 4).astype(np.int64), 0)
    csum = np.concatenate([[0], np.cumsum(w)])
    co = batch.cigar_off[:n + 1].astype(np.int64)
    return csum[co[1:]] - csum[co[:-1]]


def touched_ranges(batch: ReadBatch, shards: Sequence[Tuple[int, int]]) -> List[Tuple[int, int]]:
    """[lo, hi) columns each rank's chunk of reads can deposit into (a superset: filtered reads count too).
    Every rank computes the same list from the batch metadata, so the exchange plan needs no communication."""
    end = batch.pos[:batch.n_reads].astype(np.int64) + ref_lengths(batch)
    out = []
    for a, b in shards:
        out.append((int(batch.pos[a]), int(end[a:b].max())) if b > a else (0, 0))
    return out


def halo_plan(G: int, world_size: int, touched: Sequence[Tuple[int, int]]) -> Dict[Tuple[int, int], Tuple[int, int]]:
    """(src, dst) -> [lo, hi): the columns rank `src` deposited into that rank `dst` owns (position_slice)."""
    plan = {}
    for src, (lo, hi) in enumerate(touched):
        if hi <= lo:
            continue
        for dst in range(world_size):
            if dst == src:
                continue
            p0, p1 = position_slice(G, world_size, dst)
            a, b = max(lo, p0), min(hi, p1)
            if b > a:
                plan[(src, dst)] = (a, b)
    return plan


def _width(name: str) -> int:
    return 1 if name in ("dels", "covdiff") else 4


def _segment(name: str, a: int, b: int, G: int) -> Tuple[int, int]:
    """element range of table `name` for the columns [a, b).  The coverage difference array has G + 1 entries;
    entry i belongs to the owner of column min(i, G - 1)."""
    if name == "covdiff":
        return a, b + 1 if b == G else b
    w = _width(name)
    return a * w, b * w


def halo_exchange(tables: Dict[str, "object"], G: int, touched: Sequence[Tuple[int, int]], group=None) -> int:
    """Position-ownership exchange: rank r owns the columns position_slice(G, world, r) and keeps the history of
    those only.  After a rank deposited its chunk of reads, whatever it deposited into columns owned by another
    rank (normally a halo of one read span next to its own slice) is SENT to the owner (one message per pair),
    added there (unsigned MIN for the first-seen tables) and cleared locally.  Works on CPU tensors with gloo (tests)
    and on CUDA tensors with NCCL P2P over NVLink (production).  Returns the bytes this rank sent.

    The touched range is widened by one column so that the end marker of the coverage difference array (written
    at column end + 1) travels with it; cumulative coverage of a slice needs the prefix of the lower ranks' slices."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return 0
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    plan = halo_plan(G, world, [(lo, min(hi + 1, G)) if hi > lo else (0, 0) for lo, hi in touched])
    names = sorted(tables)
    bias = None
    ops, recvs, sends = [], [], []
    for (src, dst), (a, b) in sorted(plan.items()):
        segs = [_segment(nm, a, b, G) for nm in names]
        if src == rank:
            buf = torch.cat([tables[nm][x:y] for nm, (x, y) in zip(names, segs)])
            sends.append((buf, segs))
            ops.append(dist.P2POp(dist.isend, buf, dst, group))
        elif dst == rank:
            ref_t = tables[names[0]]
            buf = torch.empty(sum(y - x for x, y in segs), dtype=torch.int32, device=ref_t.device)
            recvs.append((buf, segs))
            ops.append(dist.P2POp(dist.irecv, buf, src, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    sent = 0
    for buf, segs in recvs:
        off = 0
        for nm, (x, y) in zip(names, segs):
            part = buf[off:off + (y - x)]
            off += y - x
            t = tables[nm]
            if nm.startswith("first"):
                if bias is None:
                    bias = torch.tensor(_BIAS, dtype=torch.int32, device=t.device)
                t[x:y] = torch.minimum(t[x:y] ^ bias, part ^ bias) ^ bias
            else:
                t[x:y] += part
    for buf, segs in sends:
        sent += buf.numel() * 4
        for nm, (x, y) in zip(names, segs):
            tables[nm][x:y] = -1 if nm.startswith("first") else 0
    return sent


def process_batch_halo(caller, batch: ReadBatch, group=None) -> int:
    """One sample, read-chunk sharding with POSITION OWNERSHIP (halo-only refinement): every rank deposits
    its contiguous chunk of the coordinate-sorted batch, then sends only the columns it touched outside its own
    slice of positions to their owners.  Each rank's tables hold the full history of ITS slice (and zeros elsewhere),
    which is the slice it genotypes.  Use either this or process_batch_sharded on a caller, not both.
    Returns the bytes this rank sent."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    h = caller._handle
    base = h.ordinal
    shards = shard_reads(batch, world)
    a, b = shards[rank]
    h.ordinal = base + a                      # first-seen ordinals are global read indices
    if b > a:
        h.push_batch(batch.slice(a, b).as_capi())
    h.ordinal = base + batch.n_reads
    sent = 0
    if world > 1:
        h.sync()
        for k in key_union([int(k) for k in h.plane_keys()], group):
            h.ensure_plane(k)
        tabs = device_tables(h)
        torch.cuda.synchronize()
        sent = halo_exchange(tabs, h.G, touched_ranges(batch, shards), group)
        torch.cuda.synchronize()
    p0, p1 = position_slice(h.G, world, rank)
    h.set_genotype_range(p0, p1)
    return sent


def make_library_comm(device: int, group=None):
    """an NCCL communicator owned by liblvc_b200.so (capi.NcclComm) spanning the ranks of the torch.distributed
    group: rank 0 draws the unique id, torch.distributed ships its 128 bytes"""
    import torch.distributed as dist
    from . import capi
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    box = [capi.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    return capi.NcclComm(device, world, rank, box[0])


def process_batch_library(caller, batch: ReadBatch, comm, mode: str = "scatter") -> int:
    """One sample, read-chunk sharding through the C-ABI exchange (lvc_reduce_tables): every rank deposits its
    contiguous chunk of the coordinate-sorted batch, then ONE grouped NCCL collective over all tables:
    mode "scatter": reduce-scatter in place onto position slices (each rank keeps the history of its slice and
    genotypes it); mode "all": all-reduce (every rank holds the complete tables; rank 0 keeps the history between
    batches, like process_batch_sharded).  Use one mode per caller.  Returns the bytes this rank fed in."""
    from . import capi
    h = caller._handle
    world, rank = comm.n_ranks, comm.rank
    base = h.ordinal
    if mode == "all" and rank != 0:
        h.reset()
    a, b = shard_reads(batch, world)[rank]
    h.ordinal = base + a                      # first-seen ordinals are global read indices
    if b > a:
        h.push_batch(batch.slice(a, b).as_capi())
    h.ordinal = base + batch.n_reads
    n = h.reduce_tables(comm, capi.REDUCE_SCATTER if mode == "scatter" else capi.REDUCE_ALL)
    if mode == "all":
        p0, p1 = h.position_slice(world, rank)
        h.set_genotype_range(p0, p1)
    return n


_NIBBLE_TO_GS = 0xFEDCBA9387625104          # lvc_common.cuh kNibbleToGS: BAM nibble -> (allele group << 2 | slot)


def batch_keys(batch: ReadBatch, min_base_quality: int) -> List[int]:
    """the (allele group << 8 | quality) plane keys a batch can deposit into (a superset: every stored base is looked at,
    soft clips and padding included) -- what peer_setup needs BEFORE the first deposit"""
    q = np.asarray(batch.qual, dtype=np.uint8)
    s = np.asarray(batch.seq4, dtype=np.uint8)
    n = min(len(q), 2 * len(s))
    nib = np.empty(2 * len(s), dtype=np.uint8)
    nib[0::2], nib[1::2] = s >> 4, s & 15
    group = np.array([((_NIBBLE_TO_GS >> (4 * k)) & 15) >> 2 for k in range(16)], dtype=np.uint16)[nib[:n]]
    keep = q[:n].astype(np.int32) >= int(min_base_quality)
    return sorted(int(k) for k in np.unique((group[keep] << 8) | q[:n][keep].astype(np.uint16)))


def peer_setup(caller, comm, keys=None, group=None) -> None:
    """Position ownership over NVLink peer memory (lvc_peer_attach): make the plane set identical on every rank (the
    union of the ranks' keys and of `keys`), exchange the CUDA IPC blobs of the tables over torch.distributed and map
    the other ranks' tables.  Call once per caller (and again if a batch brings a quality never seen before)."""
    import torch.distributed as dist
    h = caller._handle
    for k in key_union(sorted(set(int(k) for k in h.plane_keys()) | set(int(k) for k in (keys or []))), group):
        h.ensure_plane(k)
    h.sync()
    blobs = [None] * comm.n_ranks
    dist.all_gather_object(blobs, h.peer_export(), group=group)
    h.peer_attach(comm.rank, blobs)
    h.stream_barrier(comm)                      # nobody deposits before everybody has mapped everybody
    h.sync()


def process_batch_peer(caller, batch: ReadBatch, comm) -> None:
    """One sample, read-chunk sharding with position ownership and NO exchange step: every rank deposits its contiguous
    chunk of the coordinate-sorted batch; bases of columns another rank owns are reduced into that rank's tables over
    NVLink by the deposit kernel itself (peer_setup must have been called).  Two stream-ordered barriers frame the
    deposit; each rank then genotypes the slice it owns (gather_variants merges the records)."""
    h = caller._handle
    world, rank = comm.n_ranks, comm.rank
    base = h.ordinal
    a, b = shard_reads(batch, world)[rank]
    h.ordinal = base + a                      # first-seen ordinals are global read indices
    h.stream_barrier(comm)                    # the previous batch's genotype pass is complete on every rank
    if b > a:
        h.push_batch(batch.slice(a, b).as_capi())
    h.ordinal = base + batch.n_reads
    h.stream_barrier(comm)                    # every rank's reductions have landed


def gather_variants(caller, group=None) -> List[dict]:
    """records of every rank's position slice, merged in position order on every rank"""
    import torch.distributed as dist
    mine = caller.prepare_variants()
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return mine
    parts = [None] * dist.get_world_size(group)
    dist.all_gather_object(parts, mine, group=group)
    return [v for part in parts for v in part]

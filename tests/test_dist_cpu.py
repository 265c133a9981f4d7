"""CPU, world_size 2 over gloo: the host-side logic of the multi-GPU paths (sharding plans and the
table reduction), checked against the oracle.  The CUDA path itself is covered by tests/test_gpu_*."""
import os
import socket

import numpy as np
import pytest

from helpers import po, synth_small, rows_to_tuples


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, rows, ref, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "covid-spings-variant-caller_b200"), os.path.join(root, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    from lvc_b200 import packing, dist as ldist
    from oracle.c_oracle import COracle
    from helpers import rows_to_tuples as r2t
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    batch = packing.pack_reads(r2t(rows), 0)                 # keep mask computed on the WHOLE batch
    a, b = ldist.shard_reads(batch, world)[rank]
    mine = batch.slice(a, b)
    co = COracle(ref, 13, 0, max_depth=10 ** 9)              # admission already applied: replay the keep mask
    mine_kept = _apply_keep(mine)
    co.st.ordinal = a                                        # global read index of the chunk's first read
    co.process(mine_kept)
    first = co.first.astype(np.uint32).view(np.int32).reshape(-1).copy()
    tabs = {"plane_ad": torch.from_numpy(co.ad.astype(np.int32).reshape(-1).copy()),
            "dels": torch.from_numpy((co.depth.astype(np.int64) - co.ad.sum(axis=1)).astype(np.int32)),
            "covdiff": torch.from_numpy(np.diff(np.concatenate([[0], co.cov.astype(np.int64), [0]])).astype(np.int32)),
            "first0": torch.from_numpy(first)}
    keys = ldist.key_union([rank + 1, 7])
    ldist.reduce_tables(tabs)
    if rank == 0:
        q.put((keys, {k: v.numpy().copy() for k, v in tabs.items()}, ldist.assign_samples(7, world, 0),
               ldist.assign_samples(7, world, 1)))
    dist.barrier()
    dist.destroy_process_group()


def _apply_keep(b):
    """emulate 'admission happened before sharding': drop the reads whose keep bit is 0 by failing their mapq"""
    from lvc_b200 import packing
    mq = b.mapq.copy()
    fl = b.flag.copy()
    fl[(b.keep & 1) == 0] |= 0x400
    return packing.ReadBatch(b.pos, fl, mq, b.keep, b.cigar_off, b.cigar, b.seq_off, b.seq4, b.qual)


def test_read_chunk_sharding_reduces_to_the_single_process_tables(lib, golden_synth):
    import torch.multiprocessing as mp
    from lvc_b200 import packing
    from oracle.c_oracle import COracle
    g = golden_synth["amplicon_like"]
    rows, ref = g["reads"], g["ref"]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, rows, ref, q)) for r in range(2)]
    for p in procs:
        p.start()
    keys, tabs, s0, s1 = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert keys == [1, 2, 7]
    assert s0 == [0, 2, 4, 6] and s1 == [1, 3, 5]
    whole = packing.pack_reads(rows_to_tuples(rows), 0)
    co = COracle(ref, 13, 0)
    co.process(whole)
    assert np.array_equal(tabs["plane_ad"].reshape(-1, 16), co.ad.astype(np.int32))
    assert np.array_equal(tabs["dels"], (co.depth.astype(np.int64) - co.ad.sum(axis=1)).astype(np.int32))
    assert np.array_equal(np.cumsum(tabs["covdiff"])[:-1], co.cov.astype(np.int64))
    assert np.array_equal(tabs["first0"].view(np.uint32).reshape(-1, 16), co.first)


def test_shard_reads_balanced_and_contiguous(lib, golden_synth):
    from lvc_b200 import packing, dist as ldist
    b = packing.pack_reads(rows_to_tuples(golden_synth["mixed_small"]["reads"]), 0)
    for world in (1, 2, 3, 8):
        parts = ldist.shard_reads(b, world)
        assert parts[0][0] == 0 and parts[-1][1] == b.n_reads
        assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
        lq = packing.query_lengths(b.cigar_off, b.cigar)
        loads = [int(lq[a:c].sum()) for a, c in parts]
        assert max(loads) - min(loads) <= 2 * int(lq.max()) + 1
    sl = b.slice(5, 40)
    assert sl.n_reads == 35 and sl.pos.tolist() == b.pos[5:40].tolist()
    assert sl.qual[:sl.n_qual].tolist() == b.qual[int(b.seq_off[5]):int(b.seq_off[40])].tolist()


def _halo_worker(rank, world, port, rows, ref, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "covid-spings-variant-caller_b200"), os.path.join(root, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    from lvc_b200 import packing, dist as ldist
    from oracle.c_oracle import COracle
    from helpers import rows_to_tuples as r2t
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    G = len(ref)
    tabs = {"plane_ad": torch.zeros(G * 16, dtype=torch.int32), "dels": torch.zeros(G, dtype=torch.int32),
            "covdiff": torch.zeros(G + 1, dtype=torch.int32), "first0": torch.full((G * 16,), -1, dtype=torch.int32)}
    ldist_width = ldist._width
    ldist._width = lambda name: 16 if name in ("plane_ad", "first0") else ldist_width(name)   # 16 alleles per column here
    sent = 0
    ordinal = 0
    for k in range(3):                                        # three live batches: history stays with the owner
        sel = list(range(k, len(rows), 3))
        batch = packing.pack_reads(r2t([rows[i] for i in sel]), 0)
        shards = ldist.shard_reads(batch, world)
        a, b = shards[rank]
        co = COracle(ref, 13, 0, max_depth=10 ** 9)
        co.st.ordinal = ordinal + a
        co.process(_apply_keep(batch.slice(a, b)))
        ordinal += batch.n_reads
        # this batch's deposits of this rank, added to the local tables (first-seen: unsigned min)
        tabs["plane_ad"] += torch.from_numpy(co.ad.astype(np.int32).reshape(-1).copy())
        tabs["dels"] += torch.from_numpy((co.depth.astype(np.int64) - co.ad.sum(axis=1)).astype(np.int32))
        tabs["covdiff"] += torch.from_numpy(np.diff(np.concatenate([[0], co.cov.astype(np.int64), [0]])).astype(np.int32))
        f = torch.from_numpy(co.first.astype(np.uint32).view(np.int32).reshape(-1).copy())
        bias = torch.tensor(-2 ** 31, dtype=torch.int32)
        tabs["first0"] = torch.minimum(tabs["first0"] ^ bias, f ^ bias) ^ bias
        sent += ldist.halo_exchange(tabs, G, ldist.touched_ranges(batch, shards))
    q.put((rank, ldist.position_slice(G, world, rank), {k: v.numpy().copy() for k, v in tabs.items()}, sent))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_halo_exchange_leaves_every_owner_with_the_full_history_of_its_slice(lib, golden_synth, world):
    import torch.multiprocessing as mp
    from lvc_b200 import packing
    from oracle.c_oracle import COracle
    g = golden_synth["amplicon_like"]
    rows, ref = g["reads"], g["ref"]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_halo_worker, args=(r, world, port, rows, ref, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    co = COracle(ref, 13, 0)
    for k in range(3):
        co.process(packing.pack_reads(rows_to_tuples([rows[i] for i in range(k, len(rows), 3)]), 0))
    G = len(ref)
    want_cd = np.diff(np.concatenate([[0], co.cov.astype(np.int64), [0]])).astype(np.int32)
    total_cd = np.zeros(G + 1, dtype=np.int64)
    full_bytes = (16 + 1 + 16) * 4 * G
    for rank, (p0, p1), tabs, sent in got:
        ad = tabs["plane_ad"].reshape(-1, 16)
        assert np.array_equal(ad[p0:p1], co.ad.astype(np.int32)[p0:p1])
        assert not ad[:p0].any() and not ad[p1:].any()                  # nothing but the owned slice is kept
        dels = (co.depth.astype(np.int64) - co.ad.sum(axis=1)).astype(np.int32)
        assert np.array_equal(tabs["dels"][p0:p1], dels[p0:p1])
        first = tabs["first0"].view(np.uint32).reshape(-1, 16)
        assert np.array_equal(first[p0:p1], co.first[p0:p1])
        assert (first[:p0] == 0xFFFFFFFF).all() and (first[p1:] == 0xFFFFFFFF).all()
        total_cd += tabs["covdiff"]
        hi = p1 + 1 if p1 == G else p1
        assert not tabs["covdiff"][:p0].any() and not tabs["covdiff"][hi:].any()
        assert sent < full_bytes                                        # a halo, not the table
    assert np.array_equal(total_cd, want_cd)


def test_halo_plan_and_touched_ranges(lib, golden_synth):
    """pure host logic of the halo exchange: every touched column outside a rank's own slice is sent to exactly its
    owner, nothing inside the own slice is sent, and the plan is empty when chunks stay inside their slices"""
    from lvc_b200 import packing, dist as ldist
    b = packing.pack_reads(rows_to_tuples(golden_synth["amplicon_like"]["reads"]), 0)
    G = len(golden_synth["amplicon_like"]["ref"])
    for world in (2, 3, 5):
        shards = ldist.shard_reads(b, world)
        touched = ldist.touched_ranges(b, shards)
        rl = ldist.ref_lengths(b)
        for (a, c), (lo, hi) in zip(shards, touched):
            if c > a:
                assert lo == int(b.pos[a]) and hi == int((b.pos[a:c].astype(np.int64) + rl[a:c]).max())
        plan = ldist.halo_plan(G, world, touched)
        for src, (lo, hi) in enumerate(touched):
            cols = np.zeros(G, dtype=np.int32)
            for (s, d), (x, y) in plan.items():
                if s == src:
                    p0, p1 = ldist.position_slice(G, world, d)
                    assert d != src and p0 <= x < y <= p1
                    cols[x:y] += 1
            own0, own1 = ldist.position_slice(G, world, src)
            want = np.zeros(G, dtype=np.int32)
            want[lo:hi] = 1
            want[own0:own1] = 0
            assert np.array_equal(cols, want)
    # chunks that stay inside their own slices: nothing to exchange
    # slices of ceil((G + 1) / n) = 51 rows: [0, 51) and [51, 100)
    assert ldist.position_slice(100, 2, 0) == (0, 51) and ldist.position_slice(100, 2, 1) == (51, 100)
    assert ldist.halo_plan(100, 2, [(0, 51), (51, 100)]) == {}
    assert ldist.halo_plan(100, 2, [(0, 60), (51, 100)]) == {(0, 1): (51, 60)}


def test_ont_batch_fast_is_deterministic_and_well_formed(lib):
    from lvc_b200 import synth, packing
    ref = synth.random_reference(5000, 5)
    a = synth.ont_batch_fast(11, ref, depth=20.0)
    b = synth.ont_batch_fast(11, ref, depth=20.0)
    for f in ("pos", "flag", "mapq", "keep", "cigar_off", "cigar", "seq_off", "seq4", "qual"):
        assert np.array_equal(getattr(a, f), getattr(b, f)), f
    n = a.n_reads
    assert n == round(5000 * 20.0 / 400) and a.n_cigar == 21 * n
    assert (np.diff(a.pos[:n]) >= 0).all()
    lq = packing.query_lengths(a.cigar_off, a.cigar)
    assert (np.diff(a.seq_off.astype(np.int64)) == lq + (lq & 1)).all()
    from lvc_b200.dist import ref_lengths
    assert (ref_lengths(a) == 400).all()
    assert int(a.qual[:a.n_qual].max()) <= 90 and (a.keep[:n] & 2).all()


def test_batch_keys_cover_every_deposited_key(lib, golden_synth):
    """peer tables fix the plane set before the first deposit: dist.batch_keys must name every (allele group, quality)
    key the oracle deposits into (a superset is fine), and the owner of a column follows lvc_position_slice"""
    from lvc_b200 import packing, dist as ldist
    for scen in ("mixed_small", "ont_like", "amplicon_like"):
        g = golden_synth[scen]
        batch = packing.pack_reads(rows_to_tuples(g["reads"]), 0)
        for mbq in (0, 13, 30):
            keys = set(ldist.batch_keys(batch, mbq))
            oc = po.OracleCaller(g["ref"], mbq, 0, 1, 1, 0.0)
            oc.process_reads(synth_small.rows_to_reads(g["reads"]))
            nib = {c: k for k, c in enumerate("=ACMGRSVTWYHKDBN")}
            gs = [((0xFEDCBA9387625104 >> (4 * k)) & 15) for k in range(16)]
            for site in oc.memory.values():
                for allele, quals in site["snvs"].items():
                    for q in quals:
                        assert ((gs[nib[allele]] >> 2) << 8 | int(q)) in keys, (scen, mbq, allele, q)
    G, world = 1000, 3
    per = (G + 1 + world - 1) // world
    for r in range(world):
        assert ldist.position_slice(G, world, r) == (min(r * per, G), min((r + 1) * per, G))

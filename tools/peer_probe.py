"""2-GPU probe: deposit-kernel time with peer tables attached, (a) normal chunk, (b) chunk kept 400 columns inside the own slice"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "covid-spings-variant-caller_b200")):
    sys.path.insert(0, p)
import numpy as np, torch, torch.distributed as dist
import bench
from lvc_b200 import capi, records, dist as ldist
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank); dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
ref, batch = bench.config5_workload()
G = len(ref)
shards = ldist.shard_reads(batch, world)
a, b = shards[rank]
comm = ldist.make_library_comm(rank)
p0, p1 = ldist.position_slice(G, world, rank)
def run(tag, sl, attach):
    mine = batch.slice(*sl)
    db, keep = bench.to_device(torch, capi, mine, dev)
    h = capi.Handle(ref.encode("latin-1"), 30, 20, device=rank, stream=stream.cuda_stream)
    for k in ldist.key_union(ldist.batch_keys(mine, 30)): h.ensure_plane(k)
    h.sync()
    if attach:
        blobs = [None] * world
        dist.all_gather_object(blobs, h.peer_export())
        h.peer_attach(rank, blobs); h.stream_barrier(comm); h.sync()
    for it in range(3):
        h.ordinal = batch.n_reads * it + sl[0]
        h.stream_barrier(comm); h.push_batch_device_async(db); h.stream_barrier(comm)
    torch.cuda.synchronize(); dist.barrier()
    h.set_timing(True); h.get_timing(0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for it in range(5):
        h.stream_barrier(comm); h.push_batch_device_async(db); h.stream_barrier(comm)
    e1.record(stream); torch.cuda.synchronize()
    ms, n = h.get_timing(0); h.set_timing(False)
    print(f"rank {rank} {tag}: reads {mine.n_reads} kernel {ms / max(n, 1):.4f} ms x{n}, loop {e0.elapsed_time(e1) / 5:.4f} ms per step", flush=True)
    dist.barrier()
    if attach: h.peer_detach(); dist.barrier()
    h.close(); del keep
pos = np.asarray(batch.pos)
run("no attach", (a, b), False)
run("attached ", (a, b), True)
lo = int(np.searchsorted(pos, p0 + 400)); hi = int(np.searchsorted(pos, p1 - 600))
run("attached, interior only", (lo, hi), True)
comm.close(); dist.destroy_process_group()

#!/usr/bin/env python
"""Latency of the merge step: peer-window kernels (alpk_peer.cu) vs torch.distributed / NCCL, same payloads.
   torchrun --nproc-per-node N --master-addr 127.0.0.1 tools/time_peer.py   -> one JSON line on rank 0"""
import json, os, sys, time
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alplib_b200 import peer as P

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
pg = P.peer_group(None)
assert pg is not None


def timed(fn, reps):
    for _ in range(20):
        fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    host = (time.perf_counter() - t0) / reps * 1e6
    t = torch.tensor([e0.elapsed_time(e1) / reps * 1e3, host], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [round(float(v), 2) for v in t.cpu()]


out = {"world": world, "unit": "us per merge [device time, host wall incl. enqueue], max over ranks"}
x = torch.randn(52, dtype=torch.float64, device="cuda")
y = torch.empty_like(x)
buf = x.clone()
out["all_reduce_52_peer"] = timed(lambda: pg.all_reduce_sum([x], out=y), 2000)
out["all_reduce_52_nccl"] = timed(lambda: dist.all_reduce(buf), 2000)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for _ in range(10):
        pg.all_reduce_sum([x], out=y)
r = timed(g.replay, 200)
out["all_reduce_52_peer_graph_of_10"] = [round(v / 10, 2) for v in r]
n_ma, n_g = 200, 200
n_loc = (n_ma + world - 1) // world
loc = torch.randn(len(range(rank, n_ma, world)), n_g, dtype=torch.float64, device="cuda")
full = torch.empty(n_ma, n_g, dtype=torch.float64, device="cuda")
pad = torch.zeros(n_loc, n_g, dtype=torch.float64, device="cuda")
flat = torch.empty(world * n_loc, n_g, dtype=torch.float64, device="cuda")
out["all_gather_200x200_peer"] = timed(lambda: pg.all_gather_rows(loc, rank, world, n_ma, out=full), 1000)
out["all_gather_200x200_nccl"] = timed(lambda: dist.all_gather_into_tensor(flat, pad), 1000)
ep, st = pg.status()
out["status"] = st
if rank == 0:
    print(json.dumps(out))
P.reset()
dist.destroy_process_group()

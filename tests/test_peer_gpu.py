"""GPU tier, 2 ranks: the peer-window merge kernels (csrc/alpk_peer.cu, alplib_b200/peer.py) against plain arithmetic.

With >= 2 visible GPUs each rank owns one GPU (NCCL process group, windows mapped over NVLink); on a 1-GPU box both ranks
share cuda:0 (gloo process group for the handle exchange, ALPK_PEER=force): the windows are then mapped between two
processes on the same device and the kernels' push / signal / wait / combine logic is the same."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rank_data(rank, n):
    rng = np.random.RandomState(100 + rank)
    return rng.standard_normal(n) * 10.0 ** rng.randint(-20, 20, n)


def _run_rank(rank, world, port, backend, outdir):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), ALPK_PEER="force", ALPK_PEER_TIMEOUT_S="20")
    ngpu = torch.cuda.device_count()
    dev = rank % ngpu
    torch.cuda.set_device(dev)
    if backend == "nccl":
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", dev))
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from alplib_b200 import dist as D
        from alplib_b200 import peer as P
        pg = P.peer_group(None)
        assert pg is not None and pg.n_ranks == world and pg.rank == rank
        res = {}
        # all-reduce of several segments, repeated: both payload halves, the epoch logic, ranks drifting apart
        a = torch.from_numpy(_rank_data(rank, 1)).cuda()
        b = torch.from_numpy(_rank_data(rank + 10, 1)).cuda()
        h = torch.from_numpy(_rank_data(rank + 20, 50)).cuda()
        for it in range(64):
            out = pg.all_reduce_sum([a, b, h])
            if it == 7 and rank == 0:
                torch.cuda._sleep(20_000_000)            # rank 0 falls ~10 ms behind: the others wait in the kernel
        res["reduce"] = out.cpu().numpy()
        # a longer buffer (more elements than threads), one segment
        big = torch.from_numpy(_rank_data(rank + 30, 5000)).cuda()
        res["reduce_big"] = pg.all_reduce_sum([big]).cpu().numpy()
        # all-gather of rows dealt round-robin, uneven counts (5 rows over 2 ranks), then contiguous blocks
        full = np.arange(5 * 7, dtype=np.float64).reshape(5, 7) + 0.25
        loc = torch.from_numpy(np.ascontiguousarray(full[rank::world])).cuda()
        res["gather_rr"] = pg.all_gather_rows(loc, rank, world, 5).cpu().numpy()
        lo, hi = D.shard_range(5, rank, world)
        loc = torch.from_numpy(np.ascontiguousarray(full[lo:hi])).cuda()
        res["gather_blocks"] = pg.all_gather_rows(loc, lo, 1, 5).cpu().numpy()
        # the merge inside a captured CUDA graph, replayed: the epoch advances on the device
        x = torch.from_numpy(_rank_data(rank + 40, 52)).cuda()
        y = torch.empty(52, dtype=torch.float64, device="cuda")
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            pg.all_reduce_sum([x], out=y)                 # warm-up on the side stream
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            pg.all_reduce_sum([x], out=y)
        outs = []
        for k in range(5):
            x.mul_(2.0)
            g.replay()
            outs.append(y.clone())
        res["graph"] = torch.stack(outs).cpu().numpy()
        # the high-level entry point routes through the windows
        rates, hists = D.merged_rates_and_hists([a, b], [h])
        res["merged_rates"] = rates.cpu().numpy()
        res["merged_hist"] = hists[0].cpu().numpy()
        epoch, status = pg.status()
        res["epoch"], res["status"] = epoch, status
        np.savez(os.path.join(outdir, f"peer{rank}.npz"), **res)
        P.reset()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_peer_window_merges(tmp_path):
    import torch
    import torch.multiprocessing as mp
    assert torch.cuda.is_available()
    world = 2
    backend = "nccl" if torch.cuda.device_count() >= world else "gloo"
    ctx = mp.get_context("spawn")
    port = _free_port()
    procs = [ctx.Process(target=_run_rank, args=(r, world, port, backend, str(tmp_path))) for r in range(world)]
    for p in procs:
        p.start()
    try:
        for p in procs:
            p.join(timeout=280)
            assert p.exitcode == 0
    finally:                                    # never leave a rank behind (it would keep the GPU busy and block interpreter exit)
        for p in procs:
            if p.is_alive():
                p.kill()
    z = [np.load(os.path.join(str(tmp_path), f"peer{r}.npz")) for r in range(world)]
    # fixed-order sum: rank 0's values first -- exactly the left-to-right NumPy sum, identical on both ranks
    cat = [np.concatenate([_rank_data(r, 1), _rank_data(r + 10, 1), _rank_data(r + 20, 50)]) for r in range(world)]
    for r in range(world):
        assert np.array_equal(z[r]["reduce"], cat[0] + cat[1])
        assert np.array_equal(z[r]["reduce_big"], _rank_data(30, 5000) + _rank_data(31, 5000))
        full = np.arange(35, dtype=np.float64).reshape(5, 7) + 0.25
        assert np.array_equal(z[r]["gather_rr"], full) and np.array_equal(z[r]["gather_blocks"], full)
        for k in range(5):
            assert np.array_equal(z[r]["graph"][k], (_rank_data(40, 52) + _rank_data(41, 52)) * 2.0 ** (k + 1))
        assert np.array_equal(z[r]["merged_rates"], (cat[0] + cat[1])[:2]) and np.array_equal(z[r]["merged_hist"], (cat[0] + cat[1])[2:])
        assert int(z[r]["status"]) == 0 and int(z[r]["epoch"]) == 64 + 1 + 2 + 1 + 5 + 1    # (the capture itself runs nothing)

"""GPU tier, 2 ranks: a row-sharded SINGLE job on real hardware (reference loop fluxes.py:226-233).

Each rank takes its block of KEPT photon rows (dist.shard_flux), runs the one-kernel chain with the fused event-weighted
histogram and merges [rates || histogram] with ONE all-reduce (dist.merged_rates_and_hists).  The concatenation of the
ranks' per-sample arrays must be bit-identical to the single-process arrays and the merged rate / spectrum equal to 1e-12.
With >= 2 visible GPUs the ranks use one GPU each and NCCL; on a 1-GPU box both ranks share cuda:0 and the small merge
buffer goes through gloo (the kernels and the sharding logic are the same)."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GEOM = dict(det_dist=4.0, det_length=0.2, det_area=0.04)
EDGES = np.logspace(np.log10(1.5), 3, 41)
KW = dict(axion_mass=1.5, axion_coupling=1e-7, n_samples=2000, seed=20261020)


def _pf(n_bins=301):
    e = np.logspace(np.log10(0.5), 3, n_bins)
    return np.stack([e, 1e16 * e ** -1.5 * np.exp(-e / 150)], axis=1)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run_rank(rank, world, port, backend, outdir):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    ngpu = torch.cuda.device_count()
    dev = rank % ngpu
    torch.cuda.set_device(dev)
    if backend == "nccl":
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", dev))
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import alplib_b200 as A
        from alplib_b200 import dist as D
        W, Ar = A.Material("W"), A.Material("Ar")
        comp = D.shard_flux(A.FluxComptonIsotropic(_pf(), W, **KW, **GEOM))
        comp.simulate()
        rate_c, evw = comp.chain(100, 5.0, materialize=True, hist_edges=EDGES)
        prim = D.shard_flux(A.FluxPrimakoffIsotropic(_pf(), W, axion_mass=0.1, axion_coupling=1e-5, **GEOM))
        prim.simulate(); prim.propagate()
        rate_p = A.PhotonEventGenerator(prim, Ar).decays_device(100, 5.0)
        hist = comp.last_hist
        rates, hists = D.merged_rates_and_hists([rate_c, rate_p], [hist])     # (gloo: staged through the host by dist.py)
        np.savez(os.path.join(outdir, f"rank{rank}.npz"), E=comp.axion_energy, F=comp.axion_flux, D=comp.decay_axion_weight,
                 S=comp.scatter_axion_weight, evw=evw.cpu().numpy(), rates=rates.cpu().numpy(), hist=hists[0].cpu().numpy(),
                 pE=prim.axion_energy, pD=prim.decay_axion_weight, n_rows=comp._rows[1], row_offset=comp.row_offset)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_single_job_equals_one_rank(tmp_path):
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
    import alplib_b200 as A
    W, Ar = A.Material("W"), A.Material("Ar")
    one = A.FluxComptonIsotropic(_pf(), W, **KW, **GEOM)
    one.simulate()
    rate1, evw1 = one.chain(100, 5.0, materialize=True, hist_edges=EDGES)
    hist1 = one.last_hist.cpu().numpy()
    prim = A.FluxPrimakoffIsotropic(_pf(), W, axion_mass=0.1, axion_coupling=1e-5, **GEOM)
    prim.simulate(); prim.propagate()
    rp1 = A.PhotonEventGenerator(prim, Ar).decays(100, 5.0)
    parts = [np.load(os.path.join(str(tmp_path), f"rank{r}.npz")) for r in range(world)]
    # balanced: the kept rows are split evenly although the first ~20 % of the table is under the Compton threshold
    nr = [int(p["n_rows"]) for p in parts]
    assert sum(nr) == one._rows[1] and abs(nr[0] - nr[1]) <= 1 and int(parts[1]["row_offset"]) > 0
    for key, ref in (("E", one.axion_energy), ("F", one.axion_flux), ("D", one.decay_axion_weight),
                     ("S", one.scatter_axion_weight), ("evw", evw1.cpu().numpy()), ("pE", prim.axion_energy),
                     ("pD", prim.decay_axion_weight)):
        cat = np.concatenate([p[key] for p in parts])
        assert cat.shape == np.asarray(ref).shape and np.array_equal(cat, ref), key
    for p in parts:                                               # every rank holds the same merged result
        assert p["rates"][0] == pytest.approx(float(rate1.item()), rel=1e-12)
        assert p["rates"][1] == pytest.approx(float(rp1), rel=1e-12)
        assert np.allclose(p["hist"], hist1, rtol=1e-12, atol=0.0)
    assert np.array_equal(parts[0]["rates"], parts[1]["rates"]) and np.array_equal(parts[0]["hist"], parts[1]["hist"])
    assert hist1.sum() == pytest.approx(float(rate1.item()), rel=1e-9)          # edges span every E_a above threshold


def _run_example_rank(rank, world, port, outdir):
    import runpy
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ns = runpy.run_path(os.path.join(root, "examples", "ex_beam_dump_scan.py"), run_name="scan_example")
    compton, primakoff, limits = ns["main"]()
    np.savez(os.path.join(outdir, f"scan{rank}.npz"), compton=compton, primakoff=primakoff)


@pytest.mark.timeout(300)
def test_scan_example_distributed_equals_single_process(tmp_path):
    """examples/ex_beam_dump_scan.py: its torchrun branch under 2 ranks (masses dealt round-robin, one all-gather per map) gives
    the maps of the single-process run bit for bit, and the limit helpers run on them."""
    import runpy
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    port = _free_port()
    procs = [ctx.Process(target=_run_example_rank, args=(r, 2, port, str(tmp_path))) for r in range(2)]
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
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        os.environ.pop(k, None)
    ns = runpy.run_path(os.path.join(root, "examples", "ex_beam_dump_scan.py"), run_name="scan_example")
    compton, primakoff, limits = ns["main"]()
    assert compton.shape == (200, 200) and np.count_nonzero(compton) > 1000 and np.count_nonzero(primakoff) > 1000
    for r in range(2):
        z = np.load(os.path.join(str(tmp_path), f"scan{r}.npz"))
        assert np.array_equal(z["compton"], compton) and np.array_equal(z["primakoff"], primakoff)
    m, lim = limits["photon coupling (Primakoff)"]
    assert m.shape == lim.shape and m.shape[0] > 10 and m[0] == 0.0


def _run_sharded_example_rank(rank, world, port, outdir):
    import runpy
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ns = runpy.run_path(os.path.join(root, "examples", "ex_sharded_beam_dump.py"), run_name="sharded_example")
    rc, rp, spec = ns["main"]()
    np.savez(os.path.join(outdir, f"sharded{rank}.npz"), rc=rc, rp=rp, spec=spec)


@pytest.mark.timeout(300)
def test_sharded_example_equals_single_process(tmp_path):
    """examples/ex_sharded_beam_dump.py under 2 ranks == the single-process run (rates to 1e-12, spectrum to 1e-10), and both
    ranks hold the same merged bits."""
    import runpy
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    port = _free_port()
    procs = [ctx.Process(target=_run_sharded_example_rank, args=(r, 2, port, str(tmp_path))) for r in range(2)]
    for p in procs:
        p.start()
    try:
        for p in procs:
            p.join(timeout=280)
            assert p.exitcode == 0
    finally:
        for p in procs:
            if p.is_alive():
                p.kill()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        os.environ.pop(k, None)
    ns = runpy.run_path(os.path.join(root, "examples", "ex_sharded_beam_dump.py"), run_name="sharded_example")
    rc, rp, spec = ns["main"]()
    z = [np.load(os.path.join(str(tmp_path), f"sharded{r}.npz")) for r in range(2)]
    assert rc > 0 and rp > 0 and spec.sum() == pytest.approx(rc, rel=1e-9)
    for r in range(2):
        assert float(z[r]["rc"]) == pytest.approx(rc, rel=1e-12) and float(z[r]["rp"]) == pytest.approx(rp, rel=1e-12)
        assert np.allclose(z[r]["spec"], spec, rtol=1e-10, atol=0.0)
    assert np.array_equal(z[0]["spec"], z[1]["spec"]) and float(z[0]["rc"]) == float(z[1]["rc"])

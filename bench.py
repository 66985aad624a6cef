#!/usr/bin/env python
"""
bench.py -- headline benchmark of the alplib Monte-Carlo flux-and-event path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # CUDA arm (one process per GPU under torchrun)
    python bench.py --impl reference [--gpus N --steps K --warmup W]  # CPU arm: the reference algorithm (oracle port)

Workload (BASELINE.json configs[1], "ex2_beam_dump", synthesised per SURVEY.md section 8d because the reference's
example file is empty): photon spectrum of 1e4 log-spaced bins (0.5 MeV..1 GeV, Phi = 1e16 E^-1.5 exp(-E/150) /s) on a
W target; Primakoff production (ma = 0.1 MeV, g = 1e-5 /MeV, one sample per bin) and Compton production (ma = 1.5 MeV,
g = 1e-7, n_samples = 1e4 per bin); propagate to a 4 m / 0.2 m / 0.04 m^2 detector; decays(100 days, 5 MeV).

A "pass" is one run of that chain over the whole spectrum; a "sample" is one weighted ALP sample taken through production
+ propagate + decay weighting.  One pass of the fused kernel takes ~0.45 ms, so a "step" is PASSES_PER_STEP = 128
back-to-back passes: `--steps 20` times 2560 passes, a sustained region of > 1 s (clocks, power and throttle reasons are
sampled over it).  With N GPUs every rank runs the full spectrum as an independent replica with its own Philox row keys
(weak scaling: N x the Monte-Carlo statistics of the same job) and the per-pass rates are merged by one peer-window kernel per
rank over NVLink (alplib_b200.peer; an NCCL all-reduce when the ranks do not share a node), joined one pass later.

`value`: samples/s with the flux tables resident in HBM, every per-sample output (E, flux, decay_w, scatter_w, event
weight: 32 B/sample) written to HBM by the fused kernel.  `e2e`: the same chain through the public drop-in classes from
HOST arrays (H2D of the flux tables and D2H of the rates inside the timed region, every pass).  Extra keys: `e2e_eager`,
`e2e_fused` (other schedules), `e2e_pipelined` (rates read one pass late through decays_async()), `e2e_host_arrays` (per-sample arrays copied to NumPy), `c3` (BASELINE config 3: 1e8 decay
events), `c4` (config 4: e+- channels on graphite), `scan` / `scan_large` (config 5 and a 100x larger scan), `strong`
(one row-sharded job with the merged [rates || histogram] all-reduce, host-driven and as a CUDA graph), `roofline`,
`cpu_baseline`.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "weighted ALP MC samples/sec (prod+propagate+decay)"
UNIT = "samples/s"
N_BINS = 10_000
N_SAMPLES = 10_000
PASSES_PER_STEP = 128          # passes of the C2 chain per timed step (see module docstring)
E2E_PASSES_PER_STEP = 64
MA_P, G_P = 0.1, 1e-5
MA_C, G_C = 1.5, 1e-7
GEOM = dict(det_dist=4.0, det_length=0.2, det_area=0.04)
DUNE = dict(det_dist=574.0, det_length=5.0, det_area=21.0)
DAYS, THR = 100, 5.0
SEED = 0xA1B2C3D4 + 2
BYTES_PER_SAMPLE = 32          # E f64 + flux f64 + decay_w f32 + scatter_w f32 + event weight f64
FLOPS_PER_SAMPLE = 180         # expanded FP64 flops of the Compton chain (SURVEY section 8d accounting)
BYTES_PER_EVENT = 72           # decay MC: 2 x 4 x f64 + f64 weight
HIST_EDGES = np.logspace(np.log10(1.5), 3, 51)      # event-weighted E_a spectrum of the sharded job: 50 log bins


def c2_spectrum(n_bins_total, lo=0.5, hi=1e3):
    e = np.logspace(np.log10(lo), np.log10(hi), n_bins_total)
    return np.stack([e, 1e16 * e ** -1.5 * np.exp(-e / 150)], axis=1)


def workload_config(world, n_samples=N_SAMPLES):
    """Identical for the CUDA arm and the reference arm (same workload; the arms differ in how many passes make a step)."""
    return {
        "workload": "C2 ex2_beam_dump (synthetic): 1e4 photon bins/GPU x 1e4 Compton samples + 1e4 Primakoff samples, "
                    "W target, propagate 4m/0.2m/0.04m2, decays(100 d, 5 MeV)",
        "n_bins_per_gpu": N_BINS, "n_samples": n_samples, "ma_compton": MA_C, "g_compton": G_C,
        "ma_primakoff": MA_P, "g_primakoff": G_P,
        "sharding": f"{world} replica(s) of the spectrum, disjoint Philox row keys, flux/{world} per bin",
        "outputs": "materialised (32 B/sample)",
        "l2": "no flush needed: each pass writes 2.4 GB of outputs (>> 126 MB L2) and re-reads none of it",
    }


# -----------------------------------------------------------------------------------------------------------------
# clocks
# -----------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append((time.perf_counter(), ln.strip()))

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        rows = []
        for ts, ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                rows.append((ts, float(f[1]), float(f[2]), float(f[3]), f[4:8]))
            except ValueError:
                continue
        sel = [r for r in rows if (t0 is None or r[0] >= t0) and (t1 is None or r[0] <= t1)] or rows
        if not sel:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in sel for i in range(4) if r[4][i].lower().startswith("active")})
        return {"sm_mhz": float(np.median([r[1] for r in sel])), "sm_mhz_min": float(min(r[1] for r in sel)),
                "sm_max_mhz": max(r[2] for r in sel), "power_w_median": float(np.median([r[3] for r in sel])),
                "power_w_max": float(max(r[3] for r in sel)), "reasons": reasons, "samples": len(sel),
                "window_s": (t1 - t0) if (t0 is not None and t1 is not None) else None}


# -----------------------------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm (NumPy oracle port, per-bin loops with n_samples-long vector ops like the
# reference's simulate_single), fanned out over host cores by contiguous bin ranges.
# -----------------------------------------------------------------------------------------------------------------
def _cpu_chunk(args):
    pf, n_samples, seed = args
    from oracle import alp_oracle as O
    W = O.AbsCrossSection(O.Material("W"))
    ga = O.geom_acceptance(GEOM["det_area"], GEOM["det_dist"])
    n = 0
    # Primakoff: simulate + propagate + decays
    E, F = O.primakoff_simulate(pf, MA_P, G_P, 74, W)
    p = O.propagate(E, F, MA_P, O.W_gg(G_P, MA_P), 1.0, GEOM["det_dist"], GEOM["det_length"], geom_accept=ga)
    _, r1 = O.decays(E, p["decay_w"], DAYS, THR)
    n += E.shape[0]
    # Compton: simulate + propagate + decays, in row blocks to bound memory
    r2 = 0.0
    us = O.UniformSource(np.random.RandomState(seed))
    blk = max(1, min(500, 5_000_000 // max(n_samples, 1)))
    for i in range(0, pf.shape[0], blk):
        E, F = O.compton_simulate(pf[i:i + blk], MA_C, G_C, 74, W, n_samples, us)
        us.log.clear()
        p = O.propagate(E, F, MA_C, O.W_ee(G_C, MA_C), 1.0, GEOM["det_dist"], GEOM["det_length"], geom_accept=ga)
        _, r = O.decays(E, p["decay_w"], DAYS, THR)
        r2 += float(r)
        n += E.shape[0]
    return n, float(r1), r2


def cpu_pass(pf, n_samples, cores, pool=None):
    """One pass of the chain on the host; returns (samples, seconds, rates)."""
    # many more chunks than cores: rows under the Compton threshold are cheap, so equal row ranges are unequal work
    n_chunks = cores if cores == 1 else cores * 8
    chunks = np.array_split(np.arange(pf.shape[0]), n_chunks)
    jobs = [(pf[c], n_samples, 1000 + k) for k, c in enumerate(chunks) if c.size]
    t0 = time.perf_counter()
    res = pool.map(_cpu_chunk, jobs, chunksize=1) if pool is not None else [_cpu_chunk(j) for j in jobs]
    dt = time.perf_counter() - t0
    return sum(r[0] for r in res), dt, (sum(r[1] for r in res), sum(r[2] for r in res))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    n_samples = N_SAMPLES                           # the CUDA arm's configuration; one pass per step (bounded sample of a
    pf = c2_spectrum(N_BINS)                        # CUDA-arm step, which is PASSES_PER_STEP such passes)
    pf[:, 1] /= max(args.gpus, 1)
    with mp.get_context("fork").Pool(cores) as pool:
        _, dt, _ = cpu_pass(pf, 1000, cores, pool)              # probe at 1/10 size
        budget = 240.0                              # keep the whole run within a few minutes
        est = 10.0 * dt * (args.steps + args.warmup)
        if est > budget:
            n_samples = int(max(20, N_SAMPLES * budget / est))
        for _ in range(args.warmup):
            cpu_pass(pf, n_samples, cores, pool)
        n_tot, t_tot = 0, 0.0
        for _ in range(args.steps):
            n, dt, rates = cpu_pass(pf, n_samples, cores, pool)
            n_tot += n
            t_tot += dt
    v = n_tot / t_tot
    sample = (f"one pass of C2 per step (the CUDA arm's step is {PASSES_PER_STEP} passes), n_samples={n_samples} per bin, all "
              f"{N_BINS} bins: {n_tot // max(args.steps, 1)} samples/step")
    out = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / max(args.steps, 1), "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": workload_config(args.gpus, n_samples), "passes_per_step": 1, "rates": list(rates),
           "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                            "note": "NumPy restatement of the reference algorithm (oracle/alp_oracle.py, pinned "
                                    "bit-for-bit against the reference); the reference itself is Python and cannot "
                                    "travel to the GPU box.  The port keeps the reference's per-bin vector loops but "
                                    "collects arrays instead of per-sample list.append, so it is ~3-5x FASTER than the "
                                    "unmodified reference (1.6e6 samples/s/core, BASELINE.md section 2)."},
           "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


# -----------------------------------------------------------------------------------------------------------------
# CUDA arm
# -----------------------------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the CUDA arm has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import alplib_b200 as A
    from alplib_b200 import _native as NV
    from alplib_b200 import dist as D
    from alplib_b200.runtime import Runtime
    rt = Runtime.get(local)
    W, Ar = A.Material("W"), A.Material("Ar")
    legs = set(args.legs.split(","))
    want = (lambda name: "all" in legs or name in legs)
    t_start = time.perf_counter()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=rt.device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # weak scaling: every rank takes the full 1e4-bin spectrum as an independent replica (its own Philox keys through
    # row_offset, flux/world per bin), i.e. N GPUs = N x the Monte-Carlo statistics of the same beam-dump job; the
    # merged rate (sum over ranks) is the same physical answer.  The `strong` leg below shards ONE job instead.
    pf = c2_spectrum(N_BINS)
    pf[:, 1] /= world
    prim = A.FluxPrimakoffIsotropic(pf, W, axion_mass=MA_P, axion_coupling=G_P, **GEOM)
    comp = A.FluxComptonIsotropic(pf, W, axion_mass=MA_C, axion_coupling=G_C, n_samples=N_SAMPLES, seed=SEED, **GEOM)
    comp.row_offset = rank * N_BINS
    prim._stage_inputs()
    comp._stage_inputs()                       # flux tables now resident in HBM
    genP = A.PhotonEventGenerator(prim, Ar)
    n_pass = prim._inputs[2] + comp._inputs[2] * N_SAMPLES          # samples per pass, identical on every rank
    n_compton = comp._inputs[2] * N_SAMPLES
    rates = torch.zeros(2, dtype=torch.float64, device=rt.device)
    rate_bufs = [torch.zeros(2, dtype=torch.float64, device=rt.device) for _ in range(2)]   # double-buffered for the async merge
    pending = []
    pass_no = [0]
    pg = None
    if world > 1:
        from alplib_b200.peer import peer_group
        pg = peer_group(None, rates)
        merge_stream = torch.cuda.Stream()
        ev_ready = [torch.cuda.Event() for _ in range(2)]
        ev_merged = [torch.cuda.Event() for _ in range(2)]
    n_kev = min(max(args.steps, 1), 16)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_kev)]

    def one_pass(kpair=None):
        r1, _ = prim.chain(DAYS, THR, materialize=True)    # Primakoff: production -> propagate -> decays in one kernel
        comp._launch_rows()
        r2, _ = comp.chain(DAYS, THR, materialize=True, events=kpair)   # events recorded right around the kernel launch
        if world == 1:
            rates[0:1].copy_(r1)
            rates[1:2].copy_(r2)
            return
        # the path's only exchange: the merged rates of every pass.  Ranks on one node: ONE libalpk kernel per rank over NVLink
        # peer memory (alpk_peer_all_reduce_sum), issued on a side stream right behind the chain kernel and joined one pass
        # later, so its NVLink round trip overlaps the next pass's kernels; otherwise an asynchronous NCCL all-reduce
        k = pass_no[0] & 1
        pass_no[0] += 1
        if pg is not None:
            cur = torch.cuda.current_stream()
            ev_ready[k].record(cur)
            merge_stream.wait_event(ev_ready[k])
            with torch.cuda.stream(merge_stream):
                pg.all_reduce_sum([r1, r2], out=rate_bufs[k])
                ev_merged[k].record(merge_stream)
            pending.append((ev_merged[k], rate_bufs[k], (r1, r2)))        # (the inputs stay referenced until the join)
        else:
            buf = rate_bufs[k]
            buf[0:1].copy_(r1)
            buf[1:2].copy_(r2)
            pending.append((dist.all_reduce(buf, async_op=True), buf, None))
        if len(pending) > 1:
            join(*pending.pop(0))

    def join(w, b, _keep):
        if pg is not None:
            torch.cuda.current_stream().wait_event(w)      # stream-side join: the compute stream waits, the host does not
        else:
            w.wait()
        rates.copy_(b)

    def drain():
        while pending:
            join(*pending.pop(0))

    def step(kpair=None):
        for p in range(PASSES_PER_STEP):
            one_pass(kpair if p == PASSES_PER_STEP // 2 else None)

    warm = max(args.warmup, 3)
    for _ in range(warm):
        step()
    drain()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.15)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = NV.launch_count()
    t_host0 = time.perf_counter()
    ev0.record()
    for k in range(args.steps):
        step(kev[k] if k < n_kev else None)    # per-launch duration of the dominant kernel: CUDA events on its stream
    drain()                                    # the last pass's merge is inside the timed region
    ev1.record()
    barrier()
    t_host1 = time.perf_counter()
    clocks = sampler.stop(t_host0, t_host1) if rank == 0 else None
    kernel_ms = [a.elapsed_time(b) for a, b in kev[:min(args.steps, n_kev)]]
    launches = NV.launch_count() - l0
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    final_rates = [float(x) for x in rates.cpu()]
    value = world * n_pass * PASSES_PER_STEP * args.steps / (ms_total * 1e-3)

    # ---- end to end through the public drop-in API, host arrays in, rates out ------------------------------------
    def e2e_pass(fused, strict_host_arrays=False):
        # Compton first: its kernels are asynchronous, so the host-side set-up of the (small) Primakoff flux overlaps them;
        # each decays() returns a host float (the D2H read of the pass's result)
        c = A.FluxComptonIsotropic(pf, W, axion_mass=MA_C, axion_coupling=G_C, n_samples=N_SAMPLES, seed=SEED,
                                   fused=fused, **GEOM)
        c.row_offset = rank * N_BINS
        c.simulate(); c.propagate()
        gen_c = A.ElectronEventGenerator(c, Ar)
        rate_c = gen_c.decays_device(DAYS, THR)          # kernels queued, no sync yet (fused: the one chain kernel)
        p = A.FluxPrimakoffIsotropic(pf, W, axion_mass=MA_P, axion_coupling=G_P, fused=fused, **GEOM)
        p.simulate(); p.propagate()
        rp = A.PhotonEventGenerator(p, Ar).decays(DAYS, THR)
        rc = float(rate_c.item())
        return rp, rc, p, c

    e2e = {}
    e2e_steps = max(3, min(args.steps, 20))
    api_txt = {None: " (default schedule: deferred, ONE chain kernel that also materialises axion_energy, axion_flux, "
                     "decay/scatter weights in HBM: 24 B/sample; per-sample event weights are not requested by decays_device())",
               False: " with fused=False (one kernel per call, every per-sample array materialised in HBM)",
               True: " with fused=True (deferred, one reduce-only kernel)"}
    for tag, fused in (("e2e", None), ("e2e_eager", False), ("e2e_fused", True)):
        if tag != "e2e" and not want("e2e_variants"):
            continue
        # warm-up: the first eager passes pay cudaMalloc; the first deferred passes after a materialising phase show
        # allocator hiccups before settling
        for _ in range(E2E_PASSES_PER_STEP if fused is not False else 16):
            res = e2e_pass(fused)
        del res
        barrier()
        t0 = time.perf_counter()
        step_ms = []
        for _ in range(e2e_steps):
            ts = time.perf_counter()
            for _p in range(E2E_PASSES_PER_STEP):
                rp, rc, p, c = e2e_pass(fused)            # (each pass ends with the D2H read of its rates)
            step_ms.append((time.perf_counter() - ts) * 1e3)
        barrier()
        dt = max_over_ranks(time.perf_counter() - t0)
        h2d = sum(int(t.numel() * t.element_size()) for t in (p._inputs[0], p._inputs[1], c._inputs[0], c._inputs[1], c._inputs[4]))
        e2e[tag] = {"value": world * n_pass * E2E_PASSES_PER_STEP * e2e_steps / dt, "unit": UNIT,
                    "h2d_bytes_per_step": h2d * E2E_PASSES_PER_STEP, "d2h_bytes_per_step": 16 * E2E_PASSES_PER_STEP,
                    "steps": e2e_steps, "passes_per_step": E2E_PASSES_PER_STEP, "timed_region_s": dt,
                    "step_ms": {"median": float(np.median(step_ms)), "max": float(np.max(step_ms)), "min": float(np.min(step_ms))},
                    "api": "Flux*Isotropic.simulate()/propagate() + EventGenerator.decays()/decays_device()" + api_txt[fused],
                    "rates": [float(rp), float(rc)]}
        del p, c
    extras = {}

    # ---- the same passes with every rate read ONE PASS LATE (EventGenerator.decays_async): host set-up of pass k+1 overlaps the
    # kernels of pass k; every pass still uploads its host tables and brings its two rates back to the host
    if want("e2e_variants"):
        def pipelined_pass(prev):
            c = A.FluxComptonIsotropic(pf, W, axion_mass=MA_C, axion_coupling=G_C, n_samples=N_SAMPLES, seed=SEED, **GEOM)
            c.row_offset = rank * N_BINS
            c.simulate(); c.propagate()
            pc = A.ElectronEventGenerator(c, Ar).decays_async(DAYS, THR)
            p = A.FluxPrimakoffIsotropic(pf, W, axion_mass=MA_P, axion_coupling=G_P, **GEOM)
            p.simulate(); p.propagate()
            pp = A.PhotonEventGenerator(p, Ar).decays_async(DAYS, THR)
            got = (prev[0].result(), prev[1].result()) if prev is not None else None
            return (pp, pc), got
        try:
            prev = None
            for _ in range(E2E_PASSES_PER_STEP):
                prev, got = pipelined_pass(prev)
            prev[0].result(); prev[1].result()
            barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                prev = None
                for _p in range(E2E_PASSES_PER_STEP):
                    prev, got = pipelined_pass(prev)
                got = (prev[0].result(), prev[1].result())          # the step ends with its last rates on the host
            barrier()
            dt = max_over_ranks(time.perf_counter() - t0)
            extras["e2e_pipelined"] = {"value": world * n_pass * E2E_PASSES_PER_STEP * e2e_steps / dt, "unit": UNIT,
                                       "h2d_bytes_per_step": e2e["e2e"]["h2d_bytes_per_step"],
                                       "d2h_bytes_per_step": 16 * E2E_PASSES_PER_STEP, "steps": e2e_steps,
                                       "passes_per_step": E2E_PASSES_PER_STEP, "timed_region_s": dt,
                                       "api": "as `e2e`, but the rates come back through EventGenerator.decays_async() and are read "
                                              "one pass late (pinned D2H + event per rate)",
                                       "rates": [float(got[0]), float(got[1])]}
        except Exception as exc:      # noqa: BLE001
            extras["e2e_pipelined"] = {"error": repr(exc)[:300]}

    # ---- per-sample arrays all the way to NumPy (what a script that reads flux.axion_energy etc. pays) ----------------
    if want("host_arrays") and rank == 0:
        try:
            extras["e2e_host_arrays"] = host_array_legs(A, rt, pf, W, Ar, e2e_pass)
        except Exception as exc:      # noqa: BLE001 -- auxiliary legs never lose the headline
            extras["e2e_host_arrays"] = {"error": repr(exc)[:300]}
    barrier()

    # ---- BASELINE configs 3 and 4 ---------------------------------------------------------------------------------
    if want("c3") and rank == 0:
        try:
            extras["c3"] = c3_leg(A, rt, pf, W, Ar)
        except Exception as exc:      # noqa: BLE001
            extras["c3"] = {"error": repr(exc)[:300]}
    if want("c4") and rank == 0:
        try:
            extras["c4"] = c4_leg(A, rt)
        except Exception as exc:      # noqa: BLE001
            extras["c4"] = {"error": repr(exc)[:300]}
    barrier()

    # ---- ONE job, row-sharded over the ranks, merged [rates || histogram] in one all-reduce (strong scaling) --------------
    if want("strong"):
        try:
            extras["strong"] = strong_leg(A, D, rt, W, Ar, rank, world, barrier, max_over_ranks)
        except Exception as exc:      # noqa: BLE001
            extras["strong"] = {"error": repr(exc)[:300]}
        barrier()

    # ---- (m_a, g) scans: BASELINE config 5 and a 100x larger one, masses dealt round-robin over the ranks ----------------
    if want("scan"):
        for key, n_bins, ns, reps in (("scan", 1000, 100, 20), ("scan_large", 10_000, 1000, 2)):
            try:
                extras[key] = scan_leg(A, W, n_bins, ns, reps, barrier, max_over_ranks)
            except Exception as exc:          # noqa: BLE001
                extras[key] = {"error": repr(exc)[:300]}

    if rank == 0:
        roofline = roofline_object(NV, rt, torch, kernel_ms, n_compton, clocks, ms_total / (args.steps * PASSES_PER_STEP))
        # ---- CPU baseline beside it (bounded sample, 1 core) -------------------------------------------------------------
        ns_cpu = 1000
        try:
            n_cpu, dt_cpu, r_cpu = cpu_pass(c2_spectrum(N_BINS), ns_cpu, 1)
            cpu_baseline = {"value": n_cpu / dt_cpu, "unit": UNIT, "cores": 1, "kind": "port",
                            "sample": f"C2 with n_samples={ns_cpu} per bin (all {N_BINS} bins): {n_cpu} samples in {dt_cpu:.2f} s",
                            "rates": list(r_cpu)}
        except Exception as exc:      # noqa: BLE001
            cpu_baseline = {"value": None, "unit": UNIT, "cores": 1, "kind": "port", "sample": "failed: " + repr(exc)[:200]}
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
               "warmup": warm, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
               "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
               "config": workload_config(world), "clocks": clocks, "e2e": e2e["e2e"],
               "gpu_launches": int(launches), "passes_per_step": PASSES_PER_STEP, "ms_per_pass": ms_total / (args.steps * PASSES_PER_STEP),
               "samples_per_step": world * n_pass * PASSES_PER_STEP, "timed_region_s": ms_total * 1e-3, "rates": final_rates,
               "rng": f"Philox4x32-{comp.rng_rounds} production stream (alplib_b200.config.PHILOX_ROUNDS), precision={comp.precision}",
               "merge_step": ("none (one rank)" if world == 1 else
                              "merged rates of every pass: " + ("one libalpk kernel per rank over NVLink peer memory (alpk_peer_all_reduce_sum) "
                                                                "on a side stream, joined one pass later" if pg is not None else
                                                                "asynchronous NCCL all-reduce, joined one pass later")),
               "roofline": roofline, "cpu_baseline": cpu_baseline}
        for k in ("e2e_eager", "e2e_fused"):
            if k in e2e:
                out[k] = e2e[k]
        out.update(extras)
        out["bench_wall_s"] = time.perf_counter() - t_start
        print(json.dumps(out))
    if world > 1:
        from alplib_b200 import peer
        peer.reset()
        dist.destroy_process_group()


def pinned_d2h_rate(rt, torch, nbytes=1 << 30):
    """Measured pinned device -> host copy rate of this box (GB/s): one cudaMemcpyAsync of 1 GiB, CUDA events, best of 3."""
    src = torch.empty(nbytes, dtype=torch.uint8, device=rt.device)
    dst = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    best = 0.0
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dst.copy_(src, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        best = max(best, nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9)
    return best


def host_array_legs(A, rt, pf, W, Ar, e2e_pass):
    """C2: the four per-sample attributes as NumPy arrays; C3: p1.pmu, p2.pmu and the weights of 1e8 decay events.
    Timed with the host clock around the whole API sequence (host table in -> NumPy arrays out)."""
    import torch
    peak = pinned_d2h_rate(rt, torch)
    out = {"pinned_d2h_gbs_measured": peak,
           "path": "Runtime.to_host: pinned, chunked (64 MiB), double-buffered cudaMemcpyAsync on a copy stream; the host copies "
                   "chunk k out of its pinned buffer while chunk k+1 is in flight"}
    # C2
    reps, t_best, nbytes = 3, None, 0
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        rp, rc, p, c = e2e_pass(None)
        arrs = [c.axion_energy, c.axion_flux, c.decay_axion_weight, c.scatter_axion_weight]
        dt = time.perf_counter() - t0
        nbytes = sum(a.nbytes for a in arrs)
        t_best = dt if t_best is None else min(t_best, dt)
        n = arrs[0].shape[0]
        del arrs, p, c
    out["c2"] = {"samples": int(n), "d2h_bytes": int(nbytes), "seconds": t_best, "gbs": nbytes / t_best / 1e9,
                 "frac_of_pinned_rate": nbytes / t_best / 1e9 / peak, "value": n / t_best, "unit": UNIT,
                 "api": "simulate(); propagate(); decays(); then axion_energy, axion_flux, decay_axion_weight, scatter_axion_weight as NumPy"}
    # C3
    prim = A.FluxPrimakoffIsotropic(pf, W, axion_mass=MA_P, axion_coupling=G_P, **GEOM)
    prim.simulate(); prim.propagate()
    gen = A.PhotonEventGenerator(prim, Ar)
    t_best = None
    for _ in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        l1, l2, w = gen.simulate_decay_4vectors(DAYS, N_SAMPLES, seed=SEED)
        a1, a2, aw = l1.pmu, l2.pmu, np.asarray(w)
        dt = time.perf_counter() - t0
        nbytes = a1.nbytes + a2.nbytes + aw.nbytes
        n_ev = aw.shape[0]
        t_best = dt if t_best is None else min(t_best, dt)
        del l1, l2, w, a1, a2, aw
    out["c3"] = {"events": int(n_ev), "d2h_bytes": int(nbytes), "seconds": t_best, "gbs": nbytes / t_best / 1e9,
                 "frac_of_pinned_rate": nbytes / t_best / 1e9 / peak, "value": n_ev / t_best, "unit": "events/s",
                 "api": "simulate_decay_4vectors(100, 10000) then p1.pmu, p2.pmu (N x 4, transposed on the device) and the weights as NumPy"}
    return out


def c3_leg(A, rt, pf, W, Ar):
    """BASELINE config 3: simulate_decay_4vectors, 1e4 Primakoff parents x 1e4 decays = 1e8 events, lab 4-vectors + weights
    written to HBM (72 B/event), through the public API; roofline against measured HBM bandwidth."""
    import torch
    prim = A.FluxPrimakoffIsotropic(pf, W, axion_mass=MA_P, axion_coupling=G_P, **GEOM)
    prim.simulate(); prim.propagate()
    gen = A.PhotonEventGenerator(prim, Ar)
    n_par = prim.n_flux()
    for _ in range(3):
        res = gen.simulate_decay_4vectors(DAYS, N_SAMPLES, seed=SEED)
    del res
    torch.cuda.synchronize()
    reps = 40
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(reps):
        l1, l2, w = gen.simulate_decay_4vectors(DAYS, N_SAMPLES, seed=SEED + k)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    n_ev = n_par * N_SAMPLES
    # end to end: host flux table in (H2D), production + propagate + decay MC, one event read back (D2H)
    t0 = time.perf_counter()
    for k in range(10):
        p = A.FluxPrimakoffIsotropic(pf, W, axion_mass=MA_P, axion_coupling=G_P, **GEOM)
        p.simulate(); p.propagate()
        l1, l2, w = A.PhotonEventGenerator(p, Ar).simulate_decay_4vectors(DAYS, N_SAMPLES, seed=SEED + k)
        first = float(l1.device[0, 0].item())
    dt = (time.perf_counter() - t0) / 10
    peaks = _peaks()
    peak = float(peaks.get("hbm_gbs", 6650.0))
    ach = n_ev * BYTES_PER_EVENT / (ms * 1e-3) / 1e9
    e_sum = float((l1.device[0] + l2.device[0]).sum().item())
    return {"workload": "C3: simulate_decay_4vectors, 1e4 parents x 1e4 a->gamma gamma decays, FP64 lab 4-vectors + weights in HBM",
            "events": int(n_ev), "ms_per_call": ms, "value": n_ev / (ms * 1e-3), "unit": "events/s", "calls_timed": reps,
            "e2e": {"value": n_ev / dt, "unit": "events/s", "h2d_bytes_per_step": int(pf.shape[0] * 16), "d2h_bytes_per_step": 8,
                    "note": "host flux table in, Primakoff simulate+propagate, decay MC, one element read back; events stay in HBM"},
            "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                         "algorithmic_bytes_per_launch": int(n_ev * BYTES_PER_EVENT),
                         "kernel": "alpk::decay2body_kernel<INJECT=0, FAST=1> (+ alpk_decay_parent_rows), timed through the API call"},
            "checks": {"sum_E1_plus_E2_over_sum_E_parent": e_sum / (N_SAMPLES * float(np.sum(np.asarray(prim.axion_energy)))),
                       "first_event_E1": first}}


def c4_leg(A, rt):
    """BASELINE config 4: FluxBremIsotropic (1e4 e- rows x 1000 samples) + FluxResonanceIsotropic (1e6 inner samples) +
    FluxPairAnnihilationIsotropic (1e4 e+ rows x 100 samples) on graphite, DUNE-like geometry, m_a = 10 MeV, g = 1e-6:
    simulate + propagate + decays through the public classes (host tables in, rates out)."""
    import torch
    Cm, Ar = A.Material("C"), A.Material("Ar")
    E = np.linspace(20, 1.2e5, 10_000)
    ele = np.stack([E, 1e14 * np.exp(-E / 2e4)], axis=1)
    pos = np.stack([E, 0.3 * 1e14 * np.exp(-E / 2e4)], axis=1)
    kw = dict(target_length=100, axion_mass=10.0, axion_coupling=1e-6, **DUNE)

    def brem():
        f = A.FluxBremIsotropic(ele, pos, Cm, n_samples=1000, seed=5, **kw)
        f.simulate(); f.propagate()
        return A.ElectronEventGenerator(f, Ar).decays(DAYS, THR), f._rows[1] * f._rows[2]

    def res():
        f = A.FluxResonanceIsotropic(pos, Cm, n_samples=1_000_000, seed=6, **kw)
        f.simulate(); f.propagate()
        return A.ElectronEventGenerator(f, Ar).decays(DAYS, THR), 1_000_000

    def pair():
        f = A.FluxPairAnnihilationIsotropic(pos, Cm, n_samples=100, seed=7, **kw)
        f.simulate(); f.propagate()
        return A.ElectronEventGenerator(f, Ar).decays(DAYS, THR), f._rows[1] * f._rows[2]

    out = {"workload": "C4: brems (1e4 x 1000) + resonance (1e6 inner samples) + pair annihilation (1e4 x 100) on graphite, "
                       "DUNE-like 574 m / 5 m / 21 m2, m_a = 10 MeV, g = 1e-6; host tables in, rates out (end to end)"}
    tot_n, tot_s = 0, 0.0
    for name, fn in (("brem", brem), ("resonance", res), ("pairann", pair)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        reps = 20
        t0 = time.perf_counter()
        for _ in range(reps):
            rate, n = fn()
        dt = (time.perf_counter() - t0) / reps
        out[name] = {"samples": int(n), "ms_per_pass": dt * 1e3, "value": n / dt, "unit": UNIT, "rate": float(rate)}
        tot_n += n
        tot_s += dt
    out["value"] = tot_n / tot_s
    out["unit"] = UNIT
    out["samples_per_pass"] = int(tot_n)
    out["ms_per_pass"] = tot_s * 1e3
    return out


def strong_leg(A, D, rt, W, Ar, rank, world, barrier, max_over_ranks):
    """ONE beam-dump job sharded by kept flux rows over the ranks (dist.shard_flux): Primakoff + Compton chain with the fused
    event-weighted E_a histogram, then ONE all-reduce of [rate_compton, rate_primakoff || 50 histogram bins].  Reported for the
    C2 job and an 8x larger one (8e4 bins); on N > 1 rank 0 afterwards runs the whole job alone, which gives the strong-scaling
    speed-up inside one run."""
    import torch
    from alplib_b200.peer import peer_group
    merge_kind = ("one libalpk kernel per rank over NVLink peer memory (alpk_peer_all_reduce_sum: push, signal, wait, fixed-order sum)"
                  if peer_group(None, rt.empty(1)) is not None else ("none (one rank)" if world == 1 else "one NCCL all_reduce"))
    out = {"collective": "merge of [2 rates || 50 histogram bins] per job (dist.merged_rates_and_hists): " + merge_kind,
           "sharding": "contiguous blocks of KEPT rows per rank (dist.shard_flux), Philox keys = global row ids"}
    d_edges = rt.upload(HIST_EDGES)

    def make(n_bins, r, w):
        pf = c2_spectrum(n_bins)
        c = A.FluxComptonIsotropic(pf, W, axion_mass=MA_C, axion_coupling=G_C, n_samples=N_SAMPLES, seed=SEED, **GEOM)
        p = A.FluxPrimakoffIsotropic(pf, W, axion_mass=MA_P, axion_coupling=G_P, **GEOM)
        if w > 1:
            D.shard_flux(c, r, w)
            D.shard_flux(p, r, w)
        c._stage_inputs(); p._stage_inputs()
        return c, p, A.PhotonEventGenerator(p, Ar)

    def job(c, p, gp, merge):
        r1, _ = p.chain(DAYS, THR, materialize=False)
        c._launch_rows()
        hist = rt.zeros(HIST_EDGES.shape[0] - 1)
        r2, _ = c.chain(DAYS, THR, materialize=False, hist_edges=d_edges, hist=hist)
        if merge:
            return D.merged_rates_and_hists([r2, r1], [hist])
        return torch.cat([r2, r1]), [hist]

    def timed(c, p, gp, merge, reps, sync):
        for _ in range(3):
            job(c, p, gp, merge)
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            rates, hists = job(c, p, gp, merge)
        e1.record()
        sync()
        return e0.elapsed_time(e1) / reps, rates, hists

    def timed_graph(c, p, gp, merge, reps, sync):
        # the same job as ONE CUDA graph per rank (Runtime.capture): a single host call replays the ~10 launches, merge included
        graph, (rates, hists) = rt.capture(lambda: job(c, p, gp, merge))
        for _ in range(3):
            graph.replay()
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            graph.replay()
        e1.record()
        sync()
        return e0.elapsed_time(e1) / reps, rates, hists

    for key, n_bins, reps in (("c2", N_BINS, 50), ("c2x8", 8 * N_BINS, 10)):
        c, p, gp = make(n_bins, rank, world)
        ms, rates, hists = timed(c, p, gp, world > 1, reps, barrier)
        ms = max_over_ranks(ms)
        try:
            ms_g, rates_g, _ = timed_graph(c, p, gp, world > 1, reps, barrier)
            ms_g = max_over_ranks(ms_g)
            graph_rate_ok = abs(float(rates_g[0].item()) / float(rates[0].item()) - 1.0) < 1e-12
        except Exception as exc:      # noqa: BLE001
            ms_g, graph_rate_ok = None, repr(exc)[:200]
        n_job = 0
        # samples of the whole job = kept Compton rows x n_samples + kept Primakoff rows (identical on every rank)
        pf = c2_spectrum(n_bins)
        n_job = int(np.count_nonzero(A.FluxComptonIsotropic(pf, W, axion_mass=MA_C)._row_keep(pf))) * N_SAMPLES \
            + int(np.count_nonzero(pf[:, 0] >= MA_P))
        entry = {"n_bins": n_bins, "samples_per_job": n_job, "ms_per_job": ms, "value": n_job / (ms * 1e-3), "unit": UNIT,
                 "rows_this_rank": int(c._inputs[2]), "rates": [float(x) for x in rates.cpu()],
                 "hist_sum_over_rate": float(hists[0].sum().item() / max(float(rates[0].item()), 1e-300)),
                 "ms_per_job_graph": ms_g, "value_graph": (n_job / (ms_g * 1e-3)) if ms_g else None,
                 "graph_rate_equals_eager": graph_rate_ok}
        del c, p, gp
        if world > 1:
            barrier()
            if rank == 0:                                         # the same job on ONE rank, the others wait at the barrier
                c1, p1, g1 = make(n_bins, 0, 1)
                ms1, rates1, _ = timed(c1, p1, g1, False, max(3, reps // 3), torch.cuda.synchronize)
                entry["ms_per_job_one_rank"] = ms1
                entry["speedup_vs_one_rank"] = ms1 / ms
                entry["efficiency"] = ms1 / ms / world
                if ms_g:
                    try:
                        ms1g, _, _ = timed_graph(c1, p1, g1, False, max(3, reps // 3), torch.cuda.synchronize)
                        entry["ms_per_job_one_rank_graph"] = ms1g
                        entry["speedup_vs_one_rank_graph"] = ms1g / ms_g
                        entry["efficiency_graph"] = ms1g / ms_g / world
                    except Exception as exc:      # noqa: BLE001
                        entry["ms_per_job_one_rank_graph"] = repr(exc)[:200]
                entry["rate_vs_one_rank_rel_diff"] = abs(float(rates[0].item()) / float(rates1[0].item()) - 1.0)
                del c1, p1, g1
            barrier()
        out[key] = entry
    return out


def scan_leg(A, W, n_bins, n_samples, reps, barrier, max_over_ranks):
    from alplib_b200.scan import scan_rates_distributed
    pf_scan = c2_spectrum(n_bins)
    ma_grid, g_grid = np.logspace(-2, 2, 200), np.logspace(-9, -3, 200)

    def scan_once():
        kw = dict(days_exposure=DAYS, threshold=THR, seed=SEED, **GEOM)
        mc = scan_rates_distributed("compton", pf_scan, W, ma_grid, g_grid, n_samples=n_samples, **kw)
        mp_ = scan_rates_distributed("primakoff", pf_scan, W, ma_grid, g_grid, **kw)
        return mc, mp_

    for _ in range(3):
        scan_once()
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        mc, mp_ = scan_once()
    barrier()
    dt = max_over_ranks(time.perf_counter() - t0)
    evals = int(pf_scan.shape[0] * (n_samples + 1))
    return {"points_per_s": reps * ma_grid.size * g_grid.size / dt, "grid": [200, 200],
            "sample_evals_per_point": evals, "sample_evals_per_s": reps * ma_grid.size * g_grid.size * evals / dt,
            "seconds_per_map_pair": dt / reps,
            "config": f"ma=logspace(-2,2,200), g=logspace(-9,-3,200), {n_bins} bins x ({n_samples} Compton + 1 Primakoff) samples, "
                      "decays(100 d, 5 MeV); host grids in, host maps out (end to end); masses dealt round-robin over the ranks, "
                      "one all-gather per map (a libalpk kernel over NVLink peer memory when the ranks share a node); grid tables and flux columns cached on the device across calls",
            "nonzero_points": [int(np.count_nonzero(mc)), int(np.count_nonzero(mp_))]}


def _peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def roofline_object(NV, rt, torch, kernel_ms, n_compton, clocks, ms_per_pass):
    """Roofline of the dominant kernel (fused Compton production -> propagate -> decays, materialising)."""
    peaks = _peaks()
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (of fallback)"
    kms = float(np.mean(kernel_ms)) if kernel_ms else float("nan")
    achieved = n_compton * BYTES_PER_SAMPLE / (kms * 1e-3) / 1e9
    # FP64 issue-rate denominator measured here (no FP64 figure in MEASURED_PEAKS.json)
    iters, blocks = 20000, rt.sm_count * 8
    scratch = rt.empty(blocks * 256)
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    NV.call("alpk_fp64_peak", 2000, blocks, NV.ptr(scratch), rt.stream())
    torch.cuda.synchronize()
    f0.record()
    NV.call("alpk_fp64_peak", iters, blocks, NV.ptr(scratch), rt.stream())
    f1.record()
    torch.cuda.synchronize()
    fp64_peak = max(blocks * 256 * 16.0 * iters / (f0.elapsed_time(f1) * 1e-3) / 1e12, 1e-9)
    fp64_ach = n_compton * FLOPS_PER_SAMPLE / (kms * 1e-3) / 1e12
    # FP64 exp throughput: library exp and the fast-mode table exp, 4 evaluations per thread-trip
    exp_rates = {}
    for kind, name in ((0, "libdevice_exp_per_s"), (1, "table_exp_per_s")):
        NV.call("alpk_exp_peak", kind, 200, blocks, NV.ptr(scratch), rt.stream())
        torch.cuda.synchronize()
        f0.record()
        NV.call("alpk_exp_peak", kind, 2000, blocks, NV.ptr(scratch), rt.stream())
        f1.record()
        torch.cuda.synchronize()
        exp_rates[name] = blocks * 256 * 4.0 * 2000 / (f0.elapsed_time(f1) * 1e-3)
    # what bounds the kernel: evidence from the committed ncu capture of this kernel (profiles/chain_metrics_r2.json, written
    # by tools/ncu_hot.py --json from the `ncu --set full` report) -- executed instructions per sample pair, the FP64-pipe and
    # issue-slot utilisation, DRAM traffic -- next to this run's time and clock
    met = {}
    try:
        met = json.load(open(os.path.join(ROOT, "profiles", "chain_metrics_r2.json")))
    except Exception:
        pass
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    issue = None
    if met.get("fp64_instr_per_pair"):
        n_fp64, n_all = float(met["fp64_instr_per_pair"]), float(met["instr_per_pair"])
        cost = float(met["issue_cost_cycles_per_pair"])
        cyc = kms * 1e-3 * sm_mhz * 1e6 * rt.sm_count * 4 / (n_compton / 64.0)      # cycles per warp-trip (32 lanes x 2 samples)
        issue = {"instr_per_pair": n_all, "fp64_instr_per_pair": n_fp64, "issue_cost_cycles_per_pair": cost,
                 "measured_cycles_per_pair": cyc, "frac": cost / cyc, "sm_mhz": sm_mhz,
                 "cost_model": "measured issue costs per warp instruction (tools/issue_microbench.cu, profiles/issue_microbench_r2.txt): "
                               "FP64 2 cycles, IMAD.WIDE 4, MUFU.*64H 8, other 1; mixed FP64/integer code issues at the SUM of these",
                 "ncu": {k: met.get(k) for k in ("fp64_pipe_pct", "issue_active_pct", "warps_active_pct", "dram_bytes_per_launch",
                                                 "duration_ms_under_ncu", "registers_per_thread", "top_stalls")},
                 "source": "profiles/chain_metrics_r2.json (ncu source page of this kernel); time and clock: this run"}
    return {"kernel": "alpk::chain_rows_kernel<ComptonProd, STAGE=2, OUT=standard, SHORT=0, ROUNDS=7> "
                      "(production+propagate+decays, materialising; launched by alpk_compton_samples)",
            "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": met.get("dram_bytes_per_launch"), "peak_source": peak_src,
            "algorithmic_bytes_per_launch": n_compton * BYTES_PER_SAMPLE,
            "kernel_ms": kms, "kernel_ms_samples": len(kernel_ms), "kernel_share_of_pass": kms / ms_per_pass,
            "limiter": "instruction issue, not HBM: the reduce-only variant (no DRAM writes) is only ~9 % faster; `bound: hbm` is "
                       "the contract's vocabulary for the algorithmic-bytes roofline, `issue_slots` and `fp64` describe the limiter",
            "issue_slots": issue,
            "fp64": {"achieved": fp64_ach, "peak": fp64_peak, "unit": "TFLOP/s", "frac": fp64_ach / fp64_peak,
                     "flops_per_sample": FLOPS_PER_SAMPLE, **exp_rates,
                     "peak_source": "DFMA loop measured in this run (alpk_fp64_peak)",
                     "note": "flops use SURVEY section 8d's expanded accounting (div 8, sqrt 10, exp 25)"}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--legs", default="all",
                    help="comma list of extra legs: all | e2e_variants,host_arrays,c3,c4,strong,scan | none (profiling runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()

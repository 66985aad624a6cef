import os, sys, time
import numpy as np, torch, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import alplib_b200 as A
from alplib_b200 import runtime as R, _native as N
import bench
W, Ar = A.Material("W"), A.Material("Ar")
pf = bench.c2_spectrum(bench.N_BINS)
log = []
extra = []
def stage_flux_rows(self, table, rule, cut, row_offset=0):
    t = [time.perf_counter()]
    n, stride = int(table.shape[0]), int(table.shape[1])
    cap = ((n * 8 + 255) & ~255) * 2 + n * 4
    hbuf, ev, _ = self._pin_slot(cap); t.append(time.perf_counter())
    dev = torch.empty(max(cap, 1), dtype=torch.uint8, device=self.device); t.append(time.perf_counter())
    nk, ow, oi = C.c_uint64(), C.c_uint64(), C.c_uint64()
    N.call("alpk_stage_flux_rows", C.c_void_p(table.ctypes.data), n, stride, int(rule), float(cut), int(row_offset),
           C.c_void_p(hbuf.data_ptr()), int(hbuf.shape[0]), C.c_void_p(dev.data_ptr()), C.byref(nk), C.byref(ow),
           C.byref(oi), self.stream()); t.append(time.perf_counter())
    ev.record(torch.cuda.current_stream(self.device)); t.append(time.perf_counter())
    k, ow, oi = int(nk.value), int(ow.value), int(oi.value)
    r = (dev[0:k * 8].view(torch.float64), dev[ow:ow + k * 8].view(torch.float64), dev[oi:oi + k * 4].view(torch.int32), k)
    t.append(time.perf_counter())
    log.append([(b - a) * 1e3 for a, b in zip(t[:-1], t[1:])])
    return r
R.Runtime.stage_flux_rows = stage_flux_rows
def step(fused):
    p = A.FluxPrimakoffIsotropic(pf, W, axion_mass=bench.MA_P, axion_coupling=bench.G_P, fused=fused, **bench.GEOM)
    p.simulate(); p.propagate()
    rp = A.PhotonEventGenerator(p, Ar).decays(100, 5.0)
    c = A.FluxComptonIsotropic(pf, W, axion_mass=bench.MA_C, axion_coupling=bench.G_C, n_samples=bench.N_SAMPLES, seed=bench.SEED, fused=fused, **bench.GEOM)
    t0 = time.perf_counter(); c._reset_samples()
    from alplib_b200.fluxes import _as_flux_table, M_E
    ta = time.perf_counter()
    pf_ = _as_flux_table(c.photon_flux, "photon_flux"); tb = time.perf_counter()
    rt = c._rt
    r4 = rt.stage_flux_rows(pf_, 1, float((M_E + c.ma)**2), 0); tc = time.perf_counter()
    tab = rt.empty(max(r4[3] * 12, int(os.environ.get("TAB_MIN", "0")))); td = time.perf_counter()
    c._inputs = (r4[0], r4[1], r4[3], 10000, r4[2], None, c.seed); c._table = tab
    extra.append([(ta - t0) * 1e3, (tb - ta) * 1e3, (tc - tb) * 1e3, (td - tc) * 1e3])
    t1 = time.perf_counter(); c._launch_rows(); t2 = time.perf_counter(); c._finish_simulate()
    c.propagate()
    rc = A.ElectronEventGenerator(c, Ar).decays(100, 5.0)
    return (t1 - t0) * 1e3, (t2 - t1) * 1e3
for _ in range(10): step(False)
for rep in range(2):
    for i in range(40):
        log.clear()
        a, b = step(True)
        if a + b > 0.8:
            print("   pieces (reset, as_table, stage_flux_rows, table alloc):", ['%.3f' % x for x in extra[-1]])
            print(f"step {i}: stage {a:.2f} rows {b:.2f}  stage_flux_rows pieces (pin, empty, call, record, views): P {['%.3f' % x for x in log[0]]}  C {['%.3f' % x for x in log[1]]}")

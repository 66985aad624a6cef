"""Device runtime: one `Runtime` per CUDA device holding the reduction scratch and the H2D/D2H helpers.

PyTorch is used for plumbing only (device memory, streams, torch.distributed); every kernel is in libalpk.so.
"""
import contextlib
import ctypes as C

import numpy as np
import torch

from . import _native as N
from . import config


_NULL_CTX = contextlib.nullcontext()
_TORCH_DTYPE = {np.dtype(np.float64): torch.float64, np.dtype(np.float32): torch.float32,
                np.dtype(np.uint32): torch.int32, np.dtype(np.int32): torch.int32, np.dtype(np.int64): torch.int64}


_NP_DTYPE = {torch.float64: np.float64, torch.float32: np.float32, torch.int32: np.int32, torch.int64: np.int64,
             torch.uint8: np.uint8}


_COPY_THREADS = 4
_copy_pool = None


def _parallel_copy(dst, src):
    """dst[:] = src for large byte arrays on a few host threads (NumPy releases the GIL in the copy loop): one thread cannot
    keep up with a PCIe 5 x16 device -> host stream."""
    global _copy_pool, _COPY_THREADS
    n = dst.shape[0]
    if n < (8 << 20) or _COPY_THREADS == 1:
        dst[:] = src
        return
    if _copy_pool is None:
        # calibrate once on this chunk: some hosts (small VMs) copy faster on one thread
        import time
        from concurrent.futures import ThreadPoolExecutor
        pool = ThreadPoolExecutor(_COPY_THREADS)
        dst[:] = src                                            # (touches the destination pages)
        t0 = time.perf_counter(); dst[:] = src; t1 = time.perf_counter() - t0
        st = (n + _COPY_THREADS - 1) // _COPY_THREADS
        t0 = time.perf_counter()
        for f in [pool.submit(np.copyto, dst[i:i + st], src[i:i + st]) for i in range(0, n, st)]:
            f.result()
        tn = time.perf_counter() - t0
        if tn > 0.9 * t1:
            _COPY_THREADS = 1
            pool.shutdown()
            return
        _copy_pool = pool
        return
    step = (n + _COPY_THREADS - 1) // _COPY_THREADS
    futs = [_copy_pool.submit(np.copyto, dst[i:i + step], src[i:i + step]) for i in range(0, n, step)]
    for f in futs:
        f.result()


# cudaStream_t of torch's current stream on a device, as an integer: the C-level getter costs ~0.2 us against ~7 us for
# torch.cuda.current_stream() (which builds a Stream object), and every launch needs it
_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None) or (lambda index: torch.cuda.current_stream(index).cuda_stream)


class PendingValue:
    """A scalar result on its way to the host: the device -> host copy into pinned memory has been queued behind the kernels
    that produce it, `result()` waits for that copy alone (a CUDA event), not for whatever has been queued since.  Lets a
    loop launch job k+1 before it reads the rate of job k, so the host-side set-up of a job overlaps the previous job's
    kernels (EventGenerator.decays_async)."""
    __slots__ = ("_host", "_event", "_value")

    def __init__(self, tensor):
        self._host = torch.empty(1, dtype=tensor.dtype, pin_memory=True)
        self._host.copy_(tensor.reshape(-1)[:1], non_blocking=True)
        self._event = torch.cuda.Event()
        self._event.record()
        self._value = None

    def done(self):
        return self._value is not None or self._event.query()

    def result(self):
        if self._value is None:
            self._event.synchronize()
            self._value = np.float64(self._host.item())
        return self._value

    def __float__(self):
        return float(self.result())


class Runtime:
    _instances = {}

    def __init__(self, index):
        if not torch.cuda.is_available():
            raise N.AlpkError("alplib_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.index = index
        self.device = torch.device("cuda", index)
        name = C.create_string_buffer(128)
        sm, maj, mnr = C.c_int(), C.c_int(), C.c_int()
        N.call("alpk_device_info", index, name, 128, C.byref(sm), C.byref(maj), C.byref(mnr))
        self.name = name.value.decode()
        self.sm_count = sm.value
        self.cc = (maj.value, mnr.value)
        if maj.value != 10:
            raise N.AlpkError(f"libalpk.so is built for sm_100a only; device {index} ({self.name}) is sm_{maj.value}{mnr.value}")
        self._scratch = {}          # CUDA stream handle -> reduction scratch (launches sharing a scratch must be stream-ordered)

    @classmethod
    def get(cls, device=None):
        if device is None:
            if not torch.cuda.is_available():
                raise N.AlpkError("alplib_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
            index = torch.cuda.current_device()
        elif isinstance(device, int):
            index = device
        else:
            d = torch.device(device)
            index = d.index if d.index is not None else torch.cuda.current_device()
        rt = cls._instances.get(index)
        if rt is None:
            rt = cls._instances[index] = Runtime(index)
        return rt

    # ---- plumbing ---------------------------------------------------------------------------------------------
    @property
    def scratch(self):
        """Reduction scratch (ALPK_SCRATCH_DOUBLES doubles) of the CURRENT stream: the two-stage reductions write partial
        sums there, so flux objects driven from different streams must not share one."""
        key = _raw_stream(self.index)
        s = self._scratch.get(key)
        if s is None:
            s = self._scratch[key] = torch.empty(N.SCRATCH_DOUBLES, dtype=torch.float64, device=self.device)
        return s

    def stream(self):
        return C.c_void_p(_raw_stream(self.index))

    def empty(self, n, dtype=torch.float64):
        return torch.empty(int(n), dtype=dtype, device=self.device)

    def empty_block(self, n, dtype=torch.float64):
        """Like empty() for per-call work buffers of 0.1-1 MB (row tables, staging blocks): the request is padded past
        1 MiB so that PyTorch's caching allocator serves it from its large pool.  Its small pool (2 MiB segments) was
        measured to stall such allocations for 1-4 ms every other step once multi-GB blocks are cached
        (tools/dbg_fused3.py); HBM is not the scarce resource here."""
        item = torch.empty(0, dtype=dtype).element_size()
        n = int(n)
        pad = max(n, ((1 << 20) + 512 + item - 1) // item)
        return torch.empty(pad, dtype=dtype, device=self.device)[:n]

    def zeros(self, n, dtype=torch.float64):
        return torch.zeros(int(n), dtype=dtype, device=self.device)

    def upload(self, host, dtype=np.float64):
        """Host array -> device tensor (async on the current stream; small arrays go through pageable memory)."""
        a = np.ascontiguousarray(host, dtype=dtype)
        t = torch.from_numpy(a)
        return t.to(self.device, non_blocking=True)

    def upload_packed(self, arrays):
        """Several host arrays -> device tensors with ONE asynchronous H2D copy from pinned memory.

        `arrays`: list of (source ndarray -- may be a strided view --, numpy dtype).  The sources are gathered into a
        pinned staging slot (ring of slots, each guarded by a CUDA event so a slot is not rewritten while its copy is
        in flight) and land in one device buffer; the returned tensors are typed views of it."""
        offs, total = [], 0
        for a, dt in arrays:
            offs.append(total)
            total += (int(a.shape[0]) * np.dtype(dt).itemsize + 255) & ~255
        if total == 0:
            return [torch.empty(0, dtype=_TORCH_DTYPE[np.dtype(dt)], device=self.device) for _, dt in arrays]
        slot = self._pin_slot(total)
        hbuf, ev, hnp = slot
        for (a, dt), o in zip(arrays, offs):
            n = int(a.shape[0])
            if n:
                np.copyto(hnp[o:o + n * np.dtype(dt).itemsize].view(dt), a, casting="unsafe")
        dev = self.empty_block(total, torch.uint8)
        dev.copy_(hbuf[:total], non_blocking=True)
        ev.record(torch.cuda.current_stream(self.device))
        return [dev[o:o + int(a.shape[0]) * np.dtype(dt).itemsize].view(_TORCH_DTYPE[np.dtype(dt)])
                for (a, dt), o in zip(arrays, offs)]

    def _pin_slot(self, nbytes):
        """Next pinned staging slot of the ring (>= nbytes), safe to overwrite: (tensor, event, numpy view).  The ring's 8
        slots of 1 MiB are carved out of ONE pinned allocation at first use (eight separate cudaHostAlloc calls cost ~0.5 ms
        each and used to land in the first timed calls of a short run); a slot is re-allocated only for a larger request."""
        ring = self.__dict__.get("_pin_ring")
        if ring is None:
            base = torch.empty(8 << 20, dtype=torch.uint8, pin_memory=True)
            ring = self.__dict__["_pin_ring"] = []
            for k in range(8):
                buf = base[k << 20:(k + 1) << 20]
                ring.append((buf, torch.cuda.Event(), buf.numpy()))
        k = self.__dict__["_pin_next"] = (self.__dict__.get("_pin_next", -1) + 1) % len(ring)
        slot = ring[k]
        if slot[0].shape[0] < nbytes:
            slot[1].synchronize()
            buf = torch.empty(int(nbytes), dtype=torch.uint8, pin_memory=True)
            slot = ring[k] = (buf, torch.cuda.Event(), buf.numpy())
        else:
            slot[1].synchronize()                       # previous copy out of this slot has completed
        return slot

    # ---- device -> host -------------------------------------------------------------------------------------------
    D2H_CHUNK = 64 << 20        # bytes per pinned staging buffer of the chunked device -> host copy

    def to_host(self, tensor, out=None):
        """Device tensor -> NumPy array through pinned memory.

        Large arrays (> 4 MiB) are by default returned AS pinned memory (config.PINNED_RESULTS, see below); with that off, or
        when an `out` array is given:

        Small results (rates, maps, histograms) take one cudaMemcpyAsync into a pinned slot.  Large per-sample arrays are
        copied in chunks of D2H_CHUNK through TWO pinned buffers: while chunk k is copied from the pinned buffer into the
        destination array by the host, chunk k+1 is already in flight on the copy stream -- a pageable `tensor.cpu()`
        instead stages through the driver's small bounce buffer at a fraction of the PCIe rate.  One cudaMemcpyAsync per
        chunk, nothing batched."""
        t = tensor.contiguous()
        nbytes = t.numel() * t.element_size()
        res = np.empty(tuple(t.shape), dtype=_NP_DTYPE[t.dtype]) if out is None else out
        if nbytes == 0:
            return res
        flat_dev = t.view(-1).view(torch.uint8)
        flat_host = res.reshape(-1).view(np.uint8)
        cur = torch.cuda.current_stream(self.device)
        if nbytes <= (4 << 20):
            hbuf, ev, hnp = self._pin_slot(nbytes)
            hbuf[:nbytes].copy_(flat_dev, non_blocking=True)
            ev.record(cur)
            ev.synchronize()
            flat_host[:] = hnp[:nbytes]
            return res
        if out is None and config.PINNED_RESULTS:
            # large result: the NumPy array IS pinned memory (a view of a pinned torch tensor served by torch's caching host
            # allocator), filled by ONE cudaMemcpyAsync at the full PCIe rate -- no second, host-side copy and no page faults
            # on a fresh pageable array.  The first request of a size pays cudaHostAlloc; when the array is garbage-collected
            # its block goes back to the allocator's cache, so loops that re-read arrays of the same size run at DMA speed.
            try:
                dst = torch.empty(tuple(t.shape), dtype=t.dtype, pin_memory=True)
            except RuntimeError:
                dst = None                                   # (not enough lockable memory: chunked path below)
            if dst is not None:
                dst.copy_(t, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(cur)
                ev.synchronize()
                return dst.numpy()
        bufs = self._d2h_buffers()
        copy_stream = self.__dict__.setdefault("_d2h_stream", torch.cuda.Stream(self.device))
        copy_stream.wait_stream(cur)                         # the producer kernels have been queued on the current stream
        chunk = self.D2H_CHUNK
        n_chunks = (nbytes + chunk - 1) // chunk
        with torch.cuda.stream(copy_stream):
            for k in range(min(2, n_chunks)):
                lo, hi = k * chunk, min((k + 1) * chunk, nbytes)
                bufs[k][0][:hi - lo].copy_(flat_dev[lo:hi], non_blocking=True)
                bufs[k][1].record(copy_stream)
            for k in range(n_chunks):
                hb, ev, hnp = bufs[k & 1]
                lo, hi = k * chunk, min((k + 1) * chunk, nbytes)
                ev.synchronize()
                _parallel_copy(flat_host[lo:hi], hnp[:hi - lo])
                nk = k + 2
                if nk < n_chunks:                               # refill this buffer with chunk k+2 while k+1 is consumed
                    lo2, hi2 = nk * chunk, min((nk + 1) * chunk, nbytes)
                    hb[:hi2 - lo2].copy_(flat_dev[lo2:hi2], non_blocking=True)
                    ev.record(copy_stream)
        t.record_stream(copy_stream)
        return res

    def _d2h_buffers(self):
        b = self.__dict__.get("_d2h_bufs")
        if b is None:
            b = []
            for _ in range(2):
                buf = torch.empty(self.D2H_CHUNK, dtype=torch.uint8, pin_memory=True)
                b.append((buf, torch.cuda.Event(), buf.numpy()))
            self.__dict__["_d2h_bufs"] = b
        return b

    def stage_flux_rows(self, table, rule, cut, row_offset=0):
        """Row filter + SoA split + one pinned H2D copy of a (n, >=2) float64 flux table, done by libalpk
        (alpk_stage_flux_rows).  Returns (energies, weights, row ids [int32 view of uint32], n_kept) on the device."""
        n, stride = int(table.shape[0]), int(table.shape[1])
        cap = ((n * 8 + 255) & ~255) * 2 + n * 4
        hbuf, ev, _ = self._pin_slot(cap)
        dev = self.empty_block(max(cap, 1), torch.uint8)
        nk, ow, oi = C.c_uint64(), C.c_uint64(), C.c_uint64()
        N.call("alpk_stage_flux_rows", C.c_void_p(table.ctypes.data), n, stride, int(rule), float(cut), int(row_offset),
               C.c_void_p(hbuf.data_ptr()), int(hbuf.shape[0]), C.c_void_p(dev.data_ptr()), C.byref(nk), C.byref(ow),
               C.byref(oi), self.stream())
        ev.record(torch.cuda.current_stream(self.device))
        k, ow, oi = int(nk.value), int(ow.value), int(oi.value)
        return (dev[0:k * 8].view(torch.float64), dev[ow:ow + k * 8].view(torch.float64),
                dev[oi:oi + k * 4].view(torch.int32), k)

    def capture(self, fn, warmup=2):
        """Capture the launch sequence of `fn()` in a CUDA graph: (graph, fn's return value).  `graph.replay()` re-runs every
        kernel of the sequence with one host call -- for launch-bound loops (a small job repeated many times, a sharded job
        whose per-rank kernels last microseconds).  fn must be free of host synchronisation (no .item(), no host arrays out)
        and of first-use uploads: it is run `warmup` times on a side stream first, which also fills the table caches.  The
        tensors fn returns live in the graph's memory pool and are rewritten by every replay.  Peer-window merges
        (peer.PeerGroup) may be part of fn; every rank then has to capture and replay in step."""
        cur = torch.cuda.current_stream(self.device)
        side = torch.cuda.Stream(self.device)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(warmup):
                fn()
        cur.wait_stream(side)
        torch.cuda.synchronize(self.device)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = fn()
        return graph, out

    def activate(self):
        """Context manager making this runtime's device current (a no-op object when it already is)."""
        if torch.cuda.current_device() == self.index:
            return _NULL_CTX
        return torch.cuda.device(self.device)

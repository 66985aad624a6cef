"""Peer windows: the ranks' merge step (rates, histogram bins, rate-map rows) over NVLink peer memory.

One process per GPU.  `PeerGroup` gives every rank a window in device memory, exchanges the CUDA IPC handles once through
torch.distributed and maps the other ranks' windows; after that a merge is ONE kernel launch of libalpk per rank
(alpk_peer_all_reduce_sum / alpk_peer_all_gather_rows, csrc/alpk_peer.cu): push with NVLink stores, signal, wait, combine
in rank order.  No library collective, no host synchronisation, and -- the epoch being kept on the device -- the launch can
sit inside a captured CUDA graph.  `dist.py` and `scan.py` use it whenever the ranks of the process group live on one
node; otherwise (or with ALPK_PEER=0) they keep torch.distributed's all-reduce / all-gather (NCCL).
"""
import ctypes as C
import os
import socket
import weakref

import torch

from . import _native as N
from .runtime import Runtime

DEFAULT_CAPACITY = 1 << 19           # doubles per payload half (4 MiB): a 200 x 200 rate map needs 4e4, [rates || bins] a few hundred
_groups = {}                         # process group id -> PeerGroup | None (creation failed / disabled)


class PeerGroup:
    """The windows of one process group.  Collective: every rank of `group` must construct it at the same point."""

    def __init__(self, group=None, capacity=DEFAULT_CAPACITY, device=None):
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()):
            raise N.AlpkError("PeerGroup needs an initialised torch.distributed process group")
        self.group = group
        self.rank, self.n_ranks = dist.get_rank(group), dist.get_world_size(group)
        if self.n_ranks > N.PEER_MAX_RANKS:
            raise N.AlpkError(f"PeerGroup: at most {N.PEER_MAX_RANKS} ranks")
        self.rt = Runtime.get(device)
        self.capacity = int(capacity)
        self.timeout_ns = int(float(os.environ.get("ALPK_PEER_TIMEOUT_S", "10")) * 1e9)
        self._own = C.c_void_p()
        self._mapped = []
        self._closed = False
        self._fin = None
        handle = C.create_string_buffer(N.PEER_HANDLE_BYTES)
        err = None
        with self.rt.activate():
            try:
                N.call("alpk_peer_window_create", self.capacity, C.byref(self._own), handle)
            except N.AlpkError as e:          # keep going: the failure has to be agreed on by all ranks
                err = str(e)
        mine = {"host": socket.gethostname(), "pid": os.getpid(), "device": self.rt.index, "handle": bytes(handle.raw),
                "capacity": self.capacity, "error": err}
        infos = [None] * self.n_ranks
        dist.all_gather_object(infos, mine, group=group)
        problem = next((f"rank {r}: {i['error']}" for r, i in enumerate(infos) if i["error"]), None)
        if problem is None and len({i["host"] for i in infos}) != 1:
            problem = "ranks on different hosts"
        if problem is None and len({i["capacity"] for i in infos}) != 1:
            problem = "ranks disagree on the capacity"
        self._peer = N.PeerT()
        self._peer.rank, self._peer.n_ranks, self._peer.capacity = self.rank, self.n_ranks, self.capacity
        if problem is None:
            with self.rt.activate():
                for r, info in enumerate(infos):
                    if r == self.rank:
                        self._peer.windows[r] = self._own.value
                        continue
                    w = C.c_void_p()
                    try:
                        N.call("alpk_peer_window_open", info["handle"], C.byref(w))
                    except N.AlpkError as e:
                        problem = f"rank {self.rank} cannot map the window of rank {r}: {e}"
                        break
                    self._mapped.append(w)
                    self._peer.windows[r] = w.value
        # every rank learns whether every rank succeeded (a half-open group would deadlock in the first merge)
        oks = [None] * self.n_ranks
        dist.all_gather_object(oks, problem, group=group)
        bad = next((o for o in oks if o), None)
        if bad:
            self.close()
            raise N.AlpkError("PeerGroup: " + bad)
        self._fin = weakref.finalize(self, PeerGroup._release, self.rt.index, self._own.value, [w.value for w in self._mapped])

    # ---- merges -------------------------------------------------------------------------------------------------------
    def all_reduce_sum(self, parts, out=None):
        """Sum over ranks of the concatenation of `parts` (float64 CUDA tensors), added in rank order: every rank gets the
        same bits.  One kernel on the current stream; returns the [n_total] tensor."""
        parts = [p.reshape(-1) for p in parts]
        if len(parts) > N.PEER_MAX_PARTS:
            parts = [torch.cat(parts)]
        n = sum(int(p.numel()) for p in parts)
        out = self.rt.empty(n) if out is None else out
        ptrs = (C.c_void_p * len(parts))(*[p.data_ptr() for p in parts])
        lens = (C.c_uint32 * len(parts))(*[int(p.numel()) for p in parts])
        for p in parts:
            if p.dtype != torch.float64 or not p.is_cuda or not p.is_contiguous():
                raise N.AlpkError("PeerGroup.all_reduce_sum: contiguous float64 CUDA tensors")
        N.call("alpk_peer_all_reduce_sum", C.byref(self._peer), ptrs, lens, len(parts), N.ptr(out), self.timeout_ns,
               self.rt.stream())
        return out

    def all_gather_rows(self, local, row_start, row_stride, n_total_rows, out=None):
        """out[n_total_rows, row_len] assembled from every rank's local[n_local_rows, row_len], local row i being global
        row row_start + i * row_stride.  One kernel on the current stream."""
        if local.dtype != torch.float64 or not local.is_cuda:
            raise N.AlpkError("PeerGroup.all_gather_rows: float64 CUDA tensor")
        local = local.contiguous()
        row_len = int(local.shape[1])
        out = torch.empty((int(n_total_rows), row_len), dtype=torch.float64, device=local.device) if out is None else out
        N.call("alpk_peer_all_gather_rows", C.byref(self._peer), N.ptr(local) if local.numel() else None, int(local.shape[0]),
               row_len, int(row_start), int(row_stride), int(n_total_rows), N.ptr(out), self.timeout_ns, self.rt.stream())
        return out

    def status(self):
        """(merges completed on this rank, status: 0 ok / 1 a merge timed out); synchronises the current stream."""
        ep, st = C.c_uint64(), C.c_int()
        N.call("alpk_peer_status", C.byref(self._peer), C.byref(ep), C.byref(st), self.rt.stream())
        return int(ep.value), int(st.value)

    def fits_reduce(self, n):
        return int(n) * self.n_ranks <= self.capacity

    def fits_gather(self, n_total):
        return int(n_total) <= self.capacity

    # ---- teardown -----------------------------------------------------------------------------------------------------
    @staticmethod
    def _release(index, own, mapped):
        try:
            with torch.cuda.device(index):
                torch.cuda.synchronize()
                for w in mapped:
                    N.lib().alpk_peer_window_close(C.c_void_p(w))
                if own:
                    N.lib().alpk_peer_window_destroy(C.c_void_p(own))
        except Exception:       # noqa: BLE001 -- interpreter shutdown
            pass

    def close(self):
        if self._closed:
            return
        self._closed = True
        if self._fin is not None:
            self._fin.detach()
        PeerGroup._release(self.rt.index, self._own.value, [w.value for w in self._mapped])
        self._mapped, self._own = [], C.c_void_p()


def peer_group(group=None, tensor=None):
    """The cached PeerGroup of `group`, or None when peer windows do not apply: a single process, the gloo backend (CPU
    tests / host-staged runs), ALPK_PEER=0, ranks on several hosts, or IPC mapping refused (then NCCL carries the merge).
    Collective on first use."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return None
    if os.environ.get("ALPK_PEER", "1") == "0":
        return None
    if dist.get_backend(group) == "gloo" and os.environ.get("ALPK_PEER", "1") != "force":
        return None
    if tensor is not None and not tensor.is_cuda:
        return None
    if group is None:
        # the default group is a new object after destroy_process_group() + init_process_group(): windows of a group that no
        # longer exists are released here
        key = ("world", id(dist.distributed_c10d._get_default_group()))
        for k in [k for k in _groups if isinstance(k, tuple) and k[0] == "world" and k != key]:
            stale = _groups.pop(k)
            if stale is not None:
                stale.close()
    else:
        key = id(group)
    if key not in _groups:
        try:
            _groups[key] = PeerGroup(group)
        except N.AlpkError as e:
            import warnings
            warnings.warn(f"alplib_b200: peer windows unavailable ({e}); merges use torch.distributed collectives")
            _groups[key] = None
    return _groups[key]


def reset():
    """Drop the cached groups (call before destroy_process_group)."""
    for g in _groups.values():
        if g is not None:
            g.close()
    _groups.clear()

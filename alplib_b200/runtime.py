"""Device runtime: one `Runtime` per CUDA device holding the reduction scratch and the H2D/D2H helpers.

PyTorch is used for plumbing only (device memory, streams, torch.distributed); every kernel is in libalpk.so.
"""
import ctypes as C

import numpy as np
import torch

from . import _native as N


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
        self.scratch = torch.empty(N.SCRATCH_DOUBLES, dtype=torch.float64, device=self.device)

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
    def stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def empty(self, n, dtype=torch.float64):
        return torch.empty(int(n), dtype=dtype, device=self.device)

    def zeros(self, n, dtype=torch.float64):
        return torch.zeros(int(n), dtype=dtype, device=self.device)

    def upload(self, host, dtype=np.float64):
        """Host array -> device tensor (async on the current stream; small arrays go through pageable memory)."""
        a = np.ascontiguousarray(host, dtype=dtype)
        t = torch.from_numpy(a)
        return t.to(self.device, non_blocking=True)

    def activate(self):
        return torch.cuda.device(self.device)

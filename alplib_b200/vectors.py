"""Host-side 3- and 4-vector value types with the reference's interface (cross_section_mc.py:14-143), plus
`LorentzVectorArray`, the structure-of-arrays container the decay kernel fills (sequence of `LorentzVector`)."""
import numpy as np
from numpy import arccos, arctan2


class Vector3:
    def __init__(self, v1, v2, v3):
        self.set_v3(v1, v2, v3)

    def __str__(self):
        return "({0},{1},{2})".format(self.v1, self.v2, self.v3)

    def __add__(self, other):
        return Vector3(self.v1 + other.v1, self.v2 + other.v2, self.v3 + other.v3)

    def __sub__(self, other):
        return Vector3(self.v1 - other.v1, self.v2 - other.v2, self.v3 - other.v3)

    def __neg__(self):
        return Vector3(-self.v1, -self.v2, -self.v3)

    def __mul__(self, other):
        if isinstance(other, (int, float)):
            return Vector3(self.v1 * other, self.v2 * other, self.v3 * other)
        elif isinstance(other, Vector3):
            return np.dot(self.vec, other.vec)
        raise TypeError("Unsupported operand type for *: 'Vector' and '{}'".format(type(other)))

    __rmul__ = __mul__

    def unit_vec(self, eps=1e-40):
        v = self.mag()
        if v < eps:
            return Vector3(0.0, 0.0, 0.0)
        return Vector3(self.v1/v, self.v2/v, self.v3/v)

    def mag2(self):
        return np.dot(self.vec, self.vec)

    def mag(self):
        return np.sqrt(np.dot(self.vec, self.vec))

    def phi(self):
        return arctan2(self.v2, self.v1)

    def theta(self):
        return arccos(self.v3 / self.mag())

    def set_v3(self, v1, v2, v3):
        self.v1, self.v2, self.v3 = v1, v2, v3
        self.vec = np.array([v1, v2, v3])


class LorentzVector:
    """(p0, p1, p2, p3) with metric (+,-,-,-)."""
    mt = np.array([1, -1, -1, -1])

    def __init__(self, p0=0.0, p1=0.0, p2=0.0, p3=0.0):
        self.set_p4(p0, p1, p2, p3)

    def __str__(self):
        return "({0},{1},{2},{3})".format(self.p0, self.p1, self.p2, self.p3)

    def __add__(self, other):
        return LorentzVector(self.p0 + other.p0, self.p1 + other.p1, self.p2 + other.p2, self.p3 + other.p3)

    def __sub__(self, other):
        return LorentzVector(self.p0 - other.p0, self.p1 - other.p1, self.p2 - other.p2, self.p3 - other.p3)

    def __mul__(self, other):
        return np.dot(self.pmu*other.pmu, self.mt)

    __rmul__ = __mul__

    def mass2(self):
        return np.dot(self.pmu**2, self.mt)

    def mass(self):
        return np.sqrt(np.dot(self.pmu**2, self.mt))

    def energy(self):
        return self.p0

    def cosine(self):
        return self.p3 / self.momentum()

    def phi(self):
        return arctan2(self.p2, self.p1)

    def theta(self):
        return arccos(self.p3 / self.momentum())

    def momentum(self):
        return self.momentum3.mag()

    def set_p4(self, p0, p1, p2, p3):
        self.p0, self.p1, self.p2, self.p3 = p0, p1, p2, p3
        self.pmu = np.array([p0, p1, p2, p3])
        self.momentum3 = Vector3(p1, p2, p3)

    def get_3momentum(self):
        return Vector3(self.p1, self.p2, self.p3)

    def get_3velocity(self):
        return Vector3(self.p1/self.p0, self.p2/self.p0, self.p3/self.p0)


def planes_to_host_rows(planes):
    """(C, N) CUDA planes -> (N, C) host array: transposed on the device (alpk_planes_to_rows), then one pinned, chunked,
    double-buffered device -> host copy (Runtime.to_host) -- not a strided pageable gather."""
    import torch
    from . import _native as N
    from .runtime import Runtime
    c, n = int(planes.shape[0]), int(planes.shape[1])
    if planes.dtype != torch.float64 or c not in (3, 4):
        return np.ascontiguousarray(Runtime.get(planes.device).to_host(planes).T)       # (float32 mode: small)
    rt = Runtime.get(planes.device)
    with rt.activate():
        rows = torch.empty((n, c), dtype=torch.float64, device=planes.device)
        N.call("alpk_planes_to_rows", N.ptr(planes), n, c, N.ptr(rows), rt.stream())
        return rt.to_host(rows)


class LorentzVectorArray:
    """N four-vectors stored as planes (E, px, py, pz) -- what `simulate_decay_4vectors` returns.

    Behaves as a read-only sequence of `LorentzVector` (len, indexing, iteration, like the reference's list) and
    exposes the planes for vectorised use: `.pmu` is the (N, 4) host array (copied from the GPU on first use),
    `.device` the (4, N) torch CUDA tensor; `energy()`, `cosine()`, `theta()`, `phi()`, `momentum()` are
    array-valued."""

    def __init__(self, planes_device):
        self.device = planes_device             # torch tensor (4, N), float64, CUDA
        self._host = None

    @property
    def pmu(self):
        if self._host is None:
            self._host = planes_to_host_rows(self.device)
        return self._host

    def __len__(self):
        return int(self.device.shape[1])

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [LorentzVector(*row) for row in self.pmu[i]]
        return LorentzVector(*self.pmu[i])

    def __iter__(self):
        for row in self.pmu:
            yield LorentzVector(*row)

    def __array__(self, dtype=None, copy=None):
        return self.pmu if dtype is None else self.pmu.astype(dtype)

    def energy(self):
        return self.pmu[:, 0]

    def momentum(self):
        return np.sqrt(np.sum(self.pmu[:, 1:]**2, axis=1))

    def cosine(self):
        return self.pmu[:, 3] / self.momentum()

    def theta(self):
        return arccos(self.cosine())

    def phi(self):
        return arctan2(self.pmu[:, 2], self.pmu[:, 1])

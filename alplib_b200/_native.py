"""
ctypes binding of libalpk.so (include/alpk.h) -- the only route from the Python drop-in classes to the sm_100a
kernels.  There is no CPU fallback: if the shared library is missing or a call fails, an exception is raised.
"""
import ctypes as C
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ALPK_LIB_PATH") or os.path.join(_HERE, "libalpk.so")   # override: kernel-variant experiments
CSRC = os.path.join(_HERE, "csrc")
SOURCES = ["alpk_core.cu", "alpk_decay.cu", "alpk_epm.cu", "alpk_pairann.cu", "alpk_scan.cu", "alpk_misc.cu", "alpk_meson.cu",
           "alpk_scatter2.cu", "alpk_peer.cu"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-fmad=false",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "-shared"]

SCRATCH_DOUBLES = 1 << 20
COMPTON_NP = 12
PARENT_NP = 6
PARENT_GENERAL_NP = 20
BREM_NP = 6
PAIRANN_NP = 11
GL_POINTS = 32


class AlpkError(RuntimeError):
    pass


class PropT(C.Structure):
    _fields_ = [("ma", C.c_double), ("ma2", C.c_double), ("width", C.c_double), ("rescale", C.c_double),
                ("neg_dist_mbm", C.c_double), ("neg_len_mbm", C.c_double), ("geom", C.c_double),
                ("geom32", C.c_float), ("apply_geom", C.c_int), ("geom_f64", C.c_int), ("use_tof", C.c_int),
                ("tof_light", C.c_double), ("det_dist", C.c_double), ("timing_window", C.c_double),
                ("fast", C.c_int), ("k1", C.c_double), ("k2", C.c_double)]


class EventT(C.Structure):
    _fields_ = [("days_sec", C.c_double), ("days_sec32", C.c_float), ("days_f64", C.c_int),
                ("threshold", C.c_double)]


class ComptonT(C.Structure):
    _fields_ = [("ma", C.c_double), ("ma2", C.c_double), ("aa", C.c_double), ("z", C.c_double),
                ("n_samples", C.c_double), ("fast", C.c_int), ("rng_rounds", C.c_int)]


class BremT(C.Structure):
    _fields_ = [("ma", C.c_double), ("ma2", C.c_double), ("log10_ma", C.c_double), ("ep_min", C.c_double),
                ("nt", C.c_double), ("pref_num", C.c_double), ("zl", C.c_double), ("zz", C.c_double),
                ("pref_vec", C.c_double), ("ln10", C.c_double), ("n_samples", C.c_double), ("t_arg", C.c_double),
                ("vector", C.c_int), ("use_track_length", C.c_int), ("fast", C.c_int), ("rng_rounds", C.c_int)]


class PairAnnT(C.Structure):
    _fields_ = [("ma", C.c_double), ("ma2", C.c_double), ("me2", C.c_double), ("me4", C.c_double),
                ("pref", C.c_double), ("wconst", C.c_double), ("z", C.c_double), ("ge2", C.c_double),
                ("ntad", C.c_double), ("hbarc2", C.c_double), ("n_samples", C.c_double), ("fast", C.c_int),
                ("rng_rounds", C.c_int)]


class PairGammaT(C.Structure):
    _fields_ = [("ma", C.c_double), ("ma2", C.c_double), ("ma4", C.c_double), ("me2", C.c_double), ("me4", C.c_double),
                ("ep_min", C.c_double), ("nt", C.c_double), ("zc", C.c_double), ("n_samples", C.c_double),
                ("max_t", C.c_double), ("fast", C.c_int)]


class Meson3T(C.Structure):
    _fields_ = [("mm", C.c_double), ("ma", C.c_double), ("fM", C.c_double), ("ckm", C.c_double),
                ("total_width", C.c_double), ("c0", C.c_double), ("coupling", C.c_double), ("a", C.c_double),
                ("b", C.c_double), ("d", C.c_double), ("pref", C.c_double), ("n_samples", C.c_double), ("model", C.c_int)]


class M2T(C.Structure):
    _fields_ = [("model", C.c_int), ("ma2", C.c_double), ("ma4", C.c_double), ("mN2", C.c_double), ("mN4", C.c_double),
                ("z", C.c_double), ("ff_a2", C.c_double), ("pref", C.c_double), ("me2", C.c_double), ("me4", C.c_double),
                ("tmax", C.c_double)]


class Scatter2T(C.Structure):
    _fields_ = [("m2", M2T), ("s", C.c_double), ("msum", C.c_double), ("pp", C.c_double), ("ee", C.c_double),
                ("p3_cm", C.c_double), ("e3_cm", C.c_double), ("vol", C.c_double), ("p1_cm", C.c_double),
                ("n_samples", C.c_double), ("den", C.c_double), ("mat", C.c_double * 16), ("p12", C.c_double * 4),
                ("log_sampling", C.c_int)]


MESON3_NP = 13
PEER_MAX_RANKS = 16
PEER_MAX_PARTS = 8
PEER_HANDLE_BYTES = 64


class PeerT(C.Structure):      # alpk_peer_t
    _fields_ = [("windows", C.c_void_p * PEER_MAX_RANKS), ("rank", C.c_int), ("n_ranks", C.c_int), ("capacity", C.c_uint64)]


def sources():
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def build(force=False, verbose=False):
    """Compile libalpk.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
    srcs = sources()
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")] \
        + [os.path.join(_HERE, "..", "include", "alpk.h")]
    if not force and os.path.exists(LIB_PATH):
        t = os.path.getmtime(LIB_PATH)
        if all(os.path.getmtime(d) <= t for d in deps):
            return LIB_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    # one nvcc per translation unit, in parallel, then link
    # (kernel-variant experiments: ALPK_LIB_PATH picks the output, ALPK_NVCC_EXTRA adds -D flags; objects kept apart)
    objdir = os.path.join(_HERE, "build" if LIB_PATH == os.path.join(_HERE, "libalpk.so") else
                          "build_" + os.path.basename(LIB_PATH).replace(".", "_"))
    os.makedirs(objdir, exist_ok=True)
    cflags = [f for f in NVCC_FLAGS if f != "-shared"] + (["-Xptxas", "-v"] if verbose else []) \
        + os.environ.get("ALPK_NVCC_EXTRA", "").split()
    procs = []
    for src in srcs:
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        procs.append((obj, subprocess.Popen([nvcc] + cflags + ["-c", "-o", obj, src], stdout=subprocess.PIPE,
                                            stderr=subprocess.STDOUT, text=True)))
    objs, log, failed = [], "", False
    for obj, pr in procs:
        out, _ = pr.communicate()
        log += out
        failed |= pr.returncode != 0
        objs.append(obj)
    if verbose:
        sys.stderr.write(log)
    if failed:
        raise AlpkError("nvcc failed:\n" + log)
    r = subprocess.run([nvcc, "-shared", "-o", LIB_PATH] + objs, capture_output=True, text=True)
    if r.returncode != 0:
        raise AlpkError("link failed:\n" + r.stdout + r.stderr)
    return LIB_PATH


_lib = None

_P = C.c_void_p
_U64 = C.c_uint64
_U32 = C.c_uint32
_D = C.c_double
_I = C.c_int

_SIGNATURES = {
    "alpk_version": ([], _I),
    "alpk_last_error": ([], C.c_char_p),
    "alpk_launch_count": ([], C.c_uint64),
    "alpk_device_info": ([_I, C.c_char_p, _I, C.POINTER(_I), C.POINTER(_I), C.POINTER(_I)], _I),
    "alpk_set_device": ([_I], _I),
    "alpk_malloc": ([C.POINTER(C.c_void_p), _U64], _I),
    "alpk_free": ([_P], _I),
    "alpk_memcpy_h2d": ([_P, _P, _U64, _P], _I),
    "alpk_memcpy_d2h": ([_P, _P, _U64, _P], _I),
    "alpk_sync": ([_P], _I),
    "alpk_stage_flux_rows": ([_P, _U64, _U32, _I, _D, _U32, _P, _U64, _P, C.POINTER(_U64), C.POINTER(_U64),
                              C.POINTER(_U64), _P], _I),
    "alpk_propagate": ([_P, _P, _U64, C.POINTER(PropT), _P, _P, _P, _P, _P], _I),
    "alpk_decays": ([_P, _P, _U64, C.POINTER(EventT), _P, _P, _P, _P], _I),
    "alpk_scatter_events": ([_I, _P, _P, _U64, C.POINTER(_D), _D, _D, C.POINTER(EventT), _P, _P, _P, _P], _I),
    "alpk_pair_events": ([_P, _P, _U64, _P, _P, _I, _D, _D, C.POINTER(EventT), _P, _P, _P, _P], _I),
    "alpk_table_sigma": ([_I, _P, _P, _I, _P, _U64, _P, _P], _I),
    "alpk_abs_sigma_mev": ([_P, _P, _I, _P, _U64, _P, _P], _I),
    "alpk_primakoff": ([_P, _P, _U64, _P, _P, _I, _D, _D, _D, _P, _P, _P], _I),
    "alpk_primakoff_chain": ([_P, _P, _U64, _P, _P, _I, _D, _D, _D, C.POINTER(PropT), C.POINTER(EventT), _P, _P, _P, _P, _P, _P,
                              _P, _P, _P, _P], _I),
    "alpk_compton_rows": ([_P, _P, _U64, _P, _P, _I, C.POINTER(ComptonT), _P, _P], _I),
    "alpk_compton_samples": ([_P, _U64, _U32, _P, C.POINTER(ComptonT), _P, _U64, _P, _P,
                              C.POINTER(PropT), _P, _P, _P, _P, C.POINTER(EventT), _P, _P, _P, _P, _I, _P, _P], _I),
    "alpk_brem_rows": ([_P, _P, _U64, _P, _P, _U64, _P, _P, _I, _P, _P, _I, _P, _P, C.POINTER(BremT), _P, _P, _P], _I),
    "alpk_brem_samples": ([_P, _U64, _U32, _P, C.POINTER(BremT), _P, _U64, _P, _P,
                           C.POINTER(PropT), _P, _P, _P, _P, C.POINTER(EventT), _P, _P, _P, _P, _I, _P, _P], _I),
    "alpk_resonance_sum": ([_P, _P, _I, _D, _D, _D, _U64, _P, _U64, _I, _P, _P, _P], _I),
    "alpk_pairann_rows": ([_P, _P, _U64, C.POINTER(PairAnnT), _P, _P], _I),
    "alpk_pairann_samples": ([_P, _U64, _U32, _P, C.POINTER(PairAnnT), _P, _U64, _P, _P,
                              C.POINTER(PropT), _P, _P, _P, _P, C.POINTER(EventT), _P, _P, _P, _P, _I, _P, _P], _I),
    "alpk_decay_parent_rows": ([_P, _P, _U64, _D, C.POINTER(EventT), _U32, _P, _P], _I),
    "alpk_decay2body": ([_P, _U64, _U32, _P, _P, _U64, _I, _P, _P, _P, _P], _I),
    "alpk_decay2body_f32": ([_P, _U64, _U32, _P, _U64, _P, _P, _P, _P], _I),
    "alpk_decay2body_general": ([_P, _P, _U64, _D, _D, _U32, _P, _P, _U64, _P, _P, _P, _P, _P, _P], _I),
    "alpk_scan_tiles": ([_U32, _U32], _U32),
    "alpk_scan": ([_I, _P, _P, _U32, _U32, _P, _P, _I, _P, _P, _I, _P, _P, _I, _D, _D, _D, _U32, _U64, _I,
                   C.POINTER(PropT), C.POINTER(EventT), _P, _P, _U32, _P, _P], _I),
    "alpk_vol_int_chunks": ([_U64, _U32], _U32),
    "alpk_vol_int": ([_P, _P, _U64, C.POINTER(PropT), _P, _P, _P, _U32, _D, _D, _P, _P, _P, _P, _P, _P], _I),
    "alpk_nuclear": ([_P, _U64, _D, _D, _D, _D, _D, _P, _P, _P], _I),
    "alpk_pi0_flux": ([_P, _U64, _P, _U64, _D, _D, _D, _D, _D, _D, _D, _P, _P, _P, _P], _I),
    "alpk_meson3_rows": ([_I, _P, _U64, _P, _P, _U64, C.POINTER(Meson3T), _D, _D, _P, _U64, _U32, _P, _P], _I),
    "alpk_meson3_samples": ([_I, _P, _U64, _U32, C.POINTER(Meson3T), _D, _I, _P, _P, _U64, _U32, _P, _P, _P, _P, _P, _P,
                             _P, _P], _I),
    "alpk_meson3_dgamma": ([_P, _U64, _D, C.POINTER(Meson3T), _P, _P], _I),
    "alpk_meson_mixing": ([_P, _U64, _D, _D, _D, _P, _P, _P, _P, _P], _I),
    "alpk_meson3_m2": ([_D, _P, _U64, _D, C.POINTER(Meson3T), _P, _P], _I),
    "alpk_pairgamma": ([_P, _P, _U64, _U32, _P, _P, _P, _I, C.POINTER(PairGammaT), _P, _P, _U64, _P, _P, _P], _I),
    "alpk_m2_eval": ([C.POINTER(M2T), _P, _U64, _P, _U64, _D, _P, _P], _I),
    "alpk_atomic_ff2": ([_P, _U64, _D, _D, _P, _P], _I),
    "alpk_scatter2to2": ([C.POINTER(Scatter2T), _U64, _P, _U64, _U32, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P], _I),
    "alpk_planes_to_rows": ([_P, _U64, _I, _P, _P], _I),
    "alpk_scale": ([_P, _U64, _D, _P, _P], _I),
    "alpk_hist": ([_P, _P, _P, _U64, _P, _I, _P, _P], _I),
    "alpk_sum": ([_P, _U64, _P, _P, _P], _I),
    "alpk_peer_window_bytes": ([_U64], _U64),
    "alpk_peer_window_create": ([_U64, C.POINTER(C.c_void_p), C.c_char_p], _I),
    "alpk_peer_window_open": ([C.c_char_p, C.POINTER(C.c_void_p)], _I),
    "alpk_peer_window_close": ([_P], _I),
    "alpk_peer_window_destroy": ([_P], _I),
    "alpk_peer_status": ([C.POINTER(PeerT), C.POINTER(_U64), C.POINTER(_I), _P], _I),
    "alpk_peer_all_reduce_sum": ([C.POINTER(PeerT), C.POINTER(C.c_void_p), C.POINTER(_U32), _I, _P, _U64, _P], _I),
    "alpk_peer_all_gather_rows": ([C.POINTER(PeerT), _P, _U32, _U32, _U32, _U32, _U32, _P, _U64, _P], _I),
    "alpk_fp64_peak": ([_U64, _I, _P, _P], _I),
    "alpk_exp_peak": ([_I, _U64, _I, _P, _P], _I),
}


def declared_symbols():
    """Every entry point declared in include/alpk.h (parsed from the header, used by the CPU-only ABI test)."""
    import re
    hdr = open(os.path.join(_HERE, "..", "include", "alpk.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(alpk_[a-z0-9_]+)\s*\(", hdr)))


def lib():
    """Load libalpk.so (once).  Raises AlpkError if it has not been built -- never falls back to the CPU."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH) and not os.environ.get("ALPK_NO_AUTOBUILD") and not os.environ.get("ALPK_LIB_PATH"):
        # the library is a build artefact (git-ignored): compile it on first use if nvcc is there.  One builder at a
        # time (torchrun starts one process per GPU); this is still the CUDA path -- there is no CPU fallback.
        import fcntl
        import shutil
        if shutil.which(os.environ.get("NVCC", "nvcc")):
            with open(os.path.join(_HERE, ".build.lock"), "w") as lock:
                fcntl.flock(lock, fcntl.LOCK_EX)
                try:
                    if not os.path.exists(LIB_PATH):
                        build()
                finally:
                    fcntl.flock(lock, fcntl.LOCK_UN)
    if not os.path.exists(LIB_PATH):
        raise AlpkError(f"{LIB_PATH} not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(nvcc, sm_100a).  alplib_b200 has no CPU fallback.")
    try:
        L = C.CDLL(LIB_PATH)
    except OSError as e:
        raise AlpkError(f"cannot load {LIB_PATH}: {e}") from e
    for name, (argtypes, restype) in _SIGNATURES.items():
        try:
            fn = getattr(L, name)
        except AttributeError as e:
            raise AlpkError(f"{LIB_PATH} does not export {name}") from e
        fn.argtypes = argtypes
        fn.restype = restype
    _lib = L
    return L


_fns = {}


def call(name, *args):
    fn = _fns.get(name)
    if fn is None:
        fn = _fns[name] = getattr(lib(), name)
    rc = fn(*args)
    if rc != 0:
        raise AlpkError(f"{name} failed ({rc}): {lib().alpk_last_error().decode(errors='replace')}")


def launch_count():
    return int(lib().alpk_launch_count())


def ptr(t):
    """Device pointer of a torch CUDA tensor (or None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise AlpkError("alplib_b200 kernels need CUDA tensors; there is no CPU path")
    if not t.is_contiguous():
        raise AlpkError("non-contiguous tensor passed to a kernel")
    return C.c_void_p(t.data_ptr())

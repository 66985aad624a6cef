"""CPU tier: the C-ABI library builds for sm_100a, loads, and exports every symbol include/alpk.h declares (no compute
calls -- there is no GPU here), and the product path refuses to run without a CUDA device instead of falling back."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from alplib_b200 import _native as N


@pytest.fixture(scope="module")
def libpath():
    if not os.path.exists(N.LIB_PATH):
        N.build()
    return N.LIB_PATH


def test_header_symbols_exported(libpath):
    declared = N.declared_symbols()
    assert len(declared) >= 15
    L = ctypes.CDLL(libpath)
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/alpk.h but not exported by libalpk.so"
    # the ctypes binding covers the whole header
    assert sorted(N._SIGNATURES) == declared


def test_lib_binds_and_reports_version(libpath):
    L = N.lib()
    assert L.alpk_version() == 1
    assert L.alpk_last_error() is not None
    hdr = open(os.path.join(ROOT, "include", "alpk.h")).read()
    assert int(re.search(r"#define ALPK_SCRATCH_DOUBLES (\d+)", hdr).group(1)) == N.SCRATCH_DOUBLES
    assert int(re.search(r"#define ALPK_COMPTON_NP (\d+)", hdr).group(1)) == N.COMPTON_NP
    assert int(re.search(r"#define ALPK_PARENT_NP (\d+)", hdr).group(1)) == N.PARENT_NP


def test_struct_layouts_match_header():
    # field order/size of the ctypes mirrors (doubles are 8-aligned, float+3 ints pack into 16 bytes)
    assert ctypes.sizeof(N.PropT) == 7 * 8 + 16 + 3 * 8 + 8 + 16
    assert ctypes.sizeof(N.EventT) == 8 + 8 + 8
    assert ctypes.sizeof(N.ComptonT) == 5 * 8 + 8
    assert N.PropT.k1.offset == 104 and ctypes.sizeof(N.BremT) == 12 * 8 + 16 and ctypes.sizeof(N.PairAnnT) == 11 * 8 + 8
    assert N.PropT.geom32.offset == 56 and N.PropT.tof_light.offset == 72
    assert N.EventT.threshold.offset == 16


def test_peer_window_abi(libpath):
    """alpk_peer_t mirrors the header; the window size helper needs no device."""
    hdr = open(os.path.join(ROOT, "include", "alpk.h")).read()
    assert int(re.search(r"#define ALPK_PEER_MAX_RANKS (\d+)", hdr).group(1)) == N.PEER_MAX_RANKS
    assert int(re.search(r"#define ALPK_PEER_MAX_PARTS (\d+)", hdr).group(1)) == N.PEER_MAX_PARTS
    assert int(re.search(r"#define ALPK_PEER_HANDLE_BYTES (\d+)", hdr).group(1)) == N.PEER_HANDLE_BYTES
    assert ctypes.sizeof(N.PeerT) == N.PEER_MAX_RANKS * 8 + 4 + 4 + 8 and N.PeerT.capacity.offset == N.PEER_MAX_RANKS * 8 + 8
    L = N.lib()
    assert L.alpk_peer_window_bytes(1000) == 512 + 2 * 1000 * 8
    # a merge without mapped windows is refused on the host, before any launch
    peer = N.PeerT(rank=0, n_ranks=2, capacity=1000)
    ptrs = (ctypes.c_void_p * 1)(0)
    lens = (ctypes.c_uint32 * 1)(4)
    assert L.alpk_peer_all_reduce_sum(ctypes.byref(peer), ptrs, lens, 1, None, 0, None) == 2
    assert b"not mapped" in L.alpk_last_error()


def test_peer_group_not_used_without_nccl():
    """Single process / no process group: merges stay local, peer_group() is None (the gloo tiers use torch.distributed)."""
    import torch
    from alplib_b200 import peer
    assert peer.peer_group(None) is None
    assert peer.peer_group(None, torch.zeros(1)) is None


def test_library_contains_sm100a_code_only(libpath):
    out = os.popen(f"cuobjdump -lelf {libpath} 2>/dev/null").read()
    if out:
        assert "sm_100a" in out and "sm_90" not in out and "sm_80" not in out


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import alplib_b200 as A
    fl = A.FluxPrimakoffIsotropic(np.array([[100.0, 1e12]]), A.Material("W"))
    with pytest.raises(N.AlpkError, match="no CPU fallback"):
        fl.simulate()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "alplib_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f"{f} imports the oracle"
                assert "hostcheck" not in src

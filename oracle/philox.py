"""
TEST INFRASTRUCTURE ONLY -- NumPy restatement of the free-running RNG of libalpk (Philox4x32-10, Salmon et al. SC'11;
u53 construction of MT19937's genrand_res53 as used by numpy's legacy random_sample).  Lets the tests replay the
kernels' free-running uniforms on the CPU and feed them to the oracle, so free-running mode is checked sample by
sample as well as statistically.  Known-answer vectors from the Random123 distribution pin the block function.
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
TAG_COMPTON_EA, TAG_BREM_ROW, TAG_BREM_EA, TAG_RESONANCE, TAG_PAIRANN, TAG_DECAY2 = 1, 2, 3, 4, 5, 6
TAG_SCATTER2 = 8          # alpk_scatter2to2 (alpk_scatter2.cu)

# Rounds of the per-sample PRODUCTION streams (tags 1, 3, 5): 10 in the faithful arithmetic mode, 7 (Philox4x32-7) in the fast
# modes -- alplib_b200/config.py:PHILOX_ROUNDS.  Tests set this to the mode under test; every other stream is 10 rounds.
PRODUCTION_ROUNDS = 10
_PRODUCTION_TAGS = (TAG_COMPTON_EA, TAG_BREM_EA, TAG_PAIRANN)


def rounds_of(tag, rounds=None):
    if rounds is not None:
        return int(rounds)
    return PRODUCTION_ROUNDS if tag in _PRODUCTION_TAGS else 10


def philox4x32_10(c0, c1, c2, c3, k0, k1, rounds=10):
    """Philox4x32-R block function (R = 10 unless `rounds` says otherwise)."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint32) for c in np.broadcast_arrays(c0, c1, c2, c3))
    k0 = np.uint32(k0); k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(int(rounds)):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), p0.astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), p1.astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32(k0 + W0); k1 = np.uint32(k1 + W1)
    return c0, c1, c2, c3


def u53(a, b):
    return ((a >> np.uint32(5)).astype(np.float64) * 67108864.0 + (b >> np.uint32(6)).astype(np.float64)) / 9007199254740992.0


def uniform_1(seed, tag, row, n, rounds=None):
    """n uniforms of stream (seed, tag, row): sample s = counter s>>1, word pair s&1."""
    s = np.arange(n, dtype=np.uint64)
    c = s >> np.uint64(1)
    x, y, z, w = philox4x32_10((c & np.uint64(0xFFFFFFFF)).astype(np.uint32), (c >> np.uint64(32)).astype(np.uint32),
                               np.uint32(row), np.uint32(tag), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF,
                               rounds=rounds_of(tag, rounds))
    return np.where((s & np.uint64(1)) == 1, u53(z, w), u53(x, y))


def uniform_2(seed, tag, row, n):
    """n pairs of uniforms of stream (seed, tag, row): sample s = counter s; returns (u1[n], u2[n])."""
    s = np.arange(n, dtype=np.uint64)
    x, y, z, w = philox4x32_10((s & np.uint64(0xFFFFFFFF)).astype(np.uint32), (s >> np.uint64(32)).astype(np.uint32),
                               np.uint32(row), np.uint32(tag), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    return u53(x, y), u53(z, w)


def uniform_2_f32(seed, tag, row, n):
    """The 24-bit uniform pairs of the FP32 decay kernel: event s = counter s>>1, word pair s&1, (x|z, y|w) >> 8."""
    s_ = np.arange(n, dtype=np.uint64)
    c = s_ >> np.uint64(1)
    x, y, z, w = philox4x32_10((c & np.uint64(0xFFFFFFFF)).astype(np.uint32), np.uint32(0), np.uint32(row), np.uint32(tag),
                               seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    odd = (s_ & np.uint64(1)) == 1
    a = np.where(odd, z, x) >> np.uint32(8)
    b = np.where(odd, w, y) >> np.uint32(8)
    return a.astype(np.float64) * 2.0 ** -24, b.astype(np.float64) * 2.0 ** -24

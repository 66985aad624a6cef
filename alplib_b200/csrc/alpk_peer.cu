// libalpk: the path's one exchange step (merged rates / histograms / rate-map rows of the ranks) done by ONE
// kernel over NVLink peer memory instead of a library collective.
//
//   Every rank owns a WINDOW (cudaMalloc'd, exported with cudaIpcGetMemHandle, mapped by the other ranks of the node with
//   cudaIpcOpenMemHandle): a header {epoch, arrival flags[rank]} and a payload of two halves.  A merge is a single-CTA
//   kernel that
//     1. gathers the caller's segments (rates, histogram bins, map rows) and PUSHES them with plain stores into its slot of
//        every rank's window (its own included) -- NVLink writes, no copy engine, no host involvement;
//     2. __threadfence_system(), then stores the new epoch into flags[rank] of every window (release);
//     3. spins (acquire loads on its OWN window, local HBM) until every rank's flag has reached the epoch;
//     4. combines the slots in rank order -- all-reduce: a fixed-order sum, so every rank gets the same bits and the result
//        does not depend on arrival order; all-gather: a copy of the assembled rows -- into the caller's output.
//   The epoch lives in the window and is advanced by the kernel itself, so a captured CUDA graph can replay the merge.  The
//   payload halves alternate with the epoch's parity: a rank can run at most one merge ahead of the slowest rank (it cannot
//   pass step 3 of merge e+1 before everyone has signalled e+1, i.e. finished reading e), and merge e+1 writes the other half.
//   A spin that exceeds `timeout_ns` (a rank died or never launched) poisons the output with NaN and sets the window's
//   status word instead of hanging the GPU.
//
// Latency on a B200 NVSwitch node is a few microseconds (one launch + one NVLink round trip), against 20-40 us of a
// library all-reduce / all-gather of the same few hundred bytes plus its host-side enqueue.
#include <math_constants.h>
#include <string.h>

#include "../../include/alpk.h"
#include "common.cuh"

namespace alpk {

constexpr int kPeerThreads = 1024;
constexpr uint64_t kPeerHeaderBytes = 512;        // epoch, status, flags[ALPK_PEER_MAX_RANKS]

struct PeerHeader {
    unsigned long long epoch;                     // merges completed by the owner of this window
    unsigned int status;                          // 0 ok, 1 a spin timed out
    unsigned int arrive;                          // multi-CTA merges: CTAs that have pushed / finished (reset by the last one)
    unsigned long long flags[ALPK_PEER_MAX_RANKS];   // flags[r]: last epoch rank r has pushed into THIS window
};
static_assert(sizeof(PeerHeader) <= kPeerHeaderBytes, "peer header");

struct PeerArgs {
    unsigned char* win[ALPK_PEER_MAX_RANKS];
    int rank, n_ranks;
    uint64_t capacity;                            // doubles per payload half
    unsigned long long timeout_ns;
};

struct PeerParts {
    const double* ptr[ALPK_PEER_MAX_PARTS];
    uint32_t len[ALPK_PEER_MAX_PARTS];
    int n;
};

__device__ __forceinline__ double* payload(unsigned char* w) { return reinterpret_cast<double*>(w + kPeerHeaderBytes); }

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ double ld_relaxed_sys(const double* p) {
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// steps 2 + 3: every thread has pushed its data; returns false on a timeout
__device__ __forceinline__ bool peer_signal_and_wait(const PeerArgs& p, unsigned long long epoch, bool* s_ok) {
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) *s_ok = true;
    __syncthreads();
    if ((int)threadIdx.x < p.n_ranks) {
        PeerHeader* remote = reinterpret_cast<PeerHeader*>(p.win[threadIdx.x]);
        st_release_sys(&remote->flags[p.rank], epoch);
        const PeerHeader* mine = reinterpret_cast<const PeerHeader*>(p.win[p.rank]);
        const unsigned long long t0 = global_ns();
        while (ld_acquire_sys(&mine->flags[threadIdx.x]) < epoch) {
            if (global_ns() - t0 > p.timeout_ns) { *s_ok = false; break; }
        }
    }
    __syncthreads();
    return *s_ok;
}

__device__ __forceinline__ void peer_finish(const PeerArgs& p, unsigned long long epoch, bool ok) {
    if (threadIdx.x == 0) {
        PeerHeader* mine = reinterpret_cast<PeerHeader*>(p.win[p.rank]);
        mine->epoch = epoch;
        if (!ok) mine->status = 1u;
    }
}

// all-reduce (sum) of the concatenated segments; out[n_total]
__global__ void __launch_bounds__(kPeerThreads)
peer_all_reduce_kernel(const PeerArgs p, const PeerParts parts, uint32_t n_total, double* __restrict__ out) {
    __shared__ bool s_ok;
    const unsigned long long epoch = reinterpret_cast<const PeerHeader*>(p.win[p.rank])->epoch + 1ull;
    const uint64_t half = (epoch & 1ull) * p.capacity;
    // 1. push my values into slot `rank` of every window
    for (uint32_t i = threadIdx.x; i < n_total; i += blockDim.x) {
        uint32_t k = i;
        int s = 0;
        while (s < parts.n - 1 && k >= parts.len[s]) { k -= parts.len[s]; ++s; }
        const double v = parts.ptr[s][k];
        for (int r = 0; r < p.n_ranks; ++r) payload(p.win[r])[half + (uint64_t)p.rank * n_total + i] = v;
    }
    const bool ok = peer_signal_and_wait(p, epoch, &s_ok);
    // 4. fixed-order sum of the ranks' slots (the same bits on every rank)
    const double* mine = payload(p.win[p.rank]) + half;
    for (uint32_t i = threadIdx.x; i < n_total; i += blockDim.x) {
        double v = 0.0;
        for (int r = 0; r < p.n_ranks; ++r) v += ld_relaxed_sys(mine + (uint64_t)r * n_total + i);
        out[i] = ok ? v : CUDART_NAN;
    }
    __syncthreads();
    peer_finish(p, epoch, ok);
}

// all-gather of row blocks: local[n_local_rows][row_len]; local row i is global row row_start + i * row_stride of
// out[n_total_rows][row_len] (row_stride = n_ranks, row_start = rank: rows dealt round-robin; stride 1: contiguous blocks).
// A few CTAs (all resident: grid <= kGatherMaxCtas) share the push and the copy-out; the LAST CTA to have pushed sends the
// signal (every CTA fenced its stores before arriving), every CTA then waits on the local flags.
constexpr int kGatherMaxCtas = 32;

__global__ void __launch_bounds__(kPeerThreads)
peer_all_gather_rows_kernel(const PeerArgs p, const double* __restrict__ local, uint32_t n_local_rows, uint32_t row_len,
                            uint32_t row_start, uint32_t row_stride, uint32_t n_total_rows, double* __restrict__ out) {
    __shared__ bool s_ok;
    __shared__ bool s_last;
    PeerHeader* const hdr = reinterpret_cast<PeerHeader*>(p.win[p.rank]);
    const unsigned long long epoch = ld_acquire_sys(&hdr->epoch) + 1ull;
    const uint64_t half = (epoch & 1ull) * p.capacity;
    const uint64_t n_local = (uint64_t)n_local_rows * row_len;
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = tid; i < n_local; i += nthr) {
        const uint32_t row = (uint32_t)(i / row_len);
        const uint32_t col = (uint32_t)(i - (uint64_t)row * row_len);
        const double v = local[i];
        const uint64_t dst = half + (uint64_t)(row_start + row * row_stride) * row_len + col;
        for (int r = 0; r < p.n_ranks; ++r) payload(p.win[r])[dst] = v;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        s_ok = true;
        s_last = atomicAdd(&hdr->arrive, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (s_last) {                                                    // every CTA's stores are fenced and counted: signal
        __threadfence_system();
        if ((int)threadIdx.x < p.n_ranks)
            st_release_sys(&reinterpret_cast<PeerHeader*>(p.win[threadIdx.x])->flags[p.rank], epoch);
    }
    if ((int)threadIdx.x < p.n_ranks) {
        const unsigned long long t0 = global_ns();
        while (ld_acquire_sys(&hdr->flags[threadIdx.x]) < epoch) {
            if (global_ns() - t0 > p.timeout_ns) { s_ok = false; break; }
        }
    }
    __syncthreads();
    const bool ok = s_ok;
    const double* mine = payload(p.win[p.rank]) + half;
    const uint64_t n_total = (uint64_t)n_total_rows * row_len;
    for (uint64_t i = tid; i < n_total; i += nthr) out[i] = ok ? ld_relaxed_sys(mine + i) : CUDART_NAN;
    __syncthreads();
    if (threadIdx.x == 0) {
        if (!ok) hdr->status = 1u;
        __threadfence();
        // the last CTA to FINISH advances the epoch (no CTA can still be reading it) and re-arms the counter
        if (atomicAdd(&hdr->arrive, 1u) == 2u * gridDim.x - 1u) {
            hdr->arrive = 0u;
            hdr->epoch = epoch;
        }
    }
}

static int to_args(const char* what, const alpk_peer_t* h, PeerArgs& a, uint64_t timeout_ns) {
    if (!h) { set_error("%s: null peer group", what); return 2; }
    if (h->n_ranks < 1 || h->n_ranks > ALPK_PEER_MAX_RANKS || h->rank < 0 || h->rank >= h->n_ranks) {
        set_error("%s: rank %d of %d (at most %d ranks)", what, h->rank, h->n_ranks, ALPK_PEER_MAX_RANKS);
        return 2;
    }
    memset(&a, 0, sizeof(a));
    for (int r = 0; r < h->n_ranks; ++r) {
        if (!h->windows[r]) { set_error("%s: window of rank %d is not mapped", what, r); return 2; }
        a.win[r] = static_cast<unsigned char*>(h->windows[r]);
    }
    a.rank = h->rank; a.n_ranks = h->n_ranks; a.capacity = h->capacity;
    a.timeout_ns = timeout_ns ? timeout_ns : 10000000000ull;
    return 0;
}

}  // namespace alpk

using namespace alpk;

ALPK_EXPORT uint64_t alpk_peer_window_bytes(uint64_t capacity_doubles) {
    return kPeerHeaderBytes + 2ull * capacity_doubles * sizeof(double);
}

ALPK_EXPORT int alpk_peer_window_create(uint64_t capacity_doubles, void** h_window, unsigned char* h_handle) {
    if (!h_window || !h_handle || capacity_doubles == 0) { set_error("alpk_peer_window_create: bad argument"); return 2; }
    static_assert(sizeof(cudaIpcMemHandle_t) == ALPK_PEER_HANDLE_BYTES, "IPC handle size");
    void* w = nullptr;
    const uint64_t bytes = alpk_peer_window_bytes(capacity_doubles);
    cudaError_t e = cudaMalloc(&w, bytes);
    if (e != cudaSuccess) { set_error("alpk_peer_window_create: cudaMalloc(%llu): %s", (unsigned long long)bytes, cudaGetErrorString(e)); return 1; }
    e = cudaMemset(w, 0, bytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaIpcMemHandle_t hd;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&hd, w);
    if (e != cudaSuccess) { set_error("alpk_peer_window_create: %s", cudaGetErrorString(e)); cudaFree(w); cudaGetLastError(); return 1; }
    memcpy(h_handle, &hd, sizeof(hd));
    *h_window = w;
    return 0;
}

ALPK_EXPORT int alpk_peer_window_open(const unsigned char* h_handle, void** h_window) {
    if (!h_handle || !h_window) { set_error("alpk_peer_window_open: null argument"); return 2; }
    cudaIpcMemHandle_t hd;
    memcpy(&hd, h_handle, sizeof(hd));
    void* w = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&w, hd, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { set_error("alpk_peer_window_open: %s", cudaGetErrorString(e)); cudaGetLastError(); return 1; }
    *h_window = w;
    return 0;
}

ALPK_EXPORT int alpk_peer_window_close(void* d_window) {
    cudaError_t e = cudaIpcCloseMemHandle(d_window);
    if (e != cudaSuccess) { set_error("alpk_peer_window_close: %s", cudaGetErrorString(e)); cudaGetLastError(); return 1; }
    return 0;
}

ALPK_EXPORT int alpk_peer_window_destroy(void* d_window) {
    cudaError_t e = cudaFree(d_window);
    if (e != cudaSuccess) { set_error("alpk_peer_window_destroy: %s", cudaGetErrorString(e)); cudaGetLastError(); return 1; }
    return 0;
}

ALPK_EXPORT int alpk_peer_status(const alpk_peer_t* h_peer, uint64_t* h_epoch, int* h_status, void* stream) {
    PeerArgs a;
    if (to_args("alpk_peer_status", h_peer, a, 0)) return 2;
    PeerHeader hd;
    cudaError_t e = cudaMemcpyAsync(&hd, a.win[a.rank], sizeof(hd), cudaMemcpyDeviceToHost, (cudaStream_t)stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)stream);
    if (e != cudaSuccess) { set_error("alpk_peer_status: %s", cudaGetErrorString(e)); return 1; }
    if (h_epoch) *h_epoch = hd.epoch;
    if (h_status) *h_status = (int)hd.status;
    return 0;
}

ALPK_EXPORT int alpk_peer_all_reduce_sum(const alpk_peer_t* h_peer, const double* const* h_parts, const uint32_t* h_part_len,
                                         int n_parts, double* out, uint64_t timeout_ns, void* stream) {
    PeerArgs a;
    if (to_args("alpk_peer_all_reduce_sum", h_peer, a, timeout_ns)) return 2;
    if (n_parts < 1 || n_parts > ALPK_PEER_MAX_PARTS || !h_parts || !h_part_len || !out) {
        set_error("alpk_peer_all_reduce_sum: 1..%d segments", ALPK_PEER_MAX_PARTS);
        return 2;
    }
    PeerParts parts;
    memset(&parts, 0, sizeof(parts));
    uint64_t n_total = 0;
    for (int i = 0; i < n_parts; ++i) {
        if (!h_parts[i] && h_part_len[i]) { set_error("alpk_peer_all_reduce_sum: null segment %d", i); return 2; }
        parts.ptr[i] = h_parts[i]; parts.len[i] = h_part_len[i];
        n_total += h_part_len[i];
    }
    parts.n = n_parts;
    if (n_total == 0) return 0;
    if (n_total * (uint64_t)a.n_ranks > a.capacity) {
        set_error("alpk_peer_all_reduce_sum: %llu doubles x %d ranks exceed the window (%llu per half)", (unsigned long long)n_total,
                  a.n_ranks, (unsigned long long)a.capacity);
        return 2;
    }
    peer_all_reduce_kernel<<<1, kPeerThreads, 0, (cudaStream_t)stream>>>(a, parts, (uint32_t)n_total, out);
    return check_launch("alpk_peer_all_reduce_sum");
}

ALPK_EXPORT int alpk_peer_all_gather_rows(const alpk_peer_t* h_peer, const double* local, uint32_t n_local_rows, uint32_t row_len,
                                          uint32_t row_start, uint32_t row_stride, uint32_t n_total_rows, double* out,
                                          uint64_t timeout_ns, void* stream) {
    PeerArgs a;
    if (to_args("alpk_peer_all_gather_rows", h_peer, a, timeout_ns)) return 2;
    if (!out || (n_local_rows && !local)) { set_error("alpk_peer_all_gather_rows: null argument"); return 2; }
    if ((uint64_t)n_total_rows * row_len > a.capacity) {
        set_error("alpk_peer_all_gather_rows: %u x %u doubles exceed the window (%llu per half)", n_total_rows, row_len,
                  (unsigned long long)a.capacity);
        return 2;
    }
    if (n_local_rows && (uint64_t)row_start + (uint64_t)(n_local_rows - 1) * row_stride >= n_total_rows) {
        set_error("alpk_peer_all_gather_rows: local rows fall outside the %u output rows", n_total_rows);
        return 2;
    }
    if ((uint64_t)n_total_rows * row_len == 0) return 0;
    // one CTA per ~8k output doubles, all co-resident (they spin on the flags)
    const uint64_t n_out = (uint64_t)n_total_rows * row_len;
    int grid = (int)((n_out + 8191) / 8192);
    grid = grid < 1 ? 1 : (grid > kGatherMaxCtas ? kGatherMaxCtas : grid);
    peer_all_gather_rows_kernel<<<grid, kPeerThreads, 0, (cudaStream_t)stream>>>(a, local, n_local_rows, row_len, row_start,
                                                                               row_stride, n_total_rows, out);
    return check_launch("alpk_peer_all_gather_rows");
}

// desman_b200/csrc/exchange_kernel.cuh -- the per-sweep exchange of the sharded chain as ONE small kernel over NVLink peer
// memory instead of an NCCL all-reduce.  The payloads are tiny (S*G+16 statistics, then 2 words: fixed-point ll, nchange), so the
// exchange is pure latency: NCCL's launch + protocol cost ~16 us per call on 2 GPUs; a one-shot "everybody writes its
// contribution into everybody's mailbox, raises a flag, waits for the other flags, sums in rank order" is a few us.
//
// Every rank owns a mailbox (cudaMalloc + cudaIpc handle, opened by the peers at desman_comm_init):
//     slot[parity][rank][XCH_WORDS] uint64   contributions, double buffered on the parity of the exchange number
//     flag[rank]                   uint64   number of the last exchange whose contribution from `rank` is complete
// Double buffering suffices: a rank can start exchange n+2 only after every peer raised flag n+1, i.e. after every peer's
// kernel of exchange n -- the last reader of the parity-n slots -- has finished (stream order on the peer).
// The sums are taken in rank order on every rank: identical results everywhere (the payloads are integers anyway).
#pragma once
#include "common.cuh"
#include "misc_kernels.cuh"

#define XCH_MAX_RANKS 16

struct XchParams {
    unsigned long long *mail[XCH_MAX_RANKS];   // mailbox of every rank (peer pointers; mail[rank] is local)
    int rank, nranks;
    int words;                                 // payload length
    int cap_words;                             // XCH_WORDS of the mailbox layout
    unsigned long long seq;                    // number of this exchange (1, 2, ...)
    unsigned long long *data;                  // in: this rank's contribution; out: the sum over ranks
    unsigned long long *data2;                 // optional second segment (appended to the first in the mailbox), or null
    int words2;
    int *err;                                  // set to 1 if a peer never showed up (spin limit)
};

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// (all threads of the block; contains block barriers)
__device__ __forceinline__ void exchange_sum_body(const XchParams &p)
{
    const int n = p.nranks, W1 = p.words, W = p.words + p.words2;
    const size_t slot_off = ((size_t)(p.seq & 1ull) * n + p.rank) * p.cap_words;
    // 1. my contribution into every mailbox (peer stores over NVLink; the local one is a plain store)
    for (int i = threadIdx.x; i < W * n; i += blockDim.x) {
        const int r = i / W, j = i - r * W;
        p.mail[r][slot_off + j] = (j < W1) ? p.data[j] : p.data2[j - W1];
    }
    __threadfence_system();
    __syncthreads();
    // 2. raise my flag everywhere, then wait for everybody's flag in my own mailbox
    unsigned long long *flags = p.mail[p.rank] + (size_t)2 * n * p.cap_words;
    if ((int)threadIdx.x < n) {
        const int r = threadIdx.x;
        st_release_sys(p.mail[r] + (size_t)2 * n * p.cap_words + p.rank, p.seq);
        long long spins = 0;
        while (ld_acquire_sys(flags + r) < p.seq) {
            if (++spins > (1ll << 28)) { *p.err = 1; break; }      // ~ half a minute: a peer died; do not hang the GPU
            __nanosleep(64);
        }
    }
    __syncthreads();
    // 3. sum in rank order
    const unsigned long long *mine = p.mail[p.rank] + (size_t)(p.seq & 1ull) * n * p.cap_words;
    for (int j = threadIdx.x; j < W; j += blockDim.x) {
        unsigned long long acc = 0ull;
        for (int r = 0; r < n; r++) acc += mine[(size_t)r * p.cap_words + j];
        if (j < W1) p.data[j] = acc; else p.data2[j - W1] = acc;
    }
}

__global__ void __launch_bounds__(512) exchange_sum_kernel(XchParams p)
{
    pdl_enter();         // a link of the sweep's dependent chain (common.cuh): scheduled under the tail of the statistics kernel
    exchange_sum_body(p);
}

// The exchange of sweep k and the lp / stores / MAP bookkeeping of sweep k-1 that follows it in the sharded chain, as ONE launch:
// both are single-block latency kernels, and a hand-over between two launches costs more than either's work.  The words the
// bookkeeping reads (f.red_i) are the second segment of the exchange (x.data2), summed by this very block.
__global__ void __launch_bounds__(512) exchange_finalize_kernel(XchParams x, FinalParams f)
{
    __shared__ double sh[256];
    __shared__ int upd;
#if !PDL_EARLY
    pdl_enter();
#endif
    const FinalEarly e = finalize_early(f, sh);       // log-priors of gamma / eta of sweep k-1: drawn several grids ago
#if PDL_EARLY
    pdl_enter();
#endif
    exchange_sum_body(x);
    __syncthreads();                                  // (the sums written by other threads of this block: visible to thread 0)
    finalize_late(f, e, &upd);
}

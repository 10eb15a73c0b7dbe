// desman_b200/csrc/common.cuh -- shared device helpers (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define DESMAN_FULL_MASK 0xffffffffu

// Programmatic dependent launch (engine.cu launch_k): the kernels of a sweep form one dependent chain on one stream, and each
// completion -> launch hand-over costs ~4 us on B200 when left to the stream.  Every kernel of the chain starts with
// pdl_enter(): wait until the preceding grid has completed and its writes are visible, then allow the next grid of the chain
// to be scheduled (its CTAs become resident as this grid's CTAs retire and park in their own wait).  Without the launch
// attribute both instructions are no-ops.  Nothing produced by an earlier kernel may be read before pdl_enter().
// PDL_EARLY: a few kernels run the part of their prologue that reads nothing of the immediately preceding grid (shared-memory
// initialisation; operands written two or more grids earlier, which are complete because the predecessor passed its own
// pdl_enter() before this grid could be scheduled) BEFORE pdl_enter(), i.e. under the tail of the predecessor.
#ifndef PDL_EARLY
#define PDL_EARLY 1
#endif
__device__ __forceinline__ void pdl_enter()
{
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// -DKPROF build (diagnosis only, tools/kprof.py): %globaltimer stamps of every CTA's entry / exit (and of the phases of a few
// kernels) recorded in a device buffer that desman_kprof_dump() copies out.  The product build compiles none of it.
enum { KP_MAINT = 0, KP_MUB, KP_MUC, KP_DRAW, KP_TGM, KP_TAU, KP_LL, KP_FIN, KP_COPY, KP_TAU_WARP, KP_TGM_PRO, KP_MUB_WARP, KP_TC_EVT, KP_TAUO };
#ifdef KPROF
struct KRec { int kid, cta, warp, x; unsigned long long t0, t1, a, b, c, d; };
#define KREC_CAP (1 << 18)
__device__ KRec g_krec[KREC_CAP];
__device__ unsigned int g_krec_n;
__device__ __forceinline__ unsigned long long gtimer()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __noinline__ void krec_put(int kid, int cta, int warp, int x, unsigned long long t0, unsigned long long t1,
                                      unsigned long long a = 0, unsigned long long b = 0, unsigned long long c = 0, unsigned long long d = 0)
{
    const unsigned int i = atomicAdd(&g_krec_n, 1u);
    if (i < KREC_CAP) g_krec[i] = KRec{kid, cta, warp, x, t0, t1, a, b, c, d};
}
struct KProfScope {
    int kid; unsigned long long t0;
    __device__ KProfScope(int k) : kid(k), t0(gtimer()) {}
    __device__ ~KProfScope() { if (threadIdx.x == 0) krec_put(kid, (int)blockIdx.x, 0, 0, t0, gtimer()); }
};
#define KPROF_SCOPE(kid) KProfScope kprof_scope_(kid)
#else
#define KPROF_SCOPE(kid)
#endif

// counter "stage" tags of the Philox contract; c3 = stage<<28 | ...
enum { STAGE_TAU = 1, STAGE_MU = 2, STAGE_GAMMA = 3, STAGE_ETA = 4, STAGE_GAMMA_BOOST = 5, STAGE_ETA_BOOST = 6 };

// Philox4x32-10 (Salmon, Moraes, Dror, Shaw 2011).  Counter-based: every draw of the chain is a
// pure function of (seed, sweep, stage, site, sample, base, index), so results do not depend on
// grid shape, warp scheduling or GPU count.
__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                               uint32_t k0, uint32_t k1)
{
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0;
        uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}

__device__ __forceinline__ double warp_sum(double x)
{
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) x += __shfl_xor_sync(DESMAN_FULL_MASK, x, m);
    return x;  // xor butterfly: every lane ends with the bit-identical sum
}

__device__ __forceinline__ float warp_sum(float x)
{
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) x += __shfl_xor_sync(DESMAN_FULL_MASK, x, m);
    return x;
}

__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long x)
{
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) x += __shfl_xor_sync(DESMAN_FULL_MASK, x, m);
    return x;
}

__device__ __forceinline__ float warp_max(float x)
{
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) x = fmaxf(x, __shfl_xor_sync(DESMAN_FULL_MASK, x, m));
    return x;
}

// 2-bit-per-strain packed haplotype code of one site (G <= 32), built from the uint8 tau row.
__device__ __forceinline__ uint64_t load_tau_code(const uint8_t *__restrict__ tau_row, int G, int lane)
{
    uint32_t t = (lane < G) ? (uint32_t)tau_row[lane] & 3u : 0u;
    uint32_t lo = __reduce_or_sync(DESMAN_FULL_MASK, lane < 16 ? t << (2 * lane) : 0u);
    uint32_t hi = __reduce_or_sync(DESMAN_FULL_MASK, lane >= 16 ? t << (2 * (lane - 16)) : 0u);
    return ((uint64_t)hi << 32) | lo;
}

__device__ __forceinline__ int code_get(uint64_t code, int g) { return (int)((code >> (2 * g)) & 3ull); }
__device__ __forceinline__ uint64_t code_set(uint64_t code, int g, int b)
{
    return (code & ~(3ull << (2 * g))) | ((uint64_t)b << (2 * g));
}

__device__ __forceinline__ int4 ld_counts(const int4 *p)
{
    int4 r;  // streaming 128-bit load: one (v,s) cell, read once per pass
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

// Control words of the pattern groups (tau_group_kernel.cuh, table_maintain_kernel): requests and list lengths.
enum { GC_REGROUP = 0,   // regroup wanted (finalize_sweep sets it when orphans pile up; a table rebuild implies one)
       GC_HAVE = 1,      // the groups are valid for the current slot numbering
       GC_CALM = 2,      // the previous sweep flipped few sites: the screening pass pays off
       GC_NITEMS = 3, GC_NSINGLES = 4, GC_NWORK = 5, GC_CURSOR = 6,
       GC_ORPHANS = 7,   // sites that changed pattern since the last regroup (screened out by their slot mismatch)
       GC_IMG_OK = 8,    // the fp16 count image of the tensor-memory screening pass (tau_group_tc_kernel.cuh) was built and fits
       GC_IMG_ROWS = 9,  // its padded row count
       GC_WORTH = 10,    // the grouping pays: enough sites share patterns (set at every regroup from the realised groups)
       GC_COUNT = 12 };
// Is the work list of the screening pass valid for this sweep?  (One test for the screening kernels and the kernels that walk
// the list: the groups exist, the chain is calm, the realised groups are worth a table each, and -- for the tensor-memory
// form -- the count image was built.)
__device__ __forceinline__ bool grp_active(const int *gctl, int need_img)
{
    return gctl[GC_HAVE] && gctl[GC_CALM] && gctl[GC_WORTH] && (!need_img || gctl[GC_IMG_OK]);
}

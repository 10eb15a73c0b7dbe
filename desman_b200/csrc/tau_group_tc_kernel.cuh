// desman_b200/csrc/tau_group_tc_kernel.cuh -- K1t: the screening pass of the tau Gibbs update (c_sample_tau.c:130-188) on the
// Blackwell tensor path: TMA bulk copies (cp.async.bulk + mbarrier) stage the count rows in shared memory, tcgen05.mma
// contracts them with the pattern's table, the sums live in tensor memory and come back with tcgen05.ld.
//
// What is computed is exactly what tau_group_kernel.cuh describes: for a work item (one haplotype pattern, <= 128 sites) the
// log-likelihood differences of a site's G steps are a dense contraction of its count row with the pattern's table,
//     D[site][c] = sum_{s,b} n[site][s][b] * Wd[c][s][b],      c = 3 g + j,   Wd = lg2(q) - lg2(P)  (same FP32 arithmetic),
// a [sites x 4S] x [4S x 3G] product.  Here the SITES are the M dimension of the MMA (128 rows = 128 TMEM lanes, one thread of the
// epilogue per site: the 3G sums of a site sit in ONE thread's registers and the gap test needs no shuffle) and the table
// columns the N dimension.  Operands are FP16 (kind::f16, FP32 accumulation in TMEM):
//   * counts < 2048 are exact in FP16 (11-bit significand): the group-ordered copy of the count rows is kept as fp16x4 cells
//     (8 bytes per (v,s) instead of 16: the pass reads HALF the bytes of the canonical int32x4 tensor);
//   * a table entry is split as  Wd = h + l,  h = Wd truncated to 11 significant bits, l = fp16(Wd - h)
//     (|Wd - h - l| <= 2^-21 |Wd| + 2^-24: below 2^-14 both pieces are subnormal fp16 values on the 2^-24 grid); h and l are
//     separate columns of the B operand (N = 2*NC), summed by the epilogue.
// Layout (no swizzle, K-major, the canonical "interleaved" UMMA layout): 8 rows x 16 bytes form a 128-byte core matrix; core
// matrices that are neighbours along K are LBO = 128 bytes apart, 8-row groups SBO = KC*128 bytes apart (KC = 16-byte chunks
// per K block).  The regroup pass (maintain_kernel.cuh) writes the count image in exactly this order, every work item padded
// to a multiple of 8 rows, so the rows of an item for one K block are ONE contiguous span of global memory: one
// cp.async.bulk per (item, K block), no tensor map, no register staging.
// Roles (warp-specialised, one persistent CTA per SM; everything between them goes through mbarriers):
//   warps 0-3   epilogue: TMEM -> registers (warp w owns lanes 32w..32w+31 = rows of the item), gap test, work list
//   warp 4      item fetcher: item records up to 8 items ahead (8 lanes, 8 records in flight), L2 prefetch of the item's rows
//   warp 5      MMA issuer (one elected lane)
//   warp 6      copy issuer: bulk copies of the count rows (ring of stages)
//   warps 7-22  table builders: mixture P in FP64, the 12 lg2 per (strain, sample), fp16 split, stores in B-operand order
// Results are those of the FFMA / mma.sync forms draw for draw (the error model below is charged instead of theirs).
#pragma once
#include <cuda_fp16.h>
#include "common.cuh"
#include "mu_agg_kernel.cuh"
#include "tau_kernel.cuh"
#include "tau_group_kernel.cuh"

// Sites per work item = M of the MMA.  (Same-box A/B at C3: 64-row items -- 2424 instead of 1823, in a ring of 4 stages -- cost
// 53 us against 45: the pass is bound by the table build, which is per item.)
#ifndef TC_ROWS
#define TC_ROWS 128
#endif
#define TC_M 128
#define TC_EPI_WARPS (TC_ROWS / 32) // TMEM lanes = rows of an item
#define TC_MAXRING 8
#ifndef TC_BUILD_WARPS
#define TC_BUILD_WARPS 16
#endif
#define TC_THREADS ((TC_EPI_WARPS + 3 + TC_BUILD_WARPS) * 32)
#define TC_NREC 8                   // item-record ring: how far the tickets run ahead

struct TauGroupTcParams {
    const unsigned char *img;  // count image: [K block][row group][KC chunks][8 rows][16 bytes], fp16x4 cells
    const int *img_site;       // [rows] site of an image row (-1: padding)
    const float *img_nsite;    // [rows] reads of the row's site, rounded up
    long long img_rg;          // row groups (of 8 rows) per K block of the image
    const double *gamma;       // [S][G]
    const double *eta;         // [16]
    const uint32_t *words;     // MT19937 words [V*G] or nullptr (Philox)
    int V, S, G;
    int SK, nkb;               // samples per K block (multiple of 4, <= 64), K blocks
    int NC;                    // table columns padded to a multiple of 8 (3G <= NC); N of the MMA = 2*NC
    TauGroup grp;
    unsigned long long *tier_counts;
    float *dbg;                // [V][3G] the sums D (log2 units) of every screened site, or nullptr (validation only)
    int early;                 // 1: control words, items and image are at least two grids old (see the kernel's prologue)
    // sharded chain: the MAP snapshot tau_star <- tau of the previous sweep, if its bookkeeping (two grids ago) raised the flag;
    // tau is not touched by this kernel, only by the list kernels after it
    const uint8_t *star_src; uint8_t *star_dst; size_t star_n; const int *star_flag;
};

struct TcRec { int slot, count, img0; unsigned int code_lo, code_hi; int pad[3]; };

// ------------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// (the suspend-time hint lets the hardware park the thread until the phase completes instead of returning after its short
// default time-out: without it the 20-odd polling threads of a CTA issued 12 of the kernel's 20 million warp instructions --
// ncu: issue slots 56 % busy, half of it try_wait + branch -- and the table builders competed with them for issue slots)
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(bar), "r"(parity), "r"(0x989680u) : "memory");
}
// A whole warp waits: ONE lane polls (31 fewer pollers competing with the working warps for issue slots), the others park
// at the warp barrier
__device__ __forceinline__ void mbar_wait_warp(uint32_t bar, uint32_t parity, int lane)
{
    if (lane == 0) mbar_wait(bar, parity);
    __syncwarp();
}
// TMA bulk copy global -> shared, completion counted in bytes on an mbarrier; the rows are read once per pass: evict first
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(pol) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// one lane of a converged warp
__device__ __forceinline__ bool tc_elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0u;
}
// D[tmem] (+)= A[smem desc] * B[smem desc], kind::f16 (FP16 operands, FP32 accumulation)
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// 8 consecutive FP32 columns of this thread's TMEM lane
__device__ __forceinline__ void tc_ld8(uint32_t taddr, float (&v)[8])
{
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// the registers of a tcgen05.ld are defined only after tcgen05.wait::ld: re-define them (empty volatile asm, ordered after the
// wait) so that no use can be scheduled ahead of it
__device__ __forceinline__ void tc_launder8(float (&v)[8])
{
    asm volatile("" : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]), "+f"(v[7]));
}

#define TC_WL_CAP 128
// a warp's staged work-list entries -> the global list
__device__ __forceinline__ void tc_flush_worklist(const TauGroup &grp, const uint2 *wl, int n, int lane)
{
    int pos = 0;
    if (lane == 0) pos = atomicAdd(grp.gctl + GC_NWORK, n);
    pos = __shfl_sync(DESMAN_FULL_MASK, pos, 0);
    for (int k = lane; k < n; k += 32) grp.work[pos + k] = wl[k];
    __syncwarp();
}

// shared-memory matrix descriptor, no swizzle, K-major (cute::UMMA::SmemDescriptor: start >> 4 in [0,14), LBO >> 4 in [16,30),
// SBO >> 4 in [32,46), version 1 in [46,48), layout type 0 = SWIZZLE_NONE in [61,64))
__device__ __forceinline__ uint64_t tc_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) | (1ull << 46);
}

// ------------------------------------------------------------------------------------------------ sizes (host + device)
struct TcLayout {
    int Sp, KC, N, acc_stride, tmem_cols;
    int nst, ntb, nacc;                                // ring depths: count stages, table buffers, accumulators
    size_t off_gT, off_eta, off_eta32, off_gT32, off_wl, off_rec, off_soff, off_bar, off_stage, off_table, stage_bytes, table_bytes, total;
};
__host__ __device__ static inline TcLayout tc_layout(int S, int G, int SK, int nkb, int NC)
{
    TcLayout L;
    L.Sp = SK * nkb;                                   // samples incl. padding
    L.KC = SK / 2;                                     // 16-byte chunks per K block (4 samples = 2 chunks = one MMA K step)
    L.N = 2 * NC;
    L.acc_stride = L.N <= 32 ? 32 : L.N <= 64 ? 64 : L.N <= 128 ? 128 : 256;
    L.nacc = L.acc_stride <= 64 ? 4 : 2;
    L.tmem_cols = L.nacc * L.acc_stride < 32 ? 32 : L.nacc * L.acc_stride;
    size_t o = 0;
    L.off_gT = o; o += sizeof(double) * (size_t)G * L.Sp;
    L.off_eta = o; o += sizeof(double) * 16;
    L.off_eta32 = o; o += sizeof(float) * 16;
    L.off_gT32 = o; o += sizeof(float) * (size_t)G * L.Sp;
    L.off_wl = o; o += sizeof(uint2) * TC_WL_CAP * TC_EPI_WARPS;
    L.off_rec = o; o += sizeof(TcRec) * TC_NREC;
    L.off_soff = o; o += sizeof(uint32_t) * 2 * TC_MAXRING;      // start / end of the live allocations of the stage ring
    o = (o + 15) & ~(size_t)15;
    L.off_bar = o; o += 8 * (2 * TC_NREC + 6 * TC_MAXRING);
    o = (o + 1023) & ~(size_t)1023;
    L.stage_bytes = (size_t)(TC_ROWS / 8) * L.KC * 128;
    L.table_bytes = (size_t)(L.N / 8) * L.KC * 128;
    // Ring depths from what shared memory is left (227 KB per CTA, 3 KB kept back): a stage is in flight from its copy to the
    // completion of its MMA (HBM latency + transfer + hand-overs).  With TC_ROWS < 128 the MMA still reads M = 128 rows: the
    // upper part comes from whatever follows (the next stage, or `pad`: finite fp16 data either way; its sums are never read).
    const size_t budget = (size_t)224 * 1024, pad = (TC_ROWS < TC_M) ? (size_t)(TC_M - TC_ROWS) / 8 * L.KC * 128 : 0;
    L.nst = 2; L.ntb = 2;
    auto fits = [&](int nst, int ntb) { return o + (size_t)nst * L.stage_bytes + pad + (size_t)ntb * L.table_bytes <= budget; };
    while (L.nst < 4 && fits(L.nst + 1, L.ntb)) L.nst++;
    if (fits(L.nst, 3)) L.ntb = 3;
    while (L.nst < TC_MAXRING && fits(L.nst + 1, L.ntb)) L.nst++;
    L.off_stage = o; o += (size_t)L.nst * L.stage_bytes + pad;
    L.off_table = o; o += (size_t)L.ntb * L.table_bytes;
    L.total = o;
    return L;
}

// -DTC_PROFILE (diagnosis build, tools/prof_tg.py): cycles every role spends in each of its waits, printed by a few CTAs
#ifdef TC_PROFILE
#define TCW(acc, stmt) do { const long long t_ = clock64(); stmt; acc += clock64() - t_; } while (0)
#define TCP(x) x
#else
#define TCW(acc, stmt) stmt
#define TCP(x)
#endif

// -DKPROF: per-item events of every role of a few CTAs (tools/kprof.py prints the pipeline of one CTA)
// (time stamps go to shared memory and are written out when the CTA is done: a record costs a global atomic, and one per
// event would stretch the very hand-overs it is meant to show)
// role-end stamps of a few CTAs (time + SM clock); -DTC_NO_TCE leaves only these
#ifdef KPROF
#define TCR(role) do { if (blockIdx.x % 37 == 0) krec_put(KP_TC_EVT, (int)blockIdx.x, 100 + (role), 0, gtimer(), (unsigned long long)clock64()); } while (0)
#else
#define TCR(role)
#endif
#if defined(KPROF) && !defined(TC_NO_TCE)
#define TC_TCE 1
#define TCE_ITEMS 32
// (plain stores by ONE writer per event: atomics from all the builder warps stretched the table builds they were timing)
#define TCE(role, item) do { if ((item) < TCE_ITEMS && blockIdx.x % 37 == 0) tce_s[role][item] = gtimer(); } while (0)
#else
#define TCE(role, item)
#endif

// ------------------------------------------------------------------------------------------------ kernel
__global__ void __launch_bounds__(TC_THREADS, 1) tau_group_tc_kernel(TauGroupTcParams p)
{
    extern __shared__ __align__(1024) unsigned char smem[];      // (no swizzle: the operands need 16-byte alignment only)
    const int S = p.S, G = p.G, SK = p.SK, nkb = p.nkb, NC = p.NC;
    const TcLayout L = tc_layout(S, G, SK, nkb, NC);
    const int Sp = L.Sp, KC = L.KC;
    double *gT = reinterpret_cast<double *>(smem + L.off_gT);           // [G][Sp]
    double *eta_s = reinterpret_cast<double *>(smem + L.off_eta);       // [16]
    float4 *eta32 = reinterpret_cast<float4 *>(smem + L.off_eta32);     // [4]
    float *gT32 = reinterpret_cast<float *>(smem + L.off_gT32);         // [G][Sp]
    TcRec *rec = reinterpret_cast<TcRec *>(smem + L.off_rec);           // [TC_NREC]
    const uint32_t bar0 = smem_u32(smem + L.off_bar);
    // barriers: rec_full[NREC] rec_empty[NREC] cnt_full/empty[nst] tab_full/empty[ntb] acc_full/empty[nacc] (TC_MAXRING slots each)
    const uint32_t rec_full = bar0, rec_empty = bar0 + 8 * TC_NREC, cnt_full = bar0 + 8 * (2 * TC_NREC), cnt_empty = cnt_full + 8 * TC_MAXRING,
                   tab_full = cnt_full + 8 * 2 * TC_MAXRING, tab_empty = cnt_full + 8 * 3 * TC_MAXRING, acc_full = cnt_full + 8 * 4 * TC_MAXRING,
                   acc_empty = cnt_full + 8 * 5 * TC_MAXRING;
    const uint32_t nst = (uint32_t)L.nst, ntb = (uint32_t)L.ntb, nacc = (uint32_t)L.nacc;
    const uint32_t stage0 = smem_u32(smem + L.off_stage), table0 = smem_u32(smem + L.off_table);
    __shared__ uint32_t tmem_base_s;
#ifdef TC_TCE
    __shared__ unsigned long long tce_s[10][TCE_ITEMS];   // roles: 0 copy issued, 1 mma committed, 2/3 table start first/last, 4/5 table done first/last, 6 acc seen, 7 item done
    for (int i = threadIdx.x; i < 10 * TCE_ITEMS; i += TC_THREADS) tce_s[i / TCE_ITEMS][i % TCE_ITEMS] = 0ull;
#endif
    __shared__ unsigned int gmin_bits, emin_bits;
    __shared__ int unnorm;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __shared__ __align__(8) unsigned long long par_ready_s;      // mbarrier: gamma / eta staged, gmin_bits / emin_bits / unnorm final
    const uint32_t par_ready = smem_u32(&par_ready_s);
    // ---- prologue that touches nothing of the preceding grid: barriers, zeroed operand buffers, tensor memory
    if (tid == 0) {
        for (int i = 0; i < TC_NREC; i++) { mbar_init(rec_full + 8 * i, 1); mbar_init(rec_empty + 8 * i, TC_EPI_WARPS); }
        for (int i = 0; i < TC_MAXRING; i++) {
            mbar_init(cnt_full + 8 * i, 1); mbar_init(cnt_empty + 8 * i, 1);
            mbar_init(tab_full + 8 * i, TC_BUILD_WARPS); mbar_init(tab_empty + 8 * i, 1);
            mbar_init(acc_full + 8 * i, 1); mbar_init(acc_empty + 8 * i, TC_EPI_WARPS);
        }
        mbar_init(par_ready, 1);
        gmin_bits = 0x7f800000u; emin_bits = 0x7f800000u; unnorm = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    {   // stages: rows beyond an item's own stay zero / finite; tables: padded columns and samples stay zero for good
        uint4 *z = reinterpret_cast<uint4 *>(smem + L.off_stage);
        const size_t n16 = (L.off_table + (size_t)L.ntb * L.table_bytes - L.off_stage) / 16;
        for (size_t i = tid; i < n16; i += TC_THREADS) z[i] = make_uint4(0u, 0u, 0u, 0u);
        for (int i = tid; i < G * Sp; i += TC_THREADS) gT32[i] = 1.0f;    // (finite operands for the builders' warm-up pass)
        if (tid < 16) reinterpret_cast<float *>(eta32)[tid] = 0.25f;
    }
    // The group control words, the item records and the count image were written by the maintenance launch at the start of
    // the sweep; inside desman_update at least two grids (statistics, draw) lie in between, so they are complete and visible
    // when this grid starts (the grid before us passed ITS dependency wait before it let us launch): with p.early everything
    // that needs only them -- the record fetch, the first copies, the L2 prefetches -- runs under the tail of the draw kernel.
    // Only gamma / eta (the draw kernel's output) and the MT19937 words need the wait: table builders and epilogue.
    if (!p.early) pdl_enter();
    KPROF_SCOPE(KP_TGM);
    if (p.star_flag && *p.star_flag) {
        const size_t n16 = p.star_n / 16, i0 = (size_t)blockIdx.x * TC_THREADS + tid, st = (size_t)gridDim.x * TC_THREADS;
        const uint4 *s4 = reinterpret_cast<const uint4 *>(p.star_src);
        uint4 *d4 = reinterpret_cast<uint4 *>(p.star_dst);
        for (size_t i = i0; i < n16; i += st) d4[i] = s4[i];
        for (size_t i = n16 * 16 + i0; i < p.star_n; i += st) p.star_dst[i] = p.star_src[i];
    }
    const int *gctl = p.grp.gctl;
    const int nitems = grp_active(gctl, 1) ? gctl[GC_NITEMS] : 0;
    if ((int)blockIdx.x >= nitems) {
        if (p.early) pdl_enter();
        return;
    }
    if (warp == 0) {   // TMEM: accumulators of N columns (allocation: power of two >= 32), owned by warp 0
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"((uint32_t)L.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_proxy_async();               // the zero fill above is read by the tensor core (async proxy)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    const uint32_t fullG = (G >= 32) ? 0xffffffffu : ((1u << G) - 1u);
    const uint32_t sbo = (uint32_t)KC * 128u;
    const int ncol = 3 * G;
    const float LN2 = 0.69314718f;
    // what the builders and the epilogue derive from gamma / eta once they are staged (after par_ready)
    struct TcConsts { bool fast_ok; float bn_scale; };
    auto consts = [&]() {
        TcConsts k;
        const float qmin = 0.99f * __uint_as_float(gmin_bits) * __uint_as_float(emin_bits);
        k.fast_ok = qmin >= TAU_QMIN && !unnorm;
        const float mq0 = fmaxf(1.0f, 1.0f - log2f(fmaxf(qmin, TAU_QMIN)));
        // per read, log2 units: the entry model of the FFMA form (relative parts, lg2.approx floors, lg2.approx and the lq - lP
        // rounding per unit of |lg2|) + [fp16 split 2^-22 + one FP32 accumulation step per 4 samples and piece, each charged
        // 2^-20 of the running magnitude + the hi + lo/1024 add] * max|Wd|, |Wd| <= mq0
#ifdef TC_TABLE_F64
        const float e_entry = TAU_C0 + (2.3841858e-7f + 5.9604645e-8f) * (2.0f * mq0) + TAU_CANCEL(G) / fmaxf(qmin, TAU_QMIN) + 5.9604645e-8f;   // (+ 2^-24: subnormal pieces)
#else
        // FP32 table build: a candidate q + eta gamma and the mixture P carry a relative error <= (G + 2) 2^-24 each (G - 1 FMA
        // steps over non-negative terms, two operand roundings, the last FMA; 2 more kept in reserve) -> log2(e) (2G + 8) 2^-24
        // on their lg2 difference; + the two lg2.approx floors; per unit of |lg2|: lg2.approx 2^-22, the difference 2^-24
        const float e_entry = (float)(2 * G + 8) * 5.9604645e-8f * 1.4426950f + 2.0f * 2.3841858e-7f +
                              (2.3841858e-7f + 5.9604645e-8f) * (2.0f * mq0) + 5.9604645e-8f;                   // (+ 2^-24: subnormal pieces)
#endif
        const float e_mma = (float)(Sp / 2 + 12) * 9.5367432e-7f;
        k.bn_scale = (e_entry + e_mma * mq0) * LN2 * 1.0001f;
        return k;
    };

    if (warp == TC_EPI_WARPS) {
        // =============================================================== item fetcher: item records, up to TC_NREC items ahead
        // Items are dealt round robin (item k of this CTA = blockIdx + k * gridDim; they are numbered longest first, so every
        // CTA gets one item of every size band).  A record is two loads of ~1 us: eight lanes fetch eight records at a time
        // (a single thread fetching one record after another -- or worse, drawing a ticket with a global atomic first -- was the
        // period of the whole pipeline: 2.2 us per item), start the item's rows on their way into L2, and lane 0 publishes them
        // in order as ring slots come free.
        const size_t kb_stride = (size_t)p.img_rg * KC * 128;
        TCP(long long w_rec = 0; long long n_it = 0; long long rows = 0; const long long t_all = clock64(););
        bool done = false;
        for (uint32_t base = 0; !done; base += 8) {
            int4 a = make_int4(0, 0, 0, 0), b = make_int4(0, 0, 0, 0);
            const long long it = (long long)blockIdx.x + (long long)(base + (uint32_t)lane) * (long long)gridDim.x;
            if (lane < 8 && it < (long long)nitems) {
                a = p.grp.items[2 * it]; b = p.grp.items[2 * it + 1];
                const uint32_t bytes = (((uint32_t)a.z + 7u) & ~7u) * (uint32_t)KC * 16u;
                for (int kb = 0; kb < nkb; kb++) l2_prefetch_row(p.img + (size_t)kb * kb_stride + (size_t)(b.z >> 3) * KC * 128, bytes);
            }
            for (int l = 0; l < 8 && !done; l++) {
                TcRec rc;
                rc.slot = __shfl_sync(DESMAN_FULL_MASK, a.x, l); rc.count = __shfl_sync(DESMAN_FULL_MASK, a.z, l);
                rc.img0 = __shfl_sync(DESMAN_FULL_MASK, b.z, l);
                rc.code_lo = (unsigned int)__shfl_sync(DESMAN_FULL_MASK, b.x, l); rc.code_hi = (unsigned int)__shfl_sync(DESMAN_FULL_MASK, b.y, l);
                rc.pad[0] = rc.pad[1] = rc.pad[2] = 0;
                const uint32_t i = base + (uint32_t)l, r = i % TC_NREC;
                if (lane == 0) {
                    TCW(w_rec, mbar_wait(rec_empty + 8 * r, ((i / TC_NREC) & 1u) ^ 1u));
                    rec[r] = rc;
                    mbar_arrive(rec_full + 8 * r);                     // (release: the record is visible to the waiters)
                }
                __syncwarp();
                if (rc.count == 0) done = true;                        // the sentinel: no more items
                TCP(n_it++; rows += rc.count;);
            }
        }
        if (lane == 0) TCR(0);
        TCP(if (lane == 0 && blockIdx.x % 37 == 0) printf("cta %3d items: %lld rows %lld total %lld wait rec_empty %lld\n", (int)blockIdx.x, n_it, rows, clock64() - t_all, w_rec););
    } else if (warp == TC_EPI_WARPS + 2) {
        // =============================================================== copy issuer: count rows of (item, K block) into the stage ring
        // The stage area is a BYTE ring, not a ring of full-size slots: an item's rows take what they need (8 rows x KC x 16 B
        // granules; the average item of C3 is 27 of the 64 KB a 128-row item takes), so three to five copies are in flight
        // instead of two.  Allocations are released in order (the MMA commits); an MMA always READS 128 rows from its start
        // (rows beyond the item's own are whatever follows: finite fp16, sums never read), so a start must leave a full-size
        // footprint below the end of the area.
        if (lane == 0) {
            const size_t kb_stride = (size_t)p.img_rg * KC * 128;
            volatile uint32_t *soff = reinterpret_cast<volatile uint32_t *>(smem + L.off_soff);
            const uint32_t R = nst * (uint32_t)L.stage_bytes, F = (uint32_t)L.stage_bytes;
            uint32_t u = 0, u_tail = 0, head = 0;
            TCP(long long w_rec = 0; long long w_cnt = 0; const long long t_all = clock64(););
            for (uint32_t i = 0;; i++) {
                const uint32_t r = i % TC_NREC;
                TCW(w_rec, mbar_wait(rec_full + 8 * r, (i / TC_NREC) & 1u));
                const int count = rec[r].count, img0 = rec[r].img0;
                if (count == 0) break;
                const uint32_t bytes = (((uint32_t)count + 7u) & ~7u) * (uint32_t)KC * 16u;
                for (int kb = 0; kb < nkb; kb++, u++) {
                    if (head + F > R) head = 0;
                    while (u_tail < u) {                                   // make room: the oldest live allocation first
                        const uint32_t s = u_tail % TC_MAXRING;
                        const bool full = (u - u_tail) == TC_MAXRING, overlap = soff[2 * s] < head + bytes && head < soff[2 * s + 1];
                        if (!full && !overlap) break;
                        TCW(w_cnt, mbar_wait(cnt_empty + 8 * s, (u_tail / TC_MAXRING) & 1u));
                        u_tail++;
                    }
                    const uint32_t cs = u % TC_MAXRING;
                    soff[2 * cs] = head; soff[2 * cs + 1] = head + bytes;
#ifdef TC_ABL_NOCOPY
                    mbar_arrive(cnt_full + 8 * cs);
#else
                    mbar_arrive_tx(cnt_full + 8 * cs, bytes);               // (release: the offsets are visible to the MMA issuer)
                    tma_bulk_g2s(stage0 + head, p.img + (size_t)kb * kb_stride + (size_t)(img0 >> 3) * KC * 128, bytes, cnt_full + 8 * cs);
#endif
                    head += bytes;
                    TCE(0, i);
                }
            }
            TCR(1);
            TCP(if (blockIdx.x % 37 == 0) printf("cta %3d copies: total %lld wait rec_full %lld cnt_empty %lld\n", (int)blockIdx.x, clock64() - t_all, w_rec, w_cnt););
        }
    } else if (warp == TC_EPI_WARPS + 1) {
        // =============================================================== MMA issuer
        // The whole warp walks the loop (lane 0 polls the barriers), ONE ELECTED lane issues: with the branch warp-uniform and
        // the issue under elect.sync the compiler emits straight uniform-datapath code; issued from inside `if (lane == 0)` every
        // tcgen05.mma sat in its own "elect an active thread and loop" construct (ELECT / BRA.U.ANY), and the 16 MMAs of an item
        // took the thread 1.3-1.8 us -- the longest stage of the pipeline (tools/ubench/umma_issue.cu: 0.5 us for a tight loop).
        const uint32_t idesc = (1u << 4) | ((uint32_t)(L.N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);   // F16 x F16 -> F32, K-major A and B
        const int nk = KC / 2;                                            // one K step = 16 fp16 = 2 chunks = 256 bytes = 16 units
        uint32_t u = 0;
        TCP(long long w_rec = 0; long long w_acc = 0; long long w_tab = 0; long long w_cnt = 0; const long long t_all = clock64(););
        for (uint32_t i = 0;; i++) {
            const uint32_t r = i % TC_NREC;
            TCW(w_rec, mbar_wait_warp(rec_full + 8 * r, (i / TC_NREC) & 1u, lane));
            if (rec[r].count == 0) break;
            const uint32_t as = i % nacc;
            TCW(w_acc, mbar_wait_warp(acc_empty + 8 * as, ((i / nacc) & 1u) ^ 1u, lane));
            const uint32_t d = tmem_base + as * (uint32_t)L.acc_stride;
            for (int kb = 0; kb < nkb; kb++, u++) {
                const uint32_t cs = u % TC_MAXRING, ts = u % ntb;
                TCW(w_tab, mbar_wait_warp(tab_full + 8 * ts, (u / ntb) & 1u, lane));
                TCW(w_cnt, mbar_wait_warp(cnt_full + 8 * cs, (u / TC_MAXRING) & 1u, lane));
                if (kb == 0 && lane == 0) TCE(8, i);
                tc_fence_after();
                if (tc_elect_one()) {
                    uint64_t a = tc_desc(stage0 + reinterpret_cast<const volatile uint32_t *>(smem + L.off_soff)[2 * cs], 128u, sbo);
                    uint64_t b = tc_desc(table0 + ts * (uint32_t)L.table_bytes, 128u, sbo);
#ifdef TC_ABL_MMA
                    tc_mma_f16(d, a, b, idesc, kb ? 1u : 0u);
#else
                    int k = 0;
                    if (nk >= 4) {
                        tc_mma_f16(d, a, b, idesc, kb ? 1u : 0u);
                        tc_mma_f16(d, a + 16, b + 16, idesc, 1u);
                        tc_mma_f16(d, a + 32, b + 32, idesc, 1u);
                        tc_mma_f16(d, a + 48, b + 48, idesc, 1u);
                        a += 64; b += 64;
                        for (k = 4; k + 4 <= nk; k += 4, a += 64, b += 64) {
                            tc_mma_f16(d, a, b, idesc, 1u);
                            tc_mma_f16(d, a + 16, b + 16, idesc, 1u);
                            tc_mma_f16(d, a + 32, b + 32, idesc, 1u);
                            tc_mma_f16(d, a + 48, b + 48, idesc, 1u);
                        }
                    }
                    for (; k < nk; k++, a += 16, b += 16) tc_mma_f16(d, a, b, idesc, (kb | k) ? 1u : 0u);
#endif
                    tc_commit(cnt_empty + 8 * cs);
                    tc_commit(tab_empty + 8 * ts);
                    if (kb == nkb - 1) tc_commit(acc_full + 8 * as);
                }
                __syncwarp();
            }
            if (lane == 0) TCE(1, i);
        }
        if (lane == 0) TCR(2);
        TCP(if (lane == 0 && blockIdx.x % 37 == 0) printf("cta %3d mma: total %lld wait rec_full %lld acc_empty %lld tab_full %lld cnt_full(after tab) %lld\n", (int)blockIdx.x, clock64() - t_all, w_rec, w_acc, w_tab, w_cnt););
    } else if (warp >= TC_EPI_WARPS + 3) {
        // =============================================================== table builders (TC_BUILD_WARPS warps, no cross-warp dependency)
        // A warp task = 8 strains x 4 samples; a lane = one (strain, sample): its base q[b] = P[b] - eta[cur][b] gamma (FP64, one
        // rounding) serves the 3 candidates x 4 bases = 12 table entries it writes.  The mixture P[s][b] = sum_h eta[tau_h][b]
        // gamma[s][h] (FP64) of the task's 4 samples is formed by lanes 0-15 (one (sample, base) each; lanes 16-31 mirror them)
        // and handed round by shuffles.  Stores: for a fixed candidate the 8 strains of a task hit 8 different rows mod 8 (3 is
        // coprime to 8), so a half warp writes 16 distinct 8-byte pieces of 128-byte core matrices: conflict-free.
        // The tasks of a warp are the same for every item (task = warp + k * TC_BUILD_WARPS): everything but the pattern --
        // sample and strain of the lane, the addresses of its gamma entries and of its 6 stores -- is worked out once, before the
        // item loop (the first version spent 300 of its 390 instructions per task on that arithmetic: ncu source page).
        const int bw = warp - (TC_EPI_WARPS + 3);
        const int gl = (lane >> 1) & 7, shf = lane & 1, sp = lane >> 4;
        const int ps = (lane >> 2) & 3, pbb = lane & 3;                   // the (sample, base) pair this lane forms P for
        const int mys = 2 * sp + shf;                                      // this lane's sample within the task
        const int ngo = (G + 7) >> 3, noct = NC >> 3, nquad = SK >> 2, ntask = ngo * nquad;
        constexpr int TPRE = 2;                                            // tasks per warp and K block kept in registers
        struct TaskC { const double *gTp; const double *gg; const float *gf; const float *gcol; uint32_t off[3]; int g2, g; bool p_ok, ok; };
        auto task_consts = [&](int task, int kb) {
            TaskC t;
            const int go = task % ngo, sq = task / ngo;
            const int s_p = kb * SK + 4 * sq + ps, g = 8 * go + gl, sl = 4 * sq + mys, sm = kb * SK + sl;
            t.p_ok = task < ntask && s_p < S;
            t.ok = task < ntask && g < G && sm < S;
            t.gTp = gT + (t.p_ok ? s_p : 0);
            t.gg = gT + (t.ok ? g * Sp + sm : 0);
            t.gf = gT32 + (t.ok ? g * Sp + sm : 0);
            t.g2 = 2 * (t.ok ? g : 0);
            t.g = t.ok ? g : 0;
            t.gcol = gT32 + (t.ok ? sm : 0);
#pragma unroll
            for (int j = 0; j < 3; j++) {
                const int n = 3 * g + j;
                t.off[j] = (uint32_t)(n >> 3) * sbo + (uint32_t)(sl >> 1) * 128u + (uint32_t)(n & 7) * 16u + (uint32_t)shf * 8u;
            }
            return t;
        };
        const bool pre_ok = nkb == 1;                                     // (with several K blocks the constants are re-derived per block)
        TaskC pre[TPRE];
#pragma unroll
        for (int k = 0; k < TPRE; k++) pre[k] = task_consts(bw + k * TC_BUILD_WARPS, 0);
        const uint32_t lo_off = (uint32_t)noct * sbo;                     // rows [NC, 2 NC): the remainders
#ifdef TC_TABLE_F64
        // one task: the lane's 12 entries of table `tab` for pattern `code`
        auto do_task = [&](const TaskC &t, uint64_t code, unsigned char *tab) {
            double Pv = 1.0;                                                // padding samples: finite logs
#ifdef TC_ABL_PLOOP
            if (t.p_ok) Pv = 0.3 + 1e-3 * (double)(code & 15ull);
            if (false) {
#else
            if (t.p_ok) {
#endif
                double Pa = 0.0, Pb = 0.0;                                  // two chains: even and odd strains
                const double *e = eta_s + pbb;
                const double *gp = t.gTp;
                uint64_t cc = code;
                int h = 0;
                for (; h + 1 < G; h += 2, cc >>= 4, gp += 2 * Sp) {
                    Pa = fma(e[4 * (int)(cc & 3ull)], gp[0], Pa);
                    Pb = fma(e[4 * (int)((cc >> 2) & 3ull)], gp[Sp], Pb);
                }
                if (h < G) Pa = fma(e[4 * (int)(cc & 3ull)], gp[0], Pa);
                Pv = Pa + Pb;
            }
            const float lv_ = lg2_fast((float)Pv);
#ifdef TC_ABL_SHFL
            const double P0 = Pv, P1 = Pv + 1e-9, P2 = Pv + 2e-9, P3 = Pv + 3e-9;
            const float l0 = lv_, l1 = lv_ + 1e-3f, l2 = lv_ + 2e-3f, l3 = lv_ + 3e-3f;
#else
            const double P0 = __shfl_sync(DESMAN_FULL_MASK, Pv, 4 * mys + 0), P1 = __shfl_sync(DESMAN_FULL_MASK, Pv, 4 * mys + 1),
                         P2 = __shfl_sync(DESMAN_FULL_MASK, Pv, 4 * mys + 2), P3 = __shfl_sync(DESMAN_FULL_MASK, Pv, 4 * mys + 3);
            const float l0 = __shfl_sync(DESMAN_FULL_MASK, lv_, 4 * mys + 0), l1 = __shfl_sync(DESMAN_FULL_MASK, lv_, 4 * mys + 1),
                        l2 = __shfl_sync(DESMAN_FULL_MASK, lv_, 4 * mys + 2), l3 = __shfl_sync(DESMAN_FULL_MASK, lv_, 4 * mys + 3);
#endif
            if (t.ok) {
                const int cur = (int)((code >> t.g2) & 3ull);
                const double2 *ecp = reinterpret_cast<const double2 *>(eta_s + 4 * cur);
                const double2 ec01 = ecp[0], ec23 = ecp[1];
                const double gg = *t.gg;
                const float gf = *t.gf;
#ifdef TC_ABL_F64
                const float q0 = fmaxf(fmaf(-(float)ec01.x, gf, l0), 0.f), q1 = fmaxf(fmaf(-(float)ec01.y, gf, l1), 0.f),
                            q2 = fmaxf(fmaf(-(float)ec23.x, gf, l2), 0.f), q3 = fmaxf(fmaf(-(float)ec23.y, gf, l3), 0.f);
                (void)gg;
#else
                const float q0 = fmaxf((float)fma(-ec01.x, gg, P0), 0.f), q1 = fmaxf((float)fma(-ec01.y, gg, P1), 0.f),
                            q2 = fmaxf((float)fma(-ec23.x, gg, P2), 0.f), q3 = fmaxf((float)fma(-ec23.y, gg, P3), 0.f);
#endif
#pragma unroll
                for (int j = 0; j < 3; j++) {
                    const float4 ea = eta32[(cur + 1 + j) & 3];
#ifdef TC_ABL_LG2
                    const float w0 = fmaf(ea.x, gf, q0) - l0, w1 = fmaf(ea.y, gf, q1) - l1, w2 = fmaf(ea.z, gf, q2) - l2, w3 = fmaf(ea.w, gf, q3) - l3;
#else
                    const float w0 = lg2_fast(fmaf(ea.x, gf, q0)) - l0, w1 = lg2_fast(fmaf(ea.y, gf, q1)) - l1,
                                w2 = lg2_fast(fmaf(ea.z, gf, q2)) - l2, w3 = lg2_fast(fmaf(ea.w, gf, q3)) - l3;
#endif
                    // h = the entry truncated to fp16's 11 significant bits (a mask: exact in fp16 unless |w| < 2^-14, where the
                    // conversion rounds it on the 2^-24 grid), l = fp16(w - h_truncated)
                    const float t0 = __uint_as_float(__float_as_uint(w0) & 0xffffe000u), t1 = __uint_as_float(__float_as_uint(w1) & 0xffffe000u),
                                t2 = __uint_as_float(__float_as_uint(w2) & 0xffffe000u), t3 = __uint_as_float(__float_as_uint(w3) & 0xffffe000u);
                    const __half2 h01 = __floats2half2_rn(t0, t1), h23 = __floats2half2_rn(t2, t3);
                    const __half2 l01 = __floats2half2_rn(w0 - t0, w1 - t1), l23 = __floats2half2_rn(w2 - t2, w3 - t3);
                    uint2 hv, lv;
                    hv.x = *reinterpret_cast<const uint32_t *>(&h01); hv.y = *reinterpret_cast<const uint32_t *>(&h23);
                    lv.x = *reinterpret_cast<const uint32_t *>(&l01); lv.y = *reinterpret_cast<const uint32_t *>(&l23);
                    *reinterpret_cast<uint2 *>(tab + t.off[j]) = hv;                 // rows [0, NC): h
                    *reinterpret_cast<uint2 *>(tab + t.off[j] + lo_off) = lv;        // rows [NC, 2 NC): l
                }
            }
        };
#else
        // one task: the lane's 12 entries of table `tab` for pattern `code`.
        // ALL IN FP32, and on purpose: FP64 arithmetic and FP64 -> FP32 conversions stall while tcgen05.mma are in flight
        // (tools/ubench/umma_interfere.cu: DFMA and F2F.F32.F64 run 2.3 x slower next to a stream of these MMAs, FFMA / MUFU / SHFL
        // do not) -- with the FP64 mixture of the first version every builder warp sat out a 3.3 us stall once or twice per launch.
        // The base q[b] = sum over the OTHER strains of eta[tau_h][b] gamma[s][h] is formed directly (a sum of non-negative terms:
        // no cancellation, relative error <= (G+1) 2^-24 incl. the operand roundings) instead of P - eta[cur][b] gamma[s][g]; the
        // mixture the entries are taken relative to is P[b] = q[b] + eta[cur][b] gamma[s][g].
        auto do_task = [&](const TaskC &t, uint64_t code, unsigned char *tab) {
            if (!t.ok) return;
            const int cur = (int)((code >> t.g2) & 3ull);
            float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
            {
                const float *gp = t.gcol;
                uint64_t cc = code;
                for (int h = 0; h < G; h++, cc >>= 2, gp += Sp) {
                    const float4 e = eta32[(int)(cc & 3ull)];
                    const float gm = (h == t.g) ? 0.f : *gp;                 // (adds exactly 0 for the lane's own strain)
                    q0 = fmaf(e.x, gm, q0); q1 = fmaf(e.y, gm, q1); q2 = fmaf(e.z, gm, q2); q3 = fmaf(e.w, gm, q3);
                }
            }
            const float gf = *t.gf;
            const float4 ec = eta32[cur];
            const float l0 = lg2_fast(fmaf(ec.x, gf, q0)), l1 = lg2_fast(fmaf(ec.y, gf, q1)), l2 = lg2_fast(fmaf(ec.z, gf, q2)),
                        l3 = lg2_fast(fmaf(ec.w, gf, q3));
#pragma unroll
            for (int j = 0; j < 3; j++) {
                const float4 ea = eta32[(cur + 1 + j) & 3];
                const float w0 = lg2_fast(fmaf(ea.x, gf, q0)) - l0, w1 = lg2_fast(fmaf(ea.y, gf, q1)) - l1,
                            w2 = lg2_fast(fmaf(ea.z, gf, q2)) - l2, w3 = lg2_fast(fmaf(ea.w, gf, q3)) - l3;
                // h = the entry truncated to fp16's 11 significant bits (a mask: exact in fp16 unless |w| < 2^-14, where the
                // conversion rounds it on the 2^-24 grid), l = fp16(w - h_truncated)
                const float t0 = __uint_as_float(__float_as_uint(w0) & 0xffffe000u), t1 = __uint_as_float(__float_as_uint(w1) & 0xffffe000u),
                            t2 = __uint_as_float(__float_as_uint(w2) & 0xffffe000u), t3 = __uint_as_float(__float_as_uint(w3) & 0xffffe000u);
                const __half2 h01 = __floats2half2_rn(t0, t1), h23 = __floats2half2_rn(t2, t3);
                const __half2 l01 = __floats2half2_rn(w0 - t0, w1 - t1), l23 = __floats2half2_rn(w2 - t2, w3 - t3);
                uint2 hv, lv;
                hv.x = *reinterpret_cast<const uint32_t *>(&h01); hv.y = *reinterpret_cast<const uint32_t *>(&h23);
                lv.x = *reinterpret_cast<const uint32_t *>(&l01); lv.y = *reinterpret_cast<const uint32_t *>(&l23);
                *reinterpret_cast<uint2 *>(tab + t.off[j]) = hv;                 // rows [0, NC): h
                *reinterpret_cast<uint2 *>(tab + t.off[j] + lo_off) = lv;        // rows [NC, 2 NC): l
            }
        };
#endif
        const int bt = tid - (TC_EPI_WARPS + 3) * 32, nbt = TC_BUILD_WARPS * 32;
        if (p.early) {
            // warm-up pass over table buffer 0 with placeholder operands (every entry it writes is rewritten by item 0's build):
            // the first table of a launch cost 3-4.7 us instead of 0.9 on instructions that came from HBM (L2 is cold at the
            // start of a sweep); here that happens under the tail of the draw kernel
            unsigned char *tab0 = smem + L.off_table;
            if (pre_ok) {
#pragma unroll
                for (int k = 0; k < TPRE; k++)
                    if (bw + k * TC_BUILD_WARPS < ntask) do_task(pre[k], 0ull, tab0);
            } else {
                for (int task = bw; task < ntask; task += TC_BUILD_WARPS) do_task(task_consts(task, 0), 0ull, tab0);
            }
            pdl_enter();
        }
        {   // gamma / eta of this sweep (the draw kernel's output) -> shared memory; smallest entries; normalisation check
            float gmin_l = __int_as_float(0x7f800000);
            for (int i = bt; i < G * Sp; i += nbt) {
                const int g = i / Sp, s2 = i - g * Sp;
                const double x = (s2 < S) ? p.gamma[(size_t)s2 * G + g] : 0.0;
                gT[i] = x;
                gT32[i] = (float)x;
                if (s2 < S && x > 0.0) gmin_l = fminf(gmin_l, (float)x);
            }
            atomicMin(&gmin_bits, __float_as_uint(gmin_l));
            if (bt < 16) {
                eta_s[bt] = p.eta[bt];
                reinterpret_cast<float *>(eta32)[bt] = (float)p.eta[bt];
                atomicMin(&emin_bits, __float_as_uint(fmaxf((float)p.eta[bt], 0.f)));
            }
            asm volatile("bar.sync 1, %0;" ::"r"(nbt) : "memory");
            // unnormalised input (rows of gamma or eta summing to more than 1): no screening, every site goes to the per-site kernel
            for (int s2 = bt; s2 < S + 4; s2 += nbt) {
                double t = 0.0;
                if (s2 < S) for (int g = 0; g < G; g++) t += gT[g * Sp + s2];
                else for (int b = 0; b < 4; b++) t += eta_s[4 * (s2 - S) + b];
                if (!(t <= 1.0001)) unnorm = 1;
            }
            asm volatile("bar.sync 1, %0;" ::"r"(nbt) : "memory");
            if (bt == 0) mbar_arrive(par_ready);
        }
        const bool fast_ok = consts().fast_ok;
#ifdef KPROF
        if (bt == 0) krec_put(KP_TGM_PRO, (int)blockIdx.x, 0, nitems, gtimer(), 0);
#endif
        uint32_t u = 0;
        if (lane == 0 && bw == 0) TCR(7);
        TCP(long long w_rec = 0; long long w_tab = 0; long long t_work = 0; long long t_task = 0; long long t_fence = 0; const long long t_all = clock64(););
        for (uint32_t i = 0;; i++) {
            const uint32_t r = i % TC_NREC;
            TCW(w_rec, mbar_wait_warp(rec_full + 8 * r, (i / TC_NREC) & 1u, lane));
            const TcRec rc = rec[r];
            if (rc.count == 0) break;
            const uint64_t code = ((uint64_t)rc.code_hi << 32) | rc.code_lo;
            for (int kb = 0; kb < nkb; kb++, u++) {
                const uint32_t ts = u % ntb;
                TCW(w_tab, mbar_wait_warp(tab_empty + 8 * ts, ((u / ntb) & 1u) ^ 1u, lane));
                TCP(const long long tw0 = clock64(););
                if (lane == 0 && bw == 0) TCE(2, i);
                if (lane == 0 && bw == TC_BUILD_WARPS - 1) TCE(3, i);
                unsigned char *tab = smem + L.off_table + ts * L.table_bytes;
#ifndef TC_ABL_TABLE
                if (fast_ok) {                                           // (otherwise the tables stay zero and every site is listed)
                    if (pre_ok) {
#pragma unroll
                        for (int k = 0; k < TPRE; k++)
                            if (bw + k * TC_BUILD_WARPS < ntask) do_task(pre[k], code, tab);
                        for (int task = bw + TPRE * TC_BUILD_WARPS; task < ntask; task += TC_BUILD_WARPS) do_task(task_consts(task, 0), code, tab);
                    } else {
                        for (int task = bw; task < ntask; task += TC_BUILD_WARPS) do_task(task_consts(task, kb), code, tab);
                    }
                }
#endif
                TCP(const long long tw1 = clock64(););
#ifndef TC_ABL_NOFENCE
                fence_proxy_async();                                     // generic-proxy stores -> async-proxy reads of the MMA
#endif
                TCP(const long long tw2 = clock64(););
                __syncwarp();
                if (lane == 0) mbar_arrive(tab_full + 8 * ts);
                if (lane == 0 && bw == 0) TCE(4, i);
                if (lane == 0 && bw == TC_BUILD_WARPS - 1) TCE(5, i);
                TCP(t_work += clock64() - tw0; t_task += tw1 - tw0; t_fence += tw2 - tw1;);
            }
        }
        if (lane == 0 && bw == 0) TCR(3);
        TCP(if (lane == 0 && bw == 0 && blockIdx.x % 37 == 0) printf("cta %3d tables: total %lld wait rec_full %lld tab_empty %lld work %lld (task %lld fence %lld)\n", (int)blockIdx.x, clock64() - t_all, w_rec, w_tab, t_work, t_task, t_fence););
    } else {
        // =============================================================== epilogue (warps 0-3: TMEM lanes 32 w .. 32 w + 31)
        unsigned int n_decided = 0;
        const int row = warp * 32 + lane;
        uint2 *wl = reinterpret_cast<uint2 *>(smem + L.off_wl) + warp * TC_WL_CAP;
        int wl_n = 0;
        TCP(long long w_rec = 0; long long w_acc = 0; const long long t_all = clock64(););
        if (p.early) pdl_enter();                                            // (MT19937 words; orphan marks of the previous sweep)
        mbar_wait_warp(par_ready, 0u, lane);
        const TcConsts kc = consts();
        const bool fast_ok = kc.fast_ok;
        const float bn_scale = kc.bn_scale;
        // The site and read count of a row are two global loads (~1 us from HBM): they are issued ONE ITEM AHEAD (the record of
        // item i+1 is waited for before the sums of item i), so that they are not on the per-item critical path of this role.
        TCW(w_rec, mbar_wait_warp(rec_full, 0u, lane));
        TcRec rc = rec[0];
        int vraw = 0;
        float nk = 0.f;
        if (rc.count && row < rc.count) { vraw = p.img_site[rc.img0 + row]; nk = p.img_nsite[rc.img0 + row]; }
        for (uint32_t i = 0;; i++) {
            const uint32_t r = i % TC_NREC;
            if (rc.count == 0) break;
            const uint32_t r1 = (i + 1) % TC_NREC;
            TCW(w_rec, mbar_wait_warp(rec_full + 8 * r1, ((i + 1) / TC_NREC) & 1u, lane));
            const TcRec rcn = rec[r1];
            int vraw_n = 0;
            float nk_n = 0.f;
            if (rcn.count && row < rcn.count) { vraw_n = p.img_site[rcn.img0 + row]; nk_n = p.img_nsite[rcn.img0 + row]; }
            const bool have = row < rc.count;
            const bool orphan = vraw < 0;                                   // the site has left this group (tau_sample_kernel marks the row)
            const int vown = orphan ? ~vraw : vraw;
            bool zero_word = false;
            if (have && p.words) {
                const uint32_t *w = p.words + (size_t)vown * G;
                for (int g = 0; g < G; g++) zero_word |= (w[g] == 0u);      // u == 0 (c_sample_tau.c:174): reference-order path
            }
            const uint32_t as = i % nacc;
            TCW(w_acc, mbar_wait_warp(acc_full + 8 * as, (i / nacc) & 1u, lane));
            if (warp == 0 && lane == 0) TCE(6, i);
            tc_fence_after();
            const uint32_t t0 = tmem_base + as * (uint32_t)L.acc_stride + ((uint32_t)(warp * 32) << 16);
            const float bn = nk * bn_scale + 1e-6f;
            // 24 columns at a time, all six loads in flight before the one wait; a column is "open" unless its candidate trails the
            // current base by more than the gap after the bound; a strain is decided "stay" iff its three columns are closed
            unsigned long long open = 0ull;
#ifdef TC_ABL_EPI
            for (int c0 = NC; c0 < NC; c0 += 24) {
#else
            for (int c0 = 0; c0 < NC; c0 += 24) {
#endif
                float hi[24], lo[24];
#pragma unroll
                for (int q = 0; q < 3; q++) {
                    if (c0 + 8 * q < NC) {                                  // (warp-uniform: the loads are .sync.aligned)
                        tc_ld8(t0 + (uint32_t)(c0 + 8 * q), reinterpret_cast<float(&)[8]>(hi[8 * q]));
                        tc_ld8(t0 + (uint32_t)(NC + c0 + 8 * q), reinterpret_cast<float(&)[8]>(lo[8 * q]));
                    } else {
#pragma unroll
                        for (int e = 0; e < 8; e++) { hi[8 * q + e] = -1.0e30f; lo[8 * q + e] = 0.f; }
                    }
                }
                tc_wait_ld();
#pragma unroll
                for (int q = 0; q < 3; q++) {
                    tc_launder8(reinterpret_cast<float(&)[8]>(hi[8 * q]));
                    tc_launder8(reinterpret_cast<float(&)[8]>(lo[8 * q]));
                }
#pragma unroll
                for (int e = 0; e < 24; e++) {
                    const float d = hi[e] + lo[e];
                    if (!(d * LN2 + bn < -TAU_GAP)) open |= 1ull << (c0 + e);
                    if (p.dbg && have && c0 + e < ncol) p.dbg[(size_t)vown * ncol + c0 + e] = d;
                }
            }
            uint32_t mask = 0;
            for (int g = 0; g < G; g++) if ((open >> (3 * g)) & 7ull) mask |= 1u << g;
            // the sums are in registers: hand the accumulator and the record back
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { mbar_arrive(acc_empty + 8 * as); mbar_arrive(rec_empty + 8 * r); }
            bool push = false;
            if (have) {
                if (!fast_ok || orphan || zero_word) mask = fullG;            // orphan: its pattern is not this group's
                push = mask != 0u;
                if (!push) n_decided += (unsigned int)G;
            }
            // work-list entries are staged per warp in shared memory and flushed in batches: one global atomic (~1 us with its
            // return value) per ~100 entries instead of one per item
            const unsigned int bal = __ballot_sync(DESMAN_FULL_MASK, push);
            if (bal) {
                if (wl_n + __popc(bal) > TC_WL_CAP) { tc_flush_worklist(p.grp, wl, wl_n, lane); wl_n = 0; }
                if (push) wl[wl_n + __popc(bal & ((1u << lane) - 1u))] = make_uint2((unsigned int)vown, mask);
                wl_n += __popc(bal);
                __syncwarp();
            }
            if (warp == 0 && lane == 0) TCE(7, i);
            rc = rcn; vraw = vraw_n; nk = nk_n;
        }
        if (lane == 0 && warp == 0) TCR(4);
        if (wl_n) tc_flush_worklist(p.grp, wl, wl_n, lane);
        n_decided = (unsigned int)warp_sum_u64((unsigned long long)n_decided);
        if (lane == 0 && n_decided && p.tier_counts) atomicAdd(p.tier_counts, (unsigned long long)n_decided);
        if (lane == 0 && warp == 0) TCR(5);
        TCP(if (lane == 0 && warp == 0 && blockIdx.x % 37 == 0) printf("cta %3d epilogue: total %lld wait rec_full %lld acc_full %lld\n", (int)blockIdx.x, clock64() - t_all, w_rec, w_acc););
    }
    // ---- teardown: every role is done with tensor memory
    tc_fence_before();
    __syncthreads();
    if (tid == 0) TCR(6);
#ifdef TC_TCE
    if (tid == 0 && blockIdx.x % 37 == 0)
        for (int it = 0; it < TCE_ITEMS; it++)
            for (int ro = 0; ro < 10; ro++)
                if (tce_s[ro][it] != 0ull && tce_s[ro][it] != ~0ull) krec_put(KP_TC_EVT, (int)blockIdx.x, ro, it, tce_s[ro][it], 0);
#endif
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)L.tmem_cols) : "memory");
    }
}

// desman_b200/csrc/tau_group_kernel.cuh -- K1g: pattern-grouped screening pass of the tau Gibbs update
// (c_sample_tau.c:130-188), plus the bookkeeping that keeps the site groups current.
//
// Why.  In the per-site kernel (tau_kernel.cuh) every (v,g) step evaluates 12*S logs: 6e8 MUFU operations per sweep at
// BASELINE config C3, a ~150 us floor, ten times the HBM floor of the count tensor.  But the log terms of a step,
//     Wd[g][a][s][b] = log(P_sb - eta[cur_g][b]*gamma[s][g] + eta[a][b]*gamma[s][g]) - log P_sb,    P_sb = sum_h eta[tau_h][b]*gamma[s][h],
// depend on the site only through its haplotype pattern tau_v (the key of the persistent pattern table, mu_agg_kernel.cuh),
// and a converged chain has ~12*2^G patterns for V sites.  So the sites are grouped by pattern, the table Wd is built ONCE
// per group (in shared memory, same FP32/MUFU arithmetic and error model as tier 1 of the per-site kernel), and the
// log-likelihood differences of a site become a small dense contraction of its count row with that table:
//     D[v][g][a] = sum_{s,b} n[v][s][b] * Wd[g][a][s][b]           (3*G*4*S FFMA per site, no transcendental)
// As long as the pattern of a site does not change during its walk over the strains (no flip), the D of all G steps
// come from the same table.  A step whose current base leads every other candidate by more than TAU_GAP nats after the rigorous
// error bound is decided ("stay", exactly as tier 1 of the per-site kernel); a site with any undecided step is appended to a
// work list together with the bit mask of those steps, and the per-site kernel then walks only the listed sites, skipping
// the decided steps until the first flip.  Results are therefore identical to the per-site kernel's, draw for draw.
//
// Mapping.  One warp per work item (pattern, <= TG_ITEM_SITES sites); the warp owns a private Wd table in shared memory.  A pass
// contracts 16 sites: lane = (r, o), r = 4 site rows of 4 sites each (register tile), o = 8 sample groups (s = o + 8k).
// Every LDS.128 of the table feeds 16 FFMA (measured on B200: LDS.128 costs 4 LSU cycles per warp unless all 32 lanes
// read the same address, tools/ubench; 4 sites per lane balance the LSU and FMA pipes).  Count rows are read as 128-bit
// cells, 8 consecutive cells (one 128-B line) per site row and instruction, after a TMA bulk prefetch into L2 one pass
// ahead (cp.async.bulk.prefetch.L2).  Partial sums are combined over the 8 sample groups with a transposing
// butterfly (3 levels; every lane ends with the complete sums of GB/2 (site, strain) pairs).
#pragma once
#include "common.cuh"
#include "mu_agg_kernel.cuh"
#include "tau_kernel.cuh"

#ifndef TG_ITEM_SITES
#define TG_ITEM_SITES 64     // sites per work item (A/B at C3: 256: 66 us, 128: 60.8, 64: 58.3, 48: 63.4, 32: 68.3 for the screening pass)
#endif
#define TG_PASS_SITES 16
#define TG_MAX_WARPS 4

struct TauGroup {
    int *site_slot;     // [V] table slot of the site's current pattern (mu_aggregate / agg_move_site keep it)
    int *order;         // [V] sites of multi-site patterns, sorted by slot
    int *singles;       // [V] sites that are alone in their pattern (straight to the per-site kernel)
    int4 *items;        // two words per work item: {slot, first row, rows, 0}, {pattern code lo, hi, 0, 0}
    uint2 *work;        // [V] {site, mask of undecided strains}
    int *slot_cnt, *slot_fill, *slot_start, *slot_item;   // [cap_slots] regroup scratch
    int *gctl;          // [GC_COUNT]
};

struct TauGroupParams {
    const float4 *countsf;   // [rows][S] counts as FP32 (exact: counts <= 2^24), rows in group order: row pos <-> site order[pos]
    const float *nsite;      // [rows] reads of the row's site, rounded up
    const double *gamma;     // [S][G]
    const double *eta;       // [16]
    const uint32_t *words;   // MT19937 words [V*G] (a zero word means u == 0: per-site kernel), or nullptr (Philox: u > 0)
    int V, S, G;
    TauGroup grp;
    unsigned long long *tier_counts;
};

__device__ __forceinline__ float4 ld_countsf(const float4 *p)
{
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}

// TMA bulk prefetch of a contiguous block of count rows into L2 (one instruction per warp: the address is warp-uniform)
__device__ __forceinline__ void l2_prefetch_row(const void *p, uint32_t bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// one level of the transposing butterfly: N values per lane -> N/2; lanes with (lane & mask) keep the upper half
template <int N>
__device__ __forceinline__ void tg_reduce_level(float (&v)[N], bool upper, int mask)
{
#pragma unroll
    for (int i = 0; i < N / 2; i++) {
        const float send = upper ? v[i] : v[i + N / 2];
        const float keep = upper ? v[i + N / 2] : v[i];
        v[i] = keep + __shfl_xor_sync(DESMAN_FULL_MASK, send, mask);
    }
}

// Table layout: Wd[s][row] float4 (b = 0..3), row = g*3 + j, sample-major with one float4 of padding per sample so that
// the 8 lanes of a quarter warp (consecutive s) hit 8 different bank groups and the rows of one sample are immediate
// offsets of one base address.  Sp = S rounded up to 16 (the sample loop is unrolled by two 8-sample steps).
static inline size_t tg_table_bytes(int S, int G, int GB)
{
    const size_t Sp = (size_t)((S + 15) & ~15), nGB = (size_t)((G + GB - 1) / GB);
    return Sp * (nGB * GB * 3 + 1) * sizeof(float4);
}
static inline size_t tg_shared_bytes(int S, int G)
{
    const size_t Sp = (size_t)((S + 15) & ~15);
    return (size_t)G * Sp * (sizeof(double) + sizeof(float)) + 16 * sizeof(double) + 4 * sizeof(float4);
}

// acc[i][j] += n[i] . Wd[s][j]   for the 4 sites of the lane and the NJ rows of one strain block
template <int NJ>
__device__ __forceinline__ void tg_step(float (&acc)[4 * NJ], const float4 (&n)[4], const float4 *__restrict__ Ws)
{
#pragma unroll
    for (int j = 0; j < NJ; j++) {
        const float4 w = Ws[j];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            float a = acc[i * NJ + j];
            a = fmaf(n[i].x, w.x, a); a = fmaf(n[i].y, w.y, a); a = fmaf(n[i].z, w.z, a); a = fmaf(n[i].w, w.w, a);
            acc[i * NJ + j] = a;
        }
    }
}

template <int GB>
__global__ void __launch_bounds__(TG_MAX_WARPS * 32, 1) tau_group_kernel(TauGroupParams p)
{
    pdl_enter();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int *gctl = p.grp.gctl;
    if (!grp_active(gctl, 0)) return;
    const int S = p.S, G = p.G;
    const int Sp = (S + 15) & ~15, nk = Sp >> 3;
    const int nGB = (G + GB - 1) / GB;
    constexpr int NJ = GB * 3;            // table rows of one strain block
    constexpr int NV = 4 * NJ;            // partial sums per lane and pass
    constexpr int PP = GB / 2;            // (site, strain) pairs a lane ends up with
    const int STR = nGB * NJ + 1;         // float4 per sample (odd: conflict-free)

    double *gT = reinterpret_cast<double *>(smem_raw);                    // [G][Sp]
    double *eta_s = gT + (size_t)G * Sp;                                  // [16]
    float4 *eta32 = reinterpret_cast<float4 *>(eta_s + 16);               // [4]
    float *gT32 = reinterpret_cast<float *>(eta32 + 4);                   // [G][Sp]
    float4 *Wall = reinterpret_cast<float4 *>(gT32 + (size_t)G * Sp);     // [warps][Sp][STR]
    __shared__ unsigned int gmin_bits, emin_bits;

    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    if (threadIdx.x == 0) { gmin_bits = 0x7f800000u; emin_bits = 0x7f800000u; }
    __syncthreads();
    float gmin_l = __int_as_float(0x7f800000);
    for (int i = threadIdx.x; i < G * Sp; i += blockDim.x) {
        const int g = i / Sp, s = i - g * Sp;
        const double x = (s < S) ? p.gamma[(size_t)s * G + g] : 0.0;
        gT[i] = x;
        gT32[i] = (float)x;
        if (s < S && x > 0.0) gmin_l = fminf(gmin_l, (float)x);     // masked strains (gamma == 0): q = P there
    }
    atomicMin(&gmin_bits, __float_as_uint(gmin_l));
    if (threadIdx.x < 16) {
        eta_s[threadIdx.x] = p.eta[threadIdx.x];
        reinterpret_cast<float *>(eta32)[threadIdx.x] = (float)p.eta[threadIdx.x];
        atomicMin(&emin_bits, __float_as_uint(fmaxf((float)p.eta[threadIdx.x], 0.f)));
    }
    __syncthreads();
    // launch-level a-priori error bound per read (log2 units), same model as tier 1 of tau_sample_kernel:
    // every table entry is lg2(q) - lg2(P) with q, P >= qmin; the FP32 sum of one output passes through 4*nk FFMA and
    // 3 shuffle adds, and the subtraction lq - lP adds one more rounding per unit of |lg2|
    const float qmin = 0.99f * __uint_as_float(gmin_bits) * __uint_as_float(emin_bits);
    const bool fast_ok = qmin >= TAU_QMIN;
    const float mq0 = fmaxf(1.0f, 1.0f - log2f(fmaxf(qmin, TAU_QMIN)));
    const float c1 = 2.3841858e-7f + (float)(4 * nk + 9) * 5.9604645e-8f;
    const float e_read = TAU_C0 + c1 * (2.0f * mq0) + TAU_CANCEL(G) / fmaxf(qmin, TAU_QMIN);
    const float LN2 = 0.69314718f;
    const float bn_scale = e_read * LN2 * 1.0001f;

    float4 *W = Wall + (size_t)wib * Sp * STR;
    const int r = lane >> 3, o = lane & 7;
    const bool owner = (o & 1) == 0;
    const int isl = o >> 1;                                   // site (within the row) this lane decides for
    const uint32_t fullG = (G >= 32) ? 0xffffffffu : ((1u << G) - 1u);
    const uint32_t row_bytes = (uint32_t)S * 16u;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    unsigned int n_decided = 0;

    while (true) {
        int it = 0;
        if (lane == 0) it = atomicAdd(p.grp.gctl + GC_CURSOR, 1);
        it = __shfl_sync(DESMAN_FULL_MASK, it, 0);
        if (it >= gctl[GC_NITEMS]) break;
        const int4 item = p.grp.items[2 * it], item1 = p.grp.items[2 * it + 1];
        const int slot = item.x, begin = item.y, count = item.z;
        const uint64_t code = ((uint64_t)(unsigned int)item1.y << 32) | (unsigned int)item1.x;
        const float4 *rows = p.countsf + (size_t)begin * S;        // the item's count rows are contiguous
        l2_prefetch_row(rows, (uint32_t)min(count, TG_PASS_SITES) * row_bytes);

        // ---- the pattern's table: Wd[s][g*3+j][b] = lg2(base_sb + eta[a_j][b]*gamma[s][g]) - lg2(P_sb), a_j = (cur_g+1+j)&3
        __syncwarp();
        for (int s = lane; s < Sp; s += 32) {
            float4 *Ws = W + (size_t)s * STR;
            if (s < S && fast_ok) {
                double P0 = 0.0, P1 = 0.0, P2 = 0.0, P3 = 0.0;
                for (int h = 0; h < G; h++) {
                    const double2 *e = reinterpret_cast<const double2 *>(eta_s + 4 * code_get(code, h));
                    const double2 e01 = e[0], e23 = e[1];
                    const double gm = gT[h * Sp + s];
                    P0 = fma(e01.x, gm, P0); P1 = fma(e01.y, gm, P1); P2 = fma(e23.x, gm, P2); P3 = fma(e23.y, gm, P3);
                }
                const float l0 = lg2_fast((float)P0), l1 = lg2_fast((float)P1), l2 = lg2_fast((float)P2), l3 = lg2_fast((float)P3);
                for (int g = 0; g < G; g++) {
                    const int cur = code_get(code, g);
                    const double2 *ecp = reinterpret_cast<const double2 *>(eta_s + 4 * cur);
                    const double2 ec01 = ecp[0], ec23 = ecp[1];
                    const double gg = gT[g * Sp + s];
                    const float gf = gT32[g * Sp + s];
                    const float q0 = fmaxf((float)fma(-ec01.x, gg, P0), 0.f), q1 = fmaxf((float)fma(-ec01.y, gg, P1), 0.f),
                                q2 = fmaxf((float)fma(-ec23.x, gg, P2), 0.f), q3 = fmaxf((float)fma(-ec23.y, gg, P3), 0.f);
#pragma unroll
                    for (int j = 0; j < 3; j++) {
                        const float4 ea = eta32[(cur + 1 + j) & 3];
                        float4 w;
                        w.x = lg2_fast(fmaf(ea.x, gf, q0)) - l0;
                        w.y = lg2_fast(fmaf(ea.y, gf, q1)) - l1;
                        w.z = lg2_fast(fmaf(ea.z, gf, q2)) - l2;
                        w.w = lg2_fast(fmaf(ea.w, gf, q3)) - l3;
                        Ws[g * 3 + j] = w;
                    }
                }
                for (int t = G * 3; t < nGB * NJ; t++) Ws[t] = zero4;
            } else {
                for (int t = 0; t < nGB * NJ; t++) Ws[t] = zero4;
            }
        }
        __syncwarp();

        float4 na[4];                                                  // count cells of sample step 0 of the current pass
#pragma unroll
        for (int i = 0; i < 4; i++) {                                  // (row 0 stands in for the missing rows of a partial pass)
            const int idx = r * 4 + i;
            na[i] = (o < S) ? ld_countsf(rows + (size_t)(idx < count ? idx : 0) * S + o) : zero4;
        }

        for (int base = 0; base < count; base += TG_PASS_SITES) {
            const bool more = base + TG_PASS_SITES < count;
            // next pass: its rows into L2 while this pass is contracted
            if (more) l2_prefetch_row(rows + (size_t)(base + TG_PASS_SITES) * S, (uint32_t)min(count - base - TG_PASS_SITES, TG_PASS_SITES) * row_bytes);
            const int iown = base + r * 4 + isl;
            const bool have_own = iown < count;
            const int pown = begin + (have_own ? iown : 0);
            const float nown = p.nsite[pown];
            const int vown = p.grp.order[pown];
            const int sown = p.grp.site_slot[vown];
            const float4 *row[4], *rown[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int idx = base + r * 4 + i, idn = idx + TG_PASS_SITES;
                row[i] = rows + (size_t)(idx < count ? idx : 0) * S + o;
                rown[i] = rows + (size_t)(idn < count ? idn : 0) * S + o;
            }
            uint32_t mask = 0;

            for (int gb = 0; gb < nGB; gb++) {
                const float4 *Wb = W + (size_t)o * STR + gb * NJ;
                float acc[NV];
#pragma unroll
                for (int i = 0; i < NV; i++) acc[i] = 0.f;
                if (gb > 0) {
#pragma unroll
                    for (int i = 0; i < 4; i++) na[i] = (o < S) ? ld_countsf(row[i]) : zero4;
                }
                float4 nb[4];
                for (int k = 0; k < nk; k += 2) {
                    const int s1 = o + 8 * (k + 1), s2 = o + 8 * (k + 2);
#pragma unroll
                    for (int i = 0; i < 4; i++) nb[i] = (s1 < S) ? ld_countsf(row[i] + 8 * (k + 1)) : zero4;
                    tg_step<NJ>(acc, na, Wb + (size_t)(8 * k) * STR);
#pragma unroll
                    for (int i = 0; i < 4; i++) na[i] = (s2 < S) ? ld_countsf(row[i] + 8 * (k + 2)) : zero4;
                    tg_step<NJ>(acc, nb, Wb + (size_t)(8 * k + 8) * STR);
                }
                if (gb == nGB - 1 && more) {
                    // first sample step of the next pass: in flight during the reduction below
#pragma unroll
                    for (int i = 0; i < 4; i++) na[i] = (o < S) ? ld_countsf(rown[i]) : zero4;
                }
                // combine the 8 sample groups: lane o ends with values [o*NV/8, (o+1)*NV/8) = pairs (isl, gl = (o&1)*PP + q)
                tg_reduce_level<NV>(acc, (o & 4) != 0, 4);
                tg_reduce_level<NV / 2>(reinterpret_cast<float(&)[NV / 2]>(acc), (o & 2) != 0, 2);
                tg_reduce_level<NV / 4>(reinterpret_cast<float(&)[NV / 4]>(acc), (o & 1) != 0, 1);
                const float bn = nown * bn_scale + 1e-6f;
                uint32_t ml = 0;
#pragma unroll
                for (int q = 0; q < PP; q++) {
                    const float d0 = acc[3 * q], d1 = acc[3 * q + 1], d2 = acc[3 * q + 2];
                    const float top = fmaxf(fmaxf(d0, d1), d2);
                    const bool fin = (fabsf(d0) + fabsf(d1) + fabsf(d2)) < 1.0e30f;
                    const bool stay = fin && (top * LN2 + bn < -TAU_GAP);
                    if (!stay) ml |= 1u << ((o & 1) * PP + q);
                }
                ml |= __shfl_xor_sync(DESMAN_FULL_MASK, ml, 1);
                mask |= ml << (gb * GB);
            }
            mask &= fullG;
            bool push = false;
            if (owner && have_own) {
                if (!fast_ok || sown != slot) mask = fullG;                           // orphan: its pattern is not this group's
                if (p.words) {
                    const uint32_t *w = p.words + (size_t)vown * G;
                    for (int g = 0; g < G; g++) if (w[g] == 0u) mask = fullG;         // u == 0 (c_sample_tau.c:174): reference-order path
                }
                push = mask != 0u;
                if (!push) n_decided += (unsigned int)G;
            }
            const unsigned int bal = __ballot_sync(DESMAN_FULL_MASK, push);
            if (bal) {
                int pos = 0;
                if (lane == 0) pos = atomicAdd(p.grp.gctl + GC_NWORK, __popc(bal));
                pos = __shfl_sync(DESMAN_FULL_MASK, pos, 0) + __popc(bal & ((1u << lane) - 1u));
                if (push) p.grp.work[pos] = make_uint2((unsigned int)vown, mask);
            }
        }
    }
    n_decided = (unsigned int)warp_sum_u64((unsigned long long)n_decided);
    if (lane == 0 && n_decided && p.tier_counts) atomicAdd(p.tier_counts, (unsigned long long)n_decided);
}

// =====================================================================================================================
// Tensor-core form of the screening pass (used when every count is < 2048, i.e. exact in TF32).
//
// The contraction D[col][site] = sum_{s,b} Wd[s][col][b] * n[site][s][b], col = g*3 + j, is a [3G x 4S] x [4S x 16 sites]
// product per pass: warp-level mma.sync m16n8k8 (TF32 operands, FP32 accumulators) with the TABLE as the A operand
// (16 table columns per M tile) and the COUNTS as the B operand (8 sites per N tile, two N tiles per pass).  The k index
// of one MMA is laid out as k = t -> (sample s0+t, base beta), k = t+4 -> (sample s0+t, base beta+1) (t = lane & 3, beta =
// 0 or 2), so that the B fragments of the two MMAs of a 4-sample group are the two halves of ONE 128-bit count cell of
// site g (g = lane >> 2) -- no register shuffling between the load and the MMA -- and the A fragments are 128-bit words
// of the table, which is stored in fragment order: Wm[s][beta/2][tile][g] = {Wd[s][16 tile+g][beta], Wd[s][16 tile+g+8][beta],
// Wd[s][16 tile+g][beta+1], Wd[s][16 tile+g+8][beta+1]}.  The FP32 table entry is split on the fly into a TF32 head (11
// significant bits) and a remainder; two MMAs per entry recover it to ~2^-20 relative.  Counts are integers < 2^11: exact.
// Error model (on top of the per-entry model of the FFMA form): representation of Wd by head + truncated remainder
// <= 2^-20 |Wd|; every MMA accumulation step is charged 2^-20 of the running magnitude (documented tensor-core behaviour is
// exact products, alignment and truncation to >= 24 bits: 8x margin); Sp/4 groups x 2 x 2 steps per output; |Wd| is bounded a
// priori from min(gamma)*min(eta): ~2e-3 log2 units per read.
// One CTA (4 warps) per work item, items dealt round robin (long items first); the warps share the item's table and take its
// 16-site passes round robin.  Count rows are pulled into L2 by TMA bulk prefetches (cp.async.bulk.prefetch.L2) one round
// of passes / one item ahead, so the 128-bit loads of a pass find them there.
#ifndef TGM_WARPS
#define TGM_WARPS 4
#endif

static inline int tgm_tiles(int G) { return (3 * G + 15) / 16; }
static inline size_t tgm_table_bytes(int S, int G)
{
    const size_t Sp = (size_t)((S + 15) & ~15);
    return Sp * (size_t)(16 * tgm_tiles(G) + 2) * sizeof(float4) + (size_t)(TGM_WARPS - 1) * tgm_tiles(G) * 8 * 32 * sizeof(float);   // + split-pass sums
}

__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void tf32_split(const float4 &w, uint32_t (&h)[4], uint32_t (&l)[4])
{
    h[0] = __float_as_uint(w.x) & 0xffffe000u; h[1] = __float_as_uint(w.y) & 0xffffe000u;
    h[2] = __float_as_uint(w.z) & 0xffffe000u; h[3] = __float_as_uint(w.w) & 0xffffe000u;
    l[0] = __float_as_uint(w.x - __uint_as_float(h[0])); l[1] = __float_as_uint(w.y - __uint_as_float(h[1]));
    l[2] = __float_as_uint(w.z - __uint_as_float(h[2])); l[3] = __float_as_uint(w.w - __uint_as_float(h[3]));
}

// one 4-sample group: cell0 / cell1 = this lane's count cells of the sites of N tile 0 / 1
template <int MT>
__device__ __forceinline__ void tgm_group(float (&ch)[MT][2][4], float (&cl)[MT][2][4], const float4 &c0, const float4 &c1,
                                          const float4 *__restrict__ Wq)
{
    const uint32_t x0 = __float_as_uint(c0.x), y0 = __float_as_uint(c0.y), z0 = __float_as_uint(c0.z), w0 = __float_as_uint(c0.w);
    const uint32_t x1 = __float_as_uint(c1.x), y1 = __float_as_uint(c1.y), z1 = __float_as_uint(c1.z), w1 = __float_as_uint(c1.w);
#pragma unroll
    for (int mt = 0; mt < MT; mt++) {
        uint32_t h[4], l[4];
        tf32_split(Wq[8 * mt], h, l);                      // bases 0,1
        mma_tf32(ch[mt][0], h, x0, y0); mma_tf32(ch[mt][1], h, x1, y1);
        mma_tf32(cl[mt][0], l, x0, y0); mma_tf32(cl[mt][1], l, x1, y1);
        tf32_split(Wq[8 * MT + 8 * mt], h, l);             // bases 2,3
        mma_tf32(ch[mt][0], h, z0, w0); mma_tf32(ch[mt][1], h, z1, w1);
        mma_tf32(cl[mt][0], l, z0, w0); mma_tf32(cl[mt][1], l, z1, w1);
    }
}

template <int MT>
__global__ void __launch_bounds__(TGM_WARPS * 32, 16 / TGM_WARPS) tau_group_mma_kernel(TauGroupParams p)
{
#if !PDL_EARLY
    pdl_enter();
#endif
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int *gctl = p.grp.gctl;
    const int S = p.S, G = p.G;
    const int Sp = (S + 15) & ~15, nq = Sp >> 2;      // 4-sample groups
    constexpr int STR = 16 * MT + 2;                  // float4 per sample: = 2 (mod 8) -> conflict-free fragment loads

    double *gT = reinterpret_cast<double *>(smem_raw);                    // [G][Sp]
    double *eta_s = gT + (size_t)G * Sp;                                  // [16]
    float4 *eta32 = reinterpret_cast<float4 *>(eta_s + 16);               // [4]
    float *gT32 = reinterpret_cast<float *>(eta32 + 4);                   // [G][Sp]
    float4 *W = reinterpret_cast<float4 *>(gT32 + (size_t)G * Sp);        // [Sp][STR] fragment order
    float *Wf = reinterpret_cast<float *>(W);
    float *red = Wf + (size_t)Sp * STR * 4;                               // [TGM_WARPS - 1][MT * 8][32] partial sums of split passes
    __shared__ unsigned int gmin_bits, emin_bits;

    const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
    if (tid == 0) { gmin_bits = 0x7f800000u; emin_bits = 0x7f800000u; }
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = tid; i < Sp * STR; i += blockDim.x) W[i] = zero4;        // padding rows / columns / samples stay zero
#if PDL_EARLY
    pdl_enter();                                                          // (shared memory only so far)
#endif
    KPROF_SCOPE(KP_TGM);
    if (!grp_active(gctl, 0)) return;
    __syncthreads();
    float gmin_l = __int_as_float(0x7f800000);
    for (int i = tid; i < G * Sp; i += blockDim.x) {
        const int g = i / Sp, s = i - g * Sp;
        const double x = (s < S) ? p.gamma[(size_t)s * G + g] : 0.0;
        gT[i] = x;
        gT32[i] = (float)x;
        if (s < S && x > 0.0) gmin_l = fminf(gmin_l, (float)x);     // masked strains (gamma == 0): q = P there
    }
    atomicMin(&gmin_bits, __float_as_uint(gmin_l));
    if (tid < 16) {
        eta_s[tid] = p.eta[tid];
        reinterpret_cast<float *>(eta32)[tid] = (float)p.eta[tid];
        atomicMin(&emin_bits, __float_as_uint(fmaxf((float)p.eta[tid], 0.f)));
    }
    __syncthreads();
    // unnormalised input (rows of gamma or eta summing to more than 1): no screening, every site goes to the per-site kernel
    __shared__ int unnorm;
    if (tid == 0) unnorm = 0;
    __syncthreads();
    for (int s = tid; s < S + 4; s += blockDim.x) {
        double t = 0.0;
        if (s < S) for (int g = 0; g < G; g++) t += gT[g * Sp + s];
        else for (int b = 0; b < 4; b++) t += eta_s[4 * (s - S) + b];
        if (!(t <= 1.0001)) unnorm = 1;
    }
    __syncthreads();
    const float qmin = 0.99f * __uint_as_float(gmin_bits) * __uint_as_float(emin_bits);
    const bool fast_ok = qmin >= TAU_QMIN && !unnorm;
    const float mq0 = fmaxf(1.0f, 1.0f - log2f(fmaxf(qmin, TAU_QMIN)));
    // per read, log2 units: entry model (relative parts, lg2.approx floors, lg2.approx and the lq - lP rounding per unit of
    // |lg2|, FP64 cancellation) + [TF32 split + Sp accumulation steps] * 2^-20 * max|Wd| of the item
    const float e_entry = TAU_C0 + (2.3841858e-7f + 5.9604645e-8f) * (2.0f * mq0) + TAU_CANCEL(G) / fmaxf(qmin, TAU_QMIN);
    const float e_mma = (float)(Sp + 8) * 9.5367432e-7f;   // + the FP32 adds that join the partial sums of a split pass
    const float LN2 = 0.69314718f;
    // q, P <= 1 (convex combinations of eta entries), so lg2 q, lg2 P <= 0 and |Wd| = |lg2 q - lg2 P| <= mq0
    const float bn_scale = (e_entry + e_mma * mq0) * LN2 * 1.0001f;

    const int g8 = lane >> 2, t4 = lane & 3;
    const uint32_t fullG = (G >= 32) ? 0xffffffffu : ((1u << G) - 1u);
    const uint32_t row_bytes = (uint32_t)S * 16u;
    const int nitems = gctl[GC_NITEMS];
    // build tasks: (sample, chunk of strains); H chunks so that the CTA's threads are all busy
    const int H = max(1, min(G, (int)blockDim.x / Sp)), GH = (G + H - 1) / H;
    unsigned int n_decided = 0;
    constexpr int ROUND = TGM_WARPS * TG_PASS_SITES;      // sites the CTA contracts per round of passes

    // the work items of this CTA: blockIdx.x first, then tickets from the sweep's cursor (items are numbered long first, so the
    // hand-out is longest-processing-time-first: the static deal left CTAs between 28 and 57 us busy, tools/kprof.py).  The
    // ticket of the next item is drawn while the table of this one is built; its record is fetched during the contraction.
    __shared__ int next_item_s;
    int4 nxt0 = make_int4(0, 0, 0, 0), nxt1 = nxt0;
    if ((int)blockIdx.x < nitems) {
        nxt0 = p.grp.items[2 * blockIdx.x]; nxt1 = p.grp.items[2 * blockIdx.x + 1];
        if (tid == 0) l2_prefetch_row(p.countsf + (size_t)nxt0.y * S, (uint32_t)min(nxt0.z, ROUND) * row_bytes);
    }

#ifdef KPROF
    if (tid == 0) krec_put(KP_TGM_PRO, (int)blockIdx.x, 0, nitems, gtimer(), 0);
#endif
#ifdef TGM_PROFILE
    long long t_build = 0, t_pass = 0, t_sync1 = 0, t_sync2 = 0, t_all0 = clock64(), n_it = 0, n_sites = 0;
#define TGM_T(x) const long long x = clock64()
#else
#define TGM_T(x)
#endif
    int it = blockIdx.x;
    while (it < nitems) {
        TGM_T(tp0);
        int ticket = 0;
        if (tid == 0) ticket = (int)gridDim.x + atomicAdd(p.grp.gctl + GC_CURSOR, 1);
        const int4 item = nxt0, item1 = nxt1;
        const int slot = item.x, begin = item.y, count = item.z;
        const uint64_t code = ((uint64_t)(unsigned int)item1.y << 32) | (unsigned int)item1.x;
        const float4 *rows = p.countsf + (size_t)begin * S;            // the item's count rows are contiguous
        int base = wib * TG_PASS_SITES;                                 // first pass of this warp

        // ---- the pattern's table: Wd[s][g*3+j][b] = lg2(base_sb + eta[a_j][b]*gamma[s][g]) - lg2(P_sb), a_j = (cur_g+1+j)&3
        for (int task = tid; task < Sp * H; task += blockDim.x) {
            const int s = task % Sp, h = task / Sp;
            const int glo = h * GH, ghi = min(G, glo + GH);
            if (s < S && fast_ok) {
                float *Ws = Wf + (size_t)s * STR * 4;
                double P0 = 0.0, P1 = 0.0, P2 = 0.0, P3 = 0.0;
                for (int hh = 0; hh < G; hh++) {
                    const double2 *e = reinterpret_cast<const double2 *>(eta_s + 4 * code_get(code, hh));
                    const double2 e01 = e[0], e23 = e[1];
                    const double gm = gT[hh * Sp + s];
                    P0 = fma(e01.x, gm, P0); P1 = fma(e01.y, gm, P1); P2 = fma(e23.x, gm, P2); P3 = fma(e23.y, gm, P3);
                }
                const float l0 = lg2_fast((float)P0), l1 = lg2_fast((float)P1), l2 = lg2_fast((float)P2), l3 = lg2_fast((float)P3);
                for (int g = glo; g < ghi; g++) {
                    const int cur = code_get(code, g);
                    const double2 *ecp = reinterpret_cast<const double2 *>(eta_s + 4 * cur);
                    const double2 ec01 = ecp[0], ec23 = ecp[1];
                    const double gg = gT[g * Sp + s];
                    const float gf = gT32[g * Sp + s];
                    const float q0 = fmaxf((float)fma(-ec01.x, gg, P0), 0.f), q1 = fmaxf((float)fma(-ec01.y, gg, P1), 0.f),
                                q2 = fmaxf((float)fma(-ec23.x, gg, P2), 0.f), q3 = fmaxf((float)fma(-ec23.y, gg, P3), 0.f);
#pragma unroll
                    for (int j = 0; j < 3; j++) {
                        const float4 ea = eta32[(cur + 1 + j) & 3];
                        const float wx = lg2_fast(fmaf(ea.x, gf, q0)) - l0, wy = lg2_fast(fmaf(ea.y, gf, q1)) - l1,
                                    wz = lg2_fast(fmaf(ea.z, gf, q2)) - l2, ww = lg2_fast(fmaf(ea.w, gf, q3)) - l3;
                        // fragment order: float4 index (beta/2)*8*MT + (c>>4)*8 + (c&7), component ((c>>3)&1) + 2*(b&1)
                        const int c = g * 3 + j;
                        float *dst = Ws + 4 * ((c >> 4) * 8 + (c & 7)) + ((c >> 3) & 1);
                        dst[0] = wx; dst[2] = wy;
                        dst[32 * MT] = wz; dst[32 * MT + 2] = ww;
                    }
                }
            }
        }
        // first round of the next item into L2 while this one is contracted
        if (tid == 0) {
            next_item_s = ticket;
            if (ticket < nitems) {
                const int4 r = p.grp.items[2 * ticket];
                l2_prefetch_row(p.countsf + (size_t)r.y * S, (uint32_t)min(r.z, ROUND) * row_bytes);
            }
        }
        TGM_T(tp1);
        __syncthreads();
        TGM_T(tp2);
        const int it_next = next_item_s;
        if (it_next < nitems) { nxt0 = p.grp.items[2 * it_next]; nxt1 = p.grp.items[2 * it_next + 1]; }

        // ---- contraction of one 16-site pass over the 4-sample groups [qlo, qhi): D (hi + lo parts) -> acc[mt][nt][4]
        auto contract = [&](int pbase, int qlo, int qhi, float (&acc)[MT][2][4]) {
            const float4 *row0 = rows + (size_t)(pbase + g8 < count ? pbase + g8 : 0) * S + t4,
                         *row1 = rows + (size_t)(pbase + g8 + 8 < count ? pbase + g8 + 8 : 0) * S + t4;
            const float4 *Wq = W + (size_t)t4 * STR + g8;
            float ch[MT][2][4], cl[MT][2][4];
#pragma unroll
            for (int mt = 0; mt < MT; mt++)
#pragma unroll
                for (int nt = 0; nt < 2; nt++)
#pragma unroll
                    for (int i = 0; i < 4; i++) { ch[mt][nt][i] = 0.f; cl[mt][nt][i] = 0.f; }
            // four groups at a time: 8 cells in flight per lane
            for (int q = qlo; q < qhi; q += 4) {
                float4 c0[4], c1[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const bool ok = 4 * (q + u) + t4 < S;
                    c0[u] = ok ? ld_countsf(row0 + 4 * (q + u)) : zero4;
                    c1[u] = ok ? ld_countsf(row1 + 4 * (q + u)) : zero4;
                }
#pragma unroll
                for (int u = 0; u < 4; u++) tgm_group<MT>(ch, cl, c0[u], c1[u], Wq + (size_t)(4 * (q + u)) * STR);
            }
#pragma unroll
            for (int mt = 0; mt < MT; mt++)
#pragma unroll
                for (int nt = 0; nt < 2; nt++)
#pragma unroll
                    for (int i = 0; i < 4; i++) acc[mt][nt][i] = ch[mt][nt][i] + cl[mt][nt][i];
        };
        // ---- decisions of one pass from its complete sums; undecided sites go to the work list
        // accumulator (mt, nt, i): table column 16 mt + 8 (i >> 1) + g8, site 8 nt + 2 t4 + (i & 1)
        // what the decisions of a pass need besides its sums, fetched BEFORE the contraction so that the (dependent) loads are
        // not waited for after it: the site this lane decides for (lanes g8 < 4 own row k = {2 t4, 2 t4 + 1, 8 + 2 t4, 9 + 2 t4}[g8]
        // of the pass), its current slot, and the read totals of the 4 sites whose sums this lane holds
        struct PassMeta { int vown, sown; float nk[4]; };
        auto fetch_meta = [&](int pbase) {
            PassMeta mt_;
            const int kown = ((g8 & 2) << 2) + 2 * t4 + (g8 & 1);
            mt_.vown = p.grp.order[begin + (pbase + kown < count ? pbase + kown : 0)];
            mt_.sown = p.grp.site_slot[mt_.vown];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int idx = pbase + ((k & 2) << 2) + 2 * t4 + (k & 1);
                mt_.nk[k] = p.nsite[begin + (idx < count ? idx : 0)];
            }
            return mt_;
        };
        auto decide = [&](int pbase, const float (&acc)[MT][2][4], const PassMeta &pm) {
            const int kown = ((g8 & 2) << 2) + 2 * t4 + (g8 & 1);
            const bool own = g8 < 4 && pbase + kown < count;
            const int vown = pm.vown, sown = pm.sown;
            const float (&nk)[4] = pm.nk;
            // a strain is decided "stay" iff each of its three candidates trails the current base by more than TAU_GAP nats after the bound
            uint32_t m[4] = {0u, 0u, 0u, 0u};
#pragma unroll
            for (int mt = 0; mt < MT; mt++)
#pragma unroll
                for (int hf = 0; hf < 2; hf++) {
                    const int c = 16 * mt + 8 * hf + g8;
                    if (c < 3 * G) {
                        const uint32_t bit = 1u << (c / 3);
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            const float d = acc[mt][k >> 1][2 * hf + (k & 1)];
                            if (!(d * LN2 + (nk[k] * bn_scale + 1e-6f) < -TAU_GAP)) m[k] |= bit;
                        }
                    }
                }
#pragma unroll
            for (int k = 0; k < 4; k++) {
                m[k] |= __shfl_xor_sync(DESMAN_FULL_MASK, m[k], 4);
                m[k] |= __shfl_xor_sync(DESMAN_FULL_MASK, m[k], 8);
                m[k] |= __shfl_xor_sync(DESMAN_FULL_MASK, m[k], 16);
            }
            uint32_t mask = ((g8 & 3) == 0 ? m[0] : (g8 & 3) == 1 ? m[1] : (g8 & 3) == 2 ? m[2] : m[3]) & fullG;
            bool push = false;
            if (own) {
                if (!fast_ok || sown != slot) mask = fullG;                           // orphan: its pattern is not this group's
                if (p.words) {
                    const uint32_t *w = p.words + (size_t)vown * G;
                    for (int g = 0; g < G; g++) if (w[g] == 0u) mask = fullG;         // u == 0 (c_sample_tau.c:174): reference-order path
                }
                push = mask != 0u;
                if (!push) n_decided += (unsigned int)G;
            }
            const unsigned int bal = __ballot_sync(DESMAN_FULL_MASK, push);
            if (bal) {
                int pos = 0;
                if (lane == 0) pos = atomicAdd(p.grp.gctl + GC_NWORK, __popc(bal));
                pos = __shfl_sync(DESMAN_FULL_MASK, pos, 0) + __popc(bal & ((1u << lane) - 1u));
                if (push) p.grp.work[pos] = make_uint2((unsigned int)vown, mask);
            }
        };

        const int np = (count + TG_PASS_SITES - 1) / TG_PASS_SITES;
        if (np > 2) {
            // long item: whole passes, round robin over the warps
            for (; base < count; base += ROUND) {
                if (base + ROUND < count) l2_prefetch_row(rows + (size_t)(base + ROUND) * S, (uint32_t)min(count - base - ROUND, TG_PASS_SITES) * row_bytes);
                float acc[MT][2][4];
                const PassMeta pm = fetch_meta(base);
                contract(base, 0, nq, acc);
                decide(base, acc, pm);
            }
        } else {
            // short item (one or two passes): the warps split the SAMPLES of a pass (4 or 2 warps per pass) and the partial
            // sums meet in shared memory, so that all four warps work and the item's latency is a fraction of a pass
            const int kw = (np == 1) ? TGM_WARPS : TGM_WARPS / 2;
            const int pass = wib / kw, part = wib % kw;
            const int qper = ((nq / 4 + kw - 1) / kw) * 4;
            const int qlo = min(nq, part * qper), qhi = min(nq, qlo + qper);
            float acc[MT][2][4];
            const PassMeta pm = fetch_meta(pass * TG_PASS_SITES);
            contract(pass * TG_PASS_SITES, qlo, qhi, acc);
            if (part > 0) {
#pragma unroll
                for (int mt = 0; mt < MT; mt++)
#pragma unroll
                    for (int nt = 0; nt < 2; nt++)
#pragma unroll
                        for (int i = 0; i < 4; i++) red[((wib - 1) * MT * 8 + (mt * 2 + nt) * 4 + i) * 32 + lane] = acc[mt][nt][i];
            }
            __syncthreads();
            if (part == 0) {
                for (int w2 = wib + 1; w2 < wib + kw; w2++)
#pragma unroll
                    for (int mt = 0; mt < MT; mt++)
#pragma unroll
                        for (int nt = 0; nt < 2; nt++)
#pragma unroll
                            for (int i = 0; i < 4; i++) acc[mt][nt][i] += red[((w2 - 1) * MT * 8 + (mt * 2 + nt) * 4 + i) * 32 + lane];
                decide(pass * TG_PASS_SITES, acc, pm);
            }
        }
        TGM_T(tp3);
        __syncthreads();
        it = it_next;
#ifdef TGM_PROFILE
        { const long long tp4 = clock64(); t_build += tp1 - tp0; t_sync1 += tp2 - tp1; t_pass += tp3 - tp2; t_sync2 += tp4 - tp3; n_it++; n_sites += count; }
#endif
    }
#ifdef TGM_PROFILE
    if (lane == 0 && (blockIdx.x % 97 == 0 || blockIdx.x == gridDim.x - 1) && p.tier_counts)
        printf("cta %4d warp %d items %lld sites %lld build %lld sync1 %lld pass %lld sync2 %lld total %lld\n", (int)blockIdx.x, wib, n_it, n_sites,
               t_build, t_sync1, t_pass, t_sync2, clock64() - t_all0);
#endif
    n_decided = (unsigned int)warp_sum_u64((unsigned long long)n_decided);
    if (lane == 0 && n_decided && p.tier_counts) atomicAdd(p.tier_counts, (unsigned long long)n_decided);
}

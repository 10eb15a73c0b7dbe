// desman_b200/csrc/maintain_kernel.cuh -- one cooperative launch per sweep that (a) clears the per-sweep accumulators and
// (b) only when asked to, rebuilds the persistent pattern table (mu_agg_kernel.cuh) and regroups the sites by pattern for the
// screening pass of the tau update (tau_group_kernel.cuh).  In steady state (b) is a handful of flag reads: the launch
// replaces three conditional no-op launches and three memsets of the first version.
//
// Requests: ctl[0] (table rebuild: host after a state upload, finalize_sweep when stale slots pile up) and gctl[GC_REGROUP]
// (finalize_sweep when orphans pile up).  Both are read by every block BEFORE the first grid barrier and cleared after it,
// so all blocks take the same path.
#pragma once
#include <cooperative_groups.h>
#include "common.cuh"
#include "mu_agg_kernel.cuh"
#include <cuda_fp16.h>
#include "tau_group_kernel.cuh"

namespace cg = cooperative_groups;

struct MaintParams {
    MuAggParams a;               // counts, tau, shape, table
    TauGroup grp;                // grp.gctl == nullptr: no grouping
    int *blk;                    // [4 * gridDim] block totals of the regroup scan
    unsigned long long *zero64;  // per-sweep accumulators cleared here: the statistics ...
    int nzero64;
    unsigned long long *red_i;   // ... and [2] fixed-point ll | nchange
    float4 *countsf;             // [V][S] FP32 copy of the count rows in group order (row pos <-> site grp.order[pos])
    float *nsite;                // [V] reads per row, rounded up
    // deferred MAP snapshot of the previous sweep (update(): tau does not change between its finalize and this launch):
    // if (*star_flag) tau_star <- tau, in place of a separate copy_tau_if launch per sweep; star_flag == nullptr: none
    const int *star_flag;
    uint8_t *tau_star;
    int force_worth;             // 1: the caller insists on the grouping (option tau_group = 1): GC_WORTH is set whatever the groups look like
    int item_sites;              // sites per work item (multiple of 8): TG_ITEM_SITES, or TC_ROWS for the tensor-memory pass
    // count image of the tensor-memory screening pass (tau_group_tc_kernel.cuh), or img == nullptr: fp16x4 cells in the
    // K-major no-swizzle UMMA order [K block][row group][KC chunks][8 rows][16 bytes]; every item starts at a multiple of 8 rows
    unsigned char *img;
    int *img_site;               // [img_cap_rows] site of an image row (~site once the site has left its group)
    float *img_nsite;            // [img_cap_rows]
    int *site_row;               // [V] image row of a site, -1: none (singles)
    int *slot_img;               // [cap_slots] first image row of a slot's items
    long long img_cap_rows;
    int SK, nkb;                 // samples per K block, K blocks
};

#define MAINT_THREADS 256

// exclusive scan of four ints over the block (256 threads); returns the block totals in tot
__device__ __forceinline__ int4 block_excl_scan4(int4 v, int4 &tot, int (*sh)[4])
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int4 inc = v;
#pragma unroll
    for (int m = 1; m < 32; m <<= 1) {
        const int x = __shfl_up_sync(DESMAN_FULL_MASK, inc.x, m), y = __shfl_up_sync(DESMAN_FULL_MASK, inc.y, m),
                  z = __shfl_up_sync(DESMAN_FULL_MASK, inc.z, m), u = __shfl_up_sync(DESMAN_FULL_MASK, inc.w, m);
        if (lane >= m) { inc.x += x; inc.y += y; inc.z += z; inc.w += u; }
    }
    __syncthreads();
    if (lane == 31) { sh[w][0] = inc.x; sh[w][1] = inc.y; sh[w][2] = inc.z; sh[w][3] = inc.w; }
    __syncthreads();
    int4 off = make_int4(0, 0, 0, 0);
    tot = make_int4(0, 0, 0, 0);
    for (int i = 0; i < MAINT_THREADS / 32; i++) {
        if (i < w) { off.x += sh[i][0]; off.y += sh[i][1]; off.z += sh[i][2]; off.w += sh[i][3]; }
        tot.x += sh[i][0]; tot.y += sh[i][1]; tot.z += sh[i][2]; tot.w += sh[i][3];
    }
    return make_int4(off.x + inc.x - v.x, off.y + inc.y - v.y, off.z + inc.z - v.z, off.w + inc.w - v.w);
}

__global__ void __launch_bounds__(MAINT_THREADS) table_maintain_kernel(MaintParams p)
{
    pdl_enter();
    KPROF_SCOPE(KP_MAINT);
    cg::grid_group grid = cg::this_grid();
    __shared__ int sh[MAINT_THREADS / 32][4];
    const AggTable &t = p.a.t;
    int *gctl = p.grp.gctl;
    const size_t gtid = blockIdx.x * (size_t)blockDim.x + threadIdx.x, gsz = (size_t)gridDim.x * blockDim.x;

    const bool do_rebuild = t.ctl[0] != 0;
    const bool do_regroup = gctl && gctl[GC_CALM] && (do_rebuild || gctl[GC_REGROUP] || !gctl[GC_HAVE]);
    for (size_t i = gtid; i < (size_t)p.nzero64; i += gsz) p.zero64[i] = 0ull;
    if (gtid < 4) p.red_i[gtid] = 0ull;
    if (gtid >= 4 && gtid < AGG_CTL_WORDS) t.ctl[gtid] = 0;          // work cursors of mu_binomial_kernel
    if (gtid == 0 && gctl) { gctl[GC_NWORK] = 0; gctl[GC_CURSOR] = 0; }
    if (p.star_flag && *p.star_flag) {
        const size_t n = (size_t)p.a.V * p.a.G, n16 = n / 16;
        const uint4 *src = reinterpret_cast<const uint4 *>(p.a.tau);
        uint4 *dst = reinterpret_cast<uint4 *>(p.tau_star);
        for (size_t i = gtid; i < n16; i += gsz) dst[i] = src[i];
        for (size_t i = n16 * 16 + gtid; i < n; i += gsz) p.tau_star[i] = p.a.tau[i];
    }
    if (!do_rebuild && !do_regroup) return;

    const int V = p.a.V, S = p.a.S, G = p.a.G;
    const int lane = threadIdx.x & 31;
    const int gw = (int)(gtid >> 5), nw = (int)(gsz >> 5);
    if (do_rebuild) {
        // ---- free every key, zero the used part of N
        unsigned int used = *t.nslots;
        if (used > t.cap_slots) used = t.cap_slots;
        for (size_t i = gtid; i <= t.hmask; i += gsz) { t.keys[i] = MUB_EMPTY; t.ids[i] = -1; }
        const size_t n = (size_t)used * S * 4;
        for (size_t i = gtid; i < n; i += gsz) t.N[i] = 0ull;
        grid.sync();
        if (gtid == 0) { *t.nslots = 0u; t.ctl[3] = 0; t.ctl[0] = 0; if (gctl) gctl[GC_HAVE] = 0; }
        grid.sync();
        // ---- aggregation pass: one warp per site
        for (int v = gw; v < V; v += nw) {
            const unsigned long long code = load_tau_code(p.a.tau + (size_t)v * G, G, lane);
            int id = 0;
            if (lane == 0) id = agg_slot(t, code, true);
            id = __shfl_sync(DESMAN_FULL_MASK, id, 0);
            if (lane == 0 && p.grp.site_slot) p.grp.site_slot[v] = id;
            unsigned long long *dst = t.N + (size_t)id * S * 4;
            const int4 *src = p.a.counts + (size_t)v * S;
            for (int s = lane; s < S; s += 32) {
                const int4 n4 = ld_counts(src + s);
                if (n4.x) atomicAdd(dst + s * 4 + 0, (unsigned long long)n4.x);
                if (n4.y) atomicAdd(dst + s * 4 + 1, (unsigned long long)n4.y);
                if (n4.z) atomicAdd(dst + s * 4 + 2, (unsigned long long)n4.z);
                if (n4.w) atomicAdd(dst + s * 4 + 3, (unsigned long long)n4.w);
            }
        }
    }
    if (!do_regroup) return;
    grid.sync();

    // ---- regroup: counting sort of the sites by slot; slots with one site go to the singles list
    unsigned int nslu = *t.nslots;
    if (nslu > t.cap_slots) nslu = t.cap_slots;
    const int nsl = (int)nslu;
    for (size_t i = gtid; i < (size_t)nsl; i += gsz) p.grp.slot_cnt[i] = 0;
    if (gtid == 0) { gctl[GC_ORPHANS] = 0; gctl[GC_REGROUP] = 0; }
    grid.sync();
    for (size_t v = gtid; v < (size_t)V; v += gsz) atomicAdd(p.grp.slot_cnt + p.grp.site_slot[v], 1);
    grid.sync();
    // scan, step 1: every block scans its segment of the slots: sites of multi-site patterns | full items | partial items |
    // singles.  Full items (TG_ITEM_SITES sites) are numbered before all partial ones, so the dynamic schedule of the
    // screening pass hands out the long items first and its tail is made of short ones.
    const int ITEM = p.item_sites;
    const int seg = (nsl + (int)gridDim.x - 1) / (int)gridDim.x;
    const int lo = (int)blockIdx.x * seg, hi = min(nsl, lo + seg);
    {
        int4 carry = make_int4(0, 0, 0, 0);
        for (int base = lo; base < hi; base += MAINT_THREADS) {
            const int sl = base + (int)threadIdx.x;
            const int c = (sl < hi) ? p.grp.slot_cnt[sl] : 0;
            const int4 val = make_int4(c >= 2 ? c : 0, c >= 2 ? c / ITEM : 0, (c >= 2 && c % ITEM) ? 1 : 0, c == 1 ? 1 : 0);
            int4 tot;
            const int4 ex = block_excl_scan4(val, tot, sh);
            if (sl < hi) {
                p.grp.slot_start[sl] = (c == 1) ? carry.w + ex.w : carry.x + ex.x;
                p.grp.slot_item[sl] = carry.y + ex.y;           // first full item (local numbering)
                p.grp.slot_fill[sl] = carry.z + ex.z;           // partial item (local numbering); reset to 0 in step 3
            }
            carry.x += tot.x; carry.y += tot.y; carry.z += tot.z; carry.w += tot.w;
        }
        if (threadIdx.x == 0) {
            p.blk[4 * blockIdx.x] = carry.x; p.blk[4 * blockIdx.x + 1] = carry.y; p.blk[4 * blockIdx.x + 2] = carry.z;
            p.blk[4 * blockIdx.x + 3] = carry.w;
        }
        if (p.img) {   // image rows: every multi-site slot padded to a multiple of 8 rows (items are multiples of 8 but the last)
            int carry8 = 0;
            for (int base = lo; base < hi; base += MAINT_THREADS) {
                const int sl = base + (int)threadIdx.x;
                const int c = (sl < hi) ? p.grp.slot_cnt[sl] : 0;
                int4 tot;
                const int4 ex = block_excl_scan4(make_int4(c >= 2 ? (c + 7) & ~7 : 0, 0, 0, 0), tot, sh);
                if (sl < hi) p.slot_img[sl] = carry8 + ex.x;
                carry8 += tot.x;
            }
            if (threadIdx.x == 0) p.blk[4 * gridDim.x + blockIdx.x] = carry8;
        }
    }
    grid.sync();
    // step 2: offsets of the blocks (a few hundred values: every thread of a block needs only its own block's offset)
    int4 off, all;
    {
        int4 part = make_int4(0, 0, 0, 0), mine = make_int4(0, 0, 0, 0);
        for (int b = (int)threadIdx.x; b < (int)gridDim.x; b += MAINT_THREADS) {
            const int4 x = make_int4(p.blk[4 * b], p.blk[4 * b + 1], p.blk[4 * b + 2], p.blk[4 * b + 3]);
            if (b < (int)blockIdx.x) { part.x += x.x; part.y += x.y; part.z += x.z; part.w += x.w; }
            mine.x += x.x; mine.y += x.y; mine.z += x.z; mine.w += x.w;
        }
        block_excl_scan4(part, off, sh);
        __syncthreads();
        block_excl_scan4(mine, all, sh);
        if (gtid == 0) {
            gctl[GC_NITEMS] = all.y + all.z; gctl[GC_NSINGLES] = all.w; gctl[GC_HAVE] = 1;
            // a table per work item costs about what three sites cost the per-site kernel: the grouping pays when the items
            // hold at least that many sites on average and a fair share of all sites is grouped at all
            const long long grouped = (long long)V - all.w, items = (long long)all.y + all.z;
            gctl[GC_WORTH] = (p.force_worth || (grouped >= 3 * items && grouped * 8 >= (long long)V)) ? 1 : 0;
        }
        __syncthreads();
    }
    int off8 = 0;
    bool img_ok = false;
    if (p.img) {
        int4 part = make_int4(0, 0, 0, 0), o8;
        for (int b = (int)threadIdx.x; b < (int)gridDim.x; b += MAINT_THREADS) {
            const int x = p.blk[4 * gridDim.x + b];
            if (b < (int)blockIdx.x) part.x += x;
            part.y += x;
        }
        block_excl_scan4(part, o8, sh);
        __syncthreads();
        off8 = o8.x;
        img_ok = (long long)o8.y <= p.img_cap_rows;
        if (gtid == 0) { gctl[GC_IMG_OK] = img_ok ? 1 : 0; gctl[GC_IMG_ROWS] = o8.y; }
    }
    // step 3: global positions and the items of this block's slots
    for (int sl = lo + (int)threadIdx.x; sl < hi; sl += MAINT_THREADS) {
        const int c = p.grp.slot_cnt[sl];
        if (c == 1) p.grp.slot_start[sl] += off.w;
        else if (c >= 2) {
            const int st = p.grp.slot_start[sl] + off.x, it0 = p.grp.slot_item[sl] + off.y, itp = all.y + off.z + p.grp.slot_fill[sl];
            p.grp.slot_start[sl] = st;
            const int nfull = c / ITEM;
            const unsigned long long code = t.slot_code[sl];
            int img0 = 0;                                           // first image row of the slot (a multiple of 8)
            if (p.img) { img0 = p.slot_img[sl] + off8; p.slot_img[sl] = img0; }
            int4 rec1 = make_int4((int)(unsigned int)code, (int)(unsigned int)(code >> 32), 0, 0);
            for (int j = 0; j < nfull; j++) {
                p.grp.items[2 * (it0 + j)] = make_int4(sl, st + j * ITEM, ITEM, 0);
                rec1.z = img0 + j * ITEM;
                p.grp.items[2 * (it0 + j) + 1] = rec1;
            }
            if (c % ITEM) {
                p.grp.items[2 * itp] = make_int4(sl, st + nfull * ITEM, c % ITEM, 0);
                rec1.z = img0 + nfull * ITEM;
                p.grp.items[2 * itp + 1] = rec1;
            }
        }
        p.grp.slot_fill[sl] = 0;
    }
    grid.sync();
    for (size_t v = gtid; v < (size_t)V; v += gsz) {
        const int sl = p.grp.site_slot[v];
        if (p.grp.slot_cnt[sl] == 1) p.grp.singles[p.grp.slot_start[sl]] = (int)v;
        else p.grp.order[p.grp.slot_start[sl] + atomicAdd(p.grp.slot_fill + sl, 1)] = (int)v;
    }
    grid.sync();
    // the count rows of the grouped sites, converted to FP32 (exact: counts <= 2^24) and laid out in group order, so that the
    // sites of a work item are one contiguous block of rows: the screening pass streams them without an index hop
    const int nrows = V - gctl[GC_NSINGLES];
    for (int pos = gw; pos < nrows; pos += nw) {
        const int v = p.grp.order[pos];
        long long tot = 0;
        for (int s2 = lane; s2 < S; s2 += 32) {
            const int4 n = ld_counts(p.a.counts + (size_t)v * S + s2);
            p.countsf[(size_t)pos * S + s2] = make_float4((float)n.x, (float)n.y, (float)n.z, (float)n.w);
            tot += (long long)n.x + n.y + n.z + n.w;
        }
        tot = (long long)warp_sum_u64((unsigned long long)tot);
        if (lane == 0) p.nsite[pos] = __ll2float_ru(tot);
    }
    if (!p.img) return;
    // the same rows as fp16x4 cells (exact: the host enables this pass only when every count < 2048) in the operand order of
    // the tensor-memory screening pass; samples beyond S inside the last K block are written as zeros
    for (size_t v = gtid; v < (size_t)V; v += gsz) p.site_row[v] = -1;
    if (!img_ok) return;
    grid.sync();
    const int KC = p.SK / 2, Spad = p.SK * p.nkb;
    const size_t kb_stride = (size_t)(p.img_cap_rows / 8) * KC * 128;
    for (int pos = gw; pos < nrows; pos += nw) {
        const int v = p.grp.order[pos];
        const int sl = p.grp.site_slot[v];
        const int row = p.slot_img[sl] + (pos - p.grp.slot_start[sl]);
        long long tot = 0;
        for (int s2 = lane; s2 < Spad; s2 += 32) {
            int4 n = make_int4(0, 0, 0, 0);
            if (s2 < S) n = ld_counts(p.a.counts + (size_t)v * S + s2);
            tot += (long long)n.x + n.y + n.z + n.w;
            const int kb = s2 / p.SK, sl2 = s2 - kb * p.SK;
            uint2 cell;
            cell.x = (uint32_t)__half_as_ushort(__int2half_rn(n.x)) | ((uint32_t)__half_as_ushort(__int2half_rn(n.y)) << 16);
            cell.y = (uint32_t)__half_as_ushort(__int2half_rn(n.z)) | ((uint32_t)__half_as_ushort(__int2half_rn(n.w)) << 16);
            unsigned char *dst = p.img + (size_t)kb * kb_stride + ((size_t)(row >> 3) * KC + (size_t)(sl2 >> 1)) * 128 +
                                 (size_t)(row & 7) * 16 + (size_t)(sl2 & 1) * 8;
            *reinterpret_cast<uint2 *>(dst) = cell;
        }
        tot = (long long)warp_sum_u64((unsigned long long)tot);
        if (lane == 0) { p.img_site[row] = v; p.img_nsite[row] = __ll2float_ru(tot); p.site_row[v] = row; }
    }
}

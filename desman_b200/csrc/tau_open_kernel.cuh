// desman_b200/csrc/tau_open_kernel.cuh -- K1o: the listed sites of the tau update (c_sample_tau.c:130-188), one CTA per site,
// the open steps of a site evaluated speculatively in parallel.
//
// What the screening pass leaves is a few hundred sites with 1-8 undecided steps each (and the orphans / single-site
// patterns with all G).  A step is a ~2 us dependent chain and tau_sample_kernel walks them one after the other on one warp,
// so its duration was that of the site with the most open steps (20 us for 640 sites at C3).  A flip is rare (a handful per
// sweep), and as long as no earlier strain of the site flips, the step of strain g sees exactly the pattern the site came in
// with.  So: the warps of the CTA stage the site once (counts row, mixture P in FP64, K = sum n lg2 P), each warp evaluates one
// open strain against the ORIGINAL pattern (same tiers, same arithmetic, same uniforms as tau_sample_kernel), and the
// decisions are valid up to and including the first strain that flips; only then the flip is applied and every later strain
// goes through another such round against the new pattern.  Draw for draw the results are those of the sequential walk.
#pragma once
#include "tau_kernel.cuh"

#define TAUO_WARPS 4

__global__ void __launch_bounds__(TAUO_WARPS * 32, 6) tau_open_kernel(TauParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int S = p.S, G = p.G;
    const int Sp = (S + 31) & ~31, nch = Sp >> 5;
    double *gT = reinterpret_cast<double *>(smem_raw);           // [G][Sp]
    double *eta_s = gT + (size_t)G * Sp;                         // [16]
    double2 *Pw = reinterpret_cast<double2 *>(eta_s + 16);       // [Sp][2] mixture probabilities of the site
    float *gT32 = reinterpret_cast<float *>(Pw + (size_t)Sp * 2);// [G][Sp]
    float4 *eta32 = reinterpret_cast<float4 *>(gT32 + (size_t)G * Sp);   // [4]
    float *Kw = reinterpret_cast<float *>(eta32 + 4);            // [Sp] sum_b n_b lg2 P_b
    float *mlPs = Kw + Sp;                                       // [Sp] max_b |lg2 P_b|
    int4 *tile = reinterpret_cast<int4 *>(mlPs + Sp);            // [Sp] counts of the site
    uint32_t *ww = reinterpret_cast<uint32_t *>(tile + Sp);      // [32] uniform words
    int *t_s = reinterpret_cast<int *>(ww + 32);                 // [32] speculative decisions, [32] their tiers
    __shared__ unsigned int gmin_bits, emin_bits;
    __shared__ int first_flip;

    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    if (threadIdx.x == 0) { gmin_bits = 0x7f800000u; emin_bits = 0x7f800000u; }
#if !PDL_EARLY
    pdl_enter();
#endif
    __syncthreads();
    // (gamma and eta were written at least two grids ago: the screening pass sits in between)
    float gmin_l = __int_as_float(0x7f800000);
    for (int i = threadIdx.x; i < G * Sp; i += blockDim.x) {
        const int g = i / Sp, s = i - g * Sp;
        const double x = (s < S) ? p.gamma[(size_t)s * G + g] : 0.0;
        gT[i] = x;
        gT32[i] = (float)x;
        if (s < S && x > 0.0) gmin_l = fminf(gmin_l, (float)x);
    }
    atomicMin(&gmin_bits, __float_as_uint(gmin_l));
    if (threadIdx.x < 16) {
        eta_s[threadIdx.x] = p.eta[threadIdx.x];
        reinterpret_cast<float *>(eta32)[threadIdx.x] = (float)p.eta[threadIdx.x];
        atomicMin(&emin_bits, __float_as_uint(fmaxf((float)p.eta[threadIdx.x], 0.f)));
    }
    __syncthreads();
    const float qmin = 0.99f * __uint_as_float(gmin_bits) * __uint_as_float(emin_bits);
    const bool fast_ok = !p.exact_only && qmin >= TAU_QMIN;
    const float mq0 = fmaxf(1.0f, 1.0f - log2f(fmaxf(qmin, TAU_QMIN)));
    const uint32_t k0 = (uint32_t)p.seed, k1 = (uint32_t)(p.seed >> 32);
    const float c1 = TAU_C1(nch), ccan = TAU_CANCEL(G);
    const float cancel0 = ccan / fmaxf(qmin, TAU_QMIN);
    const uint32_t fullG = (G >= 32) ? 0xffffffffu : ((1u << G) - 1u);
#if PDL_EARLY
    if (blockIdx.x == gridDim.x - 1) {
        // One block (the last: it has a site only when the list is longer than the grid) walks the FP64 recompute of tier 3
        // once on placeholder counts before its dependency wait.  The step is taken by one site in a few sweeps, always on
        // instructions that nobody has touched since the L2 was last flushed: that site -- and with it the launch, and under
        // sharding every rank -- waited ~8 us for code from HBM.  Here the lines are fetched under the tail of the screening pass.
        for (int s = threadIdx.x; s < Sp; s += blockDim.x) tile[s] = make_int4(1, 1, 1, 1);
        __syncthreads();
        const double La = tau_exact_logp_cand(tile, gT, eta_s, 0ull, 0, S, Sp, G, lane, wib & 3);
        double Lw[4] = {La, La - 1.0, La - 2.0, La - 3.0};
        if (tau_exact_pick(Lw, 0.3) == 77) gmin_bits = 0u;                 // (never true: keeps the calls)
        __syncthreads();
    }
    pdl_enter();
#endif
    KPROF_SCOPE(KP_TAUO);
    if (!grp_active(p.gctl, p.need_img)) return;   // no valid list: tau_sample_kernel walks every site
    const int nwork = p.gctl[GC_NWORK], nsite = nwork + p.gctl[GC_NSINGLES];
    unsigned int flips = 0, n1 = 0, n2 = 0, n3 = 0;
#ifdef KPROF
    const unsigned long long kp_t0 = gtimer();
    unsigned long long kp_rounds = 0, kp_stage = 0, kp_steps = 0, kp_q[3] = {0, 0, 0};
    int kp_sites = 0;
#endif

    // one (v,g) step against pattern `code`: the tiers of tau_sample_kernel; *tier = 1, 2 or 3
    auto step = [&](uint64_t code, int g, bool open_by_screening, float nlane, float mlP, int *tier) -> int {
        const int cur = code_get(code, g);
        const uint32_t w = ww[g];
        const double u = p.words ? (double)w / 4294967296.0 : ((double)w + 0.5) / 4294967296.0;
        int t = -1;
        const bool usable = fast_ok && (w != 0u || !p.words);
        if (usable && open_by_screening) {
            // a step the screening pass left open: its gap test already failed on the same kind of sums: straight to the brackets
            t = tau_bracket_decide(tile, Pw, Kw, gT + g * Sp, gT32 + g * Sp, eta_s + 4 * cur, eta32, cur, nch, lane, nlane, mlP, c1, ccan, u);
            *tier = 2;
        } else if (usable) {
            const double *eta_cur = eta_s + 4 * cur;
            float E0, E1, E2, KK, mq;
            tau_fp32_terms<false>(tile, Pw, Kw, gT + g * Sp, gT32 + g * Sp, eta_cur, eta32, cur, nch, lane, E0, E1, E2, KK, mq);
            const float LN2 = 0.69314718f;
            float D0 = E0 - KK, D1 = E1 - KK, D2 = E2 - KK;
            float eb = nlane * (TAU_C0 + c1 * (mq0 + mlP) + cancel0);
            warp_sum4(D0, D1, D2, eb, lane);
            const float Bn = eb * LN2 * 1.0001f + 1e-6f;
            const float x0 = D0 * LN2, x1 = D1 * LN2, x2 = D2 * LN2;
            const float top = fmaxf(fmaxf(x0, x1), x2);
            const int jm = (x0 == top) ? 0 : (x1 == top) ? 1 : 2;
            const bool finite = (fabsf(x0) + fabsf(x1) + fabsf(x2) + Bn) < 1.0e30f;
            *tier = 1;
            if (finite) {
                if (top + Bn < -TAU_GAP) t = cur;
                else {
                    const float rest = fmaxf(fmaxf(jm == 0 ? 0.f : x0, jm == 1 ? 0.f : x1), fmaxf(jm == 2 ? 0.f : x2, 0.f));
                    if ((top - Bn) - (rest + Bn) > TAU_GAP) t = (cur + 1 + jm) & 3;
                }
            }
            if (t < 0 && finite) {
                t = tau_bracket_decide(tile, Pw, Kw, gT + g * Sp, gT32 + g * Sp, eta_cur, eta32, cur, nch, lane, nlane, mlP, c1, ccan, u);
                *tier = 2;
            }
        }
        if (t < 0) *tier = 3;       // FP64 reference-order recompute: done by the four warps together (exact3 below)
        return t;
    };
    // tier 3 of one step, the four candidates on the four warps (each sum bit-identical to tau_exact_logp's); every thread
    // returns the draw
    __shared__ double L3[4];
    auto exact3 = [&](uint64_t code, int g) -> int {
        for (int a = wib; a < 4; a += TAUO_WARPS) {
            const double La = tau_exact_logp_cand(tile, gT, eta_s, code, g, S, Sp, G, lane, a);
            if (lane == 0) L3[a] = La;
        }
        __syncthreads();
        const uint32_t w = ww[g];
        const double u = p.words ? (double)w / 4294967296.0 : ((double)w + 0.5) / 4294967296.0;
        double L[4] = {L3[0], L3[1], L3[2], L3[3]};
        const int t = tau_exact_pick(L, u);
        __syncthreads();
        return t;
    };
    // this lane's reads and max |lg2 P| over its samples (what the error bounds of a step are charged on)
    auto lane_bounds = [&](float &nlane, float &mlP) {
        nlane = 0.0f; mlP = 1.0f;
        for (int c = 0; c < nch; c++) {
            const int s = c * 32 + lane;
            const int4 n = tile[s];
            nlane += (float)(n.x + n.y + n.z + n.w);
            mlP = fmaxf(mlP, mlPs[s]);
        }
    };

    for (int i = blockIdx.x; i < nsite; i += gridDim.x) {
        int v;
        uint32_t todo = fullG;
        if (i < nwork) { const uint2 e = p.work[i]; v = (int)e.x; todo = e.y & fullG; }
        else v = p.singles[i - nwork];
        // a full mask is what orphans of their group, single-site patterns (and sites with a zero MT word) get without any test:
        // their steps are ordinary ones, mostly settled by the cheap gap test
        const bool screened = (i < nwork) && todo != fullG;
#ifdef KPROF
        const unsigned long long kp_a = gtimer();
        kp_sites++;
#endif
        const uint64_t code_in = load_tau_code(p.tau + (size_t)v * G, G, lane);
        // ---- stage the site: warp c takes the 32-sample chunk c; the last warp draws the G uniform words
        for (int c = wib; c < nch; c += TAUO_WARPS) {
            const int s = c * 32 + lane;
            int4 n = make_int4(0, 0, 0, 0);
            if (s < S) n = ld_counts(p.counts + (size_t)v * S + s);
            tile[s] = n;
            double b0 = 0.0, b1 = 0.0, b2 = 0.0, b3 = 0.0;
            for (int h = 0; h < G; h++) {
                const double2 *e = reinterpret_cast<const double2 *>(eta_s + 4 * code_get(code_in, h));
                const double2 e01 = e[0], e23 = e[1];
                const double gm = gT[h * Sp + s];
                b0 = fma(e01.x, gm, b0); b1 = fma(e01.y, gm, b1);
                b2 = fma(e23.x, gm, b2); b3 = fma(e23.y, gm, b3);
            }
            if (s >= S) { b0 = 1.0; b1 = 1.0; b2 = 1.0; b3 = 1.0; }   // padding lanes: finite logs, zero counts
            Pw[s * 2] = make_double2(b0, b1); Pw[s * 2 + 1] = make_double2(b2, b3);
            const float l0 = lg2_fast((float)b0), l1 = lg2_fast((float)b1), l2 = lg2_fast((float)b2), l3 = lg2_fast((float)b3);
            Kw[s] = fmaf((float)n.x, l0, fmaf((float)n.y, l1, fmaf((float)n.z, l2, (float)n.w * l3)));
            mlPs[s] = fmaxf(fmaxf(fabsf(l0), fabsf(l1)), fmaxf(fabsf(l2), fabsf(l3)));
        }
        if (wib == TAUO_WARPS - 1) {
            uint32_t w = 0;
            if (lane < G) {
                if (p.words) w = p.words[(size_t)v * G + lane];
                else w = philox4x32_10((uint32_t)(p.v0 + v), (uint32_t)lane, p.sweep, (uint32_t)STAGE_TAU << 28, k0, k1).x;
            }
            ww[lane] = w;
            t_s[lane] = -1; t_s[32 + lane] = 0;
        }
        if (threadIdx.x == 0) first_flip = G;
        __syncthreads();
#ifdef KPROF
        const unsigned long long kp_b = gtimer();
        kp_stage += kp_b - kp_a;
#endif
        // ---- rounds: the pending strains, one per warp, against the current pattern; valid up to the first flip, which is then
        // applied (all warps, a chunk each) and makes every later strain pending again
        uint64_t code = code_in;
        uint32_t pending = todo;
        bool from_screening = screened;
        bool first_round = true;
        while (true) {
#ifdef KPROF
            kp_rounds++;
#endif
            float nlane, mlP;
#ifdef KPROF
            const unsigned long long kq0 = gtimer();
#endif
            lane_bounds(nlane, mlP);
#ifdef KPROF
            const unsigned long long kq1 = gtimer();
#endif
            // passes of TAUO_WARPS strains in ascending order; a flip found in one pass makes every later pass void (its strains
            // come after the flip and go through the next round anyway), so the passes stop there
            const int npend = __popc(pending);
            uint32_t rest = pending;
#ifdef KPROF
            unsigned long long kq2 = kq1;
#endif
            for (int base = 0; base < npend; base += TAUO_WARPS) {
                int g = -1;
                for (int k = 0; k < TAUO_WARPS && rest; k++) {             // the k-th lowest pending strain of this pass
                    const int gl = __ffs(rest) - 1;
                    rest &= rest - 1u;
                    if (k == wib) g = gl;
                }
                if (g >= 0) {
                    int tier = 0;
                    // (after a flip, the steps the screening pass had left open skip the gap test as well: it failed on a margin of
                    // tens of nats under the old pattern, and the brackets decide either way)
                    const int t = step(code, g, from_screening || (screened && ((todo >> g) & 1u)), nlane, mlP, &tier);
                    if (lane == 0) {
                        t_s[g] = t; t_s[32 + g] = tier;
                        if (t >= 0 && t != code_get(code, g)) atomicMin(&first_flip, g);
                    }
                }
#ifdef KPROF
                kq2 = gtimer();
#endif
                __syncthreads();
                if (first_flip < G) break;                                  // (uniform: read after the barrier)
            }
#ifdef KPROF
            const unsigned long long kq3 = gtimer();
            if (threadIdx.x == 0) { kp_q[0] += kq1 - kq0; kp_q[1] += kq2 - kq1; kp_q[2] += kq3 - kq2; }
#endif
            // steps left to the FP64 recompute (t < 0), in strain order; those after a flip need not be looked at
            for (int g = 0; g < G && g < first_flip; g++) {
                if (!((pending >> g) & 1u) || t_s[g] >= 0) continue;
                const int t = exact3(code, g);                               // (uniform: every thread takes this path together)
                if (threadIdx.x == 0) {
                    t_s[g] = t;
                    if (t != code_get(code, g) && g < first_flip) first_flip = g;
                }
                __syncthreads();
            }
            const int gf = first_flip;                                       // G: no pending strain flipped
            if (wib == 0) {
                // account for the decisions that stand: pending strains up to the flip; in the first round also the strains the
                // screening pass had decided (not on the mask) below the flip
                for (int g = 0; g < G && g <= gf; g++) {
                    const bool pend = (pending >> g) & 1u;
                    if (!pend && !first_round) continue;
                    const int tier = pend ? t_s[32 + g] : 1;
                    if (tier == 1) n1++; else if (tier == 2) n2++; else n3++;
                }
            }
            if (gf >= G) break;
            // ---- the flip of strain gf: P += (eta[t][b] - eta[cur][b]) * gamma[s][gf]; K = sum_b n_b lg2 P_b
            const int t = t_s[gf], cur = code_get(code, gf);
            {
                const double *et = eta_s + 4 * t, *ec2 = eta_s + 4 * cur;
                const double e0 = et[0] - ec2[0], e1 = et[1] - ec2[1], e2 = et[2] - ec2[2], e3 = et[3] - ec2[3];
                for (int c = wib; c < nch; c += TAUO_WARPS) {
                    const int s = c * 32 + lane;
                    const double gg = gT[gf * Sp + s];
                    const double2 P01 = Pw[s * 2], P23 = Pw[s * 2 + 1];
                    const double b0 = fma(e0, gg, P01.x), b1 = fma(e1, gg, P01.y), b2 = fma(e2, gg, P23.x), b3 = fma(e3, gg, P23.y);
                    Pw[s * 2] = make_double2(b0, b1); Pw[s * 2 + 1] = make_double2(b2, b3);
                    const int4 n = tile[s];
                    const float l0 = lg2_fast((float)b0), l1 = lg2_fast((float)b1), l2 = lg2_fast((float)b2), l3 = lg2_fast((float)b3);
                    Kw[s] = fmaf((float)n.x, l0, fmaf((float)n.y, l1, fmaf((float)n.z, l2, (float)n.w * l3)));
                    mlPs[s] = fmaxf(fmaxf(fabsf(l0), fabsf(l1)), fmaxf(fabsf(l2), fabsf(l3)));
                }
            }
            code = code_set(code, gf, t);
            if (threadIdx.x == 0) {
                flips++;
                if (p.tau_cnt) {
                    const size_t vg = (size_t)v * G + gf;
                    p.tau_cnt[vg * 4 + cur] += p.iter - p.tau_last[vg];
                    p.tau_last[vg] = p.iter;
                }
            }
            pending = (gf + 1 >= 32) ? 0u : (fullG & ~((2u << gf) - 1u));   // every strain after the flip
            from_screening = false; first_round = false;
            __syncthreads();                                                 // (t_s / first_flip read by all; P, K rewritten)
            if (threadIdx.x == 0) first_flip = G;
            if (pending == 0u) break;
            __syncthreads();
        }
#ifdef KPROF
        kp_steps += gtimer() - kp_b;
#endif
        if (code != code_in && wib == 0) {
            if (lane < G) p.tau[(size_t)v * G + lane] = (uint8_t)code_get(code, lane);
            if (p.agg.N) {
                const int sn = agg_move_site(p.agg, code_in, code, tile, lane);
                if (p.site_slot && lane == 0) {
                    p.site_slot[v] = sn; atomicAdd(p.gctl + GC_ORPHANS, 1);
                    if (p.site_row) { const int r = p.site_row[v]; if (r >= 0) p.img_site[r] = ~v; }
                }
            }
        }
        __syncthreads();
    }
#ifdef KPROF
    if (threadIdx.x == 0) krec_put(KP_TAU_WARP, (int)blockIdx.x, 0, (int)(kp_sites | (n2 << 8) | (n3 << 20) | (flips << 24)), kp_t0, gtimer(), kp_rounds | (kp_q[0] << 8) | (kp_q[1] << 24) | (kp_q[2] << 44), kp_stage, kp_steps, 0);
#endif
    if (wib == 0 && lane == 0) {
        if (flips) atomicAdd(p.nchange, (unsigned long long)flips);
        if (p.tier_counts) {
            if (n1) atomicAdd(p.tier_counts + 0, (unsigned long long)n1);
            if (n2) atomicAdd(p.tier_counts + 1, (unsigned long long)n2);
            if (n3) atomicAdd(p.tier_counts + 2, (unsigned long long)n3);
        }
    }
}

static inline size_t tauo_smem_bytes(int S, int G)
{
    const size_t Sp = (size_t)((S + 31) & ~31);
    return sizeof(double) * ((size_t)G * Sp + 16 + Sp * 4) + sizeof(float) * ((size_t)G * Sp + 16 + 2 * Sp) + sizeof(int4) * Sp +
           sizeof(uint32_t) * 32 + sizeof(int) * 64;
}

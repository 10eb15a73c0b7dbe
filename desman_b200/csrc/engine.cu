// desman_b200/csrc/engine.cu -- host side of libdesman_b200.so: contexts, device memory, streams,
// the sweep drivers and the C-ABI declared in include/desman_b200.h.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <emmintrin.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <atomic>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/desman_b200.h"
#include "common.cuh"
#include "misc_kernels.cuh"
#include "mu_agg_kernel.cuh"
#include "mu_kernel.cuh"
#include "nmft_kernel.cuh"
#include "tau_kernel.cuh"
#include "tau_open_kernel.cuh"
#include "tau_group_kernel.cuh"
#include "tau_group_tc_kernel.cuh"
#include "state_kernel.cuh"
#include "maintain_kernel.cuh"
#include "exchange_kernel.cuh"

// ------------------------------------------------------------------------------------------ errors
static thread_local char g_err[1024] = "";
static int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
#define CU(call)                                                                                  \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess)                                                                    \
            return fail(DESMAN_ECUDA, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__,        \
                        cudaGetErrorString(e_));                                                  \
    } while (0)
#define RET(call)                   \
    do {                            \
        int r_ = (call);            \
        if (r_ != DESMAN_OK) return r_; \
    } while (0)

extern "C" const char *desman_last_error(void) { return g_err; }
extern "C" const char *desman_build_info(void) { return "desman_b200 0.1 (sm_100a; kernels: tau_sample, mu_stats, draw_gamma_eta, finalize_sweep, mt19937, nmft)"; }
extern "C" int desman_device_count(int *n)
{
    CU(cudaGetDeviceCount(n));
    return DESMAN_OK;
}

// ------------------------------------------------------------------------------------------ NCCL (dlopen)
// The per-sweep exchange is one small all-reduce.  NCCL is bound at run time so the single-GPU
// library has no link-time dependency and shares whatever libnccl the process already loaded.
typedef struct { char internal[128]; } nccl_uid_t;
typedef void *nccl_comm_t;
struct NcclApi {
    void *h = nullptr;
    int (*GetUniqueId)(nccl_uid_t *) = nullptr;
    int (*CommInitRank)(nccl_comm_t *, int, nccl_uid_t, int) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*CommDestroy)(nccl_comm_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
};
static NcclApi g_nccl;
enum { NCCL_INT8 = 0, NCCL_INT64 = 4, NCCL_UINT64 = 5, NCCL_FLOAT64 = 8, NCCL_SUM = 0 };

static int nccl_load()
{
    if (g_nccl.h) return DESMAN_OK;
    const char *names[] = {"libnccl.so.2", "libnccl.so", nullptr};
    void *h = nullptr;
    for (int i = 0; names[i] && !h; i++) h = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
    const char *env = getenv("DESMAN_B200_NCCL");
    if (!h && env) h = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
    if (!h) return fail(DESMAN_ECOMM, "cannot dlopen libnccl.so.2 (set DESMAN_B200_NCCL=/path/to/libnccl.so.2): %s", dlerror());
    g_nccl.GetUniqueId = (int (*)(nccl_uid_t *))dlsym(h, "ncclGetUniqueId");
    g_nccl.CommInitRank = (int (*)(nccl_comm_t *, int, nccl_uid_t, int))dlsym(h, "ncclCommInitRank");
    g_nccl.AllReduce = (int (*)(const void *, void *, size_t, int, int, nccl_comm_t, cudaStream_t))dlsym(h, "ncclAllReduce");
    g_nccl.AllGather = (int (*)(const void *, void *, size_t, int, nccl_comm_t, cudaStream_t))dlsym(h, "ncclAllGather");
    g_nccl.CommDestroy = (int (*)(nccl_comm_t))dlsym(h, "ncclCommDestroy");
    g_nccl.GetErrorString = (const char *(*)(int))dlsym(h, "ncclGetErrorString");
    g_nccl.GroupStart = (int (*)())dlsym(h, "ncclGroupStart");
    g_nccl.GroupEnd = (int (*)())dlsym(h, "ncclGroupEnd");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy)
        return fail(DESMAN_ECOMM, "libnccl is missing required symbols");
    g_nccl.h = h;
    return DESMAN_OK;
}
#define NC(call)                                                                                         \
    do {                                                                                                 \
        int e_ = (call);                                                                                 \
        if (e_ != 0)                                                                                     \
            return fail(DESMAN_ECOMM, "%s failed: %s", #call, g_nccl.GetErrorString ? g_nccl.GetErrorString(e_) : "?"); \
    } while (0)

// ------------------------------------------------------------------------------------------ context
struct desman_ctx {
    int device = 0, sm_count = 148;
    cudaStream_t stream = nullptr;
    int rng_mode = DESMAN_RNG_PHILOX;
    uint64_t seed = 0;
    uint32_t sweep = 0;            // Philox sweep counter (persists across update() calls)
    uint64_t mt_consumed = 0;      // MT19937 words consumed so far
    int mt_pos = 624;              // position inside the current 624-word block
    double alpha = 0.1, delta = 0.1, epsilon = 1e-6;

    int64_t V = 0, v0 = 0, V_total = 0;
    int S = 0, G = 0;
    // device buffers
    int4 *counts = nullptr;
    uint8_t *tau = nullptr, *tau_star = nullptr;
    double *gamma = nullptr, *eta = nullptr, *eta_new = nullptr, *gamma_star = nullptr, *eta_star = nullptr;
    unsigned long long *stats = nullptr;     // [S*G + 16] sum_mu | esum
    unsigned long long *red_base = nullptr;  // [2][4] two parities of ...
    unsigned long long *red_i = nullptr;     // [4] fixed-point sum n*log p | nchange | upkeep wishes | - of the current sweep (= red_base + 4*parity)
    double *scal = nullptr;                  // [4] lp_star, iter_star, ll, lp
    int *flag = nullptr;
    uint32_t *tau_cnt = nullptr, *tau_last = nullptr;
    uint32_t *mt_state = nullptr, *words = nullptr;
    size_t words_cap = 0;
    double ll_const = 0.0;
    bool ll_const_valid = false;
    size_t counts_cap = 0;
    size_t cap_vg = 0, cap_sg = 0;
    uint32_t last_n_iter = 0;
    int pdl = 1;                             // programmatic dependent launch of the sweep's kernels (common.cuh pdl_enter)
    int tau_exact = 0;                       // 1: FP64 reference-order path for every draw
    int fixed_tau = 0;                       // 1: update() skips the tau draw (update_fixed_tau, HaploSNP_Sampler.py:409-428)
    int mu_mode = 2;                         // 1: pattern-aggregated binomial statistics (K2b), 0: per-read categorical (K2),
                                             // 2: choose from (V, G): aggregated iff the ~12*2^G possible biallelic patterns are <= V/2
    unsigned long long *agg_keys = nullptr, *agg_code = nullptr, *agg_N = nullptr, *agg_classM = nullptr;
    int classM_G = 0, classM_S = 0;
    int *agg_ids = nullptr;
    unsigned int *agg_nslots = nullptr;
    int *agg_ctl = nullptr;                  // [AGG_CTL_WORDS] rebuild wanted | - | overflow | flips since rebuild | work cursors of K2b
    size_t agg_H = 0, agg_cap_slots = 0, agg_cap_cells = 0;
    bool agg_valid = false;                  // device table reflects the current device tau and counts
    double total_reads = 0.0, ll_scale = 1.0;
    double ll_const_total = 0.0;
    unsigned long long *tiers = nullptr;     // [3] draws decided by tier 1/2/3
    // pattern groups for the screening pass of the tau update (tau_group_kernel.cuh)
    int tau_group = 2;                       // 0: off, 1: on, 2: on iff the ~12*2^G biallelic patterns are <= V/2
    int tau_group_mma = 1;                   // 1: tensor-core form of the screening pass where it applies; 0: FFMA form
    uint2 *pack16[2] = {nullptr, nullptr};   // device staging of a chunk of 4 x uint16 count cells (desman_set_counts)
    bool star_fold = false;                  // sharded chain: the next screening launch takes the MAP snapshot (tau_star <- tau if flag)
    int xch_fuse = 1;                        // exchange + bookkeeping of the previous sweep as one launch (peer-memory exchange only)
    int tau_open = 1;                        // 1: the work list of the screening pass is walked by tau_open_kernel (one CTA per site)
    int tauo_grid = 0;
    size_t tauo_smem = 0;
    int tau_group_tc = 1;                    // 1: tcgen05 / TMEM / TMA form (tau_group_tc_kernel.cuh) where it applies; 0: mma.sync / FFMA forms
    unsigned char *img = nullptr;            // fp16x4 count image in UMMA operand order
    int *img_site = nullptr, *site_row = nullptr;
    float *img_nsite = nullptr;
    size_t img_cap_rows = 0, img_bytes = 0, img_cap_v = 0;
    unsigned long long *esum_store = nullptr;   // [n_iter][16] Esum of every sweep of the last update()
    size_t esum_store_cap = 0;
    bool counts_tf32_exact = false;
    float4 *countsf = nullptr;               // [V][S] FP32 copy of the counts
    float *nsite = nullptr;                  // [V]
    size_t countsf_cap = 0;
    int *grp_site_slot = nullptr, *grp_order = nullptr, *grp_singles = nullptr, *grp_slot4 = nullptr, *grp_gctl = nullptr,
        *grp_blk = nullptr;
    int4 *grp_items = nullptr;
    uint2 *grp_work = nullptr;
    size_t grp_cap_v = 0, grp_cap_slots = 0;
    int maint_grid = 0;
    cudaEvent_t pin_ev[2] = {nullptr, nullptr};
    // scratch
    void *scratch = nullptr;
    size_t scratch_cap = 0;
    // comm
    nccl_comm_t comm = nullptr;
    int rank = 0, nranks = 1;
    // peer-memory mailboxes of the one-shot exchange (exchange_kernel.cuh); xch_ok == false: NCCL all-reduce instead
    bool xch_ok = false;
    unsigned long long *xch_mail[XCH_MAX_RANKS] = {nullptr};
    int xch_cap_words = 0;
    unsigned long long xch_seq = 0;
    int *xch_err = nullptr;
    // profiling
    int prof_kernels = 0, prof_flush = 0;
    uint4 *flush_buf = nullptr;
    size_t flush_n = 0;
    std::vector<cudaEvent_t> ev_pool;
    size_t ev_used = 0;
    struct Span { int kind; cudaEvent_t a, b; };
    std::vector<Span> spans;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> sweep_spans;
    double elapsed_ms = 0.0, k_ms[DESMAN_K_COUNT] = {0};
    int64_t k_launch[DESMAN_K_COUNT] = {0};
};

// Device buffers come from the device's stream-ordered pool (cudaMallocAsync) with the release threshold lifted, so the large
// buffers of a closed context are handed to the next one without a trip to the driver (an engine is created per
// HaploSNP_Sampler object: tens of ms of cudaMalloc/cudaFree otherwise).  The IPC mailboxes keep cudaMalloc (IPC needs it).
template <typename T>
static cudaError_t dmalloc(desman_ctx *c, T **p, size_t bytes)
{
    return cudaMallocAsync((void **)p, bytes ? bytes : 1, c->stream);
}
static void dfree(desman_ctx *c, void *p)
{
    if (p) cudaFreeAsync(p, c->stream);
}

// Launch of a kernel of the sweep's dependent chain (every such kernel begins with pdl_enter(), common.cuh): with the
// programmatic-stream-serialization attribute the grid may be scheduled while its predecessor drains; its pdl_enter()
// still orders everything it does after the predecessor's completion.
template <typename... KA, typename... A>
static cudaError_t launch_k(desman_ctx *c, void (*kern)(KA...), unsigned int grid, unsigned int block, size_t smem, A... args)
{
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = c->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = c->pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, KA(args)...);
}

static int ensure_scratch(desman_ctx *c, size_t bytes)
{
    if (bytes <= c->scratch_cap) return DESMAN_OK;
    if (c->scratch) dfree(c, c->scratch);
    c->scratch = nullptr; c->scratch_cap = 0;
    CU(dmalloc(c, &c->scratch, bytes));
    c->scratch_cap = bytes;
    return DESMAN_OK;
}

static cudaEvent_t get_event(desman_ctx *c)
{
    if (c->ev_used == c->ev_pool.size()) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        c->ev_pool.push_back(e);
    }
    return c->ev_pool[c->ev_used++];
}

struct KSpan {  // brackets n kernel launches with events when per-kernel profiling is on; k_launch counts the launches
    desman_ctx *c; int kind; cudaEvent_t a = nullptr;
    KSpan(desman_ctx *c_, int kind_, int n = 1) : c(c_), kind(kind_)
    {
        c->k_launch[kind] += n;
        // level 1: every span; level 2: only the span around the whole tau update
        if (c->prof_kernels == 1 ? kind != DESMAN_K_TAU_UPDATE : (c->prof_kernels == 2 && kind == DESMAN_K_TAU_UPDATE)) {
            a = get_event(c); cudaEventRecord(a, c->stream);
        }
    }
    ~KSpan()
    {
        if (a) { cudaEvent_t b = get_event(c); cudaEventRecord(b, c->stream); c->spans.push_back({kind, a, b}); }
    }
};

static void timing_reset(desman_ctx *c)
{
    c->ev_used = 0; c->spans.clear(); c->sweep_spans.clear();
    c->elapsed_ms = 0.0;
    for (int i = 0; i < DESMAN_K_COUNT; i++) { c->k_ms[i] = 0.0; c->k_launch[i] = 0; }
}
static void timing_collect(desman_ctx *c)
{
    for (auto &s : c->spans) { float ms = 0; cudaEventElapsedTime(&ms, s.a, s.b); c->k_ms[s.kind] += ms; }
    for (auto &s : c->sweep_spans) { float ms = 0; cudaEventElapsedTime(&ms, s.first, s.second); c->elapsed_ms += ms; }
}
static void sweep_begin(desman_ctx *c)
{
    cudaEvent_t a = get_event(c);
    cudaEventRecord(a, c->stream);
    c->sweep_spans.push_back({a, nullptr});
}
static void sweep_end(desman_ctx *c)
{
    cudaEvent_t b = get_event(c);
    cudaEventRecord(b, c->stream);
    c->sweep_spans.back().second = b;
    if (c->prof_flush && c->flush_buf) l2_flush_kernel<<<c->sm_count * 4, 256, 0, c->stream>>>(c->flush_buf, c->flush_n);
}

extern "C" int desman_ctx_create(desman_ctx **out, int device, uint64_t seed, int rng_mode)
{
    if (!out) return fail(DESMAN_EINVAL, "desman_ctx_create: out is NULL");
    if (rng_mode != DESMAN_RNG_MT19937 && rng_mode != DESMAN_RNG_PHILOX) return fail(DESMAN_EINVAL, "bad rng_mode %d", rng_mode);
    int n = 0;
    CU(cudaGetDeviceCount(&n));
    if (device < 0 || device >= n) return fail(DESMAN_EINVAL, "device %d out of range (have %d)", device, n);
    CU(cudaSetDevice(device));
    desman_ctx *c = new desman_ctx();
    c->device = device;
    c->rng_mode = rng_mode;
    CU(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device));
    {
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            unsigned long long thr = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
        } else cudaGetLastError();
    }
    CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    {
        double inv[65];
        inv[0] = 0.0;
        for (int x = 1; x <= 64; x++) inv[x] = 1.0 / (double)x;
        CU(cudaMemcpyToSymbol(c_inv_small, inv, sizeof(inv)));
    }
    CU(dmalloc(c, &c->eta, 16 * sizeof(double)));
    CU(dmalloc(c, &c->eta_new, 16 * sizeof(double)));
    CU(dmalloc(c, &c->eta_star, 16 * sizeof(double)));
    CU(dmalloc(c, &c->red_base, 8 * sizeof(unsigned long long)));
    CU(cudaMemsetAsync(c->red_base, 0, 8 * sizeof(unsigned long long), c->stream));
    c->red_i = c->red_base;
    CU(dmalloc(c, &c->scal, 4 * sizeof(double)));
    CU(dmalloc(c, &c->flag, sizeof(int)));
    CU(dmalloc(c, &c->tiers, 3 * sizeof(unsigned long long)));
    CU(cudaMemset(c->tiers, 0, 3 * sizeof(unsigned long long)));
    { const char *pd = getenv("DESMAN_B200_PDL"); if (pd) c->pdl = atoi(pd) ? 1 : 0; }
    { const char *ex = getenv("DESMAN_B200_TAU_EXACT"); c->tau_exact = (ex && atoi(ex)) ? 1 : 0; }
    { const char *tg = getenv("DESMAN_B200_TAU_GROUP"); if (tg) c->tau_group = atoi(tg); if (c->tau_group < 0 || c->tau_group > 2) c->tau_group = 2; }
    { const char *tc = getenv("DESMAN_B200_TAU_GROUP_TC"); if (tc) c->tau_group_tc = atoi(tc) ? 1 : 0; }
    { const char *mm = getenv("DESMAN_B200_MU_MODE"); if (mm) c->mu_mode = atoi(mm); if (c->mu_mode < 0 || c->mu_mode > 2) c->mu_mode = 2; }
    CU(dmalloc(c, &c->mt_state, 624 * sizeof(uint32_t)));
    CU(cudaMemset(c->scal, 0, 4 * sizeof(double)));
    *out = c;
    return desman_set_rng(c, seed, 0, 0);
}

extern "C" int desman_ctx_destroy(desman_ctx *c)
{
    if (!c) return DESMAN_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (int r = 0; r < XCH_MAX_RANKS; r++) {
        if (!c->xch_mail[r]) continue;
        if (r == c->rank) cudaFree(c->xch_mail[r]); else cudaIpcCloseMemHandle(c->xch_mail[r]);
    }
    if (c->xch_err) cudaFree(c->xch_err);
    if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
    for (void *q : {(void *)c->img, (void *)c->img_site, (void *)c->img_nsite, (void *)c->site_row, (void *)c->esum_store, (void *)c->pack16[0],
                    (void *)c->pack16[1]}) if (q) dfree(c, q);
    void *ptrs[] = {c->counts, c->tau, c->tau_star, c->gamma, c->eta, c->eta_new, c->gamma_star, c->eta_star, c->stats,
                    c->red_base, c->agg_ctl, c->scal, c->flag, c->tau_cnt, c->tau_last, c->mt_state, c->words,
                    c->scratch, c->flush_buf, c->tiers, c->agg_keys, c->agg_code, c->agg_N, c->agg_ids, c->agg_nslots, c->agg_classM,
                    c->countsf, c->nsite, c->grp_site_slot, c->grp_order, c->grp_singles, c->grp_slot4, c->grp_gctl, c->grp_blk,
                    c->grp_items, c->grp_work};
    for (void *p : ptrs) if (p) dfree(c, p);
    for (auto e : c->ev_pool) cudaEventDestroy(e);
    for (int i = 0; i < 2; i++) if (c->pin_ev[i]) cudaEventDestroy(c->pin_ev[i]);
    cudaStreamDestroy(c->stream);
    delete c;
    return DESMAN_OK;
}

extern "C" int desman_synchronize(desman_ctx *c)
{
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    return DESMAN_OK;
}

extern "C" int desman_set_hyper(desman_ctx *c, double alpha, double delta, double epsilon)
{
    if (!(alpha > 0) || !(delta > 0) || !(epsilon >= 0)) return fail(DESMAN_EINVAL, "alpha, delta must be > 0 and epsilon >= 0");
    c->alpha = alpha; c->delta = delta; c->epsilon = epsilon;
    return DESMAN_OK;
}

// GSL gsl_rng_set for gsl_rng_mt19937 (seed 0 -> 4357), then skip `consumed` words.
static int mt_seed(desman_ctx *c, uint64_t seed, uint64_t consumed)
{
    uint32_t st[624];
    uint32_t s = (uint32_t)(seed & 0xffffffffu);
    if (s == 0) s = 4357u;
    st[0] = s;
    for (int i = 1; i < 624; i++) st[i] = 1812433253u * (st[i - 1] ^ (st[i - 1] >> 30)) + (uint32_t)i;
    CU(cudaMemcpyAsync(c->mt_state, st, sizeof(st), cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    c->mt_pos = 624;
    c->mt_consumed = 0;
    if (consumed) {  // advance without storing
        mt19937_kernel<<<1, 256, 0, c->stream>>>(c->mt_state, c->mt_pos, (size_t)consumed, 0, 0, nullptr);
        CU(cudaGetLastError());
        c->mt_pos = (int)((consumed - 1) % 624) + 1;
        c->mt_consumed = consumed;
    }
    return DESMAN_OK;
}

extern "C" int desman_set_rng(desman_ctx *c, uint64_t seed, uint32_t sweep, uint64_t mt_words_consumed)
{
    CU(cudaSetDevice(c->device));
    c->seed = seed;
    c->sweep = sweep;
    return mt_seed(c, seed, mt_words_consumed);
}

extern "C" int desman_get_rng(desman_ctx *c, uint32_t *sweep, uint64_t *mt_words_consumed)
{
    if (sweep) *sweep = c->sweep;
    if (mt_words_consumed) *mt_words_consumed = c->mt_consumed;
    return DESMAN_OK;
}

// ------------------------------------------------------------------------------------------ data
// process-wide pinned staging buffers (pinning is slow): 2 x 32 MB, used by the count upload and the large downloads
static const size_t PIN_CELLS = (size_t)2 << 20;                       // int4 cells per buffer
static int4 *g_pin[2] = {nullptr, nullptr};
static std::mutex g_pin_mu;
static int ensure_pinned()
{
    if (!g_pin[0]) {
        CU(cudaMallocHost(&g_pin[0], PIN_CELLS * sizeof(int4)));
        CU(cudaMallocHost(&g_pin[1], PIN_CELLS * sizeof(int4)));
    }
    return DESMAN_OK;
}
// host-side conversion loops run on a few threads (the arrays of the class surface are tens to hundreds of MB)
static int host_threads(size_t n)
{
    if (n < ((size_t)1 << 16)) return 1;
    unsigned int hc = std::thread::hardware_concurrency();
    if (hc == 0) hc = 4;
    return (int)(hc > 16 ? 16 : hc);
}
template <typename F>
static void parallel_ranges(size_t n, F f)
{
    const int nt = host_threads(n);
    if (nt == 1) { f((size_t)0, n); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < nt; t++) th.emplace_back(f, n * t / nt, n * (t + 1) / nt);
    for (auto &x : th) x.join();
}

extern "C" int desman_set_counts(desman_ctx *c, const int64_t *variants, int64_t V, int S, int64_t v0, int64_t V_total)
{
    if (!variants || V <= 0 || S <= 0) return fail(DESMAN_EINVAL, "desman_set_counts: need V > 0, S > 0 and a counts pointer");
    if (V > 0x7fffffff || S >= (1 << 26)) return fail(DESMAN_EINVAL, "V or S too large");
    if (V_total <= 0) V_total = V;
    if (v0 < 0 || v0 + V > V_total) return fail(DESMAN_EINVAL, "shard [v0, v0+V) outside [0, V_total)");
    CU(cudaSetDevice(c->device));
    const size_t ncell = (size_t)V * S;
    if (ncell > c->counts_cap) {
        if (c->counts) dfree(c, c->counts);
        c->counts = nullptr; c->counts_cap = 0;
        CU(dmalloc(c, &c->counts, ncell * sizeof(int4)));
        c->counts_cap = ncell;
    }
    // Repack int64 -> int32x4 on the host with a few threads straight into pinned staging buffers and stream the
    // packed cells (half the bytes of the int64 tensor) to the device, double-buffered.
    const size_t chunk_cells = PIN_CELLS;                               // 32 MB of packed cells per buffer
    std::lock_guard<std::mutex> pin_lock(g_pin_mu);
    RET(ensure_pinned());
    if (!c->pin_ev[0]) {
        CU(cudaEventCreateWithFlags(&c->pin_ev[0], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&c->pin_ev[1], cudaEventDisableTiming));
    }
    std::atomic<int> bad(0);
    std::atomic<long long> total(0), or_all(0);
    int buf = 0;
    // Counts below 2^16 (the usual case) travel as 4 x uint16 per cell -- a quarter of the int64 tensor's bytes over PCIe and
    // half the host-side writes -- and are widened to the canonical int32x4 cells by a kernel; a chunk that holds a larger
    // count is simply packed again as int32x4.  The host pass is bound by memory traffic (205 MB of int64 to read at C3).
    if (!c->pack16[0]) {
        CU(dmalloc(c, &c->pack16[0], chunk_cells * sizeof(uint2)));
        CU(dmalloc(c, &c->pack16[1], chunk_cells * sizeof(uint2)));
    }
    for (size_t off = 0; off < ncell; off += chunk_cells, buf ^= 1) {
        const size_t n = (ncell - off < chunk_cells) ? ncell - off : chunk_cells;
        CU(cudaEventSynchronize(c->pin_ev[buf]));                       // previous copy out of this buffer finished
        int4 *dst = g_pin[buf];
        const int64_t *src = variants + off * 4;
        const int nt = host_threads(n);
        std::atomic<long long> or_chunk(0), sum_chunk(0);
        auto work = [&](int t, bool narrow) {
            const size_t lo = n * t / nt, hi = n * (t + 1) / nt;
            int64_t orv = 0, sum = 0, orc = 0;
            if (narrow) {
                long long *d16 = reinterpret_cast<long long *>(dst);
                for (size_t i = lo; i < hi; i++) {
                    const int64_t a = src[4 * i], b = src[4 * i + 1], d = src[4 * i + 2], e = src[4 * i + 3];
                    orc |= a | b | d | e;
                    sum += a + b + d + e;
                    // (streaming stores: the pinned buffer is written once and read by the DMA engine -- no read-for-ownership)
                    _mm_stream_si64(d16 + i, (long long)(((uint64_t)a & 0xffffull) | (((uint64_t)b & 0xffffull) << 16) |
                                                         (((uint64_t)d & 0xffffull) << 32) | (((uint64_t)e & 0xffffull) << 48)));
                }
                or_chunk |= orc;                                        // (accounted by the caller if the whole chunk is in range)
                sum_chunk += sum;
            } else {
                for (size_t i = lo; i < hi; i++) {
                    const int64_t a = src[4 * i], b = src[4 * i + 1], d = src[4 * i + 2], e = src[4 * i + 3];
                    orc |= a | b | d | e;
                    orv |= a | b | d | e | (DESMAN_MAX_COUNT - a) | (DESMAN_MAX_COUNT - b) | (DESMAN_MAX_COUNT - d) | (DESMAN_MAX_COUNT - e);
                    sum += a + b + d + e;
                    _mm_stream_si128(reinterpret_cast<__m128i *>(dst + i), _mm_set_epi32((int)e, (int)d, (int)b, (int)a));
                }
                total += sum;
                or_all |= orc;
                if (orv < 0) bad = 1;                                   // a negative count or one above the limit
            }
            _mm_sfence();
        };
        auto run = [&](bool narrow) {
            if (nt == 1) { work(0, narrow); return; }
            std::vector<std::thread> th;
            for (int t = 0; t < nt; t++) th.emplace_back(work, t, narrow);
            for (auto &x : th) x.join();
        };
        run(true);
        const long long oc = or_chunk.load();
        if (oc >= 0 && oc < 65536) {
            total += sum_chunk.load();
            or_all |= oc;
            CU(cudaMemcpyAsync(c->pack16[buf], dst, n * sizeof(uint2), cudaMemcpyHostToDevice, c->stream));
            CU(cudaEventRecord(c->pin_ev[buf], c->stream));
            widen_counts_kernel<<<c->sm_count * 4, 256, 0, c->stream>>>(c->pack16[buf], c->counts + off, n);
        } else {                                                        // a count >= 2^16 (or a negative one): the wide form
            run(false);
            CU(cudaMemcpyAsync(c->counts + off, dst, n * sizeof(int4), cudaMemcpyHostToDevice, c->stream));
            CU(cudaEventRecord(c->pin_ev[buf], c->stream));
        }
    }
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(c->stream));
    if (bad) { c->V = 0; return fail(DESMAN_EINVAL, "counts must be in [0, %d] per (v,s,base) cell", DESMAN_MAX_COUNT); }
    if (V != c->V || S != c->S) c->G = 0;  // state must be (re)set for a new shape
    c->V = V; c->S = S; c->v0 = v0; c->V_total = V_total;
    c->ll_const_valid = false;
    c->agg_valid = false;
    c->total_reads = (double)total.load();
    c->counts_tf32_exact = or_all.load() < 2048;   // every count < 2^11: exact as a TF32 operand
    return DESMAN_OK;
}

// constant multinomial-coefficient part of the log-likelihood (Desman_Utils.py:28-33), computed on first use
static int ensure_ll_const(desman_ctx *c)
{
    if (c->ll_const_valid) return DESMAN_OK;
    const int nb = c->sm_count * 4;
    double *dpart = nullptr;
    CU(dmalloc(c, &dpart, nb * sizeof(double)));
    lgamma_const_kernel<<<nb, 256, 0, c->stream>>>(c->counts, (size_t)c->V * c->S, dpart);
    CU(cudaGetLastError());
    std::vector<double> part(nb);
    CU(cudaMemcpyAsync(part.data(), dpart, nb * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    dfree(c, dpart);
    double t = 0.0;
    for (double x : part) t += x;
    c->ll_const = t;
    c->ll_const_total = t;
    if (c->nranks > 1) {   // one-time sum over the shards
        double *d = nullptr;
        CU(dmalloc(c, &d, sizeof(double)));
        CU(cudaMemcpyAsync(d, &t, sizeof(double), cudaMemcpyHostToDevice, c->stream));
        NC(g_nccl.AllReduce(d, d, 1, NCCL_FLOAT64, NCCL_SUM, c->comm, c->stream));
        CU(cudaMemcpyAsync(&c->ll_const_total, d, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        dfree(c, d);
    }
    c->ll_const_valid = true;
    return DESMAN_OK;
}

static int ensure_state(desman_ctx *c, int G)
{
    if (c->V <= 0) return fail(DESMAN_ESTATE, "set counts before state");
    if (G < 1 || G > DESMAN_MAX_G) return fail(DESMAN_EINVAL, "G must be in [1, %d]", DESMAN_MAX_G);
    const size_t nvg = (size_t)c->V * G, nsg = (size_t)c->S * G;
    if (nvg > c->cap_vg) {
        for (void *p : {(void *)c->tau, (void *)c->tau_star, (void *)c->tau_cnt, (void *)c->tau_last}) if (p) dfree(c, p);
        c->tau = c->tau_star = nullptr; c->tau_cnt = c->tau_last = nullptr; c->cap_vg = 0;
        CU(dmalloc(c, &c->tau, nvg));
        CU(dmalloc(c, &c->tau_star, nvg));
        CU(dmalloc(c, &c->tau_cnt, nvg * 4 * sizeof(uint32_t)));
        CU(dmalloc(c, &c->tau_last, nvg * sizeof(uint32_t)));
        c->cap_vg = nvg;
    }
    if (nsg > c->cap_sg) {
        for (void *p : {(void *)c->gamma, (void *)c->gamma_star, (void *)c->stats}) if (p) dfree(c, p);
        c->gamma = c->gamma_star = nullptr; c->stats = nullptr; c->cap_sg = 0;
        CU(dmalloc(c, &c->gamma, nsg * sizeof(double)));
        CU(dmalloc(c, &c->gamma_star, nsg * sizeof(double)));
        CU(dmalloc(c, &c->stats, (nsg + 16) * sizeof(unsigned long long)));
        c->cap_sg = nsg;
    }
    if (G != c->G) {
        CU(cudaMemsetAsync(c->tau_cnt, 0, nvg * 4 * sizeof(uint32_t), c->stream));
        CU(cudaMemsetAsync(c->tau_last, 0, nvg * sizeof(uint32_t), c->stream));
    }
    if (G != c->G) c->agg_valid = false;
    c->G = G;
    return DESMAN_OK;
}

// int64 one-hot [n,4] -> uint8 index; first b with tau == 1 (c_sample_tau.c:115-123); rows without a 1 are rejected
static int onehot_to_index(const int64_t *tau, size_t n, uint8_t *idx)
{
    std::atomic<long long> bad(-1);
    parallel_ranges(n, [&](size_t lo, size_t hi) {
        long long mybad = -1;
        for (size_t i = lo; i < hi; i++) {
            const int64_t *t = tau + i * 4;
            // a valid row is {0,1}^4 with exactly one 1: index = t1 + 2 t2 + 3 t3 (branch-free); anything else falls back to
            // the reference's "first b with tau == 1" rule (c_sample_tau.c:115-123) or is rejected
            const int64_t orv = t[0] | t[1] | t[2] | t[3], sum = t[0] + t[1] + t[2] + t[3];
            int b = (int)(t[1] + 2 * t[2] + 3 * t[3]);
            if ((orv & ~(int64_t)1) != 0 || sum != 1) {
                b = t[0] == 1 ? 0 : t[1] == 1 ? 1 : t[2] == 1 ? 2 : t[3] == 1 ? 3 : -1;
                if (b < 0) { mybad = (long long)i; b = 0; }
            }
            idx[i] = (uint8_t)b;
        }
        if (mybad >= 0) bad = mybad;
    });
    if (bad >= 0)
        return fail(DESMAN_EINVAL, "tau row %lld is not one-hot (undefined behaviour in the reference, c_sample_tau.c:115-123)", bad.load());
    return DESMAN_OK;
}

extern "C" int desman_set_tau_index(desman_ctx *c, const uint8_t *tau_idx, int G)
{
    CU(cudaSetDevice(c->device));
    RET(ensure_state(c, G));
    const size_t nvg = (size_t)c->V * G;
    for (size_t i = 0; i < nvg; i++) if (tau_idx[i] > 3) return fail(DESMAN_EINVAL, "tau index %zu out of range", i);
    CU(cudaMemcpyAsync(c->tau, tau_idx, nvg, cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    c->agg_valid = false;
    return DESMAN_OK;
}

extern "C" int desman_get_tau_index(desman_ctx *c, uint8_t *tau_idx)
{
    if (!c->G) return fail(DESMAN_ESTATE, "no state");
    CU(cudaSetDevice(c->device));
    CU(cudaMemcpyAsync(tau_idx, c->tau, (size_t)c->V * c->G, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return DESMAN_OK;
}

// gamma_zero_ok: the reference ABI path (c_sample_tau) takes gamma with whole strain columns at exactly 0.0 -- the masked
// abundances of Eta_Sampler.maskGamma (Eta_Sampler.py:147-157,367), which the reference C accepts -- as long as every sample
// keeps a positive mixture; the chain drivers need log(gamma) for the prior and keep the strict test.
static int set_state_impl(desman_ctx *c, const int64_t *tau, const double *gamma, const double *eta, int G, bool gamma_zero_ok)
{
    CU(cudaSetDevice(c->device));
    RET(ensure_state(c, G));
    if (tau) {
        std::vector<uint8_t> idx((size_t)c->V * G);
        RET(onehot_to_index(tau, idx.size(), idx.data()));
        CU(cudaMemcpyAsync(c->tau, idx.data(), idx.size(), cudaMemcpyHostToDevice, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        c->agg_valid = false;
    }
    if (gamma) {
        for (int s = 0; s < c->S; s++) {
            double row = 0.0;
            for (int g = 0; g < G; g++) {
                const double x = gamma[(size_t)s * G + g];
                if (!(gamma_zero_ok ? x >= 0.0 : x > 0.0))
                    return fail(DESMAN_EINVAL, "gamma[%zu] = %g must be %s 0", (size_t)s * G + g, x, gamma_zero_ok ? ">=" : ">");
                row += x;
            }
            if (!(row > 0.0)) return fail(DESMAN_EINVAL, "gamma row %d has no positive entry", s);
        }
        CU(cudaMemcpyAsync(c->gamma, gamma, (size_t)c->S * G * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    }
    if (eta) {
        for (int i = 0; i < 16; i++) if (!(eta[i] > 0.0)) return fail(DESMAN_EINVAL, "eta[%d] = %g must be > 0", i, eta[i]);
        CU(cudaMemcpyAsync(c->eta, eta, 16 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    }
    CU(cudaStreamSynchronize(c->stream));
    return DESMAN_OK;
}

extern "C" int desman_set_state(desman_ctx *c, const int64_t *tau, const double *gamma, const double *eta, int G)
{
    return set_state_impl(c, tau, gamma, eta, G, false);
}

static void index_to_onehot(const uint8_t *idx, size_t n, int64_t *tau)
{
    parallel_ranges(n, [&](size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; i++) {
            int64_t *t = tau + i * 4;
            t[0] = t[1] = t[2] = t[3] = 0;
            t[idx[i] & 3] = 1;
        }
    });
}

extern "C" int desman_get_state(desman_ctx *c, int64_t *tau, double *gamma, double *eta)
{
    if (!c->G) return fail(DESMAN_ESTATE, "no state");
    CU(cudaSetDevice(c->device));
    std::vector<uint8_t> idx;
    if (tau) {
        idx.resize((size_t)c->V * c->G);
        CU(cudaMemcpyAsync(idx.data(), c->tau, idx.size(), cudaMemcpyDeviceToHost, c->stream));
    }
    if (gamma) CU(cudaMemcpyAsync(gamma, c->gamma, (size_t)c->S * c->G * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (eta) CU(cudaMemcpyAsync(eta, c->eta, 16 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (tau) index_to_onehot(idx.data(), idx.size(), tau);
    return DESMAN_OK;
}

// ------------------------------------------------------------------------------------------ launches
// One resident wave: sites are statically strided over the warps of the grid, so a partial second wave would
// run at a fraction of the occupancy for as long as a full one.
static int tau_grid(desman_ctx *c)
{
    int occ = 0;
    const size_t smem = tau_smem_bytes(c->S, c->G);
    cudaFuncSetAttribute(tau_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, tau_sample_kernel, TAU_WARPS * 32, smem) != cudaSuccess || occ < 1) occ = 1;
    int64_t want = (c->V + TAU_WARPS - 1) / TAU_WARPS;
    int64_t cap = (int64_t)c->sm_count * occ;
    return (int)(want < cap ? want : cap);
}

// Persistent pattern table (mu_agg_kernel.cuh).  Capacity 2.5 V slots: finalize asks for a rebuild once more than
// 1.25 V slots were handed out, and one tau pass can add at most V.
static int ensure_agg(desman_ctx *c)
{
    const size_t V = (size_t)c->V, slots = V * 5 / 2 + 64, cells = slots * c->S * 4;
    if (slots > c->agg_cap_slots || cells > c->agg_cap_cells) {
        for (void *q : {(void *)c->agg_keys, (void *)c->agg_code, (void *)c->agg_N, (void *)c->agg_ids, (void *)c->agg_nslots}) if (q) dfree(c, q);
        c->agg_keys = c->agg_code = c->agg_N = nullptr; c->agg_ids = nullptr; c->agg_nslots = nullptr;
        c->agg_cap_slots = c->agg_cap_cells = 0;
        size_t H = 64;
        while (H < 4 * slots / 3) H <<= 1;
        CU(dmalloc(c, &c->agg_keys, H * sizeof(unsigned long long)));
        CU(dmalloc(c, &c->agg_ids, H * sizeof(int)));
        CU(dmalloc(c, &c->agg_code, slots * sizeof(unsigned long long)));
        CU(dmalloc(c, &c->agg_nslots, sizeof(unsigned int)));
        CU(dmalloc(c, &c->agg_N, cells * sizeof(unsigned long long)));
        CU(cudaMemsetAsync(c->agg_N, 0, cells * sizeof(unsigned long long), c->stream));
        CU(cudaMemsetAsync(c->agg_nslots, 0, sizeof(unsigned int), c->stream));
        c->agg_H = H; c->agg_cap_slots = slots; c->agg_cap_cells = cells;
        c->agg_valid = false;
    }
    if (!c->agg_ctl) {
        CU(dmalloc(c, &c->agg_ctl, AGG_CTL_WORDS * sizeof(int)));
        CU(cudaMemsetAsync(c->agg_ctl, 0, AGG_CTL_WORDS * sizeof(int), c->stream));
    }
    // merged class totals of the statistics kernels: dense over the sets of strains
    if (c->G <= MUC_MAX_G && (c->G != c->classM_G || c->S != c->classM_S)) {
        if (c->agg_classM) dfree(c, c->agg_classM);
        c->agg_classM = nullptr; c->classM_G = c->classM_S = 0;
        const size_t n = ((size_t)1 << c->G) * c->S;
        CU(dmalloc(c, &c->agg_classM, n * sizeof(unsigned long long)));
        CU(cudaMemsetAsync(c->agg_classM, 0, n * sizeof(unsigned long long), c->stream));
        c->classM_G = c->G; c->classM_S = c->S;
    }
    // site groups of the screening pass
    if (V > c->grp_cap_v || c->agg_cap_slots > c->grp_cap_slots) {
        for (void *q : {(void *)c->grp_site_slot, (void *)c->grp_order, (void *)c->grp_singles, (void *)c->grp_slot4,
                        (void *)c->grp_items, (void *)c->grp_work}) if (q) dfree(c, q);
        c->grp_site_slot = c->grp_order = c->grp_singles = c->grp_slot4 = nullptr; c->grp_items = nullptr; c->grp_work = nullptr;
        c->grp_cap_v = c->grp_cap_slots = 0;
        CU(dmalloc(c, &c->grp_site_slot, V * sizeof(int)));
        CU(dmalloc(c, &c->grp_order, V * sizeof(int)));
        CU(dmalloc(c, &c->grp_singles, V * sizeof(int)));
        CU(dmalloc(c, &c->grp_slot4, 5 * c->agg_cap_slots * sizeof(int)));
        CU(dmalloc(c, &c->grp_items, 2 * (V / 2 + V / TG_ITEM_SITES + 64) * sizeof(int4)));
        CU(dmalloc(c, &c->grp_work, V * sizeof(uint2)));
        c->grp_cap_v = V; c->grp_cap_slots = c->agg_cap_slots;
        c->agg_valid = false;
    }
    if (!c->grp_gctl) {
        CU(dmalloc(c, &c->grp_gctl, GC_COUNT * sizeof(int)));
        CU(cudaMemsetAsync(c->grp_gctl, 0, GC_COUNT * sizeof(int), c->stream));
        CU(dmalloc(c, &c->grp_blk, 8 * 2048 * sizeof(int)));
    }
    // fixed-point scale of the log-likelihood accumulator: |sum n log p| <= reads * 88 must stay below 2^62
    const double reads = (c->total_reads > 1.0 ? c->total_reads : 1.0) * ((double)c->V_total / (double)c->V);
    int k = (int)floor(log2(4.6e18 / (reads * 88.0)));
    if (k > 40) k = 40;
    if (k < 0) k = 0;
    c->ll_scale = ldexp(1.0, k);
    return DESMAN_OK;
}

static AggTable agg_table(desman_ctx *c)
{
    AggTable t;
    t.keys = c->agg_keys; t.ids = c->agg_ids; t.hmask = (unsigned int)(c->agg_H - 1); t.slot_code = c->agg_code;
    t.nslots = c->agg_nslots; t.N = c->agg_N; t.cap_slots = (unsigned int)c->agg_cap_slots; t.S = c->S; t.ctl = c->agg_ctl;
    return t;
}

static MuAggParams agg_params(desman_ctx *c, const double *gamma, const double *eta)
{
    MuAggParams p;
    p.counts = c->counts; p.tau = c->tau; p.gamma = gamma; p.eta = eta;
    p.seed = c->seed; p.sweep = c->sweep; p.shard = (uint32_t)c->v0;
    p.V = (int)c->V; p.S = c->S; p.G = c->G;
    p.t = agg_table(c);
    p.sum_mu = c->stats; p.esum = c->stats + (size_t)c->S * c->G;
    p.ll_scale = c->ll_scale; p.ll_fx = c->red_i;
    p.upkeep = 0; p.gctl_u = nullptr; p.V_local_u = (long long)c->V; p.agg_limit_u = (unsigned int)(c->V + c->V / 4);
    p.eta_commit = nullptr;
    p.classM = (c->G <= MUC_MAX_G && c->classM_G == c->G && c->classM_S == c->S) ? c->agg_classM : nullptr;
    return p;
}

// Is the screening pass of the tau update (tau_group_kernel.cuh) worth keeping groups for?  Same rule as the statistics:
// sites share patterns when the ~12*2^G biallelic patterns are few compared with V.  *gb / *warps: strain block and warps
// per CTA of the kernel (each warp owns one table in shared memory).
static bool group_use_mma(const desman_ctx *c)
{
    return c->tau_group_mma && c->counts_tf32_exact && tgm_tiles(c->G) <= 3 &&
           tg_shared_bytes(c->S, c->G) + tgm_table_bytes(c->S, c->G) + 1024 <= 200 * 1024;
}

// Shape of the tensor-memory screening pass: K blocks of <= 64 samples (balanced, multiples of 4), table columns padded to 8
// (the widest K block of 64, 32 or 16 samples whose two count stages and two table buffers fit in shared memory)
static bool tc_shape(const desman_ctx *c, int *SK, int *nkb, int *NC)
{
    *NC = (3 * c->G + 7) & ~7;
    for (int kmax = 64; kmax >= 16; kmax >>= 1) {
        const int nb = (c->S + kmax - 1) / kmax;
        *nkb = nb;
        *SK = (((c->S + nb - 1) / nb) + 3) & ~3;
        if (tc_layout(c->S, c->G, *SK, *nkb, *NC).total + 2048 <= 227 * 1024) return true;
    }
    return false;
}
static bool group_use_tc(const desman_ctx *c)
{
    if (!c->tau_group_tc || !c->counts_tf32_exact || 3 * c->G > 64) return false;       // counts < 2048: exact in FP16
    int SK, nkb, NC;
    return tc_shape(c, &SK, &nkb, &NC);
}

static bool group_config(const desman_ctx *c, int *gb, int *warps)
{
    if (c->tau_group == 0 || c->tau_exact) return false;
    // auto: keep groups whenever the shape allows it; whether they are USED is decided on the device at every regroup from
    // the realised groups (GC_WORTH, maintain_kernel.cuh) -- at G >= 16 far fewer than the 12*2^G possible patterns occur
    if (c->tau_group == 2 && !(c->G <= 24 && c->V >= 1024)) return false;
    if (group_use_tc(c)) return true;
    if (group_use_mma(c)) return true;
    const int r = c->G % 8, GB = (r >= 1 && r <= 4) ? 4 : 8;
    const size_t table = tg_table_bytes(c->S, c->G, GB), shared = tg_shared_bytes(c->S, c->G) + 1024;
    int w = TG_MAX_WARPS;
    if (shared + (size_t)w * table > 110 * 1024) {                 // one CTA per SM: as many warps as fit
        w = (int)((220 * 1024 - shared) / table);
        if (w > TG_MAX_WARPS) w = TG_MAX_WARPS;
    }
    if (w < 1) return false;
    if (gb) *gb = GB;
    if (warps) *warps = w;
    return true;
}

static TauGroup group_ptrs(desman_ctx *c)
{
    TauGroup g;
    g.site_slot = c->grp_site_slot; g.order = c->grp_order; g.singles = c->grp_singles; g.items = c->grp_items; g.work = c->grp_work;
    g.slot_cnt = c->grp_slot4; g.slot_fill = c->grp_slot4 + c->grp_cap_slots; g.slot_start = c->grp_slot4 + 2 * c->grp_cap_slots;
    g.slot_item = c->grp_slot4 + 3 * c->grp_cap_slots;
    g.gctl = c->grp_gctl;
    return g;
}

// fp16 count image + row tables of the tensor-memory screening pass: capacity 2 V + 1024 rows (every multi-site pattern is
// padded to a multiple of 8 rows; a regroup that needs more switches the pass off for that grouping, GC_IMG_OK)
static int ensure_img(desman_ctx *c)
{
    int SK, nkb, NC;
    tc_shape(c, &SK, &nkb, &NC);
    const size_t rows = (((size_t)2 * c->V + 1024) + 7) & ~(size_t)7;
    const size_t bytes = rows * (size_t)SK * nkb * 8;
    if (rows > c->img_cap_rows || bytes > c->img_bytes || (size_t)c->V > c->img_cap_v) {
        for (void *q : {(void *)c->img, (void *)c->img_site, (void *)c->img_nsite, (void *)c->site_row}) if (q) dfree(c, q);
        c->img = nullptr; c->img_site = c->site_row = nullptr; c->img_nsite = nullptr; c->img_cap_rows = c->img_bytes = c->img_cap_v = 0;
        CU(dmalloc(c, &c->img, bytes));
        CU(dmalloc(c, &c->img_site, rows * sizeof(int)));
        CU(dmalloc(c, &c->img_nsite, rows * sizeof(float)));
        CU(dmalloc(c, &c->site_row, (size_t)c->V * sizeof(int)));
        CU(cudaMemsetAsync(c->img, 0, bytes, c->stream));
        CU(cudaMemsetAsync(c->img_site, 0, rows * sizeof(int), c->stream));
        CU(cudaMemsetAsync(c->img_nsite, 0, rows * sizeof(float), c->stream));
        c->img_cap_rows = rows; c->img_bytes = bytes; c->img_cap_v = (size_t)c->V;
        c->agg_valid = false;     // filled by the next regroup
    }
    return DESMAN_OK;
}

static int ensure_countsf(desman_ctx *c)
{
    const size_t ncell = (size_t)c->V * c->S;
    if (ncell > c->countsf_cap) {
        if (c->countsf) dfree(c, c->countsf);
        if (c->nsite) dfree(c, c->nsite);
        c->countsf = nullptr; c->nsite = nullptr; c->countsf_cap = 0;
        CU(dmalloc(c, &c->countsf, ncell * sizeof(float4)));
        CU(dmalloc(c, &c->nsite, (size_t)c->V * sizeof(float)));
        c->countsf_cap = ncell;
        c->agg_valid = false;     // filled by the next regroup
    }
    return DESMAN_OK;
}

// Start of a sweep: ONE cooperative launch that clears the per-sweep accumulators (statistics, fixed-point ll, nchange,
// work-list length) and, only when a request is pending (host: state upload; device: finalize_sweep_kernel), rebuilds the
// pattern table and regroups the sites (maintain_kernel.cuh).
static int sync_table(desman_ctx *c, bool deferred_star_copy = false)
{
    RET(ensure_agg(c));
    const bool grouping = group_config(c, nullptr, nullptr);
    const bool use_tc = grouping && group_use_tc(c);
    if (grouping) RET(ensure_countsf(c));
    if (use_tc) RET(ensure_img(c));
    if (!c->agg_valid) {
        const int zero = 0;   // new counts / state: the groups (and the row copy that follows them) are rebuilt with the table
        CU(cudaMemcpyAsync(c->grp_gctl + GC_HAVE, &zero, sizeof(int), cudaMemcpyHostToDevice, c->stream));
        const int one = 1;
        CU(cudaMemcpyAsync(c->agg_ctl, &one, sizeof(int), cudaMemcpyHostToDevice, c->stream));
        // optimistic: a freshly uploaded state is screened on its first sweep (a converged state pays off at once; after a
        // random one finalize_sweep_kernel turns the screening off until the chain has calmed down)
        CU(cudaMemcpyAsync(c->grp_gctl + GC_CALM, &one, sizeof(int), cudaMemcpyHostToDevice, c->stream));
        c->agg_valid = true;
    }
    if (!c->maint_grid) {
        int occ = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, table_maintain_kernel, MAINT_THREADS, 0) != cudaSuccess || occ < 1) occ = 1;
        if (occ > 4) occ = 4;
        c->maint_grid = c->sm_count * occ;
        if (c->maint_grid > 2048) c->maint_grid = 2048;
    }
    MaintParams p;
    p.a = agg_params(c, c->gamma, c->eta);
    p.grp = group_ptrs(c);
    if (!grouping) p.grp.gctl = nullptr;
    p.blk = c->grp_blk;
    p.zero64 = c->stats; p.nzero64 = (int)((size_t)c->S * c->G + 16);
    p.red_i = c->red_i;
    p.countsf = c->countsf; p.nsite = c->nsite;
    p.star_flag = deferred_star_copy ? c->flag : nullptr; p.tau_star = c->tau_star;
    p.item_sites = use_tc ? TC_ROWS : TG_ITEM_SITES;
    p.force_worth = c->tau_group == 1;
    p.img = nullptr; p.img_site = nullptr; p.img_nsite = nullptr; p.site_row = nullptr; p.slot_img = nullptr;
    p.img_cap_rows = 0; p.SK = 4; p.nkb = 1;
    if (use_tc) {
        int NC;
        tc_shape(c, &p.SK, &p.nkb, &NC);
        p.img = c->img; p.img_site = c->img_site; p.img_nsite = c->img_nsite; p.site_row = c->site_row;
        p.slot_img = c->grp_slot4 + 4 * c->grp_cap_slots; p.img_cap_rows = (long long)c->img_cap_rows;
    }
    void *args[] = {&p};
    {
        KSpan k(c, DESMAN_K_MAINT);
        CU(cudaLaunchCooperativeKernel((const void *)table_maintain_kernel, dim3(c->maint_grid), dim3(MAINT_THREADS), args, 0, c->stream));
    }
    return DESMAN_OK;
}

// sum n*log p of the current device tau under (gamma, eta) into red_i[0] (fixed point, cleared by sync_table)
static int launch_ll(desman_ctx *c, const double *gamma, const double *eta, double *eta_commit = nullptr, bool upkeep = false)
{
    MuAggParams p = agg_params(c, gamma, eta);
    p.eta_commit = eta_commit;
    p.upkeep = upkeep ? 1 : 0;
    p.gctl_u = group_config(c, nullptr, nullptr) ? c->grp_gctl : nullptr;
    {
        KSpan k(c, DESMAN_K_FINAL);
        CU(launch_k(c, ll_table_kernel, c->sm_count * 4, 256, 0, p));       // ~one (slot, 32 samples) item per warp: a latency chain
    }
    CU(cudaGetLastError());
    return DESMAN_OK;
}

// Draw the V*G uniform words of one c_sample_tau call from the MT19937 stream (rank slice under sharding).
static int gen_mt_words(desman_ctx *c)
{
    const size_t total = (size_t)c->V_total * c->G, lo = (size_t)c->v0 * c->G, n = (size_t)c->V * c->G;
    if (n > c->words_cap) {
        if (c->words) dfree(c, c->words);
        c->words = nullptr; c->words_cap = 0;
        CU(dmalloc(c, &c->words, n * sizeof(uint32_t)));
        c->words_cap = n;
    }
    {
        KSpan k(c, DESMAN_K_MT);
        mt19937_kernel<<<1, 256, 0, c->stream>>>(c->mt_state, c->mt_pos, total, lo, lo + n, c->words);
    }
    CU(cudaGetLastError());
    c->mt_consumed += total;
    c->mt_pos = (int)((c->mt_consumed - 1) % 624) + 1;
    return DESMAN_OK;
}

template <int NT>
static int launch_tau_group_mma_t(desman_ctx *c, const TauGroupParams &p)
{
    const size_t smem = tg_shared_bytes(c->S, c->G) + tgm_table_bytes(c->S, c->G);
    CU(cudaFuncSetAttribute(tau_group_mma_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, tau_group_mma_kernel<NT>, TGM_WARPS * 32, smem) != cudaSuccess || occ < 1) occ = 1;
    CU(launch_k(c, tau_group_mma_kernel<NT>, c->sm_count * occ, TGM_WARPS * 32, smem, p));
    CU(cudaGetLastError());
    return DESMAN_OK;
}

static int launch_tau_group_tc(desman_ctx *c, const TauGroupParams &q, float *dbg, int early = 0)
{
    TauGroupTcParams p;
    tc_shape(c, &p.SK, &p.nkb, &p.NC);
    p.img = c->img; p.img_site = c->img_site; p.img_nsite = c->img_nsite; p.img_rg = (long long)(c->img_cap_rows / 8);
    p.gamma = q.gamma; p.eta = q.eta; p.words = q.words; p.V = q.V; p.S = q.S; p.G = q.G;
    p.grp = q.grp; p.tier_counts = q.tier_counts; p.dbg = dbg;
    p.early = (early && c->pdl) ? 1 : 0;
    p.star_src = nullptr; p.star_dst = nullptr; p.star_n = 0; p.star_flag = nullptr;
    if (c->star_fold) {      // the conditional MAP snapshot of the sharded chain (the bookkeeping launch before the draw decided it)
        p.star_src = c->tau; p.star_dst = c->tau_star; p.star_n = (size_t)c->V * c->G; p.star_flag = c->flag;
        c->star_fold = false;
    }
    const size_t smem = tc_layout(c->S, c->G, p.SK, p.nkb, p.NC).total;
    CU(cudaFuncSetAttribute(tau_group_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CU(launch_k(c, tau_group_tc_kernel, c->sm_count, TC_THREADS, smem, p));
    CU(cudaGetLastError());
    return DESMAN_OK;
}

template <int GB>
static int launch_tau_group_t(desman_ctx *c, const TauGroupParams &p, int warps)
{
    const size_t smem = tg_shared_bytes(c->S, c->G) + (size_t)warps * tg_table_bytes(c->S, c->G, GB);
    CU(cudaFuncSetAttribute(tau_group_kernel<GB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, tau_group_kernel<GB>, warps * 32, smem) != cudaSuccess || occ < 1) occ = 1;
    CU(launch_k(c, tau_group_kernel<GB>, c->sm_count * occ, warps * 32, smem, p));
    CU(cudaGetLastError());
    return DESMAN_OK;
}

// One tau pass (gamma, eta: device pointers).  maintain: keep the pattern table current (it must be in sync); then the
// screening pass runs first where groups are kept, and the per-site kernel walks its work list only.
// red_i[1] (nchange) must be zero on entry (sync_table, or the caller's memset).
// will launch_tau (maintain = true) run the tensor-memory screening launch?  (the same conditions, in the same order)
static bool tau_runs_tc(desman_ctx *c)
{
    if (!c->agg_valid || !agg_table(c).N) return false;
    int gb = 8, gw = 1;
    return group_config(c, &gb, &gw) && group_use_tc(c);
}

static int launch_tau(desman_ctx *c, const double *gamma, const double *eta, bool maintain, bool count_occupancy, uint32_t iter,
                      int g_begin = 0, double *logp_out = nullptr, int early = 0)
{
    KSpan whole(c, DESMAN_K_TAU_UPDATE, 0);
    TauParams p;
    p.counts = c->counts; p.tau = c->tau; p.gamma = gamma; p.eta = eta;
    p.words = nullptr;
    if (c->rng_mode == DESMAN_RNG_MT19937) { RET(gen_mt_words(c)); p.words = c->words; }
    p.seed = c->seed; p.sweep = c->sweep; p.v0 = c->v0;
    p.V = (int)c->V; p.S = c->S; p.G = c->G;
    p.nchange = c->red_i + 1;
    const int grid = tau_grid(c);
    memset(&p.agg, 0, sizeof(p.agg));
    if (maintain && c->agg_valid) p.agg = agg_table(c);
    else c->agg_valid = false;
    p.tau_cnt = count_occupancy ? c->tau_cnt : nullptr;
    p.tau_last = c->tau_last;
    p.iter = iter;
    p.exact_only = c->tau_exact;
    p.tier_counts = c->tiers;
    p.work = nullptr; p.singles = nullptr; p.gctl = nullptr; p.site_slot = nullptr;
    p.img_site = nullptr; p.site_row = nullptr; p.need_img = 0;
    p.g_begin = g_begin; p.logp_out = logp_out; p.skip_listed = 0; p.prob_off = nullptr;
    int gb = 8, gwarps = 1;
    if (p.agg.N && group_config(c, &gb, &gwarps)) {
        TauGroupParams q;
        q.countsf = c->countsf; q.nsite = c->nsite; q.gamma = gamma; q.eta = eta; q.words = p.words;
        q.V = (int)c->V; q.S = c->S; q.G = c->G;
        q.grp = group_ptrs(c);
        q.tier_counts = c->tiers;
        {
            KSpan k(c, DESMAN_K_TAU_GROUP);
            if (group_use_tc(c)) {
                RET(launch_tau_group_tc(c, q, nullptr, early));
                p.img_site = c->img_site; p.site_row = c->site_row; p.need_img = 1;
            } else if (group_use_mma(c)) {
                switch (tgm_tiles(c->G)) {
                case 1: RET(launch_tau_group_mma_t<1>(c, q)); break;
                case 2: RET(launch_tau_group_mma_t<2>(c, q)); break;
                default: RET(launch_tau_group_mma_t<3>(c, q)); break;
                }
            } else if (gb == 4) RET(launch_tau_group_t<4>(c, q, gwarps));
            else RET(launch_tau_group_t<8>(c, q, gwarps));
        }
        p.work = c->grp_work; p.singles = c->grp_singles; p.gctl = c->grp_gctl; p.site_slot = c->grp_site_slot;
    }
    const size_t smem = tau_smem_bytes(c->S, c->G);
    if (smem > 227 * 1024) return fail(DESMAN_EINVAL, "S*G too large for the shared-memory tile (%zu bytes)", smem);
    CU(cudaFuncSetAttribute(tau_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const size_t smem_o = tauo_smem_bytes(c->S, c->G);
    const bool open_kernel = p.work != nullptr && c->tau_open && smem_o <= 200 * 1024;
    {
        KSpan k(c, DESMAN_K_TAU, open_kernel ? 2 : 1);
        if (open_kernel) {
            // the work list of the screening pass: one CTA per listed site, its open steps in parallel (tau_open_kernel.cuh);
            // tau_sample_kernel then runs only if the list is not valid (burn-in: no groups, or the chain is not calm)
            if (!c->tauo_grid) {
                CU(cudaFuncSetAttribute(tau_open_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_o));
                int occ = 0;
                if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, tau_open_kernel, TAUO_WARPS * 32, smem_o) != cudaSuccess || occ < 1) occ = 1;
                c->tauo_grid = c->sm_count * occ;
                c->tauo_smem = smem_o;
            }
            if (c->tauo_smem != smem_o) { CU(cudaFuncSetAttribute(tau_open_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_o)); c->tauo_smem = smem_o; }
            CU(launch_k(c, tau_open_kernel, c->tauo_grid, TAUO_WARPS * 32, smem_o, p));
            p.skip_listed = 1;
        }
        CU(launch_k(c, tau_sample_kernel, grid, TAU_WARPS * 32, smem, p));
    }
    CU(cudaGetLastError());
    return DESMAN_OK;
}

template <int GP>
static void launch_mu_t(desman_ctx *c, const MuParams &p, int grid)
{
    launch_k(c, mu_stats_kernel<GP>, grid, MU_WARPS * 32, 0, p);
}

// K2b: one conditional-binomial chain per (pattern, sample, base) of the (synchronised) pattern table.
// The statistics accumulators are cleared by sync_table at the start of the sweep.
static int launch_mu_agg(desman_ctx *c, const double *gamma, const double *eta)
{
    MuAggParams p = agg_params(c, gamma, eta);
    const int nch = (c->S + 31) / 32;
    const size_t smem = mub_smem_bytes(c->G);
    CU(cudaFuncSetAttribute(mu_binomial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, mu_binomial_kernel, MUB_WARPS * 32, smem) != cudaSuccess || occ < 1) occ = 1;
    int grid = c->sm_count * occ;
    while ((grid * MUB_WARPS) % nch) grid++;
    {
        KSpan k(c, DESMAN_K_MU, p.classM ? 2 : 1);
        CU(launch_k(c, mu_binomial_kernel, grid, MUB_WARPS * 32, smem, p));
        if (p.classM) {
            const size_t smem2 = muc_smem_bytes(c->G);
            const int nch8 = (c->S + 7) / 8;
            long long want = (((long long)1 << c->G) * nch8 + MUB_WARPS - 1) / MUB_WARPS;
            int grid2 = (int)(want < (long long)c->sm_count * 4 ? want : (long long)c->sm_count * 4);
            if (grid2 < 1) grid2 = 1;
            CU(launch_k(c, mu_class_kernel, grid2, MUB_WARPS * 32, smem2, p));
        }
    }
    CU(cudaGetLastError());
    return DESMAN_OK;
}

// Which statistics kernel: the aggregated form pays off when sites share patterns.  A biallelic site has one of
// ~12*2^G patterns; the rule depends on (V, G) only, so it is reproducible outside the engine (oracle, tests).
static int resolved_mu_mode(const desman_ctx *c)
{
    if (c->mu_mode != 2) return c->mu_mode;
    return (c->G <= 24 && 12.0 * ldexp(1.0, c->G) <= (double)c->V / 2.0) ? 1 : 0;
}

static int launch_mu(desman_ctx *c, const double *gamma, const double *eta)
{
    if (resolved_mu_mode(c) == 1) return launch_mu_agg(c, gamma, eta);
    MuParams p;
    p.counts = c->counts; p.tau = c->tau; p.gamma = gamma; p.eta = eta;
    p.seed = c->seed; p.sweep = c->sweep; p.v0 = c->v0;
    p.V = (int)c->V; p.S = c->S; p.G = c->G;
    p.sum_mu = c->stats; p.esum = c->stats + (size_t)c->S * c->G;
    const int nch = (c->S + 31) / 32;
    // total warps must be a multiple of the number of 32-sample chunks
    int64_t items = c->V * nch;
    int64_t blocks = (items + MU_WARPS - 1) / MU_WARPS;
    int64_t cap = (int64_t)c->sm_count * 6;
    if (blocks > cap) blocks = cap;
    blocks = ((blocks + nch - 1) / nch) * nch;
    const int grid = (int)blocks;
    {
        KSpan k(c, DESMAN_K_MU);
        const int G = c->G;
        if (G <= 2) launch_mu_t<2>(c, p, grid);
        else if (G <= 4) launch_mu_t<4>(c, p, grid);
        else if (G <= 8) launch_mu_t<8>(c, p, grid);
        else if (G <= 12) launch_mu_t<12>(c, p, grid);
        else if (G <= 16) launch_mu_t<16>(c, p, grid);
        else if (G <= 20) launch_mu_t<20>(c, p, grid);
        else if (G <= 24) launch_mu_t<24>(c, p, grid);
        else launch_mu_t<32>(c, p, grid);
    }
    CU(cudaGetLastError());
    return DESMAN_OK;
}

static int exchange_sum(desman_ctx *c, unsigned long long *data, int words, unsigned long long *data2 = nullptr, int words2 = 0)
{
    XchParams p;
    for (int r = 0; r < XCH_MAX_RANKS; r++) p.mail[r] = c->xch_mail[r];
    p.rank = c->rank; p.nranks = c->nranks; p.words = words; p.cap_words = c->xch_cap_words;
    p.data2 = data2; p.words2 = words2;
    p.seq = ++c->xch_seq; p.data = data; p.err = c->xch_err;
    CU(launch_k(c, exchange_sum_kernel, 1, 512, 0, p));
    CU(cudaGetLastError());
    return DESMAN_OK;
}

// the statistics of this sweep and, with them, the [ll, nchange] words `red2` of the previous one (or null): ONE
// synchronisation point per sweep under sharding (every exchange costs the skew between the ranks)
static int allreduce_stats(desman_ctx *c, unsigned long long *red2 = nullptr)
{
    if (c->nranks <= 1) return DESMAN_OK;
    KSpan k(c, DESMAN_K_OTHER);
    const size_t n = (size_t)c->S * c->G + 16;
    if (c->xch_ok && (int)n + 3 <= c->xch_cap_words) return exchange_sum(c, c->stats, (int)n, red2, red2 ? 3 : 0);
    if (red2 && g_nccl.GroupStart && g_nccl.GroupEnd) {
        NC(g_nccl.GroupStart());
        NC(g_nccl.AllReduce(c->stats, c->stats, n, NCCL_UINT64, NCCL_SUM, c->comm, c->stream));
        NC(g_nccl.AllReduce(red2, red2, 3, NCCL_INT64, NCCL_SUM, c->comm, c->stream));
        NC(g_nccl.GroupEnd());
        return DESMAN_OK;
    }
    NC(g_nccl.AllReduce(c->stats, c->stats, n, NCCL_UINT64, NCCL_SUM, c->comm, c->stream));
    if (red2) NC(g_nccl.AllReduce(red2, red2, 3, NCCL_INT64, NCCL_SUM, c->comm, c->stream));
    return DESMAN_OK;
}
static int allreduce_red(desman_ctx *c)
{
    if (c->nranks <= 1) return DESMAN_OK;
    KSpan k(c, DESMAN_K_OTHER);
    if (c->xch_ok) return exchange_sum(c, c->red_i, 3);      // two's complement sums: same bits as the int64 all-reduce
    NC(g_nccl.AllReduce(c->red_i, c->red_i, 3, NCCL_INT64, NCCL_SUM, c->comm, c->stream));
    return DESMAN_OK;
}

// The peer-memory exchange does not hang on a dead peer: after its spin limit it raises xch_err and carries on with what it
// has.  Every API that exchanged checks the flag once its results are on the host (stream already synchronised).
static int check_xch(desman_ctx *c)
{
    if (!c->xch_ok) return DESMAN_OK;
    int xe = 0;
    CU(cudaMemcpyAsync(&xe, c->xch_err, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (xe) return fail(DESMAN_ECOMM, "peer-memory exchange timed out waiting for another rank");
    return DESMAN_OK;
}

static int launch_draw(desman_ctx *c, const unsigned long long *stats, double *gamma_out, double *eta_out,
                       unsigned long long *esum_keep = nullptr)
{
    DrawParams p;
    p.esum_keep = esum_keep;
    p.sum_mu = stats; p.esum = stats + (size_t)c->S * c->G;
    p.S = c->S; p.G = c->G; p.alpha = c->alpha; p.delta = c->delta; p.epsilon = c->epsilon;
    p.seed = c->seed; p.sweep = c->sweep; p.gamma_out = gamma_out; p.eta_out = eta_out;
    const size_t smem = ((size_t)c->S * c->G + 16) * sizeof(double);
    if (smem > 200 * 1024) return fail(DESMAN_EINVAL, "S*G too large for draw kernel");
    CU(cudaFuncSetAttribute(draw_gamma_eta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    {
        KSpan k(c, DESMAN_K_DRAW);
        // one variate per thread where possible: the kernel is one latency chain (rejection sampler in FP64) per variate
        int threads = ((c->S * c->G + 16 + 31) / 32) * 32;
        if (threads > DRAW_MAX_THREADS) threads = DRAW_MAX_THREADS;
        if (threads < 64) threads = 64;
        CU(launch_k(c, draw_gamma_eta_kernel, 1, threads, smem, p));
    }
    CU(cudaGetLastError());
    return DESMAN_OK;
}

struct StoreBufs { double *ll = nullptr, *lp = nullptr, *nch = nullptr, *gs = nullptr, *es = nullptr; };

// ll (from the table) -> lp, stores, MAP bookkeeping.  gamma/eta: the state the likelihood is evaluated at.
static int launch_finalize_only(desman_ctx *c, const unsigned long long *red, const double *gamma, const double *eta, int it,
                                int star_mode, const StoreBufs &sb, bool store_ge, double *eta_commit = nullptr, bool copy_now = true);
// eta_commit: where the chain's eta lives when `eta` is the freshly drawn eta_new (committed by the finalize kernel: :347)
// copy_now = false: the MAP snapshot tau_star <- tau (if lp improved) is left to the next sync_table(c, true), or to the caller
static int launch_finalize(desman_ctx *c, const double *gamma, const double *eta, int it, int star_mode, const StoreBufs &sb,
                           bool store_ge, double *eta_commit = nullptr, bool copy_now = true)
{
    RET(launch_ll(c, gamma, eta, nullptr, it >= 0));
    RET(allreduce_red(c));
    return launch_finalize_only(c, c->red_i, gamma, eta, it, star_mode, sb, store_ge, eta_commit, copy_now);
}
static int launch_star_copy(desman_ctx *c)
{
    KSpan k(c, DESMAN_K_FINAL);
    CU(launch_k(c, copy_tau_if_kernel, c->sm_count, 256, 0, (const uint8_t *)c->tau, c->tau_star, (size_t)c->V * c->G, (const int *)c->flag));
    return DESMAN_OK;
}
static FinalParams final_params(desman_ctx *c, const unsigned long long *red, const double *gamma, const double *eta, int it,
                                int star_mode, const StoreBufs &sb, bool store_ge, double *eta_commit)
{
    FinalParams p;
    p.red_i = (const long long *)red; p.ll_const = c->ll_const_total; p.ll_inv_scale = 1.0 / c->ll_scale;
    p.gamma = gamma; p.eta = eta; p.eta_commit = eta_commit;
    p.S = c->S; p.G = c->G; p.V_total = (double)c->V_total; p.alpha = c->alpha; p.delta = c->delta;
    p.lg_alphaG = lgamma(c->alpha * c->G); p.lg_alpha = lgamma(c->alpha);
    p.lg_delta4 = lgamma(4.0 * c->delta); p.lg_delta = lgamma(c->delta);
    p.it = it; p.star_mode = star_mode;
    p.ll_store = sb.ll; p.lp_store = sb.lp; p.nchange_store = sb.nch;
    p.gamma_store = store_ge ? sb.gs : nullptr; p.eta_store = store_ge ? sb.es : nullptr;
    p.gamma_star = c->gamma_star; p.eta_star = c->eta_star; p.scal = c->scal; p.flag = c->flag;
    p.agg_nslots = c->agg_nslots; p.agg_ctl = c->agg_ctl; p.agg_limit = (unsigned int)(c->V + c->V / 4);
    p.gctl = group_config(c, nullptr, nullptr) ? c->grp_gctl : nullptr;
    p.V_local = (long long)c->V;
    return p;
}
// lp, stores, MAP bookkeeping from the (already summed) words red = [fixed-point ll, nchange, upkeep wishes]
static int launch_finalize_only(desman_ctx *c, const unsigned long long *red, const double *gamma, const double *eta, int it,
                                int star_mode, const StoreBufs &sb, bool store_ge, double *eta_commit, bool copy_now)
{
    const FinalParams p = final_params(c, red, gamma, eta, it, star_mode, sb, store_ge, eta_commit);
    {
        KSpan k(c, DESMAN_K_FINAL);
        CU(launch_k(c, finalize_sweep_kernel, 1, 256, 0, p));
    }
    if (copy_now) RET(launch_star_copy(c));
    CU(cudaGetLastError());
    return DESMAN_OK;
}
// The sharded chain's exchange of sweep k (statistics + the three words of sweep k-1) and the bookkeeping of sweep k-1 that
// consumes those words: one launch over peer memory (exchange_finalize_kernel), or the all-reduce followed by the finalize launch.
static int exchange_and_finalize(desman_ctx *c, unsigned long long *red_prev, const double *gamma, const double *eta, int it,
                                 const StoreBufs &sb, bool copy_now)
{
    const size_t n = (size_t)c->S * c->G + 16;
    if (c->xch_ok && (int)n + 3 <= c->xch_cap_words && c->xch_fuse) {
        XchParams x;
        for (int r = 0; r < XCH_MAX_RANKS; r++) x.mail[r] = c->xch_mail[r];
        x.rank = c->rank; x.nranks = c->nranks; x.words = (int)n; x.cap_words = c->xch_cap_words;
        x.data2 = red_prev; x.words2 = 3;
        x.seq = ++c->xch_seq; x.data = c->stats; x.err = c->xch_err;
        const FinalParams f = final_params(c, red_prev, gamma, eta, it, 0, sb, true, nullptr);
        {
            KSpan k(c, DESMAN_K_OTHER);
            CU(launch_k(c, exchange_finalize_kernel, 1, 512, 0, x, f));
        }
        if (copy_now) RET(launch_star_copy(c));
        CU(cudaGetLastError());
        return DESMAN_OK;
    }
    RET(allreduce_stats(c, red_prev));
    return launch_finalize_only(c, red_prev, gamma, eta, it, 0, sb, true, nullptr, copy_now);
}

static int require_state(desman_ctx *c)
{
    if (!c || c->V <= 0) return fail(DESMAN_ESTATE, "no counts set");
    if (c->G <= 0) return fail(DESMAN_ESTATE, "no state set");
    CU(cudaSetDevice(c->device));
    return DESMAN_OK;
}

// ------------------------------------------------------------------------------------------ single steps
extern "C" int desman_sample_tau(desman_ctx *c, int64_t *nchange)
{
    RET(require_state(c));
    CU(cudaMemsetAsync(c->red_i + 1, 0, sizeof(unsigned long long), c->stream));
    RET(launch_tau(c, c->gamma, c->eta, false, false, 0));
    if (c->rng_mode == DESMAN_RNG_PHILOX) c->sweep++;
    unsigned long long n = 0;
    CU(cudaMemcpyAsync(&n, c->red_i + 1, sizeof(n), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (c->nranks > 1) return fail(DESMAN_ESTATE, "desman_sample_tau is a single-rank call");
    if (nchange) *nchange = (int64_t)n;
    return DESMAN_OK;
}

// sampleTauFixTau (HaploSNP_Sampler.py:196-222) on the current device state: strains [H, G) are redrawn in order (Philox
// contract of the tau draws), logp [V][4] = normalised log-probabilities of strain H's bases before its draw
extern "C" int desman_sample_tau_fix(desman_ctx *c, int H, double *logp, int64_t *nchange)
{
    RET(require_state(c));
    if (H < 0 || H >= c->G) return fail(DESMAN_EINVAL, "desman_sample_tau_fix: H must be in [0, G)");
    if (c->rng_mode != DESMAN_RNG_PHILOX) return fail(DESMAN_ESTATE, "desman_sample_tau_fix draws under the Philox contract");
    if (c->nranks > 1) return fail(DESMAN_ESTATE, "desman_sample_tau_fix is a single-rank call");
    double *dl = nullptr;
    if (logp) CU(dmalloc(c, &dl, (size_t)c->V * 4 * sizeof(double)));
    CU(cudaMemsetAsync(c->red_i + 1, 0, sizeof(unsigned long long), c->stream));
    int rc = launch_tau(c, c->gamma, c->eta, false, false, 0, H, dl);
    c->sweep++;
    unsigned long long n = 0;
    if (rc == DESMAN_OK) {
        CU(cudaMemcpyAsync(&n, c->red_i + 1, sizeof(n), cudaMemcpyDeviceToHost, c->stream));
        if (logp) CU(cudaMemcpyAsync(logp, dl, (size_t)c->V * 4 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
    }
    if (dl) dfree(c, dl);
    if (nchange) *nchange = (int64_t)n;
    return rc;
}

extern "C" int desman_mu_stats(desman_ctx *c, int64_t *sum_mu, int64_t *esum)
{
    RET(require_state(c));
    RET(sync_table(c));
    RET(launch_mu(c, c->gamma, c->eta));
    RET(allreduce_stats(c));
    const size_t nsg = (size_t)c->S * c->G;
    std::vector<unsigned long long> h(nsg + 16);
    CU(cudaMemcpyAsync(h.data(), c->stats, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (sum_mu) for (size_t i = 0; i < nsg; i++) sum_mu[i] = (int64_t)h[i];
    if (esum) for (int i = 0; i < 16; i++) esum[i] = (int64_t)h[nsg + i];
    return check_xch(c);
}

extern "C" int desman_draw_gamma_eta(desman_ctx *c, const int64_t *sum_mu, const int64_t *esum, double *gamma, double *eta)
{
    RET(require_state(c));
    const size_t nsg = (size_t)c->S * c->G;
    RET(ensure_scratch(c, (nsg + 16) * 16 + 256));
    unsigned long long *st = (unsigned long long *)c->scratch;
    double *out = (double *)(st + nsg + 16);
    std::vector<unsigned long long> h(nsg + 16);
    for (size_t i = 0; i < nsg; i++) h[i] = (unsigned long long)sum_mu[i];
    for (int i = 0; i < 16; i++) h[nsg + i] = (unsigned long long)esum[i];
    CU(cudaMemcpyAsync(st, h.data(), h.size() * 8, cudaMemcpyHostToDevice, c->stream));
    RET(launch_draw(c, st, out, out + nsg));
    if (gamma) CU(cudaMemcpyAsync(gamma, out, nsg * 8, cudaMemcpyDeviceToHost, c->stream));
    if (eta) CU(cudaMemcpyAsync(eta, out + nsg, 16 * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return DESMAN_OK;
}

extern "C" int desman_loglik(desman_ctx *c, double *ll, double *lp)
{
    RET(require_state(c));
    RET(ensure_ll_const(c));
    RET(sync_table(c));
    RET(launch_ll(c, c->gamma, c->eta));
    RET(allreduce_red(c));
    long long fx = 0;
    CU(cudaMemcpyAsync(&fx, c->red_i, sizeof(fx), cudaMemcpyDeviceToHost, c->stream));
    std::vector<double> g((size_t)c->S * c->G), e(16);
    CU(cudaMemcpyAsync(g.data(), c->gamma, g.size() * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(e.data(), c->eta, 16 * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    RET(check_xch(c));
    const double h_ll = c->ll_const_total + (double)fx / c->ll_scale;
    // prior on host (tiny): Desman_Utils.py:35-44, HaploSNP_Sampler.py:448-459
    double prior = 0.0;
    for (int s = 0; s < c->S; s++) {
        double r = lgamma(c->alpha * c->G);
        for (int k = 0; k < c->G; k++) { r += (c->alpha - 1.0) * log(g[(size_t)s * c->G + k]); r -= lgamma(c->alpha); }
        prior += r;
    }
    for (int a = 0; a < 4; a++) {
        double r = lgamma(4.0 * c->delta);
        for (int k = 0; k < 4; k++) { r += (c->delta - 1.0) * log(e[a * 4 + k]); r -= lgamma(c->delta); }
        prior += r;
    }
    prior += (double)c->V_total * c->G * log(0.25);
    if (ll) *ll = h_ll;
    if (lp) *lp = h_ll + prior;
    return DESMAN_OK;
}

// ------------------------------------------------------------------------------------------ joint states, general tau
extern "C" int desman_loglik_general(desman_ctx *c, const double *tau, const double *gamma, const double *eta, int G, double *ll)
{
    if (!c || c->V <= 0) return fail(DESMAN_ESTATE, "no counts set");
    if (!tau || !gamma || !eta || !ll || G < 1 || G > DESMAN_MAX_G) return fail(DESMAN_EINVAL, "desman_loglik_general: bad argument");
    CU(cudaSetDevice(c->device));
    RET(ensure_ll_const(c));
    const size_t ntau = (size_t)c->V * G * 4, nsg = (size_t)c->S * G;
    const int nb = c->sm_count * 4;
    double *d = nullptr;
    CU(dmalloc(c, &d, (ntau + nsg + 16 + nb) * sizeof(double)));
    double *d_tau = d, *d_gamma = d + ntau, *d_eta = d_gamma + nsg, *d_part = d_eta + 16;
    CU(cudaMemcpyAsync(d_tau, tau, ntau * 8, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(d_gamma, gamma, nsg * 8, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(d_eta, eta, 16 * 8, cudaMemcpyHostToDevice, c->stream));
    loglik_general_kernel<<<nb, 256, 0, c->stream>>>(c->counts, d_tau, d_gamma, d_eta, (int)c->V, c->S, G, d_part);
    CU(cudaGetLastError());
    std::vector<double> h(nb);
    CU(cudaMemcpyAsync(h.data(), d_part, nb * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    dfree(c, d);
    if (c->nranks > 1) return fail(DESMAN_ESTATE, "desman_loglik_general is a single-rank call");
    double t = 0.0;
    for (int i = 0; i < nb; i++) t += h[i];
    *ll = c->ll_const_total + t;
    return DESMAN_OK;
}

extern "C" int desman_state_logprob(desman_ctx *c, const int64_t *variants, int64_t N, int S, const double *gamma, const double *eta,
                                    int G, const int64_t *index, double *logprob, double *lp_at_index, double *maxlp, double *lse,
                                    int64_t *argmax)
{
    if (!c) return fail(DESMAN_EINVAL, "desman_state_logprob: ctx is NULL");
    if (!variants) {
        if (c->V <= 0) return fail(DESMAN_ESTATE, "no counts set");
        N = c->V; S = c->S;
    }
    if (N <= 0 || S <= 0 || !gamma || !eta || G < 1) return fail(DESMAN_EINVAL, "desman_state_logprob: bad argument");
    if (index && !lp_at_index) return fail(DESMAN_EINVAL, "desman_state_logprob: index without lp_at_index");
    const int K = 4 * S;
    if (G > 15 || ((size_t)1 << (2 * G)) * (size_t)K * 8 > ((size_t)4 << 30))
        return fail(DESMAN_EINVAL, "4^G joint states do not fit: 4^%d * %d doubles exceed 4 GiB (the reference's own tauStates "
                                   "table, HaploSNP_Sampler.py:95-103, is out of reach well before that)", G, K);
    const long long T = 1ll << (2 * G);
    if (index) for (int64_t n = 0; n < N; n++) if (index[n] < 0 || index[n] >= T) return fail(DESMAN_EINVAL, "index[%lld] is not a state", (long long)n);
    CU(cudaSetDevice(c->device));
    const int nblk = (int)((T + ST_BN - 1) / ST_BN);
    // sites per pass: partials (24 B per site and state block) within 64 MB, the optional full rows within 256 MB
    long long Nc = ((long long)64 << 20) / ((long long)nblk * 24);
    if (logprob) { const long long m = ((long long)256 << 20) / (T * 8); if (m < Nc) Nc = m; }
    Nc = (Nc / ST_BM) * ST_BM;
    if (Nc < ST_BM) Nc = ST_BM;
    if (Nc > N) Nc = N;
    const size_t nsg = (size_t)S * G;
    double *d_gamma = nullptr, *d_ls = nullptr, *d_cd = nullptr, *d_pm = nullptr, *d_ps = nullptr, *d_lp = nullptr, *d_out = nullptr,
           *d_full = nullptr;
    long long *d_pa = nullptr, *d_idx = nullptr, *d_v64 = nullptr, *d_arg = nullptr;
    int rc = DESMAN_OK;
    auto body = [&]() -> int {
        CU(dmalloc(c, &d_gamma, (nsg + 16) * 8));
        CU(dmalloc(c, &d_ls, (size_t)K * T * 8));
        CU(dmalloc(c, &d_cd, (size_t)K * Nc * 8));
        CU(dmalloc(c, &d_pm, (size_t)Nc * nblk * 8));
        CU(dmalloc(c, &d_ps, (size_t)Nc * nblk * 8));
        CU(dmalloc(c, &d_pa, (size_t)Nc * nblk * 8));
        CU(dmalloc(c, &d_out, (size_t)Nc * 2 * 8));
        CU(dmalloc(c, &d_arg, (size_t)Nc * 8));
        CU(dmalloc(c, &d_lp, (size_t)Nc * 8));
        if (index) CU(dmalloc(c, &d_idx, (size_t)Nc * 8));
        if (variants) CU(dmalloc(c, &d_v64, (size_t)Nc * K * 8));
        if (logprob) CU(dmalloc(c, &d_full, (size_t)Nc * T * 8));
        CU(cudaMemcpyAsync(d_gamma, gamma, nsg * 8, cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(d_gamma + nsg, eta, 16 * 8, cudaMemcpyHostToDevice, c->stream));
        state_logsite_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(d_gamma, d_gamma + nsg, S, G, T, d_ls);
        CU(cudaGetLastError());
        for (long long n0 = 0; n0 < N; n0 += Nc) {
            const int nc = (int)((N - n0 < Nc) ? N - n0 : Nc);
            if (variants) CU(cudaMemcpyAsync(d_v64, variants + (size_t)n0 * K, (size_t)nc * K * 8, cudaMemcpyHostToDevice, c->stream));
            if (index) CU(cudaMemcpyAsync(d_idx, index + n0, (size_t)nc * 8, cudaMemcpyHostToDevice, c->stream));
            state_counts_kernel<<<c->sm_count * 4, 256, 0, c->stream>>>(variants ? d_v64 : nullptr, c->counts, variants ? 0 : n0, nc, K, d_cd);
            StateParams p;
            p.Cd = d_cd; p.logSiteT = d_ls; p.Nc = nc; p.K = K; p.T = T; p.nblk = nblk;
            p.part_max = d_pm; p.part_sum = d_ps; p.part_arg = d_pa;
            p.index = index ? d_idx : nullptr; p.lp_at_index = d_lp; p.logprob = d_full;
            state_logprob_kernel<<<dim3((unsigned)nblk, (unsigned)((nc + ST_BM - 1) / ST_BM)), 256, 0, c->stream>>>(p);
            state_reduce_kernel<<<(nc + 127) / 128, 128, 0, c->stream>>>(d_pm, d_ps, d_pa, nc, nblk, d_out, d_out + nc, d_arg);
            CU(cudaGetLastError());
            if (maxlp) CU(cudaMemcpyAsync(maxlp + n0, d_out, (size_t)nc * 8, cudaMemcpyDeviceToHost, c->stream));
            if (lse) CU(cudaMemcpyAsync(lse + n0, d_out + nc, (size_t)nc * 8, cudaMemcpyDeviceToHost, c->stream));
            if (argmax) CU(cudaMemcpyAsync(argmax + n0, d_arg, (size_t)nc * 8, cudaMemcpyDeviceToHost, c->stream));
            if (index) CU(cudaMemcpyAsync(lp_at_index + n0, d_lp, (size_t)nc * 8, cudaMemcpyDeviceToHost, c->stream));
            if (logprob) CU(cudaMemcpyAsync(logprob + (size_t)n0 * T, d_full, (size_t)nc * T * 8, cudaMemcpyDeviceToHost, c->stream));
            CU(cudaStreamSynchronize(c->stream));
        }
        return DESMAN_OK;
    };
    rc = body();
    void *bufs[] = {d_gamma, d_ls, d_cd, d_pm, d_ps, d_pa, d_out, d_arg, d_lp, d_idx, d_v64, d_full};
    for (void *b : bufs) if (b) dfree(c, b);
    return rc;
}

// Validation of the tensor-memory screening pass (tests only): regroup the current state and return, per site, the 3G sums
// D[v][3g+j] (log2 units) its contraction produced (NaN: the site is in no group) and the site's undecided-strain mask
// (0xffffffff: not listed = every step decided "stay").
extern "C" int desman_debug_screen(desman_ctx *c, float *D, uint32_t *mask)
{
    RET(require_state(c));
    if (!group_config(c, nullptr, nullptr) || !group_use_tc(c))
        return fail(DESMAN_ESTATE, "desman_debug_screen: the tensor-memory screening pass does not apply to this shape / option set");
    RET(sync_table(c));
    const size_t n = (size_t)c->V * 3 * c->G;
    float *dD = nullptr;
    CU(dmalloc(c, &dD, n * sizeof(float)));
    CU(cudaMemsetAsync(dD, 0xff, n * sizeof(float), c->stream));
    TauGroupParams q;
    q.countsf = c->countsf; q.nsite = c->nsite; q.gamma = c->gamma; q.eta = c->eta; q.words = nullptr;
    q.V = (int)c->V; q.S = c->S; q.G = c->G; q.grp = group_ptrs(c); q.tier_counts = nullptr;
    int rc = launch_tau_group_tc(c, q, dD);
    if (rc == DESMAN_OK && D) CU(cudaMemcpyAsync(D, dD, n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    int g[GC_COUNT];
    CU(cudaMemcpyAsync(g, c->grp_gctl, sizeof(g), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    dfree(c, dD);
    if (rc != DESMAN_OK) return rc;
    if (!g[GC_IMG_OK]) return fail(DESMAN_ESTATE, "desman_debug_screen: the count image did not fit (%d rows)", g[GC_IMG_ROWS]);
    if (mask) {
        for (size_t v = 0; v < (size_t)c->V; v++) mask[v] = 0xffffffffu;
        std::vector<uint2> w((size_t)g[GC_NWORK]);
        if (!w.empty()) CU(cudaMemcpyAsync(w.data(), c->grp_work, w.size() * sizeof(uint2), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        for (auto &e : w) mask[e.x] = e.y;
    }
    return DESMAN_OK;
}

// ------------------------------------------------------------------------------------------ chains
static int alloc_stores(desman_ctx *c, int n_iter, bool with_ge, StoreBufs *sb, const double *h_gs, const double *h_es)
{
    const size_t nsg = (size_t)c->S * c->G;
    size_t bytes = sizeof(double) * (3 * (size_t)n_iter + (with_ge || h_gs ? (size_t)n_iter * (nsg + 16) : 0)) + 256;
    RET(ensure_scratch(c, bytes));
    double *base = (double *)c->scratch;
    sb->ll = base; sb->lp = base + n_iter; sb->nch = base + 2 * (size_t)n_iter;
    if (with_ge || h_gs) { sb->gs = base + 3 * (size_t)n_iter; sb->es = sb->gs + (size_t)n_iter * nsg; }
    if (h_gs) {
        CU(cudaMemcpyAsync(sb->gs, h_gs, (size_t)n_iter * nsg * 8, cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(sb->es, h_es, (size_t)n_iter * 16 * 8, cudaMemcpyHostToDevice, c->stream));
    }
    return DESMAN_OK;
}

static int fetch_stores(desman_ctx *c, int n_iter, const StoreBufs &sb, double *gamma_store, double *eta_store,
                        double *ll_store, double *lp_store, int64_t *nchange_store)
{
    const size_t nsg = (size_t)c->S * c->G;
    std::vector<double> nch(n_iter);
    if (ll_store) CU(cudaMemcpyAsync(ll_store, sb.ll, n_iter * 8, cudaMemcpyDeviceToHost, c->stream));
    if (lp_store) CU(cudaMemcpyAsync(lp_store, sb.lp, n_iter * 8, cudaMemcpyDeviceToHost, c->stream));
    if (nchange_store) CU(cudaMemcpyAsync(nch.data(), sb.nch, n_iter * 8, cudaMemcpyDeviceToHost, c->stream));
    if (gamma_store && sb.gs) CU(cudaMemcpyAsync(gamma_store, sb.gs, (size_t)n_iter * nsg * 8, cudaMemcpyDeviceToHost, c->stream));
    if (eta_store && sb.es) CU(cudaMemcpyAsync(eta_store, sb.es, (size_t)n_iter * 16 * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (nchange_store) for (int i = 0; i < n_iter; i++) nchange_store[i] = (int64_t)llround(nch[i]);
    RET(check_xch(c));
    if (c->agg_ctl) {
        int ctl[3] = {0, 0, 0};
        CU(cudaMemcpyAsync(ctl, c->agg_ctl, sizeof(ctl), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        if (ctl[2]) { c->agg_valid = false; return fail(DESMAN_ESTATE, "pattern table overflow or inconsistency (internal error)"); }
    }
    return DESMAN_OK;
}

static int prepare_profiling(desman_ctx *c)
{
    timing_reset(c);
    if (c->prof_flush && !c->flush_buf) {
        c->flush_n = ((size_t)256 << 20) / sizeof(uint4);
        CU(dmalloc(c, &c->flush_buf, c->flush_n * sizeof(uint4)));
    }
    return DESMAN_OK;
}

// update(), HaploSNP_Sampler.py:334-365
extern "C" int desman_update(desman_ctx *c, int n_iter, double *gamma_store, double *eta_store, double *ll_store,
                             double *lp_store, int64_t *nchange_store)
{
    RET(require_state(c));
    if (n_iter < 0) return fail(DESMAN_EINVAL, "n_iter < 0");
    if (c->rng_mode != DESMAN_RNG_PHILOX)
        return fail(DESMAN_ESTATE, "desman_update needs DESMAN_RNG_PHILOX: the mu/gamma/eta draws of the reference follow numpy's "
                                   "sequential legacy stream, which has no parallel form (DESIGN.md section 4)");
    RET(ensure_ll_const(c));
    StoreBufs sb;
    RET(alloc_stores(c, n_iter > 0 ? n_iter : 1, true, &sb, nullptr, nullptr));
    if ((size_t)n_iter > c->esum_store_cap) {          // E_store[i].sum(axis=(0,1)) per sweep (chibMarginalLogLikelihood2, :557)
        if (c->esum_store) dfree(c, c->esum_store);
        c->esum_store = nullptr; c->esum_store_cap = 0;
        CU(dmalloc(c, &c->esum_store, (size_t)n_iter * 16 * sizeof(unsigned long long)));
        c->esum_store_cap = (size_t)n_iter;
    }
    RET(prepare_profiling(c));
    const size_t nvg = (size_t)c->V * c->G;
    CU(cudaMemsetAsync(c->tau_cnt, 0, nvg * 4 * sizeof(uint32_t), c->stream));
    CU(cudaMemsetAsync(c->tau_last, 0, nvg * sizeof(uint32_t), c->stream));
    // pre-sweep ll/lp and star state (:336-338)
    RET(sync_table(c));
    RET(launch_finalize(c, c->gamma, c->eta, -1, 0, sb, false));
    // Under sharding the [ll, nchange] words of sweep k travel with the statistics of sweep k+1 (one exchange, i.e. one
    // inter-rank synchronisation, per sweep instead of two), and lp / stores / MAP bookkeeping of sweep k follow that exchange --
    // still before gamma and tau of sweep k+1 change.  The words are double buffered on the parity of the sweep.
    const bool lagged = c->nranks > 1;
    for (int it = 0; it < n_iter; it++) {
        sweep_begin(c);
        unsigned long long *red_prev = c->red_i;
        if (lagged) c->red_i = c->red_base + 4 * (it & 1);
        // clears the accumulators; table upkeep when pending; the MAP snapshot of the previous sweep (single rank: under
        // sharding the bookkeeping of sweep it-1 follows the exchange below and keeps its own copy launch)
        RET(sync_table(c, !lagged && it > 0));
        RET(launch_mu(c, c->gamma, c->eta));                            // sampleMu   (:341)
        if (lagged && it > 0) {
            // ll, lp, stores, star of sweep it-1; its MAP snapshot tau_star <- tau rides in the screening launch when that runs
            c->star_fold = !c->fixed_tau && tau_runs_tc(c);
            RET(exchange_and_finalize(c, red_prev, c->gamma, c->eta, it - 1, sb, !c->star_fold));
        } else RET(allreduce_stats(c));
        RET(launch_draw(c, c->stats, c->gamma, c->eta_new, c->esum_store + (size_t)it * 16));   // sampleGamma (:342) + sampleEta's draw (:347)
        // (early: the statistics and draw launches separate the screening pass from the maintenance launch that wrote its inputs)
        if (!c->fixed_tau) RET(launch_tau(c, c->gamma, c->eta, true, true, (uint32_t)it, 0, nullptr, 1));
        if (c->star_fold) return fail(DESMAN_ESTATE, "internal: the MAP snapshot was not taken by the screening launch");       // sample_tau (:345), old eta (nchange cleared by sync_table)
        if (lagged) RET(launch_ll(c, c->gamma, c->eta_new, c->eta, true));     // eta <- new (:347); sum n*log p of sweep it, reduced with the next exchange
        else RET(launch_finalize(c, c->gamma, c->eta_new, it, 0, sb, true, c->eta, false));   // eta <- new (:347); ll, lp, stores, star (:349-358)
        sweep_end(c);
        c->sweep++;
    }
    if (lagged && n_iter > 0) {
        RET(allreduce_red(c));
        RET(launch_finalize_only(c, c->red_i, c->gamma, c->eta, n_iter - 1, 0, sb, true));
    } else if (n_iter > 0) RET(launch_star_copy(c));                    // the snapshot of the last sweep
    c->red_i = c->red_base;
    {
        KSpan k(c, DESMAN_K_OTHER);
        flush_tau_counts_kernel<<<c->sm_count * 2, 256, 0, c->stream>>>(c->tau, c->tau_cnt, c->tau_last, nvg, (uint32_t)n_iter);
    }
    CU(cudaGetLastError());
    c->last_n_iter = (uint32_t)n_iter;
    RET(fetch_stores(c, n_iter, sb, gamma_store, eta_store, ll_store, lp_store, nchange_store));
    timing_collect(c);
    return DESMAN_OK;
}

// updateTau(), HaploSNP_Sampler.py:383-407
extern "C" int desman_update_tau(desman_ctx *c, int n_iter, const double *gamma_store, const double *eta_store,
                                 double *ll_store, double *lp_store, int64_t *nchange_store)
{
    RET(require_state(c));
    if (n_iter <= 0 || !gamma_store || !eta_store) return fail(DESMAN_EINVAL, "update_tau needs n_iter > 0 and the gamma/eta stores");
    const size_t nsg = (size_t)c->S * c->G;
    for (size_t i = 0; i < (size_t)n_iter * nsg; i++) if (!(gamma_store[i] > 0.0)) return fail(DESMAN_EINVAL, "gamma_store[%zu] must be > 0", i);
    for (size_t i = 0; i < (size_t)n_iter * 16; i++) if (!(eta_store[i] > 0.0)) return fail(DESMAN_EINVAL, "eta_store[%zu] must be > 0", i);
    RET(ensure_ll_const(c));
    StoreBufs sb;
    RET(alloc_stores(c, n_iter, false, &sb, gamma_store, eta_store));
    RET(prepare_profiling(c));
    const size_t nvg = (size_t)c->V * c->G;
    CU(cudaMemsetAsync(c->tau_cnt, 0, nvg * 4 * sizeof(uint32_t), c->stream));
    CU(cudaMemsetAsync(c->tau_last, 0, nvg * sizeof(uint32_t), c->stream));
    // lp_star from (gamma_store[0], tau, eta_store[0]) (:386-388)
    RET(sync_table(c));
    RET(launch_finalize(c, sb.gs, sb.es, -1, 1, sb, false));
    for (int it = 0; it < n_iter; it++) {
        const double *gm = sb.gs + (size_t)it * nsg, *et = sb.es + (size_t)it * 16;
        sweep_begin(c);
        RET(sync_table(c));
        RET(launch_tau(c, gm, et, true, true, (uint32_t)it));
        RET(launch_finalize(c, gm, et, it, 1, sb, false));
        sweep_end(c);
        if (c->rng_mode == DESMAN_RNG_PHILOX) c->sweep++;
    }
    flush_tau_counts_kernel<<<c->sm_count * 2, 256, 0, c->stream>>>(c->tau, c->tau_cnt, c->tau_last, nvg, (uint32_t)n_iter);
    CU(cudaGetLastError());
    c->last_n_iter = (uint32_t)n_iter;
    RET(fetch_stores(c, n_iter, sb, nullptr, nullptr, ll_store, lp_store, nchange_store));
    timing_collect(c);
    return DESMAN_OK;
}

extern "C" int desman_get_star(desman_ctx *c, int64_t *tau_star, double *gamma_star, double *eta_star, double *lp_star,
                               int *iter_star)
{
    RET(require_state(c));
    std::vector<uint8_t> idx;
    double sc[4];
    if (tau_star) {
        idx.resize((size_t)c->V * c->G);
        CU(cudaMemcpyAsync(idx.data(), c->tau_star, idx.size(), cudaMemcpyDeviceToHost, c->stream));
    }
    if (gamma_star) CU(cudaMemcpyAsync(gamma_star, c->gamma_star, (size_t)c->S * c->G * 8, cudaMemcpyDeviceToHost, c->stream));
    if (eta_star) CU(cudaMemcpyAsync(eta_star, c->eta_star, 16 * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(sc, c->scal, sizeof(sc), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (tau_star) index_to_onehot(idx.data(), idx.size(), tau_star);
    if (lp_star) *lp_star = sc[0];
    if (iter_star) *iter_star = (int)sc[1];
    return DESMAN_OK;
}

// Esum[a_obs][b_true] of every sweep of the last desman_update (E_store[i].sum(axis=(0,1)), HaploSNP_Sampler.py:557)
extern "C" int desman_get_esum_store(desman_ctx *c, int64_t *esum_store)
{
    RET(require_state(c));
    const size_t n = (size_t)c->last_n_iter * 16;
    if (!esum_store || n == 0 || c->last_n_iter > c->esum_store_cap) return fail(DESMAN_ESTATE, "no update() to report");
    CU(cudaMemcpyAsync(esum_store, c->esum_store, n * sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return DESMAN_OK;
}

extern "C" int desman_get_star_index(desman_ctx *c, uint8_t *tau_star_idx)
{
    RET(require_state(c));
    CU(cudaMemcpyAsync(tau_star_idx, c->tau_star, (size_t)c->V * c->G, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return DESMAN_OK;
}

extern "C" int desman_get_tau_sum(desman_ctx *c, int64_t *tau_sum)
{
    RET(require_state(c));
    const size_t n = (size_t)c->V * c->G * 4;
    std::vector<uint32_t> h(n);
    CU(cudaMemcpyAsync(h.data(), c->tau_cnt, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    parallel_ranges(n, [&](size_t lo, size_t hi) { for (size_t i = lo; i < hi; i++) tau_sum[i] = (int64_t)h[i]; });
    return DESMAN_OK;
}

extern "C" int desman_get_tau_sum_u32(desman_ctx *c, uint32_t *tau_sum)
{
    RET(require_state(c));
    // pageable D2H copies crawl through a driver bounce buffer: go through the pinned staging buffers (double buffered) and
    // copy out on a few threads
    std::lock_guard<std::mutex> pin_lock(g_pin_mu);
    RET(ensure_pinned());
    const size_t total = (size_t)c->V * c->G * 4 * sizeof(uint32_t), chunk = PIN_CELLS * sizeof(int4);
    const char *src = (const char *)c->tau_cnt;
    char *dst = (char *)tau_sum;
    int buf = 0;
    size_t off_prev = 0, n_prev = 0;
    for (size_t off = 0; off < total || n_prev; off += chunk, buf ^= 1) {
        size_t n = 0;
        if (off < total) {
            n = (total - off < chunk) ? total - off : chunk;
            CU(cudaMemcpyAsync(g_pin[buf], src + off, n, cudaMemcpyDeviceToHost, c->stream));
        }
        if (n_prev) {          // copy out the previous chunk while this one is in flight
            const char *ps = (const char *)g_pin[buf ^ 1];
            char *pd = dst + off_prev;
            parallel_ranges(n_prev, [&](size_t lo, size_t hi) { memcpy(pd + lo, ps + lo, hi - lo); });
        }
        CU(cudaStreamSynchronize(c->stream));
        off_prev = off; n_prev = n;
    }
    return DESMAN_OK;
}

// ------------------------------------------------------------------------------------------ NMFT
extern "C" int desman_nmft_factorize(desman_ctx *c, const int64_t *snps, int64_t V, int S, int G, double *tau, double *gamma,
                                     int max_iter, double min_change, int fix_gamma, int *n_iter_done, double *div_final,
                                     double *div_trace)
{
    if (!c || !snps || !tau || !gamma || V <= 0 || S <= 0 || G < 1 || G > DESMAN_MAX_G)
        return fail(DESMAN_EINVAL, "desman_nmft_factorize: bad arguments");
    CU(cudaSetDevice(c->device));
    return nmft_factorize_impl(c->stream, c->sm_count, snps, V, S, G, tau, gamma, max_iter, min_change, fix_gamma, n_iter_done,
                               div_final, div_trace, g_err, sizeof(g_err));
}

extern "C" int desman_nmft_last_timing(double *ms, int *iters)
{
    if (ms) *ms = g_nmft_last_ms;
    if (iters) *iters = g_nmft_last_iters;
    return DESMAN_OK;
}

// ------------------------------------------------------------------------------------------ comm
// data plane of the per-sweep exchange: 0 none (one rank), 1 NCCL all-reduce, 2 one-shot peer-memory mailboxes (exchange_kernel.cuh)
extern "C" int desman_comm_kind(desman_ctx *c)
{
    return c->nranks <= 1 ? 0 : (c->xch_ok ? 2 : 1);
}

extern "C" int desman_comm_unique_id(char id[128])
{
    RET(nccl_load());
    nccl_uid_t u;
    NC(g_nccl.GetUniqueId(&u));
    memcpy(id, u.internal, 128);
    return DESMAN_OK;
}

extern "C" int desman_comm_init(desman_ctx *c, const char id[128], int rank, int nranks)
{
    if (nranks < 1 || rank < 0 || rank >= nranks) return fail(DESMAN_EINVAL, "bad rank/nranks");
    if (nranks == 1) { c->rank = 0; c->nranks = 1; return DESMAN_OK; }
    RET(nccl_load());
    CU(cudaSetDevice(c->device));
    nccl_uid_t u;
    memcpy(u.internal, id, 128);
    NC(g_nccl.CommInitRank(&c->comm, nranks, u, rank));
    c->rank = rank; c->nranks = nranks;
    {   // NCCL connects its channels lazily at the first collective (~1 s on 8 GPUs): do that here, not inside the first sweep
        long long *dw = nullptr;
        CU(cudaMalloc(&dw, sizeof(long long)));
        CU(cudaMemsetAsync(dw, 0, sizeof(long long), c->stream));
        NC(g_nccl.AllReduce(dw, dw, 1, NCCL_INT64, NCCL_SUM, c->comm, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        cudaFree(dw);
    }
    // peer-memory mailboxes for the per-sweep exchange; any failure here leaves the NCCL all-reduce in place
    // default: on from 2 ranks (DESMAN_B200_P2P=0/1 forces it; tests/test_gpu_multi.py runs both data planes at 2 and 4 ranks
    // against the oracle).  Measured per sweep at C3 per GPU, one exchange per sweep, event-timed incl. the wait for the slowest
    // rank: 2 GPUs NCCL 43.8 us; 4 GPUs peer memory 21.6 us; round 1 at 8 GPUs, two exchanges: 45.7 us vs 64.8 us with NCCL
    const char *env = getenv("DESMAN_B200_P2P");
    if (const char *ef = getenv("DESMAN_B200_XCH_FUSE")) c->xch_fuse = atoi(ef) ? 1 : 0;   // (A/B of the fused exchange + bookkeeping launch)
    if (env ? !atoi(env) : nranks < 2) return DESMAN_OK;
    if (nranks > XCH_MAX_RANKS || !g_nccl.AllGather) return DESMAN_OK;
    const int cap = 4096;                                             // words per contribution (S*G + 16 must fit)
    const size_t words = (size_t)2 * nranks * cap + nranks;
    unsigned long long *mine = nullptr;
    if (cudaMalloc(&mine, words * sizeof(unsigned long long)) != cudaSuccess) { cudaGetLastError(); return DESMAN_OK; }
    cudaMemsetAsync(mine, 0, words * sizeof(unsigned long long), c->stream);
    cudaIpcMemHandle_t hmine;
    if (cudaIpcGetMemHandle(&hmine, mine) != cudaSuccess) { cudaGetLastError(); cudaFree(mine); return DESMAN_OK; }
    // all-gather the handles (bytes) through NCCL: every rank learns every mailbox
    char *dh = nullptr;
    CU(cudaMalloc(&dh, sizeof(cudaIpcMemHandle_t) * (size_t)(nranks + 1)));
    CU(cudaMemcpyAsync(dh + sizeof(cudaIpcMemHandle_t) * (size_t)nranks, &hmine, sizeof(hmine), cudaMemcpyHostToDevice, c->stream));
    NC(g_nccl.AllGather(dh + sizeof(cudaIpcMemHandle_t) * (size_t)nranks, dh, sizeof(cudaIpcMemHandle_t), NCCL_INT8, c->comm, c->stream));
    std::vector<cudaIpcMemHandle_t> all(nranks);
    CU(cudaMemcpyAsync(all.data(), dh, sizeof(cudaIpcMemHandle_t) * (size_t)nranks, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    cudaFree(dh);
    int ok = 1;
    for (int r = 0; r < nranks; r++) {
        if (r == rank) { c->xch_mail[r] = mine; continue; }
        void *q = nullptr;
        if (cudaIpcOpenMemHandle(&q, all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; break; }
        c->xch_mail[r] = (unsigned long long *)q;
    }
    // every rank must take the same path: agree on success with a tiny all-reduce (min via sum of failures)
    long long *dflag = nullptr;
    CU(cudaMalloc(&dflag, sizeof(long long)));
    long long bad = ok ? 0 : 1;
    CU(cudaMemcpyAsync(dflag, &bad, sizeof(bad), cudaMemcpyHostToDevice, c->stream));
    NC(g_nccl.AllReduce(dflag, dflag, 1, NCCL_INT64, NCCL_SUM, c->comm, c->stream));
    CU(cudaMemcpyAsync(&bad, dflag, sizeof(bad), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    cudaFree(dflag);
    if (bad == 0) {
        CU(cudaMalloc(&c->xch_err, sizeof(int)));
        CU(cudaMemsetAsync(c->xch_err, 0, sizeof(int), c->stream));
        CU(cudaStreamSynchronize(c->stream));
        c->xch_cap_words = cap; c->xch_seq = 0; c->xch_ok = true;
    }
    return DESMAN_OK;
}

// ------------------------------------------------------------------------------------------ options
extern "C" int desman_set_option(desman_ctx *c, const char *name, int64_t value)
{
    if (!c || !name) return fail(DESMAN_EINVAL, "desman_set_option: bad arguments");
    if (!strcmp(name, "tau_exact")) { c->tau_exact = value ? 1 : 0; return DESMAN_OK; }
    if (!strcmp(name, "fixed_tau")) { c->fixed_tau = value ? 1 : 0; return DESMAN_OK; }
    if (!strcmp(name, "mu_mode")) { c->mu_mode = (value == 0 || value == 1) ? (int)value : 2; return DESMAN_OK; }
    if (!strcmp(name, "pdl")) { c->pdl = value ? 1 : 0; return DESMAN_OK; }
    // the single-step calls that draw no tau (mu_stats, draw_gamma_eta) leave the Philox sweep counter where it is: a caller
    // looping over them (chibMarginalLogLikelihood, HaploSNP_Sampler.py:621-710) moves it on itself
    if (!strcmp(name, "advance_sweep")) { c->sweep += (uint32_t)value; return DESMAN_OK; }
    // stream of the tau draws; the Philox sweep counter and the MT19937 position are both kept across a switch
    if (!strcmp(name, "rng_mode")) {
        if (value != DESMAN_RNG_MT19937 && value != DESMAN_RNG_PHILOX) return fail(DESMAN_EINVAL, "bad rng_mode %lld", (long long)value);
        c->rng_mode = (int)value; return DESMAN_OK;
    }
    if (!strcmp(name, "tau_group_mma")) { c->tau_group_mma = value ? 1 : 0; return DESMAN_OK; }
    if (!strcmp(name, "tau_open")) { c->tau_open = value ? 1 : 0; return DESMAN_OK; }
    if (!strcmp(name, "xch_fuse")) { c->xch_fuse = value ? 1 : 0; return DESMAN_OK; }
    if (!strcmp(name, "tau_group_tc")) { c->tau_group_tc = value ? 1 : 0; c->agg_valid = false; return DESMAN_OK; }
    if (!strcmp(name, "tau_group")) { c->tau_group = (value == 0 || value == 1) ? (int)value : 2; c->agg_valid = false; return DESMAN_OK; }
    return fail(DESMAN_EINVAL, "unknown option '%s'", name);
}

extern "C" int desman_get_tier_counts(desman_ctx *c, int64_t out[3], int reset)
{
    CU(cudaSetDevice(c->device));
    unsigned long long h[3];
    CU(cudaMemcpyAsync(h, c->tiers, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    if (reset) CU(cudaMemsetAsync(c->tiers, 0, sizeof(h), c->stream));
    CU(cudaStreamSynchronize(c->stream));
    for (int i = 0; i < 3; i++) out[i] = (int64_t)h[i];
    return DESMAN_OK;
}

extern "C" int desman_get_group_stats(desman_ctx *c, int64_t out[8])
{
    CU(cudaSetDevice(c->device));
    int h[GC_COUNT] = {0};
    if (c->grp_gctl) CU(cudaMemcpyAsync(h, c->grp_gctl, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    unsigned int ns = 0;
    if (c->agg_nslots) CU(cudaMemcpyAsync(&ns, c->agg_nslots, sizeof(ns), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    out[0] = h[GC_HAVE]; out[1] = h[GC_CALM]; out[2] = h[GC_NITEMS]; out[3] = h[GC_NSINGLES]; out[4] = h[GC_NWORK];
    out[5] = h[GC_ORPHANS]; out[6] = (int64_t)ns; out[7] = (group_config(c, nullptr, nullptr) ? 1 : 0) | (h[GC_WORTH] ? 2 : 0);
    return DESMAN_OK;
}

// ------------------------------------------------------------------------------------------ measurement
#ifdef KPROF
// diagnosis build only (tools/kprof.py): copy out / reset the device-side timeline records
extern "C" int desman_kprof_dump(void *out, int cap, int reset)
{
    unsigned int n = 0;
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(&n, g_krec_n, sizeof(n));
    if (n > KREC_CAP) n = KREC_CAP;
    if ((int)n > cap) n = cap;
    if (out && n) cudaMemcpyFromSymbol(out, g_krec, (size_t)n * sizeof(KRec));
    if (reset) { const unsigned int z = 0; cudaMemcpyToSymbol(g_krec_n, &z, sizeof(z)); }
    return (int)n;
}
#endif
extern "C" int desman_set_profiling(desman_ctx *c, int per_kernel_events, int flush_l2_between_sweeps)
{
    c->prof_kernels = per_kernel_events; c->prof_flush = flush_l2_between_sweeps;
    return DESMAN_OK;
}

extern "C" int desman_get_timing(desman_ctx *c, double *elapsed_ms, double kernel_ms[DESMAN_K_COUNT],
                                 int64_t kernel_launches[DESMAN_K_COUNT])
{
    if (elapsed_ms) *elapsed_ms = c->elapsed_ms;
    for (int i = 0; i < DESMAN_K_COUNT; i++) {
        if (kernel_ms) kernel_ms[i] = c->k_ms[i];
        if (kernel_launches) kernel_launches[i] = c->k_launch[i];
    }
    return DESMAN_OK;
}

// ------------------------------------------------------------------------------------------ reference ABI
// Process-global stream, like the file-static gsl_rng of c_sample_tau.c:24.
static desman_ctx *g_legacy = nullptr;

extern "C" void c_initRNG(void)
{
    if (g_legacy) return;
    int dev = 0;
    const char *env = getenv("DESMAN_B200_DEVICE");
    if (env) dev = atoi(env);
    if (desman_ctx_create(&g_legacy, dev, 0, DESMAN_RNG_MT19937) != DESMAN_OK) {
        fprintf(stderr, "desman_b200: c_initRNG failed: %s\n", g_err);
        g_legacy = nullptr;
    }
}

extern "C" void c_setRNG(unsigned long int seed)
{
    if (!g_legacy) { fail(DESMAN_ESTATE, "c_setRNG before c_initRNG"); fprintf(stderr, "desman_b200: %s\n", g_err); return; }
    if (desman_set_rng(g_legacy, (uint64_t)seed, 0, 0) != DESMAN_OK) fprintf(stderr, "desman_b200: c_setRNG failed: %s\n", g_err);
}

extern "C" void c_freeRNG(void)
{
    if (g_legacy) desman_ctx_destroy(g_legacy);
    g_legacy = nullptr;
}

extern "C" int c_sample_tau(long *anTau, double *adPi, double *adEta, long *anVariants, int nV, int nG, int nS)
{
    if (!g_legacy) { fail(DESMAN_ESTATE, "c_sample_tau before c_initRNG (NULL dereference in the reference, c_sample_tau.c:174)"); return -1; }
    desman_ctx *c = g_legacy;
    if (!anTau || !adPi || !adEta || !anVariants || nV <= 0 || nG <= 0 || nS <= 0) { fail(DESMAN_EINVAL, "c_sample_tau: bad arguments"); return -1; }
    // the reference borrows all four arrays for the call and retains nothing (c_sample_tau.c:107-114,192-195)
    if (desman_set_counts(c, (const int64_t *)anVariants, nV, nS, 0, nV) != DESMAN_OK) return -1;
    if (set_state_impl(c, (const int64_t *)anTau, adPi, adEta, nG, true) != DESMAN_OK) return -1;
    std::vector<uint8_t> before((size_t)nV * nG), after((size_t)nV * nG);
    if (desman_get_tau_index(c, before.data()) != DESMAN_OK) return -1;
    int64_t nchange = 0;
    if (desman_sample_tau(c, &nchange) != DESMAN_OK) return -1;
    if (desman_get_tau_index(c, after.data()) != DESMAN_OK) return -1;
    for (size_t i = 0; i < after.size(); i++)
        if (after[i] != before[i]) { anTau[i * 4 + before[i]] = 0; anTau[i * 4 + after[i]] = 1; }   // :178-184
    return (int)nchange;
}

// ------------------------------------------------------------------------------------------ batched reference ABI
// Eta_Sampler (Eta_Sampler.py:355-369, :430-446) calls sample_tau once per gene and iteration, each call with its own masked
// gamma and a few dozen to a few thousand sites: thousands of tiny launches, each re-uploading its counts.  A batch keeps the
// counts of all genes resident and runs ONE launch per iteration over all of them; draw for draw (one process-global MT19937
// stream, gene after gene, V_k*G words each) the results are those of the sequence of c_sample_tau calls it replaces.
struct desman_batch {
    int nprob = 0, S = 0;
    std::vector<int> off;            // [nprob + 1] first site of every problem
    int4 *counts = nullptr;
    uint8_t *tau = nullptr;
    double *gamma = nullptr, *eta = nullptr;
    int *d_off = nullptr;
    unsigned long long *nchange = nullptr;
    size_t cap_vg = 0, cap_g = 0;
    int maxV = 0;
};

extern "C" int desman_batch_create(desman_batch **out, int nprob, const int64_t *const *variants, const int *nV, int nS)
{
    if (!g_legacy) return fail(DESMAN_ESTATE, "desman_batch_create before c_initRNG");
    if (!out || nprob <= 0 || !variants || !nV || nS <= 0) return fail(DESMAN_EINVAL, "desman_batch_create: bad arguments");
    desman_ctx *c = g_legacy;
    CU(cudaSetDevice(c->device));
    desman_batch *b = new desman_batch();
    b->nprob = nprob; b->S = nS; b->off.assign(nprob + 1, 0);
    for (int k = 0; k < nprob; k++) {
        if (nV[k] < 0 || (nV[k] > 0 && !variants[k])) { delete b; return fail(DESMAN_EINVAL, "desman_batch_create: problem %d", k); }
        b->off[k + 1] = b->off[k] + nV[k];
        if (nV[k] > b->maxV) b->maxV = nV[k];
    }
    const size_t Vt = (size_t)b->off[nprob];
    if (Vt == 0) { delete b; return fail(DESMAN_EINVAL, "desman_batch_create: no sites"); }
    std::vector<int4> h(Vt * nS);
    long long bad = -1, big = 0;
    for (int k = 0; k < nprob; k++) {
        const int64_t *src = variants[k];
        int4 *dst = h.data() + (size_t)b->off[k] * nS;
        for (size_t i = 0; i < (size_t)nV[k] * nS; i++) {
            const int64_t a0 = src[4 * i], a1 = src[4 * i + 1], a2 = src[4 * i + 2], a3 = src[4 * i + 3];
            if ((a0 | a1 | a2 | a3 | (DESMAN_MAX_COUNT - a0) | (DESMAN_MAX_COUNT - a1) | (DESMAN_MAX_COUNT - a2) | (DESMAN_MAX_COUNT - a3)) < 0) bad = k;
            big |= a0 | a1 | a2 | a3;
            dst[i] = make_int4((int)a0, (int)a1, (int)a2, (int)a3);
        }
    }
    if (bad >= 0) { delete b; return fail(DESMAN_EINVAL, "counts of problem %lld must be in [0, %d] per (v,s,base) cell", bad, DESMAN_MAX_COUNT); }
    CU(cudaMalloc(&b->counts, Vt * nS * sizeof(int4)));
    CU(cudaMalloc(&b->d_off, (nprob + 1) * sizeof(int)));
    CU(cudaMalloc(&b->eta, 16 * sizeof(double)));
    CU(cudaMalloc(&b->nchange, sizeof(unsigned long long)));
    CU(cudaMemcpyAsync(b->counts, h.data(), Vt * nS * sizeof(int4), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(b->d_off, b->off.data(), (nprob + 1) * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    *out = b;
    return DESMAN_OK;
}

extern "C" int desman_batch_destroy(desman_batch *b)
{
    if (!b) return DESMAN_OK;
    for (void *q : {(void *)b->counts, (void *)b->tau, (void *)b->gamma, (void *)b->eta, (void *)b->d_off, (void *)b->nchange}) if (q) cudaFree(q);
    delete b;
    return DESMAN_OK;
}

// tau[k]: int64 one-hot [V_k][G][4], mutated in place; pi[k]: [S][G] (columns may be exactly 0.0: masked strains); eta [4][4];
// nchange[k] = flips of problem k.  Equivalent to c_sample_tau(tau[k], pi[k], eta, variants[k], ...) for k = 0 .. nprob-1.
extern "C" int desman_batch_sample_tau(desman_batch *b, int64_t *const *tau, const double *const *pi, const double *eta, int nG,
                                       int *nchange)
{
    if (!g_legacy) return fail(DESMAN_ESTATE, "desman_batch_sample_tau before c_initRNG");
    if (!b || !tau || !pi || !eta || nG < 1 || nG > DESMAN_MAX_G) return fail(DESMAN_EINVAL, "desman_batch_sample_tau: bad arguments");
    desman_ctx *c = g_legacy;
    CU(cudaSetDevice(c->device));
    const int S = b->S, np = b->nprob;
    const size_t Vt = (size_t)b->off[np], nvg = Vt * nG, ng = (size_t)np * S * nG;
    for (int i = 0; i < 16; i++) if (!(eta[i] > 0.0)) return fail(DESMAN_EINVAL, "eta[%d] = %g must be > 0", i, eta[i]);
    std::vector<double> hg(ng);
    for (int k = 0; k < np; k++) {
        if (b->off[k + 1] == b->off[k]) continue;
        for (int s2 = 0; s2 < S; s2++) {
            double row = 0.0;
            for (int g = 0; g < nG; g++) {
                const double x = pi[k][(size_t)s2 * nG + g];
                if (!(x >= 0.0)) return fail(DESMAN_EINVAL, "problem %d: gamma[%d][%d] = %g must be >= 0", k, s2, g, x);
                row += x;
                hg[((size_t)k * S + s2) * nG + g] = x;
            }
            if (!(row > 0.0)) return fail(DESMAN_EINVAL, "problem %d: gamma row %d has no positive entry", k, s2);
        }
    }
    std::vector<uint8_t> before(nvg), after(nvg);
    for (int k = 0; k < np; k++) {
        const size_t n = (size_t)(b->off[k + 1] - b->off[k]) * nG;
        if (n) RET(onehot_to_index(tau[k], n, before.data() + (size_t)b->off[k] * nG));
    }
    if (nvg > b->cap_vg) { if (b->tau) cudaFree(b->tau); b->tau = nullptr; CU(cudaMalloc(&b->tau, nvg)); b->cap_vg = nvg; }
    if (ng > b->cap_g) { if (b->gamma) cudaFree(b->gamma); b->gamma = nullptr; CU(cudaMalloc(&b->gamma, ng * sizeof(double))); b->cap_g = ng; }
    CU(cudaMemcpyAsync(b->tau, before.data(), nvg, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(b->gamma, hg.data(), ng * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(b->eta, eta, 16 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemsetAsync(b->nchange, 0, sizeof(unsigned long long), c->stream));
    // the V_total * G words of the stream, in problem order (what the sequence of calls would have consumed)
    const int64_t V_keep = c->V, Vt_keep = c->V_total, v0_keep = c->v0;
    const int G_keep = c->G;
    c->V = (int64_t)Vt; c->V_total = (int64_t)Vt; c->v0 = 0; c->G = nG;
    int rc = gen_mt_words(c);
    c->V = V_keep; c->V_total = Vt_keep; c->v0 = v0_keep; c->G = G_keep;
    RET(rc);
    TauParams p;
    memset(&p, 0, sizeof(p));
    p.counts = b->counts; p.tau = b->tau; p.gamma = b->gamma; p.eta = b->eta; p.words = c->words;
    p.seed = c->seed; p.sweep = 0; p.v0 = 0; p.V = (int)Vt; p.S = S; p.G = nG;
    p.nchange = b->nchange; p.tau_last = nullptr; p.exact_only = c->tau_exact; p.prob_off = b->d_off;
    const size_t smem = tau_smem_bytes(S, nG);
    if (smem > 227 * 1024) return fail(DESMAN_EINVAL, "S*G too large for the shared-memory tile (%zu bytes)", smem);
    CU(cudaFuncSetAttribute(tau_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int gx = (b->maxV + TAU_WARPS - 1) / TAU_WARPS;
    if (gx > 32) gx = 32;
    if (gx < 1) gx = 1;
    tau_sample_kernel<<<dim3((unsigned)gx, (unsigned)np), TAU_WARPS * 32, smem, c->stream>>>(p);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(after.data(), b->tau, nvg, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    for (int k = 0; k < np; k++) {
        int n = 0;
        int64_t *t = tau[k];
        const size_t lo = (size_t)b->off[k] * nG, hi = (size_t)b->off[k + 1] * nG;
        for (size_t i = lo; i < hi; i++)
            if (after[i] != before[i]) { t[(i - lo) * 4 + before[i]] = 0; t[(i - lo) * 4 + after[i]] = 1; n++; }   // :178-184
        if (nchange) nchange[k] = n;
    }
    return DESMAN_OK;
}


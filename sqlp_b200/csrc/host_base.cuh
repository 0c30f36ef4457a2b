// host_base.cuh -- errors, NCCL binding, device buffers, the three handles, launch + profiling helpers.
// Part of the single translation unit sqlp_api.cu (included there, in order).
#pragma once

namespace {

using namespace sqlp;

thread_local std::string g_err;

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

#define CK(call)                                                                          \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess)                                                            \
            throw Error(SQLP_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

#define REQUIRE(cond, code, msg)            \
    do {                                    \
        if (!(cond)) throw Error(code, msg); \
    } while (0)

template <class F>
int32_t guard(F &&f)
{
    try {
        f();
        return SQLP_OK;
    } catch (const Error &e) {
        g_err = e.what();
        return e.code;
    } catch (const std::bad_alloc &) {
        g_err = "host allocation failed";
        return SQLP_E_NOMEM;
    } catch (const std::exception &e) {
        g_err = e.what();
        return SQLP_E_INVALID;
    }
}

// ---------------------------------------------------------------- NCCL (dlopen) --------
// Only the scenario-sharded mode needs NCCL, so it is bound lazily; a single-GPU host never
// loads it.  Types restated from nccl.h (2.x ABI).
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess_ = 0 };
enum { ncclFloat64_ = 8 };  // ncclDataType_t: ncclDouble
struct NcclApi {
    void *h = nullptr;
    int (*GetUniqueId)(ncclUniqueId *) = nullptr;
    int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*Broadcast)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;

void load_nccl()
{
    if (g_nccl.h) return;
    const char *env = getenv("SQLP_NCCL_LIB");
    const char *names[] = {env, "libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (const char *n : names) {
        if (!n) continue;
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    REQUIRE(h, SQLP_E_NCCL, "cannot dlopen libnccl.so.2 (set SQLP_NCCL_LIB)");
    auto sym = [&](const char *s) {
        void *p = dlsym(h, s);
        REQUIRE(p, SQLP_E_NCCL, std::string("missing NCCL symbol ") + s);
        return p;
    };
    g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))sym("ncclGetUniqueId");
    g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))sym("ncclCommInitRank");
    g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))sym("ncclCommDestroy");
    g_nccl.Broadcast = (decltype(g_nccl.Broadcast))sym("ncclBroadcast");
    g_nccl.AllGather = (decltype(g_nccl.AllGather))sym("ncclAllGather");
    g_nccl.CommInitAll = (decltype(g_nccl.CommInitAll))sym("ncclCommInitAll");
    g_nccl.GroupStart = (decltype(g_nccl.GroupStart))sym("ncclGroupStart");
    g_nccl.GroupEnd = (decltype(g_nccl.GroupEnd))sym("ncclGroupEnd");
    g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))sym("ncclGetErrorString");
    g_nccl.h = h;
}
#define NK(call)                                                                            \
    do {                                                                                    \
        int r_ = (call);                                                                    \
        if (r_ != ncclSuccess_)                                                             \
            throw Error(SQLP_E_NCCL, std::string(#call) + ": " + g_nccl.GetErrorString(r_)); \
    } while (0)

// ---------------------------------------------------------------- device buffers -------
// Guard zones (SQLP_GUARD=1 in the environment, read once): compute-sanitizer is closed on this GPU pool, so an
// out-of-bounds WRITE has to be caught by the library itself.  With guards on, every device buffer is allocated
// with 512 bytes of 0xA5 in front of it and behind it; sqlp_guard_check() reads all of them back and counts the
// zones that no longer hold the pattern (tests/test_gpu_guards.py runs the smallest case of every kernel under it).
constexpr size_t SQLP_GUARD_BYTES = 512;
inline bool guards_on()
{
    static const int on = [] { const char *g = getenv("SQLP_GUARD"); return g && atoi(g) != 0 ? 1 : 0; }();
    return on != 0;
}
struct DevBuf;
inline std::vector<DevBuf *> &live_bufs()
{
    static std::vector<DevBuf *> v;
    return v;
}

struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
    bool guarded = false;
    void *base() const { return guarded ? (char *)p - SQLP_GUARD_BYTES : p; }
    ~DevBuf()
    {
        if (p) cudaFree(base());
        auto &v = live_bufs();
        v.erase(std::remove(v.begin(), v.end(), this), v.end());
    }
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    // Grow to at least `need` bytes.  keep = bytes of existing content to preserve;
    // the remainder is zero-filled.  All work is enqueued on `st`.
    void ensure(size_t need, size_t keep, cudaStream_t st, bool zero = true)
    {
        if (need <= bytes) return;
        size_t nb = std::max(need, bytes + bytes / 2);
        // stream-ordered allocation from the device's memory pool: growing a buffer in the middle of a run
        // (N and K grow every SD iteration) neither synchronises the host with the device nor stalls the
        // device the way cudaMalloc / cudaFree do; the old block goes back to the pool once the work queued
        // before this point has finished with it
        const bool g = guards_on();
        const size_t pad = g ? SQLP_GUARD_BYTES : 0;
        nb = (nb + 15) / 16 * 16;
        void *nbase = nullptr;
        cudaError_t e = cudaMallocAsync(&nbase, nb + 2 * pad, st);
        if (e != cudaSuccess)
            throw Error(SQLP_E_NOMEM, std::string("cudaMallocAsync(") + std::to_string(nb) +
                                          "): " + cudaGetErrorString(e));
        void *np = (char *)nbase + pad;
        if (g) {
            CK(cudaMemsetAsync(nbase, 0xA5, pad, st));
            CK(cudaMemsetAsync((char *)np + nb, 0xA5, pad, st));
        }
        if (keep) CK(cudaMemcpyAsync(np, p, keep, cudaMemcpyDeviceToDevice, st));
        if (zero && nb > keep) CK(cudaMemsetAsync((char *)np + keep, 0, nb - keep, st));
        if (p) CK(cudaFreeAsync(base(), st));
        else live_bufs().push_back(this);
        p = np;
        bytes = nb;
        guarded = g;
    }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

struct PinnedBuf {   // page-locked host memory: device-to-host copies into it are truly asynchronous
    void *p = nullptr;
    size_t bytes = 0;
    ~PinnedBuf() { if (p) cudaFreeHost(p); }
    PinnedBuf() = default;
    PinnedBuf(const PinnedBuf &) = delete;
    PinnedBuf &operator=(const PinnedBuf &) = delete;
    void ensure(size_t need)   // contents are not preserved; the caller has drained the stream
    {
        if (need <= bytes) return;
        if (p) cudaFreeHost(p);
        p = nullptr;
        bytes = 0;
        cudaError_t e = cudaHostAlloc(&p, need, cudaHostAllocDefault);
        if (e != cudaSuccess)
            throw Error(SQLP_E_NOMEM, std::string("cudaHostAlloc(") + std::to_string(need) + "): " + cudaGetErrorString(e));
        bytes = need;
    }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

template <class T>
void upload(DevBuf &b, const std::vector<T> &v, cudaStream_t st)
{
    b.ensure(std::max<size_t>(v.size(), 1) * sizeof(T), 0, st);
    if (!v.empty())
        CK(cudaMemcpyAsync(b.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, st));
}

}  // namespace

// ---------------------------------------------------------------- handles --------------
struct sqlp_ctx {
    int device = 0;
    int rank = 0, world = 1;
    // one host thread driving several GPUs (sqlp_ctx_create_multi): the leader (rank 0) lists the contexts of
    // ranks 1 .. world-1; every handle created on the leader carries its shards on them the same way
    std::vector<sqlp_ctx *> peers;
    bool one_process = false;
    int sm_count = 148;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    ncclComm_t comm = nullptr;
    int64_t launches = 0;
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    unsigned profile = 0;         // bit cls: event scopes around the launches of kernel class cls
    bool pdl = true;              // programmatic dependent launch of every kernel (SQLP_PDL=0: plain stream order)
    bool twins = true;            // leave score-equivalent vertices out of the sweep (SQLP_TWINS=0: every vertex)
    struct ProfEvent { cudaEvent_t e0, e1; int cls; double per_k; int k_quantum; long long *k_seen; };
    std::vector<long long *> prof_kslots;   // pinned blocks of 1024 pool sizes, one slot per event scope
    std::vector<ProfEvent> prof_events;
    size_t prof_used = 0;
    double prof_work[SQLP_PROF_CLASSES] = {0, 0, 0, 0, 0, 0, 0, 0};   // flops (contraction, screening) or algorithmic bytes
    bool smem_attr[3] = {false, false, false};
    // contraction plan: which kernel, forced grid (tests), piece buffers of the even split
    int contract_mode = 0;        // 0 = automatic, 1 = streaming only, 2 = resident (or streaming), 3 = warp-specialised first
    int contract_lag_ns = 2000;
    int ws_smem_set[3] = {0, 0, 0};
    int contract_grid = 0;        // > 0: force this many CTAs (tests of the span split)
    int contract_prefetch = 0;    // > 0: items the copies run ahead (tuning knob, environment)
    int smem_per_sm = 0, smem_optin = 0;
    bool delta_smem_set = false;
    int res_smem_set[3] = {0, 0, 0};
    DevBuf d_piece_val, d_piece_idx;
    // screening pass (kernels_screen.cuh): 0 = off, 1 = automatic (default), 2 = whenever the shape allows it
    int screen_mode = 1;
    bool screen_smem_set[3] = {false, false, false};
    bool screen_fadd2 = false;    // SQLP_FADD2=1: the epilogue adds two scores per instruction (add.f32x2).  Measured on one box,
                                  // three rounds each: +2 % on the round-1 pool (fast path), -2.5 % on storm's real pool (166 registers)
    bool screen_seed = true;      // start the scan from the previous winners' scores (SQLP_SEED=0: from -Inf)
    bool screen_centre = true;    // bf16 operands relative to the centre of the pool / of the scenarios (SQLP_CENTRE=0: raw)
    bool hist_fx = true;          // per-vertex weight sums in fixed point (SQLP_HIST=float: the ordered FP64 sums)
    bool hist_fx_set[3] = {false, false, false};
    DevBuf d_hfx;
    bool resolve_rows = true;     // exact decision moves whole rows through shared memory (SQLP_RESOLVE=lanes: a lane per row)
    size_t decide_smem_set[3] = {0, 0, 0};
    bool resolve_fma = false;     // exact decision by DFMA lanes (set when the device check DMMA == DFMA chain passed)
    int reduce_mode = 0;          // cut reduction: 0 automatic, 1 per-scenario gather only, 2 per-vertex weight sums whenever possible
    bool hist_smem_set[3] = {false, false, false};
    DevBuf d_lseed;               // [NX][npad] warm-start bounds of the running pass
    DevBuf d_cand, d_cnt, d_lfin; // candidate lists of the running pass (shared by the context's epigraphs: stream order)
    DevBuf d_cell, d_cellg;       // sharded job: the rows of a call's epigraphs back to back, and their all-gather
    DevBuf d_step;                // sqlp_cell_sd_step: the step's scenario values on the device
    PinnedBuf h_step;             //                    its results on the way back
    void bind() const { CK(cudaSetDevice(device)); }
};

// Every kernel goes through here.  With ctx->pdl (default; SQLP_PDL=0 turns it off) the launch carries
// the programmatic-stream-serialization attribute: the kernel may be scheduled before its predecessor
// has drained and synchronises with it by griddep_sync() (common.cuh), its first statement.
template <class... KArgs, class... Args>
void launch_kernel(sqlp_ctx *c, const char *name, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                   Args &&...args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = c->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = c->pdl ? 1 : 0;
    cudaError_t err = cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
    if (err != cudaSuccess)
        throw Error(SQLP_E_CUDA, std::string("launch of ") + name + " (grid " + std::to_string(grid.x) + "x" +
                                     std::to_string(grid.y) + ", block " + std::to_string(block.x) + ", " +
                                     std::to_string(smem) + " B of dynamic shared memory): " + cudaGetErrorString(err));
    ++c->launches;
}
#define LAUNCH(ctx, kernel, grid, block, smem, ...) \
    launch_kernel((ctx), #kernel, kernel, dim3(grid), dim3(block), (size_t)(smem), __VA_ARGS__)

struct PoolView {   // the pool restricted to one set of stochastic rows, in tile layout
    std::vector<int> rows;
    int n_rows = 0, s_pad = 0;
    DevBuf d_rows, d_piS;
    DevBuf d_piR;            // the same view row-major [column][s_pad]: what the exact decision gathers (one contiguous row per candidate)
    int64_t synced_lo = 0;   // vertices < synced_lo are final in d_piS
    // score-equivalent vertices (kernels_pool.cuh, "twins"): with rows that can never enter a score the view holds
    // one column per CLASS of vertices equal on the relevant rows; d_act[v] = pool slot of the first vertex of class v
    std::vector<int> rel;    // relevant rows: stochastic rows + rows of rbar != 0 + rows of Tbar entries
    bool twins = false;      // rel is a proper subset of the rows: classification is on
    DevBuf d_rel, d_twin, d_act, d_hk, d_tkey, d_trep, d_tflag;
    int64_t twin_cap = 0;    // pool capacity the tables are sized for
    int64_t twin_epoch = -1, scr_epoch = -1;   // pool push epochs the classification / the bf16 operands are up to
    unsigned int tmask = 0;
    const long long *d_Kv(const sqlp_pool *p) const;   // columns of the view (device resident)
    const int *act() const { return twins ? d_act.as<int>() : nullptr; }
    // the same rows as bf16 hi / lo operands of the screening pass, built on first use
    int sp = 0;              // row slots padded to a multiple of 16
    DevBuf d_piB, d_pn, d_pnmax, d_vbad, d_scr_lo;
    DevBuf d_ctr;                             // centre the bf16 operands are taken relative to (sp values + its norm)
    int64_t ctr_cols = 0;                     // columns it was the mean of (0: none yet)
    int64_t ctr_epoch = 0;                    // counts the re-centrings (an epigraph's ctr . d_i follow it)
    int64_t scr_synced_lo = 0, scr_cap = 0;   // vertices final in d_piB / capacity (multiple of 256)
};

struct sqlp_pool {
    sqlp_ctx *ctx = nullptr;
    int64_t m2 = 0;
    int64_t cap = 0;        // vertex capacity (multiple of 128)
    int64_t K = 0;          // confirmed size
    int64_t pending = 0;    // enqueued pushes whose outcome the host has not read yet
    int64_t push_epoch = 0; // bumped by every enqueued push: views compare it with the epoch they last synced at
    DevBuf d_pi, d_hash, d_K, d_scratch, d_vnew, d_vr, d_results;
    std::vector<PoolView *> views;
    std::vector<sqlp_epi *> epis;
    std::vector<sqlp_pool *> peers;   // multi-GPU context: the replicas on ranks 1 .. world-1
    int64_t upper() const { return K + pending; }
};

inline const long long *PoolView::d_Kv(const sqlp_pool *p) const
{
    return twins ? &d_twin.as<TwinState>()->Kv : p->d_K.as<long long>();
}

struct sqlp_epi {
    sqlp_ctx *ctx = nullptr;
    sqlp_pool *pool = nullptr;
    PoolView *view = nullptr;
    std::vector<sqlp_epi *> peers;    // multi-GPU context: the shards on ranks 1 .. world-1
    int64_t m2 = 0, n1 = 0, s = 0;
    int n_T = 0;
    // template coefficients
    std::vector<double> h_rbar;
    std::vector<int64_t> h_colptr;
    std::vector<int> h_rowval;
    std::vector<double> h_nzval;
    std::vector<int> h_pos_row, h_pos_col, h_elem_j, h_elem_t;
    DevBuf d_rbar, d_colptr, d_rowval, d_nzval;
    DevBuf d_rptr, d_rcol, d_rval;        // CSR copy of Tbar (columns ascending per row) -> k_base
    DevBuf d_ridx, d_rnz;                 // non-zeros of rbar, index order -> k_epi_tables
    int r_nnz = 0;
    DevBuf d_slot_elem, d_t_elem, d_elem_base;
    DevBuf d_tj, d_tcol, d_tslot;         // T elements sorted by (row slot, col)  -> k_delta_x
    DevBuf d_cc, d_cj, d_cslot;           // T elements sorted by col              -> reduce
    DevBuf d_mcol, d_mrow, d_mslot;       // T elements sorted by (col, row)       -> eval_dual
    DevBuf d_ovals, d_ocdf, d_ocnt;       // outcome tables
    int mo = 0;
    DevBuf d_kind, d_par_a, d_par_b;      // continuous elements (NORMAL / UNIFORM)
    bool has_kinds = false, all_continuous = false;
    // scenario store
    int64_t n_global = 0, n_local = 0, cap_tiles = 0;
    double total_weight = 0.0;
    double w_absmax = 0.0;                    // largest |weight| seen (k_cut_hist_fx scales by it); NaN once a weight was NaN
    DevBuf d_D, d_dT, d_w, d_Dx;
    DevBuf d_DR;                              // d_D row-major [scenario][s_pad] (epigraphs without random T entries)
    // per-vertex tables (rho, tau)
    DevBuf d_rt;
    int64_t rt_cap = 0, rt_synced_lo = 0;
    // work buffers
    DevBuf d_cpart, d_spart;              // k_cut_hist: per-block weight sums per view column, and scalar sums
    DevBuf d_x2, d_base, d_bias, d_best_val, d_best_idx, d_partial, d_partial2, d_out, d_gather,
        d_flags, d_stage, d_scratch;
    int64_t bias_stride = 0, out_stride = 0;
    double *cur_bias = nullptr;    // bias vectors of the running call: this epigraph's, or those of a twin (same
    int64_t cur_bias_stride = 0;   //   template, pool and points) earlier in the same call
    int tmpl_id = 0;               // epigraphs of a pool with equal (rbar, Tbar) share an id
    // the cut list on the device (kernels_cuts.cuh): rows (alpha, beta[n1], weight_mark)
    double objective_weight = 1.0, lower_bound = 0.0;
    DevBuf d_cuts, d_cuts_tmp, d_inc, d_prev_inc, d_keep, d_eval, d_rows;
    int64_t n_cuts = 0, cuts_cap = 0, n_last = 0;
    bool has_inc = false, has_prev_inc = false;
    int last_nx = 0;   // points of the last cut formation whose result is still in d_out
    // screening pass: bf16 scenario operands, per-call control block, what the host has learnt
    DevBuf d_DB, d_dnu, d_dn, d_dnall, d_ebad, d_b32c, d_ctl;
    DevBuf d_prev;                            // [n_local][2]: 1 + view column selected at the previous pass, per point
    DevBuf d_prevdot, d_cd;                   // [n_local][2] the dots P_k . d_i of those columns; [n_local] ctr . d_i
    int64_t prev_cap = 0, cd_synced = 0, cd_epoch = -1;
    bool prev_valid = false;
    DevBuf d_dbar, d_pdb;                     // centre of the scenarios (sp values + norm); P_k . dbar per view column
    int64_t dbar_n = 0;                       // scenarios dbar was the mean of
    int64_t scr_synced = 0, scr_units_cap = 0;
    PinnedBuf h_ctl;               // the control block of the last pass, copied back without a synchronisation
    cudaEvent_t ctl_event = nullptr;
    bool ctl_pending = false;
    int scr_skip = 0, scr_backoff = 0;   // calls to leave the pass out after it fell back (doubles up to 1024)
    int64_t scr_runs = 0, scr_fallbacks = 0, scr_unprofitable = 0;
    int scr_nx = 0;                // points of the pass whose control block is on its way back
    ScreenCtl scr_last = {};       // statistics of the last pass the host has seen
};

namespace {

cudaStream_t S(sqlp_ctx *c) { return c->stream; }

struct ProfScope {   // CUDA events around the launch(es) of one kernel class when profiling is on
    sqlp_ctx *c;
    cudaEvent_t e1 = nullptr;
    // work: flops or algorithmic bytes of the scope.  With `pool` the work is PER VERTEX and is multiplied, when the
    // profile is read, by the pool size the launch actually saw (rounded up to `quantum`): the size lives on the
    // device (pushes are enqueued, not awaited), so it is copied into a pinned slot right behind the launch.
    ProfScope(sqlp_ctx *c_, int cls, double work, sqlp_pool *pool = nullptr, int quantum = 1, PoolView *view = nullptr)
        : c(c_)
    {
        kptr_ = pool ? (view ? view->d_Kv(pool) : pool->d_K.as<long long>()) : nullptr;
        if (!((c->profile >> cls) & 1u)) return;
        if (c->prof_used == c->prof_events.size()) {
            cudaEvent_t a = nullptr, b = nullptr;
            CK(cudaEventCreate(&a));
            CK(cudaEventCreate(&b));
            const size_t idx = c->prof_events.size();
            if (idx / 1024 == c->prof_kslots.size()) {
                long long *blk = nullptr;
                CK(cudaHostAlloc((void **)&blk, 1024 * sizeof(long long), cudaHostAllocDefault));
                c->prof_kslots.push_back(blk);
            }
            c->prof_events.push_back({a, b, cls, 0.0, 1, c->prof_kslots[idx / 1024] + idx % 1024});
        }
        sqlp_ctx::ProfEvent &pe = c->prof_events[c->prof_used++];
        pe.cls = cls;
        pe.per_k = pool ? work : 0.0;
        pe.k_quantum = quantum;
        pool_ = pool;
        idx_ = c->prof_used - 1;
        e1 = pe.e1;
        if (!pool) c->prof_work[cls] += work;
        CK(cudaEventRecord(pe.e0, c->stream));
    }
    sqlp_pool *pool_ = nullptr;
    const long long *kptr_ = nullptr;     // the size the launch sees: view columns (twins) or pool vertices
    size_t idx_ = 0;
    void stop();
    ~ProfScope();
};



inline void ProfScope::stop()
{
    if (!e1) return;
    CK(cudaEventRecord(e1, c->stream));
    if (pool_) CK(cudaMemcpyAsync(c->prof_events[idx_].k_seen, kptr_, 8, cudaMemcpyDeviceToHost, c->stream));
    e1 = nullptr;
}
inline ProfScope::~ProfScope()
{
    if (!e1) return;
    cudaEventRecord(e1, c->stream);
    if (pool_) cudaMemcpyAsync(c->prof_events[idx_].k_seen, kptr_, 8, cudaMemcpyDeviceToHost, c->stream);
}

int64_t round_up(int64_t a, int64_t b) { return (a + b - 1) / b * b; }

}  // namespace

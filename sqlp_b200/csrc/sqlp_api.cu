// sqlp_api.cu -- host side of libsqlp_b200.so: handles, memory, launches, NCCL plumbing.
// The C ABI is declared in include/sqlp_b200.h; kernels live in the kernels_*.cuh files.
// No CPU fallback: every compute entry point needs a CUDA device and fails otherwise.
#include <dlfcn.h>

#include <algorithm>
#include <array>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/sqlp_b200.h"
#include "common.cuh"
// contraction variant (overridable with -D for experiments)
#ifndef SQLP_VARIANT_MI
#define SQLP_VARIANT_MI 8
#define SQLP_VARIANT_STAGES 4
#define SQLP_VARIANT_PREFETCH 2
#define SQLP_VARIANT_CTAS 2
#endif
#ifndef SQLP_VARIANT_KG
#define SQLP_VARIANT_KG 4
#endif
// resident variant (production): warp rows, m8n8 blocks per warp, k-groups per item, CTAs per SM
#ifndef SQLP_RES_WR
#define SQLP_RES_WR 2
#define SQLP_RES_MI 8
#define SQLP_RES_KG 6
#define SQLP_RES_CTAS 1
#endif
#include "kernels_contract.cuh"
#include "kernels_contract_res.cuh"
#ifndef SQLP_WS_KG
#define SQLP_WS_KG 6
#endif
#include "kernels_contract_ws.cuh"
#include "kernels_delta.cuh"
#include "kernels_pool.cuh"
#include "kernels_reduce.cuh"
#include "kernels_cuts.cuh"
#include "kernels_screen.cuh"

#include "host_base.cuh"
#include "host_pool.cuh"
#include "host_epi.cuh"
#include "host_smps.cuh"

// ================================================================ C ABI =================
extern "C" {

#ifndef SQLP_BUILD_ID
#define SQLP_BUILD_ID "unidentified..."
#endif
// "sqlp-build-id:<hash of the sources>" is searched for in the binary by sqlp_b200/_lib.py: a stale .so is refused
const char *sqlp_version(void) { return "sqlp_b200 0.2 (sm_100a) sqlp-build-id:" SQLP_BUILD_ID; }
const char *sqlp_last_error(void) { return g_err.c_str(); }

static void ctx_init(sqlp_ctx *c, int32_t device)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        throw Error(SQLP_E_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e) +
                                     " (libsqlp_b200 has no CPU fallback)");
    REQUIRE(device >= 0 && device < n, SQLP_E_INVALID, "device ordinal out of range");
    c->device = device;
    c->bind();
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
    c->smem_per_sm = (int)prop.sharedMemPerMultiprocessor;
    c->smem_optin = (int)prop.sharedMemPerBlockOptin;
    if (const char *m = getenv("SQLP_CONTRACT")) {   // experiments / tests of the fallback
        if (!strcmp(m, "stream")) c->contract_mode = 1;
        else if (!strcmp(m, "resident")) c->contract_mode = 2;
        else if (!strcmp(m, "ws")) c->contract_mode = 3;
    }
    if (const char *g = getenv("SQLP_PDL")) c->pdl = atoi(g) != 0;
    if (const char *g = getenv("SQLP_SCREEN")) c->screen_mode = std::max(0, std::min(2, atoi(g)));
    if (const char *g = getenv("SQLP_CENTRE")) c->screen_centre = atoi(g) != 0;
    if (const char *g = getenv("SQLP_SEED")) c->screen_seed = atoi(g) != 0;
    if (const char *g = getenv("SQLP_FADD2")) c->screen_fadd2 = atoi(g) != 0;
    if (const char *g = getenv("SQLP_HIST")) c->hist_fx = strcmp(g, "float") != 0;
    if (const char *g = getenv("SQLP_TWINS")) c->twins = atoi(g) != 0;
    if (const char *g = getenv("SQLP_REDUCE")) c->reduce_mode = std::max(0, std::min(2, atoi(g)));
    if (const char *g = getenv("SQLP_CONTRACT_GRID")) c->contract_grid = atoi(g);
    if (const char *g = getenv("SQLP_CONTRACT_PREFETCH")) c->contract_prefetch = atoi(g);
    if (const char *g = getenv("SQLP_CONTRACT_LAG_NS")) c->contract_lag_ns = atoi(g);
    {   // device buffers come from the default memory pool (DevBuf::ensure); keep freed blocks in it
        cudaMemPool_t mp = nullptr;
        CK(cudaDeviceGetDefaultMemPool(&mp, device));
        unsigned long long keep_all = ~0ull;
        CK(cudaMemPoolSetAttribute(mp, cudaMemPoolAttrReleaseThreshold, &keep_all));
    }
    CK(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    c->stream = c->own_stream;
    CK(cudaEventCreate(&c->t0));
    CK(cudaEventCreate(&c->t1));
    {   // may the exact decision score candidates with plain DFMA chains?  Only if, on THIS device, the DMMA of the
        // FP64 sweep is that chain bit for bit (it is on sm_100a; checked, not assumed).  SQLP_RESOLVE=dmma keeps DMMA.
        const char *g = getenv("SQLP_RESOLVE");
        if (g && !strcmp(g, "lanes")) c->resolve_rows = false;
        if (!g || strcmp(g, "dmma")) {
            unsigned int *d_bad = nullptr, h_bad = 1;
            CK(cudaMalloc(&d_bad, 4));
            CK(cudaMemsetAsync(d_bad, 0, 4, c->stream));
            k_dmma_is_fma_chain<<<2048 / 8, 256, 0, c->stream>>>(32, 2048, d_bad);
            CK(cudaMemcpyAsync(&h_bad, d_bad, 4, cudaMemcpyDeviceToHost, c->stream));
            CK(cudaStreamSynchronize(c->stream));
            CK(cudaFree(d_bad));
            c->resolve_fma = (h_bad == 0);
        }
    }
}

int32_t sqlp_guard_check(int64_t *buffers, int64_t *damaged)
{
    return guard([&] {
        REQUIRE(buffers && damaged, SQLP_E_INVALID, "null argument");
        *buffers = 0;
        *damaged = 0;
        CK(cudaDeviceSynchronize());
        std::vector<unsigned char> h(SQLP_GUARD_BYTES);
        for (DevBuf *b : live_bufs()) {
            if (!b->p || !b->guarded) continue;
            ++*buffers;
            for (int side = 0; side < 2; ++side) {
                const void *src = side == 0 ? b->base() : (const void *)((const char *)b->p + b->bytes);
                cudaPointerAttributes at;
                if (cudaPointerGetAttributes(&at, src) != cudaSuccess) { cudaGetLastError(); continue; }
                int cur = 0;
                CK(cudaGetDevice(&cur));
                if (at.device != cur) CK(cudaSetDevice(at.device));
                CK(cudaMemcpy(h.data(), src, SQLP_GUARD_BYTES, cudaMemcpyDeviceToHost));
                if (at.device != cur) CK(cudaSetDevice(cur));
                for (unsigned char c : h)
                    if (c != 0xA5) { ++*damaged; break; }
            }
        }
    });
}

int32_t sqlp_ctx_create(int32_t device, sqlp_ctx **out)
{
    return guard([&] {
        REQUIRE(out, SQLP_E_INVALID, "null out");
        sqlp_ctx *c = new sqlp_ctx();
        try { ctx_init(c, device); } catch (...) { delete c; throw; }
        *out = c;
    });
}

int32_t sqlp_nccl_unique_id(void *out128)
{
    return guard([&] {
        REQUIRE(out128, SQLP_E_INVALID, "null out");
        load_nccl();
        ncclUniqueId id;
        NK(g_nccl.GetUniqueId(&id));
        memcpy(out128, &id, sizeof id);
    });
}

int32_t sqlp_ctx_create_dist(int32_t device, int32_t rank, int32_t world, const void *nccl_id,
                             sqlp_ctx **out)
{
    return guard([&] {
        REQUIRE(out && nccl_id, SQLP_E_INVALID, "null argument");
        REQUIRE(world >= 1 && rank >= 0 && rank < world, SQLP_E_INVALID, "bad rank/world");
        sqlp_ctx *c = new sqlp_ctx();
        try {
            ctx_init(c, device);
            c->rank = rank;
            c->world = world;
            if (world > 1) {
                load_nccl();
                ncclUniqueId id;
                memcpy(&id, nccl_id, sizeof id);
                NK(g_nccl.CommInitRank(&c->comm, world, id, rank));
            }
        } catch (...) { delete c; throw; }
        *out = c;
    });
}

int32_t sqlp_ctx_create_multi(int32_t n_gpus, const int32_t *devices, sqlp_ctx **out)
{
    return guard([&] {
        REQUIRE(out && n_gpus >= 1 && n_gpus <= 64, SQLP_E_INVALID, "bad argument");
        std::vector<int> dev((size_t)n_gpus);
        for (int i = 0; i < n_gpus; ++i) {
            dev[(size_t)i] = devices ? devices[i] : i;
            for (int j = 0; j < i; ++j) REQUIRE(dev[(size_t)j] != dev[(size_t)i], SQLP_E_INVALID, "device listed twice");
        }
        std::vector<sqlp_ctx *> cs;
        try {
            for (int i = 0; i < n_gpus; ++i) {
                cs.push_back(new sqlp_ctx());
                ctx_init(cs.back(), dev[(size_t)i]);
                cs.back()->rank = i;
                cs.back()->world = n_gpus;
                cs.back()->one_process = true;
            }
            if (n_gpus > 1) {
                load_nccl();
                std::vector<ncclComm_t> comms((size_t)n_gpus);
                NK(g_nccl.CommInitAll(comms.data(), n_gpus, dev.data()));
                for (int i = 0; i < n_gpus; ++i) cs[(size_t)i]->comm = comms[(size_t)i];
            }
        } catch (...) {
            for (sqlp_ctx *c : cs) delete c;
            throw;
        }
        cs[0]->peers.assign(cs.begin() + 1, cs.end());
        cs[0]->bind();
        *out = cs[0];
    });
}

int32_t sqlp_ctx_destroy(sqlp_ctx *c)
{
    if (c && !c->peers.empty()) {
        std::vector<sqlp_ctx *> peers;
        peers.swap(c->peers);
        for (sqlp_ctx *q : peers) sqlp_ctx_destroy(q);
    }
    return guard([&] {
        if (!c) return;
        cudaSetDevice(c->device);
        cudaStreamSynchronize(c->stream);
        if (c->comm) g_nccl.CommDestroy(c->comm);
        for (auto &pr : c->prof_events) { cudaEventDestroy(pr.e0); cudaEventDestroy(pr.e1); }
        for (long long *blk : c->prof_kslots) cudaFreeHost(blk);
        if (c->t0) cudaEventDestroy(c->t0);
        if (c->t1) cudaEventDestroy(c->t1);
        if (c->own_stream) cudaStreamDestroy(c->own_stream);
        delete c;
    });
}

int32_t sqlp_ctx_set_stream(sqlp_ctx *c, void *stream)
{
    return guard([&] {
        REQUIRE(c, SQLP_E_INVALID, "null ctx");
        c->bind();
        CK(cudaStreamSynchronize(c->stream));
        c->stream = stream ? (cudaStream_t)stream : c->own_stream;
    });
}

int32_t sqlp_ctx_synchronize(sqlp_ctx *c)
{
    return guard([&] {
        REQUIRE(c, SQLP_E_INVALID, "null ctx");
        for (int s = n_shards(c) - 1; s >= 0; --s) {
            shard(c, s)->bind();
            CK(cudaStreamSynchronize(shard(c, s)->stream));
        }
    });
}

int32_t sqlp_ctx_launch_count(sqlp_ctx *c, int64_t *n)
{
    return guard([&] {
        REQUIRE(c && n, SQLP_E_INVALID, "null argument");
        *n = c->launches;
    });
}

int32_t sqlp_ctx_timer_start(sqlp_ctx *c)
{
    return guard([&] { REQUIRE(c, SQLP_E_INVALID, "null ctx"); c->bind(); CK(cudaEventRecord(c->t0, c->stream)); });
}
int32_t sqlp_ctx_timer_stop(sqlp_ctx *c)
{
    return guard([&] { REQUIRE(c, SQLP_E_INVALID, "null ctx"); c->bind(); CK(cudaEventRecord(c->t1, c->stream)); });
}
int32_t sqlp_ctx_timer_elapsed_ms(sqlp_ctx *c, double *ms)
{
    return guard([&] {
        REQUIRE(c && ms, SQLP_E_INVALID, "null argument");
        c->bind();
        CK(cudaEventSynchronize(c->t1));
        float f = 0;
        CK(cudaEventElapsedTime(&f, c->t0, c->t1));
        *ms = f;
    });
}

int32_t sqlp_ctx_profile(sqlp_ctx *c, int32_t enable)
{
    return guard([&] {
        REQUIRE(c, SQLP_E_INVALID, "null ctx");
        c->profile = enable == 1 ? ~0u : (unsigned)enable >> 1;
        // the contraction's mask bit also covers what can stand in for it: screening, exact decision, gated sweep
        if ((c->profile >> SQLP_PROF_CONTRACT) & 1u)
            c->profile |= (1u << SQLP_PROF_SCREEN) | (1u << SQLP_PROF_RESOLVE) | (1u << SQLP_PROF_FALLBACK);
    });
}

int32_t sqlp_ctx_set_screen(sqlp_ctx *c, int32_t mode)
{
    return guard([&] {
        REQUIRE(c, SQLP_E_INVALID, "null ctx");
        REQUIRE(mode >= 0 && mode <= 2, SQLP_E_INVALID, "screen mode must be 0 (off), 1 (automatic) or 2 (always)");
        for (int s = 0; s < n_shards(c); ++s) shard(c, s)->screen_mode = mode;
    });
}

int32_t sqlp_epi_screen_stats(sqlp_epi *e, int64_t *out /*[8]*/)
{
    return guard([&] {
        REQUIRE(e && out, SQLP_E_INVALID, "null argument");
        e->ctx->bind();
        CK(cudaStreamSynchronize(S(e->ctx)));
        screen_learn(e);
        const ScreenCtl &h = e->scr_last;
        out[0] = e->scr_runs;
        out[1] = e->scr_fallbacks;
        out[2] = (int64_t)h.n_emit;
        out[3] = (int64_t)h.n_eval;
        out[4] = (int64_t)h.overflow;
        out[5] = h.bad + 2 * e->scr_unprofitable;   // bit 0: bad operands; the rest: passes judged unprofitable
        out[6] = h.live[0];
        out[7] = h.live[1];
    });
}

int32_t sqlp_ctx_profile_classes(sqlp_ctx *c, int32_t reset, double *ms, int64_t *launches, double *work)
{
    return guard([&] {
        REQUIRE(c, SQLP_E_INVALID, "null ctx");
        c->bind();
        CK(cudaStreamSynchronize(c->stream));
        double t[SQLP_PROF_CLASSES] = {}, w[SQLP_PROF_CLASSES] = {};
        int64_t n[SQLP_PROF_CLASSES] = {};
        for (size_t i = 0; i < c->prof_used; ++i) {
            const sqlp_ctx::ProfEvent &pe = c->prof_events[i];
            float f = 0;
            CK(cudaEventElapsedTime(&f, pe.e0, pe.e1));
            t[pe.cls] += f;
            ++n[pe.cls];
            // work per vertex x the pool size THIS launch saw (copied back right behind it)
            if (pe.per_k > 0.0) w[pe.cls] += pe.per_k * (double)round_up(*pe.k_seen, pe.k_quantum);
        }
        for (int k = 0; k < SQLP_PROF_CLASSES; ++k) {
            if (ms) ms[k] = t[k];
            if (launches) launches[k] = n[k];
            if (work) work[k] = c->prof_work[k] + w[k];
        }
        if (reset) {
            c->prof_used = 0;
            for (double &w : c->prof_work) w = 0;
        }
    });
}

int32_t sqlp_ctx_profile_read(sqlp_ctx *c, int32_t reset, double *ms, int64_t *launches, double *flops)
{
    double t[SQLP_PROF_CLASSES], w[SQLP_PROF_CLASSES];
    int64_t n[SQLP_PROF_CLASSES];
    int32_t rc = sqlp_ctx_profile_classes(c, reset, t, n, w);
    if (rc != SQLP_OK) return rc;
    if (ms) *ms = t[SQLP_PROF_CONTRACT];
    if (launches) *launches = n[SQLP_PROF_CONTRACT];
    if (flops) *flops = w[SQLP_PROF_CONTRACT];
    return SQLP_OK;
}

// ---------------------------------------------------------------- pool -------------------
int32_t sqlp_pool_create(sqlp_ctx *c, int64_t m2, sqlp_pool **out)
{
    int32_t st = guard([&] {
        REQUIRE(c && out, SQLP_E_INVALID, "null argument");
        REQUIRE(m2 >= 1 && m2 < (1 << 24), SQLP_E_INVALID, "bad m2");
        c->bind();
        sqlp_pool *p = new sqlp_pool();
        try {
            p->ctx = c;
            p->m2 = m2;
            p->d_K.ensure(8, 0, S(c));
            p->d_scratch.ensure(sizeof(PushScratch), 0, S(c));
            PushScratch idle;                                     // what k_pool_push expects and leaves behind
            idle.hash = 0ull;
            idle.done = 0u;
            for (int i = 0; i < SQLP_PUSH_BATCH; ++i) idle.match[i] = 0x7fffffff;
            CK(cudaMemcpyAsync(p->d_scratch.p, &idle, sizeof idle, cudaMemcpyHostToDevice, S(c)));
            p->d_vr.ensure((size_t)m2 * 8, 0, S(c));
            pool_reserve(p, 1024);
            CK(cudaStreamSynchronize(S(c)));
        } catch (...) { delete p; throw; }
        *out = p;
    });
    if (st != SQLP_OK || c->peers.empty()) return st;
    for (sqlp_ctx *q : c->peers) {   // a multi-GPU context: the replicas of the pool on the other GPUs
        sqlp_pool *pp = nullptr;
        st = sqlp_pool_create(q, m2, &pp);
        if (st != SQLP_OK) return st;
        (*out)->peers.push_back(pp);
    }
    cudaSetDevice(c->device);
    return SQLP_OK;
}

int32_t sqlp_pool_destroy(sqlp_pool *p)
{
    if (p && !p->peers.empty() && p->epis.empty()) {
        std::vector<sqlp_pool *> peers;
        peers.swap(p->peers);
        for (sqlp_pool *q : peers) sqlp_pool_destroy(q);
    }
    return guard([&] {
        if (!p) return;
        // an epigraph keeps pointers into its pool (vertex rows, views): it has to go first
        REQUIRE(p->epis.empty(), SQLP_E_INVALID, "destroy the epigraphs bound to this pool before the pool");
        cudaSetDevice(p->ctx->device);
        cudaStreamSynchronize(p->ctx->stream);
        for (PoolView *v : p->views) delete v;
        delete p;
    });
}

int32_t sqlp_pool_push_batch(sqlp_pool *p, int64_t n, const double *v, int32_t *inserted, int64_t *index)
{
    return guard([&] {
        REQUIRE(p, SQLP_E_INVALID, "null pool");
        REQUIRE(n >= 0, SQLP_E_INVALID, "negative count");
        if (n == 0) return;
        sqlp_ctx *c = p->ctx;
        REQUIRE(v || (c->world > 1 && c->rank != 0), SQLP_E_INVALID, "null vector");
        for (int s = n_shards(c) - 1; s >= 1; --s) {   // multi-GPU context: the same vectors to every GPU's replica
            shard(c, s)->bind();
            pool_push_enqueue(shard(p, s), n, v, nullptr);
        }
        c->bind();
        pool_push_enqueue(p, n, v, nullptr);
        std::vector<PushResult> res((size_t)n);
        CK(cudaMemcpyAsync(res.data(), p->d_results.p, (size_t)n * sizeof(PushResult),
                           cudaMemcpyDeviceToHost, S(c)));
        pool_confirm(p);
        sync_all_shards(c, p);
        for (int64_t i = 0; i < n; ++i) {
            if (inserted) inserted[i] = res[i].inserted;
            if (index) index[i] = res[i].index;
        }
    });
}

int32_t sqlp_pool_push(sqlp_pool *p, const double *v, int32_t *inserted, int64_t *index)
{
    return sqlp_pool_push_batch(p, 1, v, inserted, index);
}

int32_t sqlp_pool_push_dev(sqlp_pool *p, int64_t n, const double *d_v)
{
    return guard([&] {
        REQUIRE(p, SQLP_E_INVALID, "null pool");
        REQUIRE(n >= 0, SQLP_E_INVALID, "negative count");
        if (n == 0) return;
        REQUIRE(d_v || (p->ctx->world > 1 && p->ctx->rank != 0), SQLP_E_INVALID, "null vector");
        REQUIRE(p->ctx->peers.empty(), SQLP_E_UNSUPPORTED, "device pointers belong to one GPU: use the host variants with a multi-GPU context");
        p->ctx->bind();
        pool_push_enqueue(p, n, nullptr, d_v);
    });
}

int32_t sqlp_pool_size(sqlp_pool *p, int64_t *K)
{
    return guard([&] {
        REQUIRE(p && K, SQLP_E_INVALID, "null argument");
        p->ctx->bind();
        pool_confirm(p);
        *K = p->K;
    });
}

int32_t sqlp_pool_get(sqlp_pool *p, int64_t index, double *out)
{
    return guard([&] {
        REQUIRE(p && out, SQLP_E_INVALID, "null argument");
        p->ctx->bind();
        pool_confirm(p);
        REQUIRE(index >= 0 && index < p->K, SQLP_E_RANGE, "vertex index out of range");
        CK(cudaMemcpyAsync(out, p->d_pi.as<double>() + index * p->m2, (size_t)p->m2 * 8,
                           cudaMemcpyDeviceToHost, S(p->ctx)));
        CK(cudaStreamSynchronize(S(p->ctx)));
    });
}

int32_t sqlp_pool_hash(sqlp_pool *p, const double *v, uint64_t *hash)
{
    return guard([&] {
        REQUIRE(p && v && hash, SQLP_E_INVALID, "null argument");
        sqlp_ctx *c = p->ctx;
        c->bind();
        p->d_vnew.ensure((size_t)p->m2 * 8, 0, S(c));
        CK(cudaMemcpyAsync(p->d_vnew.p, v, (size_t)p->m2 * 8, cudaMemcpyHostToDevice, S(c)));
        REQUIRE(p->m2 <= 6000, SQLP_E_UNSUPPORTED, "sqlp_pool_hash (debug) handles m2 <= 6000");
        LAUNCH(c, k_pool_prepare, 1, 256, (size_t)p->m2 * 8, p->d_vnew.as<double>(), (int)p->m2,
               p->d_vr.as<double>(), p->d_scratch.as<PushScratch>());
        PushScratch sc;
        CK(cudaMemcpyAsync(&sc, p->d_scratch.p, sizeof sc, cudaMemcpyDeviceToHost, S(c)));
        CK(cudaStreamSynchronize(S(c)));
        *hash = sc.hash;
    });
}

// ---------------------------------------------------------------- epigraph ---------------
int32_t sqlp_epi_create(sqlp_ctx *c, sqlp_pool *p, int64_t m2, int64_t n1, int64_t r_nnz,
                        const int64_t *r_idx, const double *r_val, const int64_t *T_colptr,
                        const int64_t *T_rowval, const double *T_nzval, int64_t s,
                        const int32_t *pos_row, const int32_t *pos_col, sqlp_epi **out)
{
    int32_t st = guard([&] {
        REQUIRE(c && p && out, SQLP_E_INVALID, "null argument");
        REQUIRE(p->ctx == c, SQLP_E_INVALID, "pool belongs to another context");
        REQUIRE(m2 == p->m2, SQLP_E_INVALID, "m2 differs from the pool's vertex length");
        REQUIRE(n1 >= 0 && s >= 0 && r_nnz >= 0, SQLP_E_INVALID, "negative size");
        REQUIRE(T_colptr || n1 == 0, SQLP_E_INVALID, "null T_colptr");
        c->bind();
        sqlp_epi *e = new sqlp_epi();
        try {
            e->ctx = c; e->pool = p; e->m2 = m2; e->n1 = n1; e->s = s;
            e->h_rbar.assign((size_t)m2, 0.0);
            for (int64_t q = 0; q < r_nnz; ++q) {
                REQUIRE(r_idx[q] >= 0 && r_idx[q] < m2, SQLP_E_RANGE, "rbar index out of range");
                e->h_rbar[(size_t)r_idx[q]] = r_val[q];
            }
            e->h_colptr.assign((size_t)n1 + 1, 0);
            for (int64_t j = 0; j <= n1 && n1 > 0; ++j) e->h_colptr[(size_t)j] = T_colptr[j];
            int64_t nnz = n1 > 0 ? T_colptr[n1] : 0;
            REQUIRE(nnz >= 0, SQLP_E_INVALID, "bad T_colptr");
            e->h_rowval.resize((size_t)nnz);
            e->h_nzval.resize((size_t)nnz);
            for (int64_t j = 0; j < n1; ++j) {
                REQUIRE(T_colptr[j] <= T_colptr[j + 1], SQLP_E_INVALID, "T_colptr not monotone");
                for (int64_t q = T_colptr[j]; q < T_colptr[j + 1]; ++q) {
                    REQUIRE(T_rowval[q] >= 0 && T_rowval[q] < m2, SQLP_E_RANGE, "T row out of range");
                    REQUIRE(q == T_colptr[j] || T_rowval[q] > T_rowval[q - 1], SQLP_E_INVALID,
                            "T rows must ascend within a column");
                    e->h_rowval[(size_t)q] = (int)T_rowval[q];
                    e->h_nzval[(size_t)q] = T_nzval[q];
                }
            }
            auto T_at = [&](int row, int col) {
                for (int64_t q = e->h_colptr[col]; q < e->h_colptr[col + 1]; ++q)
                    if (e->h_rowval[(size_t)q] == row) return e->h_nzval[(size_t)q];
                return 0.0;
            };
            // stochastic rows S = sorted distinct rows of the position table
            std::vector<int> rows;
            for (int64_t q = 0; q < s; ++q) {
                REQUIRE(pos_row[q] >= 0 && pos_row[q] < m2, SQLP_E_RANGE, "position row out of range");
                REQUIRE(pos_col[q] >= -1 && pos_col[q] < n1, SQLP_E_RANGE, "position column out of range");
                for (int64_t r = 0; r < q; ++r)
                    REQUIRE(pos_row[r] != pos_row[q] || pos_col[r] != pos_col[q], SQLP_E_INVALID,
                            "duplicate stochastic position");
                rows.push_back(pos_row[q]);
            }
            std::sort(rows.begin(), rows.end());
            rows.erase(std::unique(rows.begin(), rows.end()), rows.end());
            // relevant rows: the only rows through which a vertex enters a score or a cut coefficient
            std::vector<int> rel(rows);
            for (int64_t j = 0; j < m2; ++j)
                if (e->h_rbar[(size_t)j] != 0.0) rel.push_back((int)j);
            for (int64_t q = 0; q < nnz; ++q) rel.push_back(e->h_rowval[(size_t)q]);
            std::sort(rel.begin(), rel.end());
            rel.erase(std::unique(rel.begin(), rel.end()), rel.end());
            // share the pool view with any epigraph that has the same row sets
            for (PoolView *v : p->views)
                if (v->rows == rows && v->rel == rel) e->view = v;
            if (!e->view) {
                PoolView *v = new PoolView();
                v->rows = rows;
                v->rel = rel;
                v->twins = c->twins && (int64_t)rel.size() < m2;   // some row can never matter: classes may exist
                if (v->twins) upload(v->d_rel, rel, S(c));
                v->n_rows = (int)rows.size();
                v->s_pad = (int)std::max<int64_t>(SQLP_BK, round_up(v->n_rows, SQLP_BK));
                upload(v->d_rows, rows, S(c));
                size_t per_chunk = (size_t)v->s_pad * SQLP_TILE * 8;
                v->d_piS.ensure((size_t)(p->cap / SQLP_TILE) * per_chunk, 0, S(c));
                v->d_piR.ensure((size_t)(p->cap / SQLP_TILE) * per_chunk, 0, S(c));
                p->views.push_back(v);
                e->view = v;
            }
            std::vector<double> base((size_t)s);
            e->h_pos_row.assign(pos_row, pos_row + s);
            e->h_pos_col.assign(pos_col, pos_col + s);
            e->h_elem_j.resize((size_t)s);
            e->h_elem_t.resize((size_t)s);
            struct TE { int j, col, row, slot; };
            std::vector<TE> te;
            for (int64_t q = 0; q < s; ++q) {
                int j = (int)(std::lower_bound(rows.begin(), rows.end(), pos_row[q]) - rows.begin());
                e->h_elem_j[(size_t)q] = j;
                if (pos_col[q] < 0) {
                    e->h_elem_t[(size_t)q] = -1;
                    base[(size_t)q] = e->h_rbar[(size_t)pos_row[q]];
                } else {
                    e->h_elem_t[(size_t)q] = e->n_T;
                    base[(size_t)q] = T_at(pos_row[q], pos_col[q]);
                    te.push_back({j, pos_col[q], pos_row[q], e->n_T});
                    ++e->n_T;
                }
            }
            upload(e->d_rbar, e->h_rbar, S(c));
            {
                std::vector<int> ri;
                std::vector<double> rv;
                for (int64_t j = 0; j < m2; ++j)
                    if (e->h_rbar[(size_t)j] != 0.0) { ri.push_back((int)j); rv.push_back(e->h_rbar[(size_t)j]); }
                e->r_nnz = (int)ri.size();
                upload(e->d_ridx, ri, S(c));
                upload(e->d_rnz, rv, S(c));
            }
            upload(e->d_colptr, e->h_colptr, S(c));
            upload(e->d_rowval, e->h_rowval, S(c));
            upload(e->d_nzval, e->h_nzval, S(c));
            {   // CSR copy: walking the columns in order keeps each row's entries column-ascending
                std::vector<int> rptr((size_t)m2 + 1, 0), rcol((size_t)nnz);
                std::vector<double> rval((size_t)nnz);
                for (int64_t q = 0; q < nnz; ++q) ++rptr[(size_t)e->h_rowval[(size_t)q] + 1];
                for (int64_t j = 0; j < m2; ++j) rptr[(size_t)j + 1] += rptr[(size_t)j];
                std::vector<int> fill(rptr.begin(), rptr.end() - 1);
                for (int64_t col = 0; col < n1; ++col)
                    for (int64_t q = e->h_colptr[(size_t)col]; q < e->h_colptr[(size_t)col + 1]; ++q) {
                        int at = fill[(size_t)e->h_rowval[(size_t)q]]++;
                        rcol[(size_t)at] = (int)col;
                        rval[(size_t)at] = e->h_nzval[(size_t)q];
                    }
                upload(e->d_rptr, rptr, S(c));
                upload(e->d_rcol, rcol, S(c));
                upload(e->d_rval, rval, S(c));
            }
            std::vector<int> slot_elem(rows.size(), -1), t_elem;
            for (int64_t q = 0; q < s; ++q) {
                if (pos_col[q] < 0) slot_elem[(size_t)e->h_elem_j[(size_t)q]] = (int)q;
                else t_elem.push_back((int)q);
            }
            upload(e->d_slot_elem, slot_elem, S(c));
            upload(e->d_t_elem, t_elem, S(c));
            upload(e->d_elem_base, base, S(c));
            auto up3 = [&](std::vector<TE> v, DevBuf &a, DevBuf &b, DevBuf &d, int key) {
                std::sort(v.begin(), v.end(), [&](const TE &x, const TE &y) {
                    if (key == 0) return x.j != y.j ? x.j < y.j : x.col < y.col;      // (row slot, col)
                    if (key == 1) return x.col != y.col ? x.col < y.col : x.row < y.row;  // (col, row)
                    return x.col != y.col ? x.col < y.col : x.row < y.row;
                });
                std::vector<int> A, B, D;
                for (auto &t : v) {
                    A.push_back(key == 0 ? t.j : t.col);
                    B.push_back(key == 0 ? t.col : (key == 1 ? t.j : t.row));
                    D.push_back(t.slot);
                }
                upload(a, A, S(c)); upload(b, B, S(c)); upload(d, D, S(c));
            };
            up3(te, e->d_tj, e->d_tcol, e->d_tslot, 0);
            up3(te, e->d_cc, e->d_cj, e->d_cslot, 1);
            up3(te, e->d_mcol, e->d_mrow, e->d_mslot, 2);
            e->d_rt.ensure((size_t)p->cap * (n1 + 1) * 8, 0, S(c));
            e->rt_cap = p->cap;
            e->d_scratch.ensure((size_t)(2 * m2 + 2 + rows.size() + 1) * 8, 0, S(c));
            CK(cudaStreamSynchronize(S(c)));
            e->tmpl_id = (int)p->epis.size() + 1;
            for (sqlp_epi *o : p->epis)
                if (o->n1 == e->n1 && o->h_rbar == e->h_rbar && o->h_colptr == e->h_colptr && o->h_rowval == e->h_rowval &&
                    o->h_nzval == e->h_nzval) { e->tmpl_id = o->tmpl_id; break; }
            p->epis.push_back(e);
        } catch (...) { delete e; throw; }
        *out = e;
    });
    if (st != SQLP_OK || c->peers.empty()) return st;
    for (size_t q = 0; q < c->peers.size(); ++q) {   // a multi-GPU context: this epigraph's shards on the other GPUs
        sqlp_epi *pe = nullptr;
        st = sqlp_epi_create(c->peers[q], p->peers[q], m2, n1, r_nnz, r_idx, r_val, T_colptr, T_rowval, T_nzval, s,
                             pos_row, pos_col, &pe);
        if (st != SQLP_OK) return st;
        (*out)->peers.push_back(pe);
    }
    cudaSetDevice(c->device);
    return SQLP_OK;
}

int32_t sqlp_epi_destroy(sqlp_epi *e)
{
    if (e && !e->peers.empty()) {
        std::vector<sqlp_epi *> peers;
        peers.swap(e->peers);
        for (sqlp_epi *q : peers) sqlp_epi_destroy(q);
    }
    return guard([&] {
        if (!e) return;
        cudaSetDevice(e->ctx->device);
        cudaStreamSynchronize(e->ctx->stream);
        auto &v = e->pool->epis;
        v.erase(std::remove(v.begin(), v.end(), e), v.end());
        if (e->ctl_event) cudaEventDestroy(e->ctl_event);
        delete e;
    });
}

int32_t sqlp_epi_add_scenarios(sqlp_epi *e, int64_t n_new, const double *values, const double *weights)
{
    if (e) for (sqlp_epi *q : e->peers) {   // multi-GPU context: every shard
        int32_t st = sqlp_epi_add_scenarios(q, n_new, values, weights);
        if (st != SQLP_OK) return st;
    }
    return guard([&] {
        REQUIRE(e, SQLP_E_INVALID, "null epigraph");
        REQUIRE(n_new >= 0, SQLP_E_INVALID, "negative count");
        REQUIRE(values || n_new == 0 || e->s == 0, SQLP_E_INVALID, "null values");
        e->ctx->bind();
        epi_add(e, n_new, values, nullptr, weights, false, 0, 0);
        CK(cudaStreamSynchronize(S(e->ctx)));
    });
}

int32_t sqlp_epi_add_scenarios_dev(sqlp_epi *e, int64_t n_new, const double *d_values,
                                   const double *weights_host)
{
    return guard([&] {
        REQUIRE(e, SQLP_E_INVALID, "null epigraph");
        REQUIRE(n_new >= 0, SQLP_E_INVALID, "negative count");
        REQUIRE(d_values || n_new == 0 || e->s == 0, SQLP_E_INVALID, "null values");
        REQUIRE(e->peers.empty(), SQLP_E_UNSUPPORTED, "device pointers belong to one GPU: use sqlp_epi_add_scenarios with a multi-GPU context");
        e->ctx->bind();
        epi_add(e, n_new, nullptr, d_values, weights_host, false, 0, 0);
    });
}

int32_t sqlp_epi_set_outcomes(sqlp_epi *e, int64_t mo, const double *vals, const double *cdf,
                              const int32_t *cnt)
{
    if (e) for (sqlp_epi *q : e->peers) {   // multi-GPU context: every shard
        int32_t st = sqlp_epi_set_outcomes(q, mo, vals, cdf, cnt);
        if (st != SQLP_OK) return st;
    }
    return guard([&] {
        REQUIRE(e && vals && cdf && cnt, SQLP_E_INVALID, "null argument");
        REQUIRE(mo >= 1, SQLP_E_INVALID, "max_outcomes < 1");
        sqlp_ctx *c = e->ctx;
        c->bind();
        for (int64_t q = 0; q < e->s; ++q)
            REQUIRE(cnt[q] >= 1 && cnt[q] <= mo, SQLP_E_INVALID, "outcome count out of range");
        std::vector<double> v(vals, vals + e->s * mo), f(cdf, cdf + e->s * mo);
        std::vector<int> n(cnt, cnt + e->s);
        // only cnt[e] entries of a row are valid: whatever pads the rest must never count as "<= u"
        for (int64_t q = 0; q < e->s; ++q)
            for (int64_t o = cnt[q]; o < mo; ++o) f[(size_t)(q * mo + o)] = INFINITY;
        upload(e->d_ovals, v, S(c));
        upload(e->d_ocdf, f, S(c));
        upload(e->d_ocnt, n, S(c));
        e->mo = (int)mo;
        CK(cudaStreamSynchronize(S(c)));
    });
}

int32_t sqlp_epi_set_distributions(sqlp_epi *e, const int32_t *kind, const double *par_a,
                                   const double *par_b)
{
    if (e) for (sqlp_epi *q : e->peers) {   // multi-GPU context: every shard
        int32_t st = sqlp_epi_set_distributions(q, kind, par_a, par_b);
        if (st != SQLP_OK) return st;
    }
    return guard([&] {
        REQUIRE(e && kind && par_a && par_b, SQLP_E_INVALID, "null argument");
        sqlp_ctx *c = e->ctx;
        c->bind();
        bool all = true;
        for (int64_t q = 0; q < e->s; ++q) {
            REQUIRE(kind[q] >= 0 && kind[q] <= 2, SQLP_E_INVALID, "distribution kind must be 0, 1 or 2");
            REQUIRE(kind[q] != 1 || par_b[q] >= 0.0, SQLP_E_INVALID, "negative variance");
            all = all && kind[q] != 0;
        }
        std::vector<int> k(kind, kind + e->s);
        std::vector<double> a(par_a, par_a + e->s), b(par_b, par_b + e->s);
        upload(e->d_kind, k, S(c));
        upload(e->d_par_a, a, S(c));
        upload(e->d_par_b, b, S(c));
        e->has_kinds = true;
        e->all_continuous = all;
        CK(cudaStreamSynchronize(S(c)));
    });
}

int32_t sqlp_epi_sample_scenarios(sqlp_epi *e, int64_t n_new, uint64_t seed, uint64_t weight_seed)
{
    if (e) for (sqlp_epi *q : e->peers) {   // multi-GPU context: every shard
        int32_t st = sqlp_epi_sample_scenarios(q, n_new, seed, weight_seed);
        if (st != SQLP_OK) return st;
    }
    return guard([&] {
        REQUIRE(e, SQLP_E_INVALID, "null epigraph");
        REQUIRE(n_new >= 0, SQLP_E_INVALID, "negative count");
        REQUIRE(e->mo > 0 || e->s == 0 || e->all_continuous, SQLP_E_INVALID,
                "call sqlp_epi_set_outcomes (and sqlp_epi_set_distributions) first");
        e->ctx->bind();
        epi_add(e, n_new, nullptr, nullptr, nullptr, true, seed, weight_seed);
    });
}

int32_t sqlp_epi_counts(sqlp_epi *e, int64_t *n_global, int64_t *n_local, double *total_weight)
{
    return guard([&] {
        REQUIRE(e, SQLP_E_INVALID, "null epigraph");
        if (n_global) *n_global = e->n_global;
        if (n_local) *n_local = e->peers.empty() ? e->n_local : e->n_global;   // one process, all GPUs: everything is local
        if (total_weight) *total_weight = e->total_weight;
    });
}

int32_t sqlp_epi_view_columns(sqlp_epi *e, int64_t *columns, int64_t *relevant_rows)
{
    return guard([&] {
        REQUIRE(e, SQLP_E_INVALID, "null epigraph");
        sqlp_ctx *c = e->ctx;
        c->bind();
        view_sync(e->pool, e->view);
        long long kv = 0;
        CK(cudaMemcpyAsync(&kv, e->view->d_Kv(e->pool), 8, cudaMemcpyDeviceToHost, S(c)));
        CK(cudaStreamSynchronize(S(c)));
        if (columns) *columns = kv;
        if (relevant_rows) *relevant_rows = (int64_t)e->view->rel.size();
    });
}

int32_t sqlp_epi_delta(sqlp_epi *e, int64_t i, double *delta_rhs, double *delta_T)
{
    return guard([&] {
        REQUIRE(e && delta_rhs, SQLP_E_INVALID, "null argument");
        REQUIRE(i >= 0 && i < e->n_local, SQLP_E_RANGE, "scenario index out of range");
        sqlp_ctx *c = e->ctx;
        c->bind();
        PoolView *v = e->view;
        std::vector<double> col((size_t)std::max(v->n_rows, 1)), trow((size_t)std::max(e->n_T, 1));
        const double *src = e->d_D.as<double>() + (i >> 7) * (int64_t)v->s_pad * SQLP_TILE;
        if (v->n_rows) {
            e->d_scratch.ensure((size_t)(2 * e->m2 + 2 + v->n_rows) * 8, 0, S(c));
            double *tmp = e->d_scratch.as<double>() + 2 * e->m2 + 2;
            LAUNCH(c, k_gather_column, 1, 128, 0, src, (int)(i & 127), v->n_rows, tmp);
            CK(cudaMemcpyAsync(col.data(), tmp, (size_t)v->n_rows * 8, cudaMemcpyDeviceToHost, S(c)));
        }
        if (e->n_T)
            CK(cudaMemcpyAsync(trow.data(), e->d_dT.as<double>() + i * e->n_T, (size_t)e->n_T * 8,
                               cudaMemcpyDeviceToHost, S(c)));
        CK(cudaStreamSynchronize(S(c)));
        for (int64_t r = 0; r < e->m2; ++r) delta_rhs[r] = 0.0;
        for (int64_t q = 0; q < e->s; ++q) {
            if (e->h_elem_t[(size_t)q] < 0) {
                delta_rhs[e->h_pos_row[(size_t)q]] = col[(size_t)e->h_elem_j[(size_t)q]];
                if (delta_T) delta_T[q] = 0.0;
            } else if (delta_T) {
                delta_T[q] = trow[(size_t)e->h_elem_t[(size_t)q]];
            }
        }
    });
}

int32_t sqlp_epi_argmax(sqlp_epi *e, const double *x, int32_t sense, double *max_val, int64_t *max_idx)
{
    return guard([&] {
        REQUIRE(e && (x || e->n1 == 0), SQLP_E_INVALID, "null argument");
        check_sense(sense);
        sqlp_ctx *c = e->ctx;
        c->bind();
        if (!e->peers.empty()) {
            // one process, all GPUs: every shard sweeps its scenarios; results come back in GLOBAL scenario order
            REQUIRE(e->n_global == 0 || (max_val && max_idx), SQLP_E_INVALID, "null output");
            const int ns = n_shards(c);
            std::vector<std::vector<double>> hv((size_t)ns);
            std::vector<std::vector<int>> hi((size_t)ns);
            for (int s = 0; s < ns; ++s) {
                sqlp_epi *es = shard(e, s);
                sqlp_ctx *cs = es->ctx;
                cs->bind();
                if (es->n_local == 0) continue;
                es->cur_bias = nullptr;
                epi_cuts_enqueue(es, 1, x, nullptr, false);
                if (es->view->twins)
                    LAUNCH(cs, k_unmap_idx, (int)std::min<int64_t>((es->n_local + 255) / 256, 4 * cs->sm_count), 256, 0,
                           es->d_best_idx.as<int>(), (long long)es->n_local, es->view->act());
                hv[(size_t)s].resize((size_t)es->n_local);
                hi[(size_t)s].resize((size_t)es->n_local);
                CK(cudaMemcpyAsync(hv[(size_t)s].data(), es->d_best_val.p, (size_t)es->n_local * 8, cudaMemcpyDeviceToHost, S(cs)));
                CK(cudaMemcpyAsync(hi[(size_t)s].data(), es->d_best_idx.p, (size_t)es->n_local * 4, cudaMemcpyDeviceToHost, S(cs)));
            }
            for (int s = 0; s < ns; ++s) {
                sqlp_epi *es = shard(e, s);
                es->ctx->bind();
                pool_confirm(es->pool);
                CK(cudaStreamSynchronize(S(es->ctx)));
                for (int64_t l = 0; l < es->n_local; ++l) {
                    const int64_t g = ((l / SQLP_TILE) * ns + s) * SQLP_TILE + l % SQLP_TILE;   // inverse of local_of
                    max_val[g] = hv[(size_t)s][(size_t)l];
                    max_idx[g] = hi[(size_t)s][(size_t)l];
                }
            }
            c->bind();
            return;
        }
        if (e->n_local == 0) return;
        REQUIRE(max_val && max_idx, SQLP_E_INVALID, "null output");
        e->cur_bias = nullptr;
        epi_cuts_enqueue(e, 1, x, nullptr, false);
        std::vector<int> idx((size_t)e->n_local);
        if (e->view->twins)   // the sweep works on view columns: back to pool slots (the Refs the reference returns)
            LAUNCH(c, k_unmap_idx, (int)std::min<int64_t>((e->n_local + 255) / 256, 4 * c->sm_count), 256, 0,
                   e->d_best_idx.as<int>(), (long long)e->n_local, e->view->act());
        CK(cudaMemcpyAsync(max_val, e->d_best_val.p, (size_t)e->n_local * 8, cudaMemcpyDeviceToHost, S(c)));
        CK(cudaMemcpyAsync(idx.data(), e->d_best_idx.p, (size_t)e->n_local * 4, cudaMemcpyDeviceToHost, S(c)));
        pool_confirm(e->pool);
        CK(cudaStreamSynchronize(S(c)));
        for (int64_t i = 0; i < e->n_local; ++i) max_idx[i] = idx[(size_t)i];
    });
}

static void finish_cut(sqlp_epi *e, int NX, const CutHost &h, double *alpha, double *beta,
                       double *weight_mark, double *val)
{
    const int n1 = (int)e->n1, NC = n1 + 2;
    REQUIRE(!(h.flags & 1), SQLP_E_NO_ARGMAX,
            "no dual vertex beat -Inf for some scenario (empty pool or all scores NaN/-Inf)");
    for (int x = 0; x < NX; ++x) {
        alpha[x] = h.out[(size_t)x * NC];
        for (int j = 0; j < n1; ++j) beta[(size_t)x * n1 + j] = h.out[(size_t)x * NC + 1 + j];
        if (val) val[x] = h.out[(size_t)x * NC + n1 + 1];
    }
    if (weight_mark) *weight_mark = e->total_weight;   // epigraph.jl:145
}

int32_t sqlp_epi_build_cut(sqlp_epi *e, const double *x, double *alpha, double *beta,
                           double *weight_mark, double *val)
{
    return guard([&] {
        REQUIRE(e && alpha && (beta || e->n1 == 0) && (x || e->n1 == 0), SQLP_E_INVALID, "null argument");
        sqlp_ctx *c = e->ctx;
        c->bind();
        if (!c->peers.empty()) {
            std::vector<std::vector<double>> xs(1, std::vector<double>(x, x + e->n1));
            cell_cuts_all_shards(c, 1, 1, &e, xs);
        } else {
            e->cur_bias = nullptr;
            CellGather cg(c, 1, 1, &e);
            epi_cuts_enqueue(e, 1, x, nullptr, true, nullptr, cg.slot(0));
            cg.run();
        }
        CutHost h;
        epi_cuts_fetch(e, 1, h);
        pool_confirm(e->pool);
        CK(cudaStreamSynchronize(S(c)));
        sync_all_shards(c, e->pool);
        finish_cut(e, 1, h, alpha, beta, weight_mark, val);
    });
}

int32_t sqlp_cell_build_cuts2(int32_t n_epi, sqlp_epi *const *epi, const double *x_cand,
                              const double *x_inc, double *alpha, double *beta, double *weight_mark,
                              double *val)
{
    return guard([&] {
        REQUIRE(n_epi >= 0 && (epi || n_epi == 0), SQLP_E_INVALID, "bad epigraph list");
        if (n_epi == 0) return;
        REQUIRE(alpha && weight_mark, SQLP_E_INVALID, "null output");
        sqlp_ctx *c = epi[0]->ctx;
        c->bind();
        std::vector<CutHost> h((size_t)n_epi);
        std::vector<double> x2;
        for (int i = 0; i < n_epi; ++i) {
            REQUIRE(epi[i] && epi[i]->ctx == c, SQLP_E_INVALID, "epigraphs must share one context");
            epi[i]->cur_bias = nullptr;
        }
        if (!c->peers.empty()) {
            std::vector<std::vector<double>> xs((size_t)n_epi);
            for (int i = 0; i < n_epi; ++i) {
                sqlp_epi *e = epi[i];
                REQUIRE((x_cand && x_inc && beta) || e->n1 == 0, SQLP_E_INVALID, "null argument");
                xs[(size_t)i].assign((size_t)2 * e->n1, 0.0);
                for (int64_t j = 0; j < e->n1; ++j) { xs[(size_t)i][(size_t)j] = x_cand[j]; xs[(size_t)i][(size_t)(e->n1 + j)] = x_inc[j]; }
            }
            cell_cuts_all_shards(c, 2, n_epi, epi, xs);
        } else {
            CellGather cg(c, 2, n_epi, epi);
            for (int i = 0; i < n_epi; ++i) {
                sqlp_epi *e = epi[i];
                REQUIRE((x_cand && x_inc && beta) || e->n1 == 0, SQLP_E_INVALID, "null argument");
                x2.assign((size_t)2 * e->n1, 0.0);
                for (int64_t j = 0; j < e->n1; ++j) { x2[(size_t)j] = x_cand[j]; x2[(size_t)(e->n1 + j)] = x_inc[j]; }
                epi_cuts_enqueue(e, 2, x2.data(), nullptr, true, bias_twin(i, epi), cg.slot(i));   // x2 is pageable: staged now
            }
            cg.run();
        }
        for (int i = 0; i < n_epi; ++i) epi_cuts_fetch(epi[i], 2, h[(size_t)i]);
        for (int i = 0; i < n_epi; ++i) pool_confirm(epi[i]->pool);
        CK(cudaStreamSynchronize(S(c)));
        sync_all_shards(c, epi[0]->pool);
        int64_t boff = 0;
        for (int i = 0; i < n_epi; ++i) {
            sqlp_epi *e = epi[i];
            finish_cut(e, 2, h[(size_t)i], alpha + 2 * i, beta + boff, weight_mark + i,
                       val ? val + 2 * i : nullptr);
            boff += 2 * e->n1;
        }
    });
}

int32_t sqlp_cell_sd_step(int32_t n_epi, sqlp_epi *const *epi, const double *values, const double *weights,
                          int64_t n_vertices, const double *vertices, int32_t *inserted, int64_t *index,
                          const double *x_cand, const double *x_inc, double *alpha, double *beta,
                          double *weight_mark, double *val)
{
    return guard([&] {
        REQUIRE(n_epi >= 1 && epi && epi[0], SQLP_E_INVALID, "bad epigraph list");
        REQUIRE(n_vertices >= 0 && (vertices || n_vertices == 0), SQLP_E_INVALID, "bad vertex list");
        REQUIRE(alpha && weight_mark, SQLP_E_INVALID, "null output");
        sqlp_ctx *c = epi[0]->ctx;
        sqlp_pool *p = epi[0]->pool;
        REQUIRE(c->world == 1 || c->rank != 0 || vertices || n_vertices == 0, SQLP_E_INVALID, "null vertices");
        c->bind();
        size_t n_values = 0, n_out = 0;
        for (int i = 0; i < n_epi; ++i) {
            sqlp_epi *e = epi[i];
            REQUIRE(e && e->ctx == c && e->pool == p, SQLP_E_INVALID, "epigraphs must share one context and one pool");
            REQUIRE((x_cand && x_inc && beta) || e->n1 == 0, SQLP_E_INVALID, "null argument");
            n_values += (size_t)e->s;
            n_out += 2 * ((size_t)e->n1 + 2) + 1;                 // two cuts (alpha, beta, val) + the flags word
        }
        REQUIRE(values || n_values == 0, SQLP_E_INVALID, "null values");
        // the drained stream of the previous call makes the two reusable buffers free
        c->h_step.ensure(((size_t)n_vertices * sizeof(PushResult) / 8 + n_out + 1) * 8);
        // algorithm.jl:45-46  add_scenario!(epi, scenario, weight) for every epigraph -- on every GPU of a multi-GPU
        // context (each keeps the scenarios it owns), the leader last so that it stays the bound device
        for (int sh = n_shards(c) - 1; sh >= 0; --sh) {
            sqlp_ctx *cs = shard(c, sh);
            cs->bind();
            cs->d_step.ensure(std::max<size_t>(n_values, 1) * 8, 0, S(cs), false);
            if (n_values)
                CK(cudaMemcpyAsync(cs->d_step.p, values, n_values * 8, cudaMemcpyHostToDevice, S(cs)));
            size_t voff = 0;
            for (int i = 0; i < n_epi; ++i) {
                epi_add(shard(epi[i], sh), 1, nullptr, cs->d_step.as<double>() + voff, weights ? weights + i : nullptr, false, 0, 0);
                voff += (size_t)epi[i]->s;
            }
            if (sh > 0 && n_vertices) pool_push_enqueue(shard(p, sh), n_vertices, vertices, nullptr);
        }
        // algorithm.jl:50,54  push!(dual_vertices, dual) for the vertices found at the candidate and the incumbent
        PushResult *h_res = c->h_step.as<PushResult>();
        double *h_out = c->h_step.as<double>() + (size_t)n_vertices * sizeof(PushResult) / 8;
        long long *h_K = reinterpret_cast<long long *>(h_out + n_out);
        if (n_vertices) {
            pool_push_enqueue(p, n_vertices, vertices, nullptr);
            CK(cudaMemcpyAsync(h_res, p->d_results.p, (size_t)n_vertices * sizeof(PushResult), cudaMemcpyDeviceToHost, S(c)));
        }
        // algorithm.jl:79-85  the candidate cut and the regenerated incumbent cut of every epigraph
        std::vector<double> x2;
        size_t ooff = 0;
        for (int i = 0; i < n_epi; ++i) epi[i]->cur_bias = nullptr;
        if (!c->peers.empty()) {
            std::vector<std::vector<double>> xs((size_t)n_epi);
            for (int i = 0; i < n_epi; ++i) {
                xs[(size_t)i].assign((size_t)2 * epi[i]->n1, 0.0);
                for (int64_t j = 0; j < epi[i]->n1; ++j) { xs[(size_t)i][(size_t)j] = x_cand[j]; xs[(size_t)i][(size_t)(epi[i]->n1 + j)] = x_inc[j]; }
            }
            cell_cuts_all_shards(c, 2, n_epi, epi, xs);           // every GPU's chain side by side, one grouped all-gather
        } else {
            CellGather cg(c, 2, n_epi, epi);
            for (int i = 0; i < n_epi; ++i) {
                sqlp_epi *e = epi[i];
                x2.assign((size_t)2 * e->n1, 0.0);
                for (int64_t j = 0; j < e->n1; ++j) { x2[(size_t)j] = x_cand[j]; x2[(size_t)(e->n1 + j)] = x_inc[j]; }
                epi_cuts_enqueue(e, 2, x2.data(), nullptr, true, bias_twin(i, epi), cg.slot(i));   // x2 is pageable: staged now
            }
            cg.run();                                             // sharded job: one all-gather for the whole cell
        }
        for (int i = 0; i < n_epi; ++i) {
            sqlp_epi *e = epi[i];
            const size_t NC = (size_t)e->n1 + 2;
            CK(cudaMemcpyAsync(h_out + ooff, e->d_out.p, 2 * NC * 8, cudaMemcpyDeviceToHost, S(c)));
            CK(cudaMemcpyAsync(h_out + ooff + 2 * NC, e->d_flags.p, 4, cudaMemcpyDeviceToHost, S(c)));
            ooff += 2 * NC + 1;
        }
        CK(cudaMemcpyAsync(h_K, p->d_K.p, 8, cudaMemcpyDeviceToHost, S(c)));
        CK(cudaStreamSynchronize(S(c)));                          // the step's only synchronisation (per GPU)
        p->K = *h_K;
        p->pending = 0;
        sync_all_shards(c, p);
        for (int64_t v = 0; v < n_vertices; ++v) {
            if (inserted) inserted[v] = h_res[v].inserted;
            if (index) index[v] = h_res[v].index;
        }
        ooff = 0;
        int64_t boff = 0;
        for (int i = 0; i < n_epi; ++i) {
            sqlp_epi *e = epi[i];
            const size_t NC = (size_t)e->n1 + 2;
            CutHost h;
            h.out.assign(h_out + ooff, h_out + ooff + 2 * NC);
            h.flags = *reinterpret_cast<const int *>(h_out + ooff + 2 * NC);
            finish_cut(e, 2, h, alpha + 2 * i, beta + boff, weight_mark + i, val ? val + 2 * i : nullptr);
            ooff += 2 * NC + 1;
            boff += 2 * e->n1;
        }
    });
}

int32_t sqlp_epi_build_cuts2(sqlp_epi *e, const double *x_cand, const double *x_inc, double alpha[2],
                             double *beta, double *weight_mark, double *val)
{
    double wm = 0;
    int32_t st = sqlp_cell_build_cuts2(1, &e, x_cand, x_inc, alpha, beta, &wm, val);
    if (st == SQLP_OK && weight_mark) *weight_mark = wm;
    return st;
}

int32_t sqlp_epi_build_cuts2_dev(sqlp_epi *e, const double *d_x2, double *d_out)
{
    return guard([&] {
        REQUIRE(e && d_out && (d_x2 || e->n1 == 0), SQLP_E_INVALID, "null argument");
        sqlp_ctx *c = e->ctx;
        c->bind();
        e->cur_bias = nullptr;
        CellGather cg(c, 2, 1, &e);
        epi_cuts_enqueue(e, 2, nullptr, d_x2, true, nullptr, cg.slot(0));
        cg.run();
        CK(cudaMemcpyAsync(d_out, e->d_out.p, (size_t)2 * (e->n1 + 2) * 8, cudaMemcpyDeviceToDevice, S(c)));
    });
}

int32_t sqlp_cell_build_cuts2_dev(int32_t n_epi, sqlp_epi *const *epi, const double *d_x2, double *d_out)
{
    return guard([&] {
        REQUIRE(n_epi >= 1 && epi && epi[0] && d_out, SQLP_E_INVALID, "bad argument");
        sqlp_ctx *c = epi[0]->ctx;
        REQUIRE(c->peers.empty(), SQLP_E_UNSUPPORTED, "device pointers belong to one GPU: use sqlp_cell_build_cuts2 with a multi-GPU context");
        c->bind();
        for (int i = 0; i < n_epi; ++i) {
            REQUIRE(epi[i] && epi[i]->ctx == c && epi[i]->n1 == epi[0]->n1, SQLP_E_INVALID,
                    "the epigraphs of a cell share a context and a first stage");
            REQUIRE(d_x2 || epi[i]->n1 == 0, SQLP_E_INVALID, "null points");
            epi[i]->cur_bias = nullptr;
        }
        CellGather cg(c, 2, n_epi, epi);
        for (int i = 0; i < n_epi; ++i)
            epi_cuts_enqueue(epi[i], 2, nullptr, d_x2, true, bias_twin(i, epi), cg.slot(i));
        cg.run();                                       // sharded job: one all-gather for the whole cell
        const size_t row = (size_t)2 * ((size_t)epi[0]->n1 + 2);
        for (int i = 0; i < n_epi; ++i)
            CK(cudaMemcpyAsync(d_out + (size_t)i * row, epi[i]->d_out.p, row * 8, cudaMemcpyDeviceToDevice, S(c)));
    });
}

// ---------------------------------------------------------------- cut list (N1 / N3) ----
int32_t sqlp_epi_set_weights(sqlp_epi *e, double objective_weight, double lower_bound)
{
    if (e) for (sqlp_epi *q : e->peers) {   // multi-GPU context: every shard
        int32_t st = sqlp_epi_set_weights(q, objective_weight, lower_bound);
        if (st != SQLP_OK) return st;
    }
    return guard([&] {
        REQUIRE(e, SQLP_E_INVALID, "null epigraph");
        e->objective_weight = objective_weight;
        e->lower_bound = lower_bound;
    });
}

int32_t sqlp_epi_cuts_push(sqlp_epi *e, double alpha, const double *beta, double weight_mark)
{
    return guard([&] {
        REQUIRE(e && (beta || e->n1 == 0), SQLP_E_INVALID, "null argument");
        sqlp_ctx *c = e->ctx;
        c->bind();
        const size_t RS = (size_t)e->n1 + 2;
        cuts_reserve(e, e->n_cuts + 1);
        std::vector<double> row(RS);
        row[0] = alpha;
        for (int64_t j = 0; j < e->n1; ++j) row[(size_t)j + 1] = beta[j];
        row[RS - 1] = weight_mark;
        CK(cudaMemcpyAsync(e->d_cuts.as<double>() + (size_t)e->n_cuts * RS, row.data(), RS * 8,
                           cudaMemcpyHostToDevice, S(c)));
        CK(cudaStreamSynchronize(S(c)));
        ++e->n_cuts;
    });
}

int32_t sqlp_epi_cuts_set_incumbent(sqlp_epi *e, double alpha, const double *beta, double weight_mark)
{
    return guard([&] {
        REQUIRE(e, SQLP_E_INVALID, "null epigraph");
        sqlp_ctx *c = e->ctx;
        c->bind();
        if (!beta && e->n1 > 0) { e->has_inc = false; return; }   // incumbent_cut = nothing
        const size_t RS = (size_t)e->n1 + 2;
        cuts_reserve(e, 1);
        std::vector<double> row(RS);
        row[0] = alpha;
        for (int64_t j = 0; j < e->n1; ++j) row[(size_t)j + 1] = beta[j];
        row[RS - 1] = weight_mark;
        CK(cudaMemcpyAsync(e->d_inc.p, row.data(), RS * 8, cudaMemcpyHostToDevice, S(c)));
        CK(cudaStreamSynchronize(S(c)));
        e->has_inc = true;
    });
}

int32_t sqlp_epi_cuts_commit(sqlp_epi *e, int32_t with_incumbent)
{
    return guard([&] {
        REQUIRE(e, SQLP_E_INVALID, "null epigraph");
        REQUIRE(e->last_nx >= 1, SQLP_E_INVALID, "no cut has been formed since the last commit");
        REQUIRE(!with_incumbent || e->last_nx == 2, SQLP_E_INVALID,
                "the incumbent cut needs sqlp_epi_build_cuts2 / sqlp_cell_build_cuts2");
        sqlp_ctx *c = e->ctx;
        c->bind();
        const size_t RS = (size_t)e->n1 + 2;
        cuts_reserve(e, e->n_cuts + 1);
        // snapshot f_{k-1} = the list before the new cuts (algorithm.jl:76, sdEpigraphInfo)
        e->n_last = e->n_cuts;
        e->has_prev_inc = e->has_inc;
        if (e->has_inc)
            CK(cudaMemcpyAsync(e->d_prev_inc.p, e->d_inc.p, RS * 8, cudaMemcpyDeviceToDevice, S(c)));
        LAUNCH(c, k_cut_store, 1, 128, 0, e->d_out.as<double>(), e->total_weight, (int)e->n1,
               e->d_cuts.as<double>() + (size_t)e->n_cuts * RS);                    // push!(epi.cuts, new_cut)
        ++e->n_cuts;
        if (with_incumbent) {                                                       // epi.incumbent_cut = ...
            LAUNCH(c, k_cut_store, 1, 128, 0, e->d_out.as<double>() + RS, e->total_weight, (int)e->n1,
                   e->d_inc.as<double>());
            e->has_inc = true;
        }
        e->last_nx = 0;
    });
}

int32_t sqlp_epi_cuts_delete(sqlp_epi *e, int64_t n, const int64_t *idx)
{
    return guard([&] {
        REQUIRE(e && (idx || n == 0) && n >= 0, SQLP_E_INVALID, "bad argument");
        if (n == 0) return;
        sqlp_ctx *c = e->ctx;
        c->bind();
        std::vector<char> drop((size_t)e->n_cuts, 0);
        for (int64_t q = 0; q < n; ++q) {
            REQUIRE(idx[q] >= 0 && idx[q] < e->n_cuts, SQLP_E_RANGE, "cut index out of range");
            REQUIRE(q == 0 || idx[q] > idx[q - 1], SQLP_E_INVALID, "indices must ascend (deleteat!)");
            drop[(size_t)idx[q]] = 1;
        }
        std::vector<int> keep;
        for (int64_t j = 0; j < e->n_cuts; ++j) if (!drop[(size_t)j]) keep.push_back((int)j);
        const int RS = (int)e->n1 + 2;
        if (!keep.empty()) {
            upload(e->d_keep, keep, S(c));
            e->d_cuts_tmp.ensure(keep.size() * RS * 8, 0, S(c), false);
            LAUNCH(c, k_cuts_gather, (int)((keep.size() * RS + 255) / 256), 256, 0, e->d_cuts.as<double>(),
                   e->d_keep.as<int>(), (int)keep.size(), RS, e->d_cuts_tmp.as<double>());
            CK(cudaMemcpyAsync(e->d_cuts.p, e->d_cuts_tmp.p, keep.size() * RS * 8, cudaMemcpyDeviceToDevice, S(c)));
        }
        CK(cudaStreamSynchronize(S(c)));   // `keep` is read by the upload
        e->n_cuts = (int64_t)keep.size();
        e->n_last = std::min(e->n_last, e->n_cuts);
    });
}

int32_t sqlp_epi_cuts_count(sqlp_epi *e, int64_t *n_cuts, int32_t *has_incumbent)
{
    return guard([&] {
        REQUIRE(e, SQLP_E_INVALID, "null epigraph");
        if (n_cuts) *n_cuts = e->n_cuts;
        if (has_incumbent) *has_incumbent = e->has_inc ? 1 : 0;
    });
}

int32_t sqlp_epi_cuts_get(sqlp_epi *e, int64_t index, double *alpha, double *beta, double *weight_mark)
{
    return guard([&] {
        REQUIRE(e, SQLP_E_INVALID, "null epigraph");
        REQUIRE(index >= -1 && index < e->n_cuts, SQLP_E_RANGE, "cut index out of range");
        REQUIRE(index >= 0 || e->has_inc, SQLP_E_RANGE, "there is no incumbent cut");
        sqlp_ctx *c = e->ctx;
        c->bind();
        const size_t RS = (size_t)e->n1 + 2;
        std::vector<double> row(RS);
        const double *src = index < 0 ? e->d_inc.as<double>() : e->d_cuts.as<double>() + (size_t)index * RS;
        CK(cudaMemcpyAsync(row.data(), src, RS * 8, cudaMemcpyDeviceToHost, S(c)));
        CK(cudaStreamSynchronize(S(c)));
        if (alpha) *alpha = row[0];
        if (beta) for (int64_t j = 0; j < e->n1; ++j) beta[j] = row[(size_t)j + 1];
        if (weight_mark) *weight_mark = row[RS - 1];
    });
}

int32_t sqlp_epi_evaluate(sqlp_epi *e, const double *x, int32_t which, double *out)
{
    return guard([&] {
        REQUIRE(e && out && (x || e->n1 == 0), SQLP_E_INVALID, "null argument");
        REQUIRE(which == 0 || which == 1, SQLP_E_INVALID, "which must be 0 (current) or 1 (snapshot)");
        sqlp_ctx *c = e->ctx;
        c->bind();
        e->d_x2.ensure((size_t)2 * std::max<int64_t>(e->n1, 1) * 8, 0, S(c));
        e->d_eval.ensure(64, 0, S(c));
        CK(cudaMemcpyAsync(e->d_x2.p, x, (size_t)e->n1 * 8, cudaMemcpyHostToDevice, S(c)));
        cuts_evaluate_enqueue(e, e->d_x2.as<double>(), 1, 2, e->d_eval.as<double>());
        double h[2];
        CK(cudaMemcpyAsync(h, e->d_eval.p, 16, cudaMemcpyDeviceToHost, S(c)));
        CK(cudaStreamSynchronize(S(c)));
        *out = h[which];
    });
}

int32_t sqlp_epi_master_rows(sqlp_epi *e, double *rows, int64_t *n_rows)
{
    return guard([&] {
        REQUIRE(e && n_rows, SQLP_E_INVALID, "null argument");
        sqlp_ctx *c = e->ctx;
        c->bind();
        const int64_t nr = e->n_cuts + (e->has_inc ? 1 : 0);
        *n_rows = nr;
        if (!rows || nr == 0) return;
        const size_t RW = (size_t)e->n1 + 1;
        cuts_reserve(e, 1);
        e->d_rows.ensure((size_t)nr * RW * 8, 0, S(c), false);
        LAUNCH(c, k_cuts_master_rows, (int)std::min<size_t>((nr * RW + 255) / 256, 1024), 256, 0, cut_list(e, false),
               (int)e->n1, e->total_weight, e->lower_bound, e->d_rows.as<double>());
        CK(cudaMemcpyAsync(rows, e->d_rows.p, (size_t)nr * RW * 8, cudaMemcpyDeviceToHost, S(c)));
        CK(cudaStreamSynchronize(S(c)));
    });
}

int32_t sqlp_cell_check_improvement(int32_t n_epi, sqlp_epi *const *epi, const double *x_cand,
                                    const double *x_inc, const double *cost, double q_factor, double *out4)
{
    return guard([&] {
        REQUIRE(n_epi >= 1 && epi && x_cand && x_inc && cost && out4, SQLP_E_INVALID, "null argument");
        sqlp_epi *e0 = epi[0];
        sqlp_ctx *c = e0->ctx;
        c->bind();
        const int64_t n1 = e0->n1;
        for (int32_t i = 0; i < n_epi; ++i)
            REQUIRE(epi[i] && epi[i]->ctx == c && epi[i]->n1 == n1, SQLP_E_INVALID,
                    "the epigraphs of a cell share a context and a first stage");
        // x2 = [cand | inc], cost, est[n_epi][4], out[4] in one scratch buffer of the first epigraph
        const size_t need = (size_t)(3 * std::max<int64_t>(n1, 1) + 4 * n_epi + 4) * 8;
        e0->d_eval.ensure(need, 0, S(c));
        double *d_x2 = e0->d_eval.as<double>(), *d_cost = d_x2 + 2 * n1, *d_est = d_cost + n1,
               *d_out = d_est + 4 * n_epi;
        CK(cudaMemcpyAsync(d_x2, x_cand, (size_t)n1 * 8, cudaMemcpyHostToDevice, S(c)));
        CK(cudaMemcpyAsync(d_x2 + n1, x_inc, (size_t)n1 * 8, cudaMemcpyHostToDevice, S(c)));
        CK(cudaMemcpyAsync(d_cost, cost, (size_t)n1 * 8, cudaMemcpyHostToDevice, S(c)));
        for (int32_t i = 0; i < n_epi; ++i)
            cuts_evaluate_enqueue(epi[i], d_x2, 2, 2, d_est + 4 * i);   // {cur@cand, cur@inc, last@cand, last@inc}
        LAUNCH(c, k_improvement, 1, 32, 0, d_est, (int)n_epi, d_cost, d_x2, (int)n1, q_factor, d_out);
        CK(cudaMemcpyAsync(out4, d_out, 32, cudaMemcpyDeviceToHost, S(c)));
        CK(cudaStreamSynchronize(S(c)));
    });
}

// ---------------------------------------------------------------- SMPS reader (row N4) ---
int32_t sqlp_smps_load(const char *cor_path, const char *tim_path, const char *sto_path, sqlp_smps **out)
{
    return guard([&] {
        REQUIRE(out, SQLP_E_INVALID, "null argument");
        *out = nullptr;
        std::unique_ptr<sqlp_smps> p(new sqlp_smps());
        smps_read_cor(*p, cor_path);
        smps_read_tim(*p, tim_path);
        if (sto_path && *sto_path) smps_read_sto(*p, sto_path);
        smps_stage2(*p);
        *out = p.release();
    });
}

int32_t sqlp_smps_destroy(sqlp_smps *p)
{
    return guard([&] { delete p; });
}

int32_t sqlp_smps_dims(sqlp_smps *p, int64_t *dims)
{
    return guard([&] {
        REQUIRE(p && dims, SQLP_E_INVALID, "null argument");
        int64_t cor_nnz = 0;
        for (auto &c : p->col_entries) cor_nnz += (int64_t)c.size();
        int64_t d[SQLP_SMPS_NDIMS] = {(int64_t)p->rows.size(), (int64_t)p->cols.size(), cor_nnz, p->n1, p->n2, p->m2,
                                      (int64_t)p->T_nzval.size(), (int64_t)p->W_nzval.size(), (int64_t)p->r_val.size(),
                                      (int64_t)p->elems.size(), p->max_outcomes, (int64_t)p->periods.size()};
        std::copy(d, d + SQLP_SMPS_NDIMS, dims);
    });
}

int32_t sqlp_smps_name(sqlp_smps *p, int32_t what, int64_t index, char *buf, int64_t buflen)
{
    return guard([&] {
        REQUIRE(p && buf && buflen > 0, SQLP_E_INVALID, "null argument");
        const std::string *s = nullptr;
        auto at = [&](size_t n) { REQUIRE(index >= 0 && (size_t)index < n, SQLP_E_RANGE, "name index out of range"); };
        switch (what) {
        case SQLP_SMPS_COR_NAME: s = &p->name; break;
        case SQLP_SMPS_TIM_NAME: s = &p->tim_name; break;
        case SQLP_SMPS_STO_NAME: s = &p->sto_name; break;
        case SQLP_SMPS_ROW_NAME: at(p->rows.size()); s = &p->rows[index]; break;
        case SQLP_SMPS_COL_NAME: at(p->cols.size()); s = &p->cols[index]; break;
        case SQLP_SMPS_PERIOD_NAME: at(p->periods.size()); s = &p->periods[index][0]; break;
        case SQLP_SMPS_PERIOD_COL: at(p->periods.size()); s = &p->periods[index][1]; break;
        case SQLP_SMPS_PERIOD_ROW: at(p->periods.size()); s = &p->periods[index][2]; break;
        case SQLP_SMPS_ELEM_COL: at(p->elems.size()); s = &p->elems[index].col; break;
        case SQLP_SMPS_ELEM_ROW: at(p->elems.size()); s = &p->elems[index].row; break;
        default: throw Error(SQLP_E_INVALID, "unknown name selector");
        }
        REQUIRE((int64_t)s->size() < buflen, SQLP_E_RANGE, "name buffer too small");
        memcpy(buf, s->c_str(), s->size() + 1);
    });
}

int32_t sqlp_smps_cor(sqlp_smps *p, char *directions, double *rhs, double *lower, double *upper,
                      int64_t *colptr, int64_t *rowval, double *nzval)
{
    return guard([&] {
        REQUIRE(p, SQLP_E_INVALID, "null argument");
        if (directions) std::copy(p->dir.begin(), p->dir.end(), directions);
        if (rhs) std::copy(p->rhs.begin(), p->rhs.end(), rhs);
        if (lower) std::copy(p->lower.begin(), p->lower.end(), lower);
        if (upper) std::copy(p->upper.begin(), p->upper.end(), upper);
        int64_t k = 0;
        if (colptr) colptr[0] = 0;
        for (size_t j = 0; j < p->col_entries.size(); ++j) {
            for (auto &kv : p->col_entries[j]) {
                if (rowval) rowval[k] = kv.first;
                if (nzval) nzval[k] = kv.second;
                ++k;
            }
            if (colptr) colptr[j + 1] = k;
        }
    });
}

int32_t sqlp_smps_stage2(sqlp_smps *p, double *rbar, int64_t *T_colptr, int64_t *T_rowval, double *T_nzval,
                         int64_t *W_colptr, int64_t *W_rowval, double *W_nzval, double *cost, double *x_cost)
{
    return guard([&] {
        REQUIRE(p, SQLP_E_INVALID, "null argument");
        if (rbar) std::copy(p->rhs.begin() + p->r2, p->rhs.end(), rbar);
        if (T_colptr) std::copy(p->T_colptr.begin(), p->T_colptr.end(), T_colptr);
        if (T_rowval) std::copy(p->T_rowval.begin(), p->T_rowval.end(), T_rowval);
        if (T_nzval) std::copy(p->T_nzval.begin(), p->T_nzval.end(), T_nzval);
        if (W_colptr) std::copy(p->W_colptr.begin(), p->W_colptr.end(), W_colptr);
        if (W_rowval) std::copy(p->W_rowval.begin(), p->W_rowval.end(), W_rowval);
        if (W_nzval) std::copy(p->W_nzval.begin(), p->W_nzval.end(), W_nzval);
        if (cost) std::copy(p->cost.begin(), p->cost.end(), cost);
        if (x_cost) std::copy(p->x_cost.begin(), p->x_cost.end(), x_cost);
    });
}

int32_t sqlp_smps_elements(sqlp_smps *p, int32_t *pos_row, int32_t *pos_col, int32_t *kind, double *par_a,
                           double *par_b, int32_t *cnt, double *vals, double *probs)
{
    return guard([&] {
        REQUIRE(p, SQLP_E_INVALID, "null argument");
        const int64_t mo = p->max_outcomes;
        for (size_t e = 0; e < p->elems.size(); ++e) {
            const SmpsElement &el = p->elems[e];
            if (pos_row) pos_row[e] = p->pos_row[e];
            if (pos_col) pos_col[e] = p->pos_col[e];
            if (kind) kind[e] = el.kind;
            if (par_a) par_a[e] = el.a;
            if (par_b) par_b[e] = el.b;
            if (cnt) cnt[e] = (int32_t)el.val.size();
            for (int64_t o = 0; o < mo; ++o) {
                bool in = o < (int64_t)el.val.size();
                if (vals) vals[e * mo + o] = in ? el.val[o] : 0.0;
                if (probs) probs[e * mo + o] = in ? el.prob[o] : 0.0;
            }
        }
    });
}

int32_t sqlp_epi_create_smps(sqlp_ctx *c, sqlp_pool *pool, sqlp_smps *p, sqlp_epi **out)
{
    if (!(c && pool && p && out)) {
        g_err = "null argument";
        return SQLP_E_INVALID;
    }
    *out = nullptr;
    sqlp_epi *e = nullptr;
    int32_t st = sqlp_epi_create(c, pool, p->m2, p->n1, (int64_t)p->r_val.size(), p->r_idx.data(), p->r_val.data(),
                                 p->T_colptr.data(), p->T_rowval.data(), p->T_nzval.data(), (int64_t)p->elems.size(),
                                 p->pos_row.data(), p->pos_col.data(), &e);
    if (st != SQLP_OK) return st;
    const int64_t s = (int64_t)p->elems.size(), mo = std::max<int64_t>(p->max_outcomes, 1);
    bool any_cont = false;
    std::vector<double> vals((size_t)(s * mo), 0.0), cdf((size_t)(s * mo), 0.0), a((size_t)s), b((size_t)s);
    std::vector<int32_t> cnt((size_t)s, 1), kind((size_t)s);
    for (int64_t q = 0; q < s; ++q) {
        const SmpsElement &el = p->elems[(size_t)q];
        kind[q] = el.kind; a[q] = el.a; b[q] = el.b;
        if (el.kind != 0) { any_cont = true; continue; }
        cnt[q] = (int32_t)el.val.size();
        double run = 0.0;
        for (size_t o = 0; o < el.val.size(); ++o) {       // cdf = running sum of the probabilities
            run += el.prob[o];
            vals[q * mo + o] = el.val[o];
            cdf[q * mo + o] = run;
        }
    }
    if (s > 0) st = sqlp_epi_set_outcomes(e, mo, vals.data(), cdf.data(), cnt.data());
    if (st == SQLP_OK && any_cont) st = sqlp_epi_set_distributions(e, kind.data(), a.data(), b.data());
    if (st != SQLP_OK) {
        std::string keep = g_err;
        sqlp_epi_destroy(e);
        g_err = keep;
        return st;
    }
    *out = e;
    return SQLP_OK;
}

int32_t sqlp_eval_dual(sqlp_epi *e, int64_t i, int64_t vertex, const double *x, double *out)
{
    return guard([&] {
        REQUIRE(e && out && (x || e->n1 == 0), SQLP_E_INVALID, "null argument");
        sqlp_ctx *c = e->ctx;
        sqlp_pool *p = e->pool;
        c->bind();
        pool_confirm(p);
        REQUIRE(i >= 0 && i < e->n_local, SQLP_E_RANGE, "scenario index out of range");
        REQUIRE(vertex >= 0 && vertex < p->K, SQLP_E_RANGE, "vertex index out of range");
        e->d_x2.ensure((size_t)2 * std::max<int64_t>(e->n1, 1) * 8, 0, S(c));
        if (e->n1) CK(cudaMemcpyAsync(e->d_x2.p, x, (size_t)e->n1 * 8, cudaMemcpyHostToDevice, S(c)));
        EvalArgs a;
        a.pi_row = p->d_pi.as<double>() + vertex * p->m2;
        a.rbar = e->d_rbar.as<double>();
        a.T_colptr = e->d_colptr.as<long long>();
        a.T_rowval = e->d_rowval.as<int>();
        a.T_nzval = e->d_nzval.as<double>();
        a.m2 = (int)e->m2; a.n1 = (int)e->n1; a.s_pad = e->view->s_pad; a.n_rows = e->view->n_rows;
        a.n_T = e->n_T;
        a.s_rows = e->view->d_rows.as<int>();
        a.Dtile = e->d_D.as<double>() + (i >> 7) * (int64_t)e->view->s_pad * SQLP_TILE;
        a.dcol = (int)(i & 127);
        a.dTrow = e->n_T ? e->d_dT.as<double>() + i * e->n_T : nullptr;
        a.mc_col = e->d_mcol.as<int>(); a.mc_row = e->d_mrow.as<int>(); a.mc_slot = e->d_mslot.as<int>();
        a.x = e->d_x2.as<double>();
        a.scratch = e->d_scratch.as<double>();
        a.out = e->d_scratch.as<double>() + 2 * e->m2;
        LAUNCH(c, k_eval_dual, 1, 32, 0, a);
        CK(cudaMemcpyAsync(out, a.out, 8, cudaMemcpyDeviceToHost, S(c)));
        CK(cudaStreamSynchronize(S(c)));
    });
}

}  // extern "C"

// host_pool.cuh -- host side of the dual-vertex pool: capacity, enqueued pushes, view / table sync.
// Part of the single translation unit sqlp_api.cu (included there, in order).
#pragma once

namespace {

void pool_reserve(sqlp_pool *p, int64_t need)
{
    if (need <= p->cap) return;
    sqlp_ctx *c = p->ctx;
    int64_t ncap = std::max<int64_t>(round_up(need, SQLP_TILE), std::max<int64_t>(1024, p->cap * 2));
    int64_t used = p->upper();
    p->d_pi.ensure((size_t)ncap * p->m2 * 8, (size_t)used * p->m2 * 8, S(c));
    p->d_hash.ensure((size_t)ncap * 8, (size_t)used * 8, S(c));
    for (PoolView *v : p->views) {
        size_t per_chunk = (size_t)v->s_pad * SQLP_TILE * 8;
        v->d_piS.ensure((size_t)(ncap / SQLP_TILE) * per_chunk,
                        (size_t)((used + SQLP_TILE - 1) / SQLP_TILE) * per_chunk, S(c));
        v->d_piR.ensure((size_t)(ncap / SQLP_TILE) * per_chunk,
                        (size_t)((used + SQLP_TILE - 1) / SQLP_TILE) * per_chunk, S(c));
    }
    for (sqlp_epi *e : p->epis) {
        e->d_rt.ensure((size_t)ncap * (e->n1 + 1) * 8, (size_t)used * (e->n1 + 1) * 8, S(c));
        e->rt_cap = ncap;
    }
    p->cap = ncap;
}

// Bring K up to date on the host (one small D2H + sync) if pushes are outstanding.
void pool_confirm(sqlp_pool *p)
{
    if (!p->pending) return;
    long long K = 0;
    CK(cudaMemcpyAsync(&K, p->d_K.p, 8, cudaMemcpyDeviceToHost, S(p->ctx)));
    CK(cudaStreamSynchronize(S(p->ctx)));
    p->K = K;
    p->pending = 0;
}

void pool_push_enqueue(sqlp_pool *p, int64_t n, const double *v_host, const double *v_dev)
{
    sqlp_ctx *c = p->ctx;
    // enqueue-only hosts (sqlp_pool_push_dev + sqlp_epi_build_cuts2_dev) never read the outcome of a push, and
    // duplicates do not grow the pool: without this the host's upper bound K + pending -- and every buffer
    // and grid sized from it -- would grow without limit.  One small read-back every 64 pushes.
    if (p->pending >= 64) pool_confirm(p);
    pool_reserve(p, p->upper() + n);
    p->d_results.ensure((size_t)n * sizeof(PushResult), 0, S(c));
    const double *src = v_dev;
    if (!v_dev || c->world > 1) {
        p->d_vnew.ensure((size_t)n * p->m2 * 8, 0, S(c));
        // one process per GPU: rank 0's vectors count and are broadcast.  One process for all GPUs
        // (sqlp_ctx_create_multi): the host hands the same vectors to every GPU, no collective needed.
        const bool mine = c->world == 1 || c->rank == 0 || c->one_process;
        if (v_host && mine)
            CK(cudaMemcpyAsync(p->d_vnew.p, v_host, (size_t)n * p->m2 * 8, cudaMemcpyHostToDevice, S(c)));
        else if (v_dev && mine)
            CK(cudaMemcpyAsync(p->d_vnew.p, v_dev, (size_t)n * p->m2 * 8, cudaMemcpyDeviceToDevice, S(c)));
        else
            REQUIRE(c->world > 1 && c->rank != 0, SQLP_E_INVALID, "push: null vector");
        if (c->world > 1 && !c->one_process)   // each new dual vertex is broadcast from rank 0 over NCCL/NVLink
            NK(g_nccl.Broadcast(p->d_vnew.p, p->d_vnew.p, (size_t)n * p->m2, ncclFloat64_, 0,
                                c->comm, S(c)));
        src = p->d_vnew.as<double>();
    }
    // algorithmic bytes: the hash scan (8 K per push) + the pushed vector in and, if new, out
    ProfScope prof(c, SQLP_PROF_POOL, (double)n * (8.0 * (double)p->upper() + 16.0 * (double)p->m2));
    // SQLP_PUSH_BATCH vectors per launch (fewer when their 1-norm terms would not fit in shared memory)
    const int64_t per = p->m2 > SQLP_PUSH_SMEM_DOUBLES
        ? 1 : std::max<int64_t>(1, std::min<int64_t>(SQLP_PUSH_BATCH, SQLP_PUSH_SMEM_DOUBLES / p->m2));
    for (int64_t i = 0; i < n; i += per) {
        const int nb = (int)std::min<int64_t>(per, n - i);
        int64_t ku = p->upper() + i;
        int grid = (int)std::min<int64_t>(std::max<int64_t>((ku + 255) / 256, 1), 2 * c->sm_count);
        size_t smem = (size_t)nb * p->m2 <= SQLP_PUSH_SMEM_DOUBLES ? (size_t)nb * p->m2 * 8 : 0;
        LAUNCH(c, k_pool_push, grid, 256, smem, p->d_pi.as<double>(), p->d_hash.as<unsigned long long>(),
               p->d_K.as<long long>(), (int)p->m2, src + i * p->m2, nb, p->d_scratch.as<PushScratch>(),
               p->d_results.as<PushResult>() + i);
    }
    p->pending += n;
    ++p->push_epoch;
}

// Bring a view / an epigraph's (rho, tau) tables up to the current pool contents.
void view_sync(sqlp_pool *p, PoolView *v)
{
    sqlp_ctx *c = p->ctx;
    int64_t hi = p->upper();
    if (v->twins) {
        // classify the new vertices on the device (the device keeps its own "synced" mark, so nothing here
        // depends on which pushes the host has confirmed) and fill the columns of the new classes
        if (v->twin_cap != p->cap) {
            // first use, or the pool's capacity grew: size the tables for it and start over
            size_t tsize = 1;
            while (tsize < (size_t)4 * (size_t)p->cap) tsize <<= 1;
            v->tmask = (unsigned int)(tsize - 1);
            v->d_twin.ensure(sizeof(TwinState), 0, S(c));
            CK(cudaMemsetAsync(v->d_twin.p, 0, sizeof(TwinState), S(c)));
            v->d_act.ensure((size_t)p->cap * 4, 0, S(c), false);
            v->d_hk.ensure((size_t)p->cap * 8, 0, S(c), false);
            v->d_tflag.ensure((size_t)p->cap, 0, S(c), false);
            v->d_tkey.ensure(tsize * 8, 0, S(c), false);
            v->d_trep.ensure(tsize * 4, 0, S(c), false);
            CK(cudaMemsetAsync(v->d_tkey.p, 0xFF, tsize * 8, S(c)));
            CK(cudaMemsetAsync(v->d_trep.p, 0x7F, tsize * 4, S(c)));      // 0x7F7F7F7F: above every pool index
            v->twin_cap = p->cap;
            v->synced_lo = 0;
            v->twin_epoch = -1;
        }
        if (v->twin_epoch != p->push_epoch && hi > 0) {
            // the kernels take their range [synced, K) from the device; the host only needs an upper bound of its length
            const int64_t work = std::max<int64_t>(1, hi - std::min(v->synced_lo, hi));
            const int wgrid = (int)std::min<int64_t>(std::max<int64_t>((work + 7) / 8, 1), 8 * c->sm_count);
            TwinState *st = v->d_twin.as<TwinState>();
            LAUNCH(c, k_twin_hash, wgrid, 256, 0, p->d_pi.as<double>(), (int)p->m2, v->d_rel.as<int>(), (int)v->rel.size(),
                   p->d_K.as<long long>(), (const TwinState *)st, v->d_hk.as<unsigned long long>(),
                   v->d_tkey.as<unsigned long long>(), v->d_trep.as<int>(), v->tmask);
            LAUNCH(c, k_twin_mark, wgrid, 256, 0, p->d_pi.as<double>(), (int)p->m2, v->d_rel.as<int>(), (int)v->rel.size(),
                   p->d_K.as<long long>(), (const TwinState *)st, (const unsigned long long *)v->d_hk.as<unsigned long long>(),
                   (const unsigned long long *)v->d_tkey.as<unsigned long long>(), (const int *)v->d_trep.as<int>(), v->tmask,
                   v->d_tflag.as<unsigned char>());
            LAUNCH(c, k_twin_compact, 1, 1024, 0, p->d_K.as<long long>(), st, (const unsigned char *)v->d_tflag.as<unsigned char>(),
                   v->d_act.as<int>());
            const int64_t fwork = work * v->n_rows;
            const int fgrid = (int)std::min<int64_t>(std::max<int64_t>((fwork + 255) / 256, 1), 8 * c->sm_count);
            LAUNCH(c, k_view_fill, fgrid, 256, 0, p->d_pi.as<double>(), (int)p->m2, v->d_rows.as<int>(), v->n_rows, v->s_pad,
                   v->d_piS.as<double>(), v->d_piR.as<double>(), (const TwinState *)st, (const int *)v->d_act.as<int>());
        }
        v->synced_lo = p->K;                // confirmed vertices: a lower bound of the device's mark
        v->twin_epoch = p->push_epoch;
        return;
    }
    if (hi > v->synced_lo) {
        int64_t work = (hi - v->synced_lo) * v->n_rows;
        int grid = (int)std::min<int64_t>(std::max<int64_t>((work + 255) / 256, 1), 8 * c->sm_count);
        LAUNCH(c, k_view_sync, grid, 256, 0, p->d_pi.as<double>(), (int)p->m2, v->d_rows.as<int>(),
               v->n_rows, v->s_pad, v->d_piS.as<double>(), v->d_piR.as<double>(), (long long)v->synced_lo,
               p->d_K.as<long long>());
    }
    v->synced_lo = p->K;   // only confirmed vertices are final
}

void epi_tables_sync(sqlp_epi *e)
{
    sqlp_pool *p = e->pool;
    sqlp_ctx *c = e->ctx;
    int64_t hi = p->upper();
    if (hi > e->rt_synced_lo) {
        int grid = (int)std::min<int64_t>(hi - e->rt_synced_lo, 16 * c->sm_count);
        const int stage_rho = e->r_nnz > 0 && e->r_nnz <= 4096;
        LAUNCH(c, k_epi_tables, grid, 128, stage_rho ? (size_t)e->r_nnz * 8 : 0, p->d_pi.as<double>(), (int)p->m2,
               e->d_ridx.as<int>(), e->d_rnz.as<double>(), e->r_nnz, e->d_colptr.as<long long>(), e->d_rowval.as<int>(),
               e->d_nzval.as<double>(), (int)e->n1, e->d_rt.as<double>(), (long long)e->rt_synced_lo,
               p->d_K.as<long long>(), stage_rho);
    }
    e->rt_synced_lo = p->K;
}

}  // namespace

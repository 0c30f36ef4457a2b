// host_epi.cuh -- host side of an epigraph: scenario store, contraction plan, cut formation, cut list.
// Part of the single translation unit sqlp_api.cu (included there, in order).
#pragma once

namespace {

DeltaTables delta_tables(sqlp_epi *e)
{
    DeltaTables tb;
    tb.s = (int)e->s;
    tb.n_T = e->n_T;
    tb.n_rows = e->view->n_rows;
    tb.slot_elem = e->d_slot_elem.as<int>();
    tb.t_elem = e->d_t_elem.as<int>();
    tb.elem_base = e->d_elem_base.as<double>();
    tb.out_vals = e->d_ovals.as<double>();
    tb.out_cdf = e->d_ocdf.as<double>();
    tb.out_cnt = e->d_ocnt.as<int>();
    tb.mo = e->mo;
    tb.kind = e->has_kinds ? e->d_kind.as<int>() : nullptr;
    tb.par_a = e->d_par_a.as<double>();
    tb.par_b = e->d_par_b.as<double>();
    return tb;
}

void epi_reserve_scenarios(sqlp_epi *e, int64_t n_local_new)
{
    sqlp_ctx *c = e->ctx;
    int64_t tiles = (n_local_new + SQLP_TILE - 1) / SQLP_TILE;
    if (tiles <= e->cap_tiles) return;
    int64_t ncap = std::max<int64_t>(tiles, std::max<int64_t>(8, e->cap_tiles * 2));
    int64_t used_tiles = (e->n_local + SQLP_TILE - 1) / SQLP_TILE;
    size_t per_tile = (size_t)e->view->s_pad * SQLP_TILE * 8;
    e->d_D.ensure(ncap * per_tile, used_tiles * per_tile, S(c));
    if (!e->n_T) e->d_DR.ensure(ncap * per_tile, used_tiles * per_tile, S(c));
    e->d_w.ensure((size_t)ncap * SQLP_TILE * 8, (size_t)used_tiles * SQLP_TILE * 8, S(c));
    if (e->n_T)
        e->d_dT.ensure((size_t)ncap * SQLP_TILE * e->n_T * 8,
                       (size_t)used_tiles * SQLP_TILE * e->n_T * 8, S(c));
    e->cap_tiles = ncap;
}

// add_scenario! for a batch: values on host (v_host), on device (v_dev) or sampled.
void epi_add(sqlp_epi *e, int64_t n_new, const double *v_host, const double *v_dev,
             const double *w_host, bool sample, uint64_t seed, uint64_t wseed)
{
    if (n_new <= 0) return;
    sqlp_ctx *c = e->ctx;
    const int64_t g0 = e->n_global;
    const int64_t g1 = g0 + n_new;
    const int64_t nl1 = local_count(g1, c->rank, c->world);
    epi_reserve_scenarios(e, nl1);
    // epigraph.jl:89  total_scenario_weight += weight, in scenario order
    for (int64_t i = 0; i < n_new; ++i) {
        double w = 1.0;
        if (sample) { if (wseed) w = 0.5 + u01(wseed, (uint64_t)(g0 + i)); }
        else if (w_host) w = w_host[i];
        e->total_weight += w;
        e->w_absmax = (std::isnan(w) || std::isnan(e->w_absmax)) ? NAN : std::max(e->w_absmax, std::fabs(w));
    }
    DeltaTables tb = delta_tables(e);
    // algorithmic bytes: 8 s in + 8 s out per scenario (sampled: out only), this rank's share
    ProfScope prof(c, SQLP_PROF_DELTA, (double)(nl1 - e->n_local) * e->s * (sample ? 8.0 : 16.0));
    // host values go through a 64 MB staging buffer piece by piece; device-resident or sampled
    // values need no staging, so the whole batch is one launch
    const int64_t piece = (sample || v_dev) ? std::max<int64_t>(SQLP_TILE, (n_new / SQLP_TILE + 2) * SQLP_TILE)
        : std::max<int64_t>(SQLP_TILE, ((int64_t)(64 << 20) / (8 * std::max<int64_t>(e->s, 1))) / SQLP_TILE * SQLP_TILE);
    for (int64_t off = 0; off < n_new;) {
        // cut pieces at 128-aligned global ordinals so a tile is never split mid-copy
        int64_t end = std::min<int64_t>(n_new, ((g0 + off) / SQLP_TILE) * SQLP_TILE + piece - g0);
        if (end <= off) end = std::min<int64_t>(n_new, off + piece);
        const int64_t cnt = end - off;
        const double *vals = nullptr, *wts = nullptr;
        if (!sample) {
            if (v_dev) {
                vals = v_dev + off * e->s;
            } else {
                e->d_stage.ensure((size_t)cnt * (e->s + 1) * 8, 0, S(c), false);
                if (e->s)
                    CK(cudaMemcpyAsync(e->d_stage.p, v_host + off * e->s, (size_t)cnt * e->s * 8,
                                       cudaMemcpyHostToDevice, S(c)));
                vals = e->d_stage.as<double>();
            }
            if (w_host) {
                if (v_dev) e->d_stage.ensure((size_t)cnt * 8, 0, S(c), false);   // weights only
                double *dw = e->d_stage.as<double>() + (v_dev ? 0 : cnt * e->s);
                CK(cudaMemcpyAsync(dw, w_host + off, (size_t)cnt * 8, cudaMemcpyHostToDevice, S(c)));
                wts = dw;
            }
        }
        const int64_t gs = g0 + off;
        const int blocks = 2 * (int)((gs + cnt - 1) / SQLP_TILE - gs / SQLP_TILE + 1);   // half tiles
        const int slab = std::min(e->view->s_pad, SQLP_DELTA_SLAB);
        const size_t dsmem = (size_t)SQLP_DELTA_COLS * delta_stride(slab) * 8;
        if (!c->delta_smem_set) {
            const int mx = SQLP_DELTA_COLS * delta_stride(SQLP_DELTA_SLAB) * 8;
            CK(cudaFuncSetAttribute(k_delta_build<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
            CK(cudaFuncSetAttribute(k_delta_build<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
            c->delta_smem_set = true;
        }
        if (sample)
            LAUNCH(c, k_delta_build<true>, blocks, SQLP_DELTA_THREADS, dsmem, tb, vals, (long long)gs,
                   (long long)cnt, c->rank, c->world, e->view->s_pad, e->d_D.as<double>(),
                   e->d_dT.as<double>(), e->d_w.as<double>(), wts, (unsigned long long)seed,
                   (unsigned long long)wseed, e->d_DR.as<double>());
        else
            LAUNCH(c, k_delta_build<false>, blocks, SQLP_DELTA_THREADS, dsmem, tb, vals, (long long)gs,
                   (long long)cnt, c->rank, c->world, e->view->s_pad, e->d_D.as<double>(),
                   e->d_dT.as<double>(), e->d_w.as<double>(), wts, 0ull, 0ull, e->d_DR.as<double>());
        if (!sample && !v_dev) CK(cudaStreamSynchronize(S(c)));   // staging buffer is reused
        off = end;
    }
    e->n_global = g1;
    e->n_local = nl1;
}

// The contraction variant used in production (see DESIGN.md for the measurements behind it).
template <int NX>
using ContractVariant = ContractCfg<NX, SQLP_VARIANT_MI, SQLP_VARIANT_STAGES, SQLP_VARIANT_PREFETCH, SQLP_VARIANT_CTAS,
                                    SQLP_VARIANT_KG>;

template <int NX>
using ResidentVariant = ResidentCfg<NX, SQLP_RES_WR, SQLP_RES_MI, SQLP_RES_KG, SQLP_RES_CTAS>;

// Resident-scenario kernel: returns false when one unit of scenarios plus a two-stage pool
// ring does not fit in shared memory (very wide stochastic row sets).
template <int NX>
bool launch_contract_resident(sqlp_epi *e, ContractArgs &a)
{
    using Cfg = ResidentVariant<NX>;
    sqlp_ctx *c = e->ctx;
    const size_t fixed = Cfg::fixed_bytes(a.s_pad), stage = Cfg::stage_bytes();
    int ctas = Cfg::CTAS;
    size_t budget = 0;
    int stages = 0;
    for (; ctas >= 1; --ctas) {
        budget = std::min<size_t>((size_t)c->smem_optin, ((size_t)c->smem_per_sm - 1024u * ctas) / ctas);
        stages = budget > fixed ? (int)std::min<size_t>((budget - fixed) / stage, SQLP_RES_MAX_STAGES) : 0;
        if (stages >= 2) break;
    }
    if (stages < 2) return false;
    const size_t smem = fixed + stage * stages;
    // The attribute belongs to the function on this device, not to the context: every context sets the same
    // (maximal) value, so a context created later can never lower what an earlier one relies on.
    if (!c->res_smem_set[NX]) {
        CK(cudaFuncSetAttribute(k_contract_resident<Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize, c->smem_optin));
        c->res_smem_set[NX] = 1;
    }
    const long long nunits = (long long)Cfg::UNITS_PER_TILE * a.ntiles;
    const long long nchunks_ub = (e->pool->upper() + SQLP_TILE - 1) / SQLP_TILE;   // host upper bound
    int grid = ctas * c->sm_count;
    if (c->contract_grid > 0) grid = c->contract_grid;
    grid = (int)std::max<long long>(1, std::min<long long>(grid, std::max<long long>(nunits * nchunks_ub, nunits)));
    const size_t pieces = (size_t)grid * 2 * NX * Cfg::ROWS;
    c->d_piece_val.ensure(pieces * 8, 0, S(c), false);
    c->d_piece_idx.ensure(pieces * 4, 0, S(c), false);
    a.nstages = stages;
    a.prefetch = std::max(1, stages - 2);
    if (c->contract_prefetch > 0) a.prefetch = std::min(c->contract_prefetch, stages - 1);
    a.piece_val = c->d_piece_val.as<double>();
    a.piece_idx = c->d_piece_idx.as<int>();
    LAUNCH(c, k_contract_resident<Cfg>, grid, Cfg::THREADS, smem, a);
    auto fixup = k_argmax_fixup<Cfg::ROWS, NX>;
    LAUNCH(c, fixup, grid, 128, 0, a, grid);
    return true;
}

// Warp-specialised kernel (one CTA per SM, producer warp + two consumer row groups).
template <int NX>
bool launch_contract_ws(sqlp_epi *e, ContractArgs &a)
{
    using Cfg = WsCfg<NX, SQLP_WS_KG>;
    sqlp_ctx *c = e->ctx;
    const size_t fixed = Cfg::fixed_bytes(a.s_pad), stage = Cfg::stage_bytes();
    const size_t budget = std::min<size_t>((size_t)c->smem_optin, (size_t)c->smem_per_sm - 1024u);
    const int stages = budget > fixed ? (int)std::min<size_t>((budget - fixed) / stage, SQLP_RES_MAX_STAGES) : 0;
    if (stages < 3) return false;
    const size_t smem = fixed + stage * stages;
    if (!c->ws_smem_set[NX]) {   // the same maximal value from every context (see launch_contract_resident)
        CK(cudaFuncSetAttribute(k_contract_ws<Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize, c->smem_optin));
        c->ws_smem_set[NX] = 1;
    }
    const long long nunits = (long long)Cfg::UNITS_PER_TILE * a.ntiles;
    const long long nchunks_ub = (e->pool->upper() + SQLP_TILE - 1) / SQLP_TILE;
    int grid = c->contract_grid > 0 ? c->contract_grid : c->sm_count;
    grid = (int)std::max<long long>(1, std::min<long long>(grid, std::max<long long>(nunits * nchunks_ub, nunits)));
    const size_t pieces = (size_t)grid * 2 * NX * Cfg::ROWS;
    c->d_piece_val.ensure(pieces * 8, 0, S(c), false);
    c->d_piece_idx.ensure(pieces * 4, 0, S(c), false);
    a.nstages = stages;
    a.prefetch = 0;
    a.lag_ns = c->contract_lag_ns;
    a.piece_val = c->d_piece_val.as<double>();
    a.piece_idx = c->d_piece_idx.as<int>();
    LAUNCH(c, k_contract_ws<Cfg>, grid, SQLP_WS_THREADS, smem, a);
    auto fixup = k_argmax_fixup<Cfg::ROWS, NX>;
    LAUNCH(c, fixup, grid, 128, 0, a, grid);
    return true;
}

template <int NX>
void launch_contract(sqlp_epi *e, const double *D, const double *bias, double *bv, int *bi,
                     const ScreenCtl *gate = nullptr)
{
    using Cfg = ContractVariant<NX>;
    sqlp_ctx *c = e->ctx;
    ContractArgs a;
    a.D = D;
    a.PiS = e->view->d_piS.as<double>();
    a.bias = bias;
    a.bias_stride = e->cur_bias_stride;
    a.d_K = e->view->d_Kv(e->pool);             // columns of the view (= pool size without twins)
    a.s_pad = e->view->s_pad;
    a.ntiles = (int)((e->n_local + SQLP_TILE - 1) / SQLP_TILE);
    a.n_local = e->n_local;
    a.best_val = bv;
    a.best_idx = bi;
    a.out_stride = e->out_stride;
    a.nstages = a.prefetch = a.lag_ns = 0;
    a.piece_val = nullptr;
    a.piece_idx = nullptr;
    a.gate = gate;
    // behind a screening pass the sweep only runs if that pass fell back: its time goes to its own class
    ProfScope prof(c, gate ? SQLP_PROF_FALLBACK : SQLP_PROF_CONTRACT,
                   gate ? 0.0 : 2.0 * (double)e->view->n_rows * (double)e->n_local, gate ? nullptr : e->pool, 1, e->view);
    bool done = false;
    // automatic: warp-specialised -> resident (fewer ring stages suffice) -> streaming (any s_pad)
    if (c->contract_mode == 0 || c->contract_mode == 3) done = launch_contract_ws<NX>(e, a);
    if (!done && c->contract_mode != 1) done = launch_contract_resident<NX>(e, a);

    if (!done) {   // streaming kernel: both operands flow through the ring
        size_t smem = Cfg::smem_bytes();
        if (!c->smem_attr[NX]) {   // per device, so per context
            CK(cudaFuncSetAttribute(k_contract_argmax<Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)smem));
            c->smem_attr[NX] = true;
        }
        const long long nunits = (long long)Cfg::UNITS_PER_TILE * a.ntiles;
        int grid = (int)std::min<long long>(nunits, (long long)Cfg::CTAS * c->sm_count);
        LAUNCH(c, k_contract_argmax<Cfg>, grid, SQLP_CT_THREADS, smem, a);
    }
    prof.stop();
}

// ---- screening pass (kernels_screen.cuh) --------------------------------------------------------
// What the host has learnt from the control block of an earlier pass (copied back without a synchronisation).
void screen_learn(sqlp_epi *e)
{
    if (!e->ctl_pending || cudaEventQuery(e->ctl_event) != cudaSuccess) return;
    e->ctl_pending = false;
    const ScreenCtl &h = *e->h_ctl.as<ScreenCtl>();
    e->scr_last = h;
    ++e->scr_runs;
    // A pass that succeeded may still not have paid: every exact evaluation is a gather of two operand rows
    // (measured on B200: the cost of ~50 pairs of the FP64 sweep), and a pool of near-ties -- storm's real duals
    // leave ~23 candidates per scenario and point among 3 600 classes -- makes the pass as slow as the sweep it
    // stands in for.  Past eight evaluations per scenario-point the sweep is the better plan.
    const bool unprofitable = (double)h.n_eval > 8.0 * (double)e->n_local * (double)std::max(1, e->scr_nx);
    if (unprofitable) ++e->scr_unprofitable;
    if (h.bad != 0 || h.overflow > h.ovf_limit || unprofitable) {
        // it fell back to the FP64 sweep (non-finite operands, or candidate lists overflowing: a pool of
        // near-ties such as storm's real duals): leave the pass out for a while, longer every time
        ++e->scr_fallbacks;
        e->scr_backoff = std::min(1024, std::max(16, e->scr_backoff * 2));
        e->scr_skip = e->scr_backoff;
    } else {
        e->scr_backoff = 0;
    }
}

// Is the pass worth trying for this call?  Needs: no random element in Tbar (d does not depend on x), operands
// that fit the kernel's shared memory, and enough work to amortise its fixed cost.
bool screen_wanted(sqlp_epi *e, int NX)
{
    sqlp_ctx *c = e->ctx;
    if (c->screen_mode == 0 || e->n_T != 0 || e->n_local == 0 || e->view->n_rows == 0) return false;
    const int sp = (int)round_up(e->view->n_rows, 16);
    if (scr_smem_bytes(sp, 3, NX) > (size_t)c->smem_optin) return false;
    screen_learn(e);
    if (c->screen_mode == 2) return true;
    if (e->scr_skip > 0) { --e->scr_skip; return false; }
    const int64_t K = e->pool->upper();
    return K >= 1024 && e->n_local >= 1024 && (double)K * (double)e->n_local >= (double)(1 << 24);
}

void screen_sync_operands(sqlp_epi *e)
{
    sqlp_ctx *c = e->ctx;
    sqlp_pool *p = e->pool;
    PoolView *v = e->view;
    if (!v->sp) {
        v->sp = (int)round_up(v->n_rows, 16);
        v->d_vbad.ensure(16, 0, S(c));
    }
    const int J = v->sp / 16;
    const int64_t hi = p->upper();
    if (hi > v->scr_cap) {
        const int64_t ncap = round_up(std::max<int64_t>(hi, std::max<int64_t>(1024, v->scr_cap * 2)), SCR_NB);
        const size_t per_chunk = (size_t)J * SCR_STAGE_BYTES;
        v->d_piB.ensure((size_t)(ncap / SCR_NB) * per_chunk, (size_t)(v->scr_cap / SCR_NB) * per_chunk, S(c));
        v->d_pn.ensure((size_t)ncap * 4, (size_t)v->scr_cap * 4, S(c));
        v->d_pnmax.ensure((size_t)(ncap / SCR_NB) * 4, (size_t)(v->scr_cap / SCR_NB) * 4, S(c));
        v->scr_cap = ncap;
    }
    v->d_scr_lo.ensure(16, 0, S(c));         // device-side mark: view columns below it are final in d_piB
    v->d_ctr.ensure((size_t)(v->sp + 1) * 8, 0, S(c));
    // The centre (kernels_screen.cuh, "Centred operands"): mean of the first <= 1 024 columns, taken again -- and
    // every column split again -- when four times as many are there.  Any centre is a correct one.
    const int64_t ctr_want = std::min<int64_t>(hi, SCR_CENTRE_COLS);
    if (c->screen_centre && ctr_want > 0 && ctr_want >= 4 * v->ctr_cols && v->ctr_cols < SCR_CENTRE_COLS) {
        LAUNCH(c, k_screen_centre, 1, 256, 0, p->d_pi.as<double>(), (int)p->m2, v->d_rows.as<int>(), v->n_rows, v->sp,
               v->act(), v->d_Kv(p), (long long)ctr_want, v->d_ctr.as<double>());
        CK(cudaMemsetAsync(v->d_scr_lo.p, 0, 8, S(c)));
        CK(cudaMemsetAsync(v->d_pnmax.p, 0, (size_t)(v->scr_cap / SCR_NB) * 4, S(c)));
        v->ctr_cols = ctr_want;
        ++v->ctr_epoch;
        v->scr_synced_lo = 0;
        v->scr_epoch = -1;
    }
    if (v->scr_epoch != p->push_epoch && hi > 0) {
        const int64_t work = std::max<int64_t>(1, hi - std::min(v->scr_synced_lo, hi));
        const int grid = (int)std::min<int64_t>(std::max<int64_t>((work + 7) / 8, 1), 8 * c->sm_count);
        LAUNCH(c, k_screen_view_sync, grid, 256, 0, p->d_pi.as<double>(), (int)p->m2, v->d_rows.as<int>(), v->n_rows, v->sp,
               v->d_piB.as<__nv_bfloat16>(), v->d_pn.as<float>(), v->d_pnmax.as<float>(), v->d_vbad.as<int>(),
               v->d_scr_lo.as<long long>(), v->d_Kv(p), v->act(), (const double *)v->d_ctr.as<double>());
        LAUNCH(c, k_screen_mark, 1, 32, 0, v->d_scr_lo.as<long long>(), v->d_Kv(p));
    }
    v->scr_synced_lo = p->K;                // confirmed vertices: a lower bound of the device's mark (in pool slots)
    v->scr_epoch = p->push_epoch;
    const int64_t units = (e->n_local + SCR_UNIT - 1) / SCR_UNIT;
    if (units > e->scr_units_cap) {
        const int64_t ncap = std::max<int64_t>(units, std::max<int64_t>(8, e->scr_units_cap * 2));
        e->d_DB.ensure((size_t)ncap * 512 * v->sp, (size_t)e->scr_units_cap * 512 * v->sp, S(c));
        e->d_dnu.ensure((size_t)ncap * 4, (size_t)e->scr_units_cap * 4, S(c));
        e->d_dn.ensure((size_t)ncap * SCR_UNIT * 4, (size_t)e->scr_units_cap * SCR_UNIT * 4, S(c));
        e->d_dnall.ensure(16, 0, S(c));
        e->d_ebad.ensure(16, 0, S(c));
        e->scr_units_cap = ncap;
    }
    e->d_dbar.ensure((size_t)(v->sp + 1) * 8, 0, S(c));
    const int64_t dbar_want = std::min<int64_t>(e->n_local, SCR_CENTRE_SCEN);
    if (c->screen_centre && dbar_want >= 4 * e->dbar_n && e->dbar_n < SCR_CENTRE_SCEN) {     // as for the view's centre
        LAUNCH(c, k_screen_dbar, 1, 256, 0, e->d_D.as<double>(), v->s_pad, v->sp, (long long)dbar_want,
               e->d_dbar.as<double>());
        CK(cudaMemsetAsync(e->d_dnu.p, 0, (size_t)e->scr_units_cap * 4, S(c)));
        CK(cudaMemsetAsync(e->d_dnall.p, 0, 4, S(c)));
        e->dbar_n = dbar_want;
        e->scr_synced = 0;
    }
    if (e->n_local > e->scr_synced) {
        const int64_t work = e->n_local - e->scr_synced;
        const int grid = (int)std::min<int64_t>(std::max<int64_t>((work + 7) / 8, 1), 16 * c->sm_count);
        LAUNCH(c, k_screen_scen_sync, grid, 256, 0, e->d_D.as<double>(), v->s_pad, v->sp, e->d_DB.as<__nv_bfloat16>(),
               e->d_dnu.as<float>(), e->d_dnall.as<float>(), e->d_ebad.as<int>(), (long long)e->scr_synced,
               (long long)e->n_local, (const double *)e->d_dbar.as<double>(), e->d_dn.as<float>());
        e->scr_synced = e->n_local;
    }
}

// Screening + exact decision for NX points, enqueued behind k_bias.  Returns the control block the FP64
// sweep is gated on (it still has to be launched: it runs if, and only if, this pass fell back).
template <int NX>
const ScreenCtl *screen_enqueue(sqlp_epi *e)
{
    sqlp_ctx *c = e->ctx;
    sqlp_pool *p = e->pool;
    PoolView *v = e->view;
    screen_sync_operands(e);
    const int64_t ku = p->upper();
    const int64_t nch = (ku + SCR_NB - 1) / SCR_NB;
    const int64_t nunits = (e->n_local + SCR_UNIT - 1) / SCR_UNIT, npad = nunits * SCR_UNIT;
    constexpr int BF = scr_bias_floats<NX>();
    e->d_b32c.ensure((size_t)std::max<int64_t>(nch, 1) * BF * 4, 0, S(c), false);
    e->d_ctl.ensure(sizeof(ScreenCtl), 0, S(c));
    // split a unit's sweep into K-ranges when there are too few units to fill the GPU
    // (one unit per SM or more: no split -- the row-staging decision and the warm start need whole sweeps, and both
    // are worth more than the last quarter of a wave)
    int R = nunits >= c->sm_count ? 1
            : (int)std::min<int64_t>(std::max<int64_t>(1, (2 * c->sm_count + nunits - 1) / nunits), std::max<int64_t>(nch, 1));
    const size_t nslots = (size_t)NX * R * 2 * npad;
    c->d_cand.ensure(nslots * SCR_CAP * sizeof(int2), 0, S(c), false);
    c->d_cnt.ensure(nslots * 4, 0, S(c), false);
    c->d_lfin.ensure(nslots * 4, 0, S(c), false);
    // more overflowing lists than this and the pass gives up: each one costs a full sweep of its scenario
    const unsigned ovf_limit = (unsigned)std::min<int64_t>(e->n_local / 16 + 16, 1 << 20);
    e->d_pdb.ensure((size_t)std::max<int64_t>(nch, 1) * SCR_NB * 8, 0, S(c), false);
    LAUNCH(c, k_screen_pdbar, (int)std::min<int64_t>(std::max<int64_t>((ku + 7) / 8, 1), 8 * c->sm_count), 256, 0,
           (const double *)v->d_piR.as<double>(), v->s_pad, (const double *)e->d_dbar.as<double>(), v->d_Kv(p),
           e->d_pdb.as<double>());
    LAUNCH(c, k_screen_prep<NX>, 1, 1024, 0, e->cur_bias, (long long)e->cur_bias_stride, v->d_pn.as<float>(),
           v->d_pnmax.as<float>(), e->d_dnall.as<float>(), v->d_vbad.as<int>(), e->d_ebad.as<int>(),
           v->d_Kv(p), v->sp, ovf_limit, e->d_b32c.as<float>(), e->d_ctl.as<ScreenCtl>(),
           (const double *)e->d_pdb.as<double>(), (const double *)v->d_ctr.as<double>(),
           (const double *)e->d_dbar.as<double>());
    // warm start of the scan from the winners of the previous pass (k_screen_seed)
    const float *lseed = nullptr;
    const bool rows_decide = c->resolve_rows && e->d_DR.p && scr_decide_smem(v->s_pad) <= (size_t)c->smem_optin;
    const bool seeding = c->screen_seed && rows_decide;
    if (seeding) {
        if (npad > e->prev_cap) {
            const int64_t ncap = std::max<int64_t>(npad, e->prev_cap * 2);
            e->d_prev.ensure((size_t)ncap * 8, (size_t)e->prev_cap * 8, S(c));
            e->d_prevdot.ensure((size_t)ncap * 16, (size_t)e->prev_cap * 16, S(c));
            e->d_cd.ensure((size_t)ncap * 8, (size_t)e->prev_cap * 8, S(c));
            e->prev_cap = ncap;
        }
        if (e->cd_epoch != v->ctr_epoch) { e->cd_synced = 0; e->cd_epoch = v->ctr_epoch; }
        if (e->n_local > e->cd_synced) {
            const int64_t work = e->n_local - e->cd_synced;
            LAUNCH(c, k_screen_cd, (int)std::min<int64_t>((work + 7) / 8, 16 * c->sm_count), 256, 0,
                   (const double *)e->d_DR.as<double>(), v->s_pad, (const double *)v->d_ctr.as<double>(),
                   (long long)e->cd_synced, (long long)e->n_local, e->d_cd.as<double>());
            e->cd_synced = e->n_local;
        }
        if (e->prev_valid) {
            c->d_lseed.ensure((size_t)NX * npad * 4, 0, S(c), false);
            SeedArgs sd;
            sd.bias = e->cur_bias;
            sd.bias_stride = e->cur_bias_stride;
            sd.prev = e->d_prev.as<int>();
            sd.prevdot = e->d_prevdot.as<double>();
            sd.cd = e->d_cd.as<double>();
            sd.d_K = v->d_Kv(p);
            sd.ctl = e->d_ctl.as<ScreenCtl>();
            sd.n_local = e->n_local;
            sd.npad = npad;
            sd.lseed = c->d_lseed.as<float>();
            LAUNCH(c, k_screen_seed<NX>, (int)std::min<int64_t>((npad + 255) / 256, 8 * c->sm_count), 256, 0, sd);
            lseed = c->d_lseed.as<float>();
        }
    }
    int nstages = SCR_MAX_STAGES;
    while (nstages > 3 && scr_smem_bytes(v->sp, nstages, NX) > (size_t)c->smem_optin) --nstages;
    const size_t smem = scr_smem_bytes(v->sp, nstages, NX);
    if (!c->screen_smem_set[NX]) {
        CK(cudaFuncSetAttribute((k_screen<NX, 0>), cudaFuncAttributeMaxDynamicSharedMemorySize, c->smem_optin));
        CK(cudaFuncSetAttribute((k_screen<NX, 1>), cudaFuncAttributeMaxDynamicSharedMemorySize, c->smem_optin));
        c->screen_smem_set[NX] = true;
    }
    ScreenArgs sa;
    sa.DB = e->d_DB.as<__nv_bfloat16>();
    sa.PiB = v->d_piB.as<__nv_bfloat16>();
    sa.b32c = e->d_b32c.as<float>();
    sa.dnmax_unit = e->d_dnu.as<float>();
    sa.dn = e->d_dn.as<float>();
    sa.ctl = e->d_ctl.as<ScreenCtl>();
    sa.d_K = v->d_Kv(p);
    sa.sp = v->sp;
    sa.nunits = (int)nunits;
    sa.R = R;
    sa.nstages = nstages;
    sa.n_local = e->n_local;
    sa.npad = npad;
    sa.cand = c->d_cand.as<int2>();
    sa.cnt = c->d_cnt.as<int>();
    sa.lfin = c->d_lfin.as<float>();
    sa.lseed = lseed;
    sa.dbg = nullptr;
    sa.desc_mode = 0;
    const int grid = (int)std::min<int64_t>(c->sm_count, nunits * R);
    {
        // executed bf16 flops: three products over sp slots for every (vertex of a whole chunk, scenario of a whole unit)
        ProfScope prof(c, SQLP_PROF_SCREEN, 6.0 * v->sp * (double)npad, p, SCR_NB, v);
        if (c->screen_fadd2) LAUNCH(c, (k_screen<NX, 1>), grid, SCR_THREADS, smem, sa);
        else LAUNCH(c, (k_screen<NX, 0>), grid, SCR_THREADS, smem, sa);
    }
    ResolveArgs ra;
    ra.D = e->d_D.as<double>();
    ra.PiS = v->d_piS.as<double>();
    ra.PiR = v->d_piR.as<double>();
    ra.bias = e->cur_bias;
    ra.bias_stride = e->cur_bias_stride;
    ra.s_pad = v->s_pad;
    ra.d_K = v->d_Kv(p);
    ra.n_local = e->n_local;
    ra.npad = npad;
    ra.R = R;
    ra.cand = c->d_cand.as<int2>();
    ra.cnt = c->d_cnt.as<int>();
    ra.lfin = c->d_lfin.as<float>();
    ra.best_val = e->d_best_val.as<double>();
    ra.best_idx = e->d_best_idx.as<int>();
    ra.out_stride = e->out_stride;
    ra.ctl = e->d_ctl.as<ScreenCtl>();
    ra.force_full = 0;
    ra.DR = e->d_DR.as<double>();
    ra.prev = nullptr;
    ra.prevdot = nullptr;
    {
        ProfScope prof(c, SQLP_PROF_RESOLVE, (double)NX * (double)e->n_local);
        if (rows_decide && R == 1) {
            // whole rows through shared memory (k_screen_decide); it also leaves the winners for the next warm start
            const size_t dsmem = scr_decide_smem(v->s_pad);
            if (dsmem > c->decide_smem_set[NX]) {
                CK(cudaFuncSetAttribute(k_screen_decide<NX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsmem));
                c->decide_smem_set[NX] = dsmem;
            }
            if (seeding) {
                ra.prev = e->d_prev.as<int>();
                ra.prevdot = e->d_prevdot.as<double>();
                e->prev_valid = true;
            }
            const int rgrid = (int)std::min<int64_t>(std::max<int64_t>((e->n_local + SCR_DEC_WARPS - 1) / SCR_DEC_WARPS, 1),
                                                     20 * c->sm_count);
            LAUNCH(c, k_screen_decide<NX>, rgrid, 32 * SCR_DEC_WARPS, dsmem, ra);
        } else {
            const int rgrid = (int)std::min<int64_t>(std::max<int64_t>((e->n_local + 7) / 8, 1), 8 * c->sm_count);
            if (c->resolve_fma) LAUNCH(c, (k_screen_resolve<NX, 1>), rgrid, 256, 0, ra);
            else LAUNCH(c, (k_screen_resolve<NX, 0>), rgrid, 256, 0, ra);
        }
    }
    e->scr_nx = NX;
    // the control block goes back to the host asynchronously; screen_learn() reads it at a later call
    if (!e->ctl_event) CK(cudaEventCreateWithFlags(&e->ctl_event, cudaEventDisableTiming));
    if (!e->ctl_pending) {
        e->h_ctl.ensure(sizeof(ScreenCtl));
        CK(cudaMemcpyAsync(e->h_ctl.p, e->d_ctl.p, sizeof(ScreenCtl), cudaMemcpyDeviceToHost, S(c)));
        CK(cudaEventRecord(e->ctl_event, S(c)));
        e->ctl_pending = true;
    }
    return e->d_ctl.as<ScreenCtl>();
}

// Everything of build_sasa_cut for NX points, enqueued on the stream.  x on host or device.
// Result lands in e->d_out as [NX][n1 + 2] = (alpha, beta[n1], val).
// bias_from: an epigraph of the same call with the same template, pool and points whose bias vectors are
// reused instead of recomputed.  cell_slot (sharded jobs): where this epigraph's row (NX * (n1 + 2) doubles + the
// flag word) goes in the cell's gather buffer; the caller then runs cell_gather() once for the whole cell.
void epi_cuts_enqueue(sqlp_epi *e, int NX, const double *x_host, const double *x_dev, bool want_cut,
                      sqlp_epi *bias_from = nullptr, double *cell_slot = nullptr)
{
    sqlp_ctx *c = e->ctx;
    sqlp_pool *p = e->pool;
    const int n1 = (int)e->n1, m2 = (int)e->m2;
    const int NC = n1 + 2;
    e->d_x2.ensure((size_t)2 * std::max(n1, 1) * 8, 0, S(c));
    e->d_out.ensure((size_t)2 * NC * 8, 0, S(c));
    e->d_flags.ensure(16, 0, S(c));
    if (x_host)
        CK(cudaMemcpyAsync(e->d_x2.p, x_host, (size_t)NX * n1 * 8, cudaMemcpyHostToDevice, S(c)));
    else
        CK(cudaMemcpyAsync(e->d_x2.p, x_dev, (size_t)NX * n1 * 8, cudaMemcpyDeviceToDevice, S(c)));
    // flags are cleared by k_base, the first kernel of the chain, and the reduction writes every entry of
    // d_out: the two memsets are only needed when there is nothing to launch
    if (e->n_local == 0) {
        CK(cudaMemsetAsync(e->d_flags.p, 0, 4, S(c)));
        CK(cudaMemsetAsync(e->d_out.p, 0, (size_t)2 * NC * 8, S(c)));
    }

    view_sync(p, e->view);
    if (want_cut) epi_tables_sync(e);

    const int64_t ku = p->upper();
    const int64_t kpad = round_up(std::max<int64_t>(ku, 1), SQLP_TILE);
    if (kpad > e->bias_stride) {
        e->bias_stride = round_up(kpad * 2, SQLP_TILE);
        e->d_bias.ensure((size_t)2 * e->bias_stride * 8, 0, S(c), false);
    }
    const int64_t ntiles = (e->n_local + SQLP_TILE - 1) / SQLP_TILE;
    if (ntiles * SQLP_TILE > e->out_stride) {
        e->out_stride = round_up(ntiles * SQLP_TILE * 2, SQLP_TILE);
        e->d_best_val.ensure((size_t)2 * e->out_stride * 8, 0, S(c), false);
        e->d_best_idx.ensure((size_t)2 * e->out_stride * 4, 0, S(c), false);
    }
    e->d_base.ensure((size_t)2 * m2 * 8, 0, S(c));

    if (ntiles > 0 && bias_from && bias_from->bias_stride >= kpad) {
        // the same template, pool and points as an earlier epigraph of this call: its bias vectors are ours.
        // k_base, the kernel that clears the flag word, is skipped with them.
        CK(cudaMemsetAsync(e->d_flags.p, 0, 4, S(c)));
        e->cur_bias = bias_from->cur_bias;
        e->cur_bias_stride = bias_from->cur_bias_stride;
    } else if (ntiles > 0) {
        e->cur_bias = e->d_bias.as<double>();
        e->cur_bias_stride = e->bias_stride;
        ProfScope prof_bias(c, SQLP_PROF_BIAS, 8.0 * (double)ku * m2 + 8.0 * NX * (double)ku);
        BaseArgs ba{e->d_rbar.as<double>(), e->d_rptr.as<int>(), e->d_rcol.as<int>(), e->d_rval.as<double>(),
                    e->d_x2.as<double>(), n1, e->d_flags.as<int>()};
        const int bgrid = (int)((kpad + 7) / 8);
        const size_t base_smem = (size_t)NX * m2 * 8;
        if (base_smem <= 40 * 1024) {
            // base_x = rbar - Tbar x is rebuilt by every block of k_bias in shared memory: one launch less
            if (NX == 2)
                LAUNCH(c, (k_bias<2, true>), bgrid, 256, base_smem, p->d_pi.as<double>(), m2, (const double *)nullptr,
                       e->view->d_Kv(p), (long long)kpad, e->d_bias.as<double>(), (long long)e->bias_stride, e->view->act(), ba);
            else
                LAUNCH(c, (k_bias<1, true>), bgrid, 256, base_smem, p->d_pi.as<double>(), m2, (const double *)nullptr,
                       e->view->d_Kv(p), (long long)kpad, e->d_bias.as<double>(), (long long)e->bias_stride, e->view->act(), ba);
        } else {
            LAUNCH(c, k_base, dim3((m2 + 127) / 128, NX), 128, 0, e->d_rbar.as<double>(), m2, n1,
                   e->d_rptr.as<int>(), e->d_rcol.as<int>(), e->d_rval.as<double>(),
                   e->d_x2.as<double>(), e->d_base.as<double>(), e->d_flags.as<int>());
            if (NX == 2)
                LAUNCH(c, (k_bias<2, false>), bgrid, 256, 0, p->d_pi.as<double>(), m2, (const double *)e->d_base.as<double>(),
                       e->view->d_Kv(p), (long long)kpad, e->d_bias.as<double>(), (long long)e->bias_stride, e->view->act(), ba);
            else
                LAUNCH(c, (k_bias<1, false>), bgrid, 256, 0, p->d_pi.as<double>(), m2, (const double *)e->d_base.as<double>(),
                       e->view->d_Kv(p), (long long)kpad, e->d_bias.as<double>(), (long long)e->bias_stride, e->view->act(), ba);
        }

        prof_bias.stop();
    }
    if (ntiles > 0) {

        if (e->n_T == 0) {
            // screening on the 5th-generation tensor cores + exact decision among the candidates; the FP64 sweep
            // is launched behind it and runs only if the pass fell back (a device-side decision, no host round trip)
            const ScreenCtl *gate = nullptr;
            if (screen_wanted(e, NX)) gate = NX == 2 ? screen_enqueue<2>(e) : screen_enqueue<1>(e);
            if (NX == 2)
                launch_contract<2>(e, e->d_D.as<double>(), e->cur_bias,
                                   e->d_best_val.as<double>(), e->d_best_idx.as<int>(), gate);
            else
                launch_contract<1>(e, e->d_D.as<double>(), e->cur_bias,
                                   e->d_best_val.as<double>(), e->d_best_idx.as<int>(), gate);
        } else {
            // some element perturbs Tbar: d(x) = delta_rhs - delta_T x is rebuilt per point
            size_t bytes = (size_t)ntiles * e->view->s_pad * SQLP_TILE * 8;
            e->d_Dx.ensure(bytes, 0, S(c), false);
            TransferList tl{e->n_T, e->d_tj.as<int>(), e->d_tcol.as<int>(), e->d_tslot.as<int>()};
            for (int x = 0; x < NX; ++x) {
                CK(cudaMemcpyAsync(e->d_Dx.p, e->d_D.p, bytes, cudaMemcpyDeviceToDevice, S(c)));
                LAUNCH(c, k_delta_x, (int)((e->n_local + 255) / 256), 256, 0, tl,
                       e->d_x2.as<double>() + (size_t)x * n1, (long long)e->n_local, e->view->s_pad,
                       e->d_D.as<double>(), e->d_dT.as<double>(), e->d_Dx.as<double>());
                launch_contract<1>(e, e->d_Dx.as<double>(), e->cur_bias + x * e->cur_bias_stride,
                                   e->d_best_val.as<double>() + x * e->out_stride,
                                   e->d_best_idx.as<int>() + x * e->out_stride);
            }
        }
    }
    if (!want_cut) return;
    e->last_nx = NX;

    const int width = NX * NC;
    if (ntiles > 0) {
        e->d_partial.ensure((size_t)ntiles * width * 8, 0, S(c), false);
        ReduceArgs r;
        r.D = e->d_D.as<double>();
        r.dT = e->d_dT.as<double>();
        r.w = e->d_w.as<double>();
        r.PiS = e->view->d_piS.as<double>();
        r.rt = e->d_rt.as<double>();
        r.act = e->view->act();
        r.bias = e->n_T == 0 ? e->cur_bias : nullptr;
        r.bias_stride = e->cur_bias_stride;
        r.best_val = e->d_best_val.as<double>();
        r.best_idx = e->d_best_idx.as<int>();
        r.out_stride = e->out_stride;
        r.n_local = e->n_local;
        r.s_pad = e->view->s_pad;
        r.n_rows = e->view->n_rows;
        r.n1 = n1;
        r.total_weight = e->total_weight;
        r.n_T = e->n_T;
        r.tc_col = e->d_cc.as<int>();
        r.tc_j = e->d_cj.as<int>();
        r.tc_slot = e->d_cslot.as<int>();
        r.partial = e->d_partial.as<double>();
        r.flags = e->d_flags.as<int>();
        // algorithmic bytes: per point N (idx + weight + winning dot) + the (rho, tau)
        // table + the cut (with delta_T != 0 the kernel also re-reads D to recompute the winning dot)
        ProfScope prof_red(c, SQLP_PROF_REDUCE,
                           NX * (24.0 * (double)e->n_local + 8.0 * (double)ku * (n1 + 1) + 8.0 * (n1 + 1)));
        // tiles and groups are split into 8 sub-ranges summed side by side when the sub-sums fit in the
        // default 48 KB of dynamic shared memory (n1 <= 382 for two points), else not at all
        const int nsub = (size_t)8 * width * 8 <= 48 * 1024 ? 8 : 1;
        const size_t sub_smem = nsub > 1 ? (size_t)nsub * width * 8 : 0;
        r.nsub = nsub;
        // sharded job: the row and its flag word go straight into the cell's gather buffer
        double *fin = cell_slot ? cell_slot : e->d_out.as<double>();
        const int group = 64;
        // many scenarios, no random element in Tbar: regroup by the selected vertex (k_cut_hist / k_cut_fold) so that
        // the (rho, tau) table is read once instead of one row per scenario and point
        const int64_t kc = round_up(std::max<int64_t>(ku, 1), 32);
        const size_t hist_smem = (size_t)kc * 8 + (size_t)SQLP_HIST_SUB * 12;
        if (e->n_T == 0 && c->reduce_mode != 1 && (c->reduce_mode == 2 || e->n_local >= 131072) &&
            hist_smem + 1024 <= (size_t)c->smem_optin) {
            // fixed-point weight sums (k_cut_hist_fx) when the scale exists: |p_i| <= max|w| / |total| < 2^eb
            int sh1 = 0;
            const size_t fx_cap = (size_t)c->smem_optin - 2048;
            bool fx = c->hist_fx && (size_t)kc * 12 <= fx_cap;
            if (fx) {
                const double pb = e->w_absmax / std::fabs(e->total_weight);
                fx = std::isfinite(pb) && pb > 0.0;
                if (fx) {
                    int eb = 0;
                    std::frexp(pb, &eb);
                    sh1 = 60 - eb;
                    fx = std::abs(sh1) < 900;
                }
            }
            const size_t fx_smem = std::min<size_t>((size_t)NX * kc * 12, fx_cap);
            const int64_t nblk = fx ? std::max<int64_t>(1, (e->n_local + SQLP_HISTFX_SEG - 1) / SQLP_HISTFX_SEG)
                                    : std::max<int64_t>(1, std::min<int64_t>(c->sm_count, (e->n_local + 4095) / 4096));
            HistArgs h;
            h.w = r.w; h.rt = r.rt; h.act = r.act; h.bias = r.bias; h.bias_stride = r.bias_stride;
            h.best_val = r.best_val; h.best_idx = r.best_idx; h.out_stride = r.out_stride; h.n_local = r.n_local;
            h.seg = fx ? SQLP_HISTFX_SEG : round_up((e->n_local + nblk - 1) / nblk, SQLP_TILE);
            h.d_Kv = e->view->d_Kv(p);
            h.kc = (int)kc; h.n1 = n1; h.total_weight = r.total_weight;
            h.D = r.D; h.PiS = r.PiS; h.s_pad = r.s_pad; h.flags = r.flags;
            h.hfx = nullptr;
            h.sc1 = std::ldexp(1.0, sh1); h.inv1 = std::ldexp(1.0, -sh1);
            if (!fx) e->d_cpart.ensure((size_t)nblk * NX * kc * 8, 0, S(c), false);
            e->d_spart.ensure((size_t)nblk * NX * 2 * 8, 0, S(c), false);
            h.cpart = e->d_cpart.as<double>();
            h.spart = e->d_spart.as<double>();
            const int nchunk = (int)((kc + SQLP_FOLD_COLS - 1) / SQLP_FOLD_COLS);
            e->d_partial.ensure((size_t)nchunk * width * 8, 0, S(c), false);
            if (fx) {
                c->d_hfx.ensure((size_t)2 * kc * 24, 0, S(c), false);
                CK(cudaMemsetAsync(c->d_hfx.p, 0, (size_t)NX * kc * 24, S(c)));
                h.hfx = c->d_hfx.as<long long>();
                if (!c->hist_fx_set[NX]) {
                    if (NX == 2) CK(cudaFuncSetAttribute(k_cut_hist_fx<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, c->smem_optin - 1024));
                    else CK(cudaFuncSetAttribute(k_cut_hist_fx<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, c->smem_optin - 1024));
                    c->hist_fx_set[NX] = true;
                }
                if (NX == 2) LAUNCH(c, k_cut_hist_fx<2>, (int)nblk, SQLP_HIST_THREADS, fx_smem, h, (int)fx_smem);
                else LAUNCH(c, k_cut_hist_fx<1>, (int)nblk, SQLP_HIST_THREADS, fx_smem, h, (int)fx_smem);
            } else {
                if (!c->hist_smem_set[NX]) {
                    // (the kernel also has 512 B of static shared memory: the dynamic part cannot be the whole opt-in size)
                    if (NX == 2) CK(cudaFuncSetAttribute(k_cut_hist<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, c->smem_optin - 1024));
                    else CK(cudaFuncSetAttribute(k_cut_hist<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, c->smem_optin - 1024));
                    c->hist_smem_set[NX] = true;
                }
                if (NX == 2) LAUNCH(c, k_cut_hist<2>, (int)nblk, SQLP_HIST_THREADS, hist_smem, h);
                else LAUNCH(c, k_cut_hist<1>, (int)nblk, SQLP_HIST_THREADS, hist_smem, h);
            }
            if (NX == 2) LAUNCH(c, k_cut_fold<2>, nchunk, 256, 0, h, (int)nblk, e->d_partial.as<double>());
            else LAUNCH(c, k_cut_fold<1>, nchunk, 256, 0, h, (int)nblk, e->d_partial.as<double>());
            const int64_t ng = (nchunk + group - 1) / group;
            e->d_partial2.ensure((size_t)ng * width * 8, 0, S(c), false);
            LAUNCH(c, k_sum_groups, (int)ng, 256, sub_smem, e->d_partial.as<double>(), (long long)nchunk, group, nsub,
                   width, e->d_partial2.as<double>(), (unsigned int *)(e->d_flags.as<int>() + 1), fin,
                   (const int *)e->d_flags.as<int>(), cell_slot != nullptr);
        } else {
            if (NX == 2) LAUNCH(c, k_cut_partial<2>, (int)ntiles, 256, sub_smem, r);
            else LAUNCH(c, k_cut_partial<1>, (int)ntiles, 256, sub_smem, r);
            int64_t ng = (ntiles + group - 1) / group;
            e->d_partial2.ensure((size_t)ng * width * 8, 0, S(c), false);
            LAUNCH(c, k_sum_groups, (int)ng, 256, sub_smem, e->d_partial.as<double>(), (long long)ntiles, group, nsub,
                   width, e->d_partial2.as<double>(), (unsigned int *)(e->d_flags.as<int>() + 1), fin,
                   (const int *)e->d_flags.as<int>(), cell_slot != nullptr);
        }
        prof_red.stop();
    } else if (cell_slot) {
        CK(cudaMemsetAsync(cell_slot, 0, (size_t)(width + 1) * 8, S(c)));   // this rank holds no scenario
    }
}

// Sharded job: ONE all-gather for all the epigraphs of a call (their rows sit back to back in the context's
// cell buffer, each followed by its "no argmax" word), then the rank-ordered sum of every row into the
// epigraph's d_out and the OR of the flag words into its d_flags -- the same bits on every rank.
struct CellGather {
    sqlp_ctx *c;
    int NX;
    std::vector<sqlp_epi *> epis;
    std::vector<size_t> off;
    size_t total = 0;
    CellGather(sqlp_ctx *c_, int NX_, int n_epi, sqlp_epi *const *epi) : c(c_), NX(NX_)
    {
        if (c->world <= 1) return;
        for (int i = 0; i < n_epi; ++i) {
            epis.push_back(epi[i]);
            off.push_back(total);
            total += (size_t)NX * ((size_t)epi[i]->n1 + 2) + 1;
        }
        c->d_cell.ensure(total * 8, 0, S(c), false);
        c->d_cellg.ensure((size_t)c->world * total * 8, 0, S(c), false);
    }
    double *slot(int i) const { return c->world > 1 ? c->d_cell.as<double>() + off[(size_t)i] : nullptr; }
    void run() { gather(); sum(); }
    void gather()
    {
        if (c->world <= 1 || epis.empty()) return;
        NK(g_nccl.AllGather(c->d_cell.p, c->d_cellg.p, total, ncclFloat64_, c->comm, S(c)));
    }
    void sum()
    {
        if (c->world <= 1 || epis.empty()) return;
        for (size_t i = 0; i < epis.size(); ++i) {
            sqlp_epi *e = epis[i];
            const int width = NX * ((int)e->n1 + 2);
            LAUNCH(c, k_rank_sum, (width + 1 + 127) / 128, 128, 0, c->d_cellg.as<double>() + off[i], c->world,
                   (long long)total, width, e->d_out.as<double>(), e->d_flags.as<int>());
        }
    }
};

// ---- one host thread, several GPUs (sqlp_ctx_create_multi) -----------------------------------------------
// Shard s of a handle: s = 0 the leader's own, s >= 1 the replica on rank s.
int n_shards(const sqlp_ctx *c) { return 1 + (int)c->peers.size(); }
sqlp_ctx *shard(sqlp_ctx *c, int s) { return s == 0 ? c : c->peers[(size_t)s - 1]; }
sqlp_pool *shard(sqlp_pool *p, int s) { return s == 0 ? p : p->peers[(size_t)s - 1]; }
sqlp_epi *shard(sqlp_epi *e, int s) { return s == 0 ? e : e->peers[(size_t)s - 1]; }

sqlp_epi *bias_twin(int i, sqlp_epi *const *epi);

// The cuts of a cell at NX points on every GPU of the context: each GPU's chain is enqueued on its own stream (so
// the GPUs work side by side), then ONE grouped all-gather, then the rank-ordered sums -- every GPU ends with the
// same bits in d_out / d_flags, the leader's are what the caller reads.  x2[i] = the points of epigraph i (host).
void cell_cuts_all_shards(sqlp_ctx *c, int NX, int n_epi, sqlp_epi *const *epi, const std::vector<std::vector<double>> &x2)
{
    const int ns = n_shards(c);
    std::vector<std::vector<sqlp_epi *>> es((size_t)ns);
    std::vector<std::unique_ptr<CellGather>> cg;
    for (int s = 0; s < ns; ++s) {
        sqlp_ctx *cs = shard(c, s);
        cs->bind();
        for (int i = 0; i < n_epi; ++i) {
            es[(size_t)s].push_back(shard(epi[i], s));
            es[(size_t)s].back()->cur_bias = nullptr;
        }
        cg.emplace_back(new CellGather(cs, NX, n_epi, es[(size_t)s].data()));
        for (int i = 0; i < n_epi; ++i)
            epi_cuts_enqueue(es[(size_t)s][(size_t)i], NX, x2[(size_t)i].data(), nullptr, true,
                             bias_twin(i, es[(size_t)s].data()), cg.back()->slot(i));
    }
    if (ns > 1) {
        NK(g_nccl.GroupStart());          // one thread, several communicators: the collective is issued as a group
        for (int s = 0; s < ns; ++s) { shard(c, s)->bind(); cg[(size_t)s]->gather(); }
        NK(g_nccl.GroupEnd());
        for (int s = 0; s < ns; ++s) { shard(c, s)->bind(); cg[(size_t)s]->sum(); }
    } else {
        cg[0]->run();
    }
    c->bind();
}

// Drain every GPU of the context; pools of the peers learn their size.
void sync_all_shards(sqlp_ctx *c, sqlp_pool *p)
{
    for (int s = 1; s < n_shards(c); ++s) {
        sqlp_ctx *cs = shard(c, s);
        cs->bind();
        if (p) pool_confirm(shard(p, s));
        CK(cudaStreamSynchronize(S(cs)));
    }
    c->bind();
}

// Among the epigraphs of one call: the first earlier one whose bias vectors this one can reuse.
sqlp_epi *bias_twin(int i, sqlp_epi *const *epi)
{
    for (int j = 0; j < i; ++j)
        if (epi[j]->pool == epi[i]->pool && epi[j]->tmpl_id == epi[i]->tmpl_id && epi[j]->n_local > 0 &&
            epi[j]->cur_bias != nullptr)
            return epi[j];
    return nullptr;
}

struct CutHost {
    std::vector<double> out;
    int flags = 0;
};

void epi_cuts_fetch(sqlp_epi *e, int NX, CutHost &h)
{
    const int NC = (int)e->n1 + 2;
    h.out.resize((size_t)NX * NC);
    CK(cudaMemcpyAsync(h.out.data(), e->d_out.p, (size_t)NX * NC * 8, cudaMemcpyDeviceToHost, S(e->ctx)));
    CK(cudaMemcpyAsync(&h.flags, e->d_flags.p, 4, cudaMemcpyDeviceToHost, S(e->ctx)));
}

// ---- device cut list -------------------------------------------------------------------------
void cuts_reserve(sqlp_epi *e, int64_t need)
{
    if (need <= e->cuts_cap) return;
    const size_t RS = (size_t)e->n1 + 2;
    int64_t ncap = std::max<int64_t>(need, std::max<int64_t>(64, e->cuts_cap * 2));
    e->d_cuts.ensure((size_t)ncap * RS * 8, (size_t)e->n_cuts * RS * 8, S(e->ctx));
    e->d_inc.ensure(RS * 8, 0, S(e->ctx));
    e->d_prev_inc.ensure(RS * 8, 0, S(e->ctx));
    e->cuts_cap = ncap;
}

CutList cut_list(sqlp_epi *e, bool last)
{
    CutList L;
    L.cuts = e->d_cuts.as<double>();
    L.n = (int)(last ? e->n_last : e->n_cuts);
    L.inc = last ? (e->has_prev_inc ? e->d_prev_inc.as<double>() : nullptr)
                 : (e->has_inc ? e->d_inc.as<double>() : nullptr);
    return L;
}

// est[0 .. nlists * NX): weighted value of the current (and snapshot) approximation at NX device points
void cuts_evaluate_enqueue(sqlp_epi *e, const double *d_x, int NX, int nlists, double *d_est)
{
    sqlp_ctx *c = e->ctx;
    cuts_reserve(e, 1);
    LAUNCH(c, k_cuts_evaluate, 1, 256, 0, cut_list(e, false), cut_list(e, true), nlists, d_x, NX, (int)e->n1,
           e->total_weight, e->lower_bound, e->objective_weight, d_est);
}

void check_sense(int32_t sense)
{
    REQUIRE(sense == SQLP_MIN_SENSE || sense == SQLP_MAX_SENSE, SQLP_E_INVALID, "bad sense");
    REQUIRE(sense == SQLP_MIN_SENSE, SQLP_E_UNSUPPORTED,
            "MAX_SENSE is unsupported: the reference's MAX branch never selects a vertex");
}

}  // namespace

// conic_debug.cu -- host-side interpreter of the gather programs built by conic_symbolic.h.
// TEST HOOK ONLY: it executes the assembly / factorisation / substitution programs for ONE seed on
// the CPU so that the index programs can be unit-tested without a GPU (tests/test_conic_symbolic.py).
// It is not reachable from any product entry point (scpb_cone_solve always launches k_ipm_solve).
#include "../../include/scpb.h"
#include "conic_symbolic.h"
#define SN_EMULATE
#include "conic_sn.cuh"

extern "C" int32_t scpb_debug_kkt_solve(int32_t n, int32_t p, int32_t m, const int32_t *A_rp, const int32_t *A_ci,
                                        const int32_t *G_rp, const int32_t *G_ci, int32_t l, int32_t nsoc,
                                        const int32_t *soc_dims, const int32_t *perm, const double *Av,
                                        const double *Gv, const double *wm, double delta, double delta_dyn,
                                        const double *rhs, double *sol, int64_t *info)
{
    ConeSymbolic S;
    static const int zero = 0;
    if (!cone_symbolic_build(S, n, p, m, A_rp, A_ci ? A_ci : &zero, G_rp, G_ci ? G_ci : &zero, l, nsoc, soc_dims, perm))
        return SCPB_ERR_ARG;
    const int nk = S.nk, ntgt = S.nnzL + nk;
    std::vector<double> Y(ntgt), Ls(S.nnzL + 1), invD(nk), v(nk);
    for (int t = 0; t < ntgt; t++) {  // kkt_assemble
        double acc = delta * S.as_sign[t];
        if (S.as_src[t] >= 0) acc += Av[S.as_src[t]];
        for (int k = S.as_ptr[t]; k < S.as_ptr[t + 1]; k++) acc += Gv[S.as_a[k]] * Gv[S.as_b[k]] * wm[S.as_c[k]];
        Y[t] = acc;
    }
    for (int lv = 0; lv < S.nlevels; lv++) {  // kkt_factor: balanced program (phase A items, then phase B items)
        const int R = S.fa_R[lv];
        for (int w = S.fa_lvl[lv]; w < S.fa_lvl[lv + 1]; w++) {
            const int *it = &S.fa_item[4 * (size_t)w];
            if (it[2] - it[1] > R * CONIC_FACTOR_PF) return SCPB_ERR_ARG;
            double part = 0.0;
            for (int k = it[1]; k < it[2]; k++) part += Y[S.ft_op[2 * (size_t)k]] * Ls[S.Lr_pos[S.ft_op[2 * (size_t)k + 1]]];
            Y[it[0]] -= part;
        }
        for (int w = S.fb_lvl[lv]; w < S.fb_lvl[lv + 1]; w++) {   // every item regularises its own copy of the pivot
            const int *it = &S.fb_item[4 * (size_t)w];
            const double sgn = (it[3] & 1) ? 1.0 : -1.0;
            double d = Y[S.nnzL + it[1]];
            if (!(sgn * d > delta_dyn)) d = sgn * delta_dyn;
            if (it[3] & 2) invD[it[1]] = 1.0 / d;
            else Ls[it[0]] = Y[it[0]] * (1.0 / d);
        }
    }
    for (int i = 0; i < nk; i++) v[S.iperm[i]] = rhs[i];
    // kkt_ldl_solve_smem: the balanced (split-item) substitution programs, level by level; within a level every
    // item reads v as it was when the level started, except for the target it accumulates into
    std::vector<double> Lrow(S.nnzL + 1);
    for (int k = 0; k < S.nnzL; k++) Lrow[k] = Ls[S.Lr_pos[k]];
    for (int lv = 0; lv < S.nlevels; lv++) {
        const int R = S.fwp_R[lv];
        for (int w = S.fwp_lvl[lv]; w < S.fwp_lvl[lv + 1]; w++) {
            const int *it = &S.fwp_item[4 * (size_t)w];
            if (it[2] - it[1] > R * CONIC_SOLVE_PF) return SCPB_ERR_ARG;
            double part = 0.0;
            for (int k = it[1]; k < it[2]; k++) part += Lrow[k] * v[S.Lr_col[k]];
            v[it[0]] -= part;
        }
    }
    for (int i = 0; i < nk; i++) v[i] *= invD[i];
    for (int lv = S.nlevels - 1; lv >= 0; lv--) {
        const int R = S.bwp_R[lv];
        for (int w = S.bwp_lvl[lv]; w < S.bwp_lvl[lv + 1]; w++) {
            const int *it = &S.bwp_item[4 * (size_t)w];
            if (it[2] - it[1] > R * CONIC_SOLVE_PF) return SCPB_ERR_ARG;
            double part = 0.0;
            for (int k = it[1]; k < it[2]; k++) part += Ls[k] * v[S.L_ri[k]];
            v[it[0]] -= part;
        }
    }
    for (int i = 0; i < nk; i++) sol[i] = v[S.iperm[i]];
    if (info) { info[0] = S.nnzL; info[1] = S.nlevels; info[2] = S.factor_ops; info[3] = (int64_t)S.as_a.size(); }
    return SCPB_OK;
}

// Same test hook for the SUPERNODAL program (conic_symbolic.h, "supernodal program"): dense panels, one supernodal
// level at a time, right-looking updates scattered through the precomputed maps.  This is the executable specification
// of the round-2 device kernels; nothing on the product path calls it.
extern "C" int32_t scpb_debug_kkt_solve_sn(int32_t n, int32_t p, int32_t m, const int32_t *A_rp, const int32_t *A_ci,
                                           const int32_t *G_rp, const int32_t *G_ci, int32_t l, int32_t nsoc,
                                           const int32_t *soc_dims, const int32_t *perm, const double *Av,
                                           const double *Gv, const double *wm, double delta, double delta_dyn,
                                           const double *rhs, double *sol, int64_t *info)
{
    ConeSymbolic S;
    static const int zero = 0;
    if (!cone_symbolic_build(S, n, p, m, A_rp, A_ci ? A_ci : &zero, G_rp, G_ci ? G_ci : &zero, l, nsoc, soc_dims, perm))
        return SCPB_ERR_ARG;
    const int nk = S.nk, ntgt = S.nnzL + nk, ns = (int)S.sn_first.size();
    std::vector<double> P((size_t)S.sn_panel_size, 0.0), invD(nk), v(nk);
    for (int t = 0; t < ntgt; t++) {  // assembly straight into the panels
        double acc = delta * S.as_sign[t];
        if (S.as_src[t] >= 0) acc += Av[S.as_src[t]];
        for (int k = S.as_ptr[t]; k < S.as_ptr[t + 1]; k++) acc += Gv[S.as_a[k]] * Gv[S.as_b[k]] * wm[S.as_c[k]];
        P[(size_t)S.sn_pos_of_target[t]] = acc;
    }
    for (int lv = 0; lv < S.sn_nlevels; lv++)
        for (int w_ = S.sn_lvl_ptr[lv]; w_ < S.sn_lvl_ptr[lv + 1]; w_++) {
            const int s = S.sn_lvl_nodes[w_], a = S.sn_first[s], w = S.sn_width[s], R = S.sn_nrows[s];
            double *Q = &P[(size_t)S.sn_panel_off[s]];
            for (int c = 0; c < w; c++) {   // dense LDL' of the panel, right-looking inside the supernode
                double d = Q[c + (size_t)R * c];
                const double sgn = (double)S.as_sign[S.nnzL + a + c];
                if (!(sgn * d > delta_dyn)) d = sgn * delta_dyn;
                Q[c + (size_t)R * c] = d;
                invD[a + c] = 1.0 / d;
                for (int r = c + 1; r < R; r++) Q[r + (size_t)R * c] /= d;
                for (int c2 = c + 1; c2 < w; c2++) {
                    const double f = Q[c2 + (size_t)R * c] * d;
                    for (int r = c2; r < R; r++) Q[r + (size_t)R * c2] -= Q[r + (size_t)R * c] * f;
                }
            }
            int k = S.sn_upd_ptr[s];   // Schur update of the ancestors' panels
            for (int y = 0; y < R - w; y++)
                for (int x = y; x < R - w; x++) {
                    double u = 0.0;
                    for (int c = 0; c < w; c++) u += Q[w + x + (size_t)R * c] * Q[c + (size_t)R * c] * Q[w + y + (size_t)R * c];
                    P[(size_t)S.sn_upd_dst[k++]] -= u;
                }
        }
    for (int i = 0; i < nk; i++) v[S.iperm[i]] = rhs[i];
    for (int lv = 0; lv < S.sn_nlevels; lv++)   // forward: column oriented, updates scattered to the rows below
        for (int w_ = S.sn_lvl_ptr[lv]; w_ < S.sn_lvl_ptr[lv + 1]; w_++) {
            const int s = S.sn_lvl_nodes[w_], w = S.sn_width[s], R = S.sn_nrows[s];
            const double *Q = &P[(size_t)S.sn_panel_off[s]];
            const int *rows = &S.sn_rows[S.sn_rows_ptr[s]];
            for (int c = 0; c < w; c++) {
                const double xc = v[rows[c]];
                for (int r = c + 1; r < R; r++) v[rows[r]] -= Q[r + (size_t)R * c] * xc;
            }
        }
    for (int i = 0; i < nk; i++) v[i] *= invD[i];
    for (int lv = S.sn_nlevels - 1; lv >= 0; lv--)   // backward: gathers from the rows below
        for (int w_ = S.sn_lvl_ptr[lv]; w_ < S.sn_lvl_ptr[lv + 1]; w_++) {
            const int s = S.sn_lvl_nodes[w_], w = S.sn_width[s], R = S.sn_nrows[s];
            const double *Q = &P[(size_t)S.sn_panel_off[s]];
            const int *rows = &S.sn_rows[S.sn_rows_ptr[s]];
            for (int c = w - 1; c >= 0; c--) {
                double acc = v[rows[c]];
                for (int r = c + 1; r < R; r++) acc -= Q[r + (size_t)R * c] * v[rows[r]];
                v[rows[c]] = acc;
            }
        }
    for (int i = 0; i < nk; i++) sol[i] = v[S.iperm[i]];
    if (info) {
        info[0] = ns; info[1] = S.sn_nlevels; info[2] = S.sn_panel_size; info[3] = (int64_t)S.sn_upd_dst.size();
        int wmax = 0, rmax = 0;
        for (int s = 0; s < ns; s++) { wmax = std::max(wmax, S.sn_width[s]); rmax = std::max(rmax, S.sn_nrows[s]); }
        info[4] = wmax; info[5] = rmax; info[6] = S.nlevels; info[7] = S.nnzL;
    }
    (void)ns;
    return SCPB_OK;
}

// The same system once more, this time through the WARP routines of conic_sn.cuh compiled in lane-emulation mode
// (SN_EMULATE): the code that k_ipm_solve runs per (supernode, seed) item, executed here one item at a time.
extern "C" int32_t scpb_debug_kkt_solve_sn_emu(int32_t n, int32_t p, int32_t m, const int32_t *A_rp, const int32_t *A_ci,
                                               const int32_t *G_rp, const int32_t *G_ci, int32_t l, int32_t nsoc,
                                               const int32_t *soc_dims, const int32_t *perm, const double *Av,
                                               const double *Gv, const double *wm, double delta, double delta_dyn,
                                               const double *rhs, double *sol, int64_t *info)
{
    ConeSymbolic S;
    static const int zero = 0;
    if (!cone_symbolic_build(S, n, p, m, A_rp, A_ci ? A_ci : &zero, G_rp, G_ci ? G_ci : &zero, l, nsoc, soc_dims, perm))
        return SCPB_ERR_ARG;
    const int nk = S.nk, ntgt = S.nnzL + nk, ns = (int)S.sn_first.size();
    static_assert(SN_SCRATCH == 384 && SN_MAXROWS == 64, "conic_symbolic.h (cls_of) assumes these scratch sizes");
    if (!S.sn_fits) return SCPB_ERR_UNSUPPORTED;
    SnProgram Q{};
    Q.first = S.sn_first.data(); Q.width = S.sn_width.data(); Q.nrows = S.sn_nrows.data();
    Q.rows_ptr = S.sn_rows_ptr.data(); Q.rows = S.sn_rows.data(); Q.lvl_ptr = S.sn_lvl_ptr.data();
    Q.lvl_nodes = S.sn_lvl_nodes.data(); Q.upd_xy = S.sn_upd_xy.data(); Q.sign = S.sn_sign.data();
    Q.panel_off = S.sn_panel_off.data(); Q.upd_ptr = S.sn_upd_ptr.data(); Q.upd_dst = S.sn_upd_dst.data();
    Q.nlevels = S.sn_nlevels; Q.cls_ptr = S.sn_cls_ptr.data();
    std::vector<double> P((size_t)S.sn_panel_size, 0.0), invD(nk), v(nk), scr(SN_SCRATCH), xs(SN_MAXROWS + 32);
    // per level the three lane-group classes, exactly as the kernel walks them (each group with its share of the scratch)
    auto each = [&](int lv, auto &&f8, auto &&f16, auto &&f32) {
        const int *cp_ = &S.sn_cls_ptr[4 * (size_t)lv];
        for (int w_ = cp_[0]; w_ < cp_[1]; w_++) f8(S.sn_lvl_nodes[w_]);
        for (int w_ = cp_[1]; w_ < cp_[2]; w_++) f16(S.sn_lvl_nodes[w_]);
        for (int w_ = cp_[2]; w_ < cp_[3]; w_++) f32(S.sn_lvl_nodes[w_]);
    };
    for (int t = 0; t < ntgt; t++) {
        double acc = delta * S.as_sign[t];
        if (S.as_src[t] >= 0) acc += Av[S.as_src[t]];
        for (int k = S.as_ptr[t]; k < S.as_ptr[t + 1]; k++) acc += Gv[S.as_a[k]] * Gv[S.as_b[k]] * wm[S.as_c[k]];
        P[(size_t)S.sn_pos_of_target[t]] = acc;
    }
    for (int lv = 0; lv < S.sn_nlevels; lv++)
        each(lv, [&](int s) { sn_factor_item<8>(Q, s, P.data(), invD.data(), 1, 0, delta_dyn, scr.data() + 3 * (SN_SCRATCH / 4)); },
             [&](int s) { sn_factor_item<16>(Q, s, P.data(), invD.data(), 1, 0, delta_dyn, scr.data() + SN_SCRATCH / 2); },
             [&](int s) { sn_factor_item<32>(Q, s, P.data(), invD.data(), 1, 0, delta_dyn, scr.data()); });
    for (int i = 0; i < nk; i++) v[S.iperm[i]] = rhs[i];
    const int XS = SN_MAXROWS + 32;
    for (int lv = 0; lv < S.sn_nlevels; lv++)
        each(lv, [&](int s) { sn_forward_item<8>(Q, s, P.data(), v.data(), 1, 0, xs.data() + 3 * (XS / 4)); },
             [&](int s) { sn_forward_item<16>(Q, s, P.data(), v.data(), 1, 0, xs.data() + XS / 2); },
             [&](int s) { sn_forward_item<32>(Q, s, P.data(), v.data(), 1, 0, xs.data()); });
    for (int i = 0; i < nk; i++) v[i] *= invD[i];
    for (int lv = S.sn_nlevels - 1; lv >= 0; lv--)
        each(lv, [&](int s) { sn_backward_item<8>(Q, s, P.data(), v.data(), 1, 0, xs.data() + 3 * (XS / 4)); },
             [&](int s) { sn_backward_item<16>(Q, s, P.data(), v.data(), 1, 0, xs.data() + XS / 2); },
             [&](int s) { sn_backward_item<32>(Q, s, P.data(), v.data(), 1, 0, xs.data()); });
    for (int i = 0; i < nk; i++) sol[i] = v[S.iperm[i]];
    if (info) { info[0] = ns; info[1] = S.sn_nlevels; info[2] = S.sn_panel_size; info[3] = (int64_t)S.sn_upd_dst.size(); }
    return SCPB_OK;
}

// conic_debug.cu -- host-side interpreter of the gather programs built by conic_symbolic.h.
// TEST HOOK ONLY: it executes the assembly / factorisation / substitution programs for ONE seed on
// the CPU so that the index programs can be unit-tested without a GPU (tests/test_conic_symbolic.py).
// It is not reachable from any product entry point (scpb_cone_solve always launches k_ipm_solve).
#include "../../include/scpb.h"
#include "conic_symbolic.h"

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

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
    for (int lv = 0; lv < S.nlevels; lv++) {  // kkt_factor
        for (int w = S.ft_lvl_ptr[lv]; w < S.ft_lvl_ptr[lv + 1]; w++) {
            const int t = S.ft_target[w];
            double acc = Y[t];
            for (int k = S.ft_op_ptr[w]; k < S.ft_op_ptr[w + 1]; k++) acc -= Y[S.ft_op_a[k]] * Ls[S.ft_op_b[k]];
            if (t >= S.nnzL) {  // same dynamic regularisation rule as kkt_factor (conic_ipm.cuh)
                const double sgn = (double)S.as_sign[t];
                if (!(sgn * acc > delta_dyn)) acc = sgn * delta_dyn;
                invD[t - S.nnzL] = 1.0 / acc;
            }
            Y[t] = acc;
        }
        for (int w = S.sc_lvl_ptr[lv]; w < S.sc_lvl_ptr[lv + 1]; w++) Ls[S.sc_pos[w]] = Y[S.sc_pos[w]] * invD[S.sc_col[w]];
    }
    for (int i = 0; i < nk; i++) v[S.iperm[i]] = rhs[i];
    for (int lv = 0; lv < S.nlevels; lv++)  // kkt_ldl_solve
        for (int w = S.lvl_ptr[lv]; w < S.lvl_ptr[lv + 1]; w++) {
            const int i = S.lvl_nodes[w];
            double acc = v[i];
            for (int k = S.Lr_rp[i]; k < S.Lr_rp[i + 1]; k++) acc -= Ls[S.Lr_pos[k]] * v[S.Lr_col[k]];
            v[i] = acc;
        }
    for (int i = 0; i < nk; i++) v[i] *= invD[i];
    for (int lv = S.nlevels - 1; lv >= 0; lv--)
        for (int w = S.lvl_ptr[lv]; w < S.lvl_ptr[lv + 1]; w++) {
            const int j = S.lvl_nodes[w];
            double acc = v[j];
            for (int k = S.L_cp[j]; k < S.L_cp[j + 1]; k++) acc -= Ls[k] * v[S.L_ri[k]];
            v[j] = acc;
        }
    for (int i = 0; i < nk; i++) sol[i] = v[S.iperm[i]];
    if (info) { info[0] = S.nnzL; info[1] = S.nlevels; info[2] = S.factor_ops; info[3] = (int64_t)S.as_a.size(); }
    return SCPB_OK;
}

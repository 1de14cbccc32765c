// conic_sn.cuh -- supernodal LDL' : warp-synchronous per-panel routines.
//
// One lane GROUP (8, 16 or 32 lanes of a warp, by panel height) owns one (supernode, seed) item of a supernodal
// level (conic_symbolic.h, "supernodal program"): a dense column-major R x w panel, D on its diagonal, unit-lower L
// below.  Lane r of the group holds ROW r of the panel in registers (w <= CONIC_SN_WMAX doubles), so the panel is read
// from HBM/L2 exactly once per use, fully coalesced (the panels of ONE seed are contiguous), and the elimination of
// its w columns costs no memory round trip and no CTA barrier -- only warp shuffles:
//   factor   : pivot broadcast from lane c, rank-1 update of the columns to the right, then the Schur complement
//              of the rows below is scattered into the ancestors' panels with fire-and-forget atomics;
//   forward  : x_S = L_SS^-1 x_S by shuffles, the rows below receive -L_below,S x_S (shared-memory atomics);
//   backward : x_S = L_SS^-T (x_S - L_below,S' x_below), one group reduction per column.
// Single-column panels of at most CONIC_SN_R1MAX rows (the thousands of leaves of the elimination tree) are worked on
// by ONE thread each (sn1_* below): no shuffles, all loads of an item issued together.
// The scalar level-scheduled programs of conic_ipm.cuh need one barrier and two memory round trips per COLUMN of a
// dense separator block; here a block of w columns is one level (bench KKT: 22 levels instead of 88).
// The executable specification of the panel program is the CPU interpreter scpb_debug_kkt_solve_sn (conic_debug.cu).
#pragma once
#include "conic_symbolic.h"

#define SN_WMAX CONIC_SN_WMAX

struct SnProgram {          // device view of the supernodal program (all arrays shared by the batch)
    const int *first, *width, *nrows, *rows_ptr, *rows, *lvl_ptr, *lvl_nodes, *upd_xy, *sign;
    const int *cls_ptr;     // [nlevels][5]: inside a level the supernodes are sorted by lane-group size 1 | 8 | 16 | 32
    const int *panel_off, *upd_ptr, *upd_dst;
    const int4 *desc;       // two int4 per supernode IN LEVEL ORDER (the order of lvl_nodes): {first, width, nrows,
                            // panel_off}, {rows_ptr, upd_ptr, -, -}; an item costs two loads, requested one item ahead
    int nlevels;
};

// top supernodes of the hybrid program (see the hy_* routines below)
struct HyProgram {
    const int4 *desc;       // two int4 per top supernode, in top-level order
    const int *tl_ptr;      // ntl + 1
    const int *upd_dst;     // below-row pairs (x >= y) -> target id
    const int *rows;        // panel row -> node (the supernodal program's sn_rows)
    int ntl;
};

#ifdef __CUDACC__
#define SN_R1 CONIC_SN_R1MAX

template <int GS>
__device__ __forceinline__ unsigned sn_group_mask()
{
    return GS == 32 ? 0xffffffffu : (((1u << GS) - 1u) << ((threadIdx.x & 31) & ~(GS - 1)));
}

// ---- numeric factorisation of one panel + Schur update of the ancestors ------------------------------------
// P: the panels of THIS seed (entry e at P[e]); invD group-blocked (entry i of seed sg at invD[i*G + sg]).
// A pivot with sgn*d <= tau is replaced by sgn*rho (dynamic regularisation); when the rejected value is not small
// (|d| > bad_abs or NaN) the inertia was lost to cancellation and *bad is raised: the caller escalates the static
// regularisation of that seed and factors again.
template <int GS>
__device__ __forceinline__ void sn_factor_panel(const SnProgram &S, const int4 d0, const int4 d1, double *P, double *invD,
                                                int G, int sg, double tau, double rho, double bad_abs, int *bad)
{
    const unsigned gm = sn_group_mask<GS>();
    const int lg = (int)(threadIdx.x & (GS - 1));
    const int a = d0.x, w = d0.y, R = d0.z, off = d0.w;
    const int sgn_l = (lg < w) ? S.sign[a + lg] : 1;
    double prow[SN_WMAX];
#pragma unroll
    for (int c = 0; c < SN_WMAX; c++)
        prow[c] = (c < w && lg < R && lg >= c) ? __ldcg(&P[(size_t)off + lg + (size_t)R * c]) : 0.0;
#pragma unroll
    for (int c = 0; c < SN_WMAX; c++) {
        if (c < w) {
            double d = __shfl_sync(gm, prow[c], c, GS);
            const double sgn = (double)__shfl_sync(gm, sgn_l, c, GS);
            if (!(sgn * d > tau)) {
                if (lg == 0 && !(fabs(d) <= bad_abs)) *bad = 1;
                d = sgn * rho;
            }
            const double inv = 1.0 / d;
            const double l = (lg > c && lg < R) ? prow[c] * inv : 0.0;
            const double ld = l * d;
#pragma unroll
            for (int c2 = c + 1; c2 < SN_WMAX; c2++) {
                if (c2 < w) {
                    const double lc2 = __shfl_sync(gm, l, c2, GS);
                    if (lg >= c2) prow[c2] = fma(-ld, lc2, prow[c2]);
                }
            }
            if (lg == c) { prow[c] = d; invD[(size_t)(a + c) * G + sg] = inv; }
            else if (lg > c) prow[c] = l;
        }
    }
#pragma unroll
    for (int c = 0; c < SN_WMAX; c++)
        if (c < w && lg < R && lg >= c) P[(size_t)off + lg + (size_t)R * c] = prow[c];
    // Schur complement of the rows below: entry (x, y), x >= y, receives -sum_c L[x,c] d_c L[y,c];
    // four rows y per round: their scatter indices are requested before the shuffles of the round
    const int nb = R - w;
    if (nb > 0) {
        double ldv[SN_WMAX];
#pragma unroll
        for (int c = 0; c < SN_WMAX; c++) {
            const double dc = (c < w) ? __shfl_sync(gm, prow[c], c, GS) : 0.0;
            ldv[c] = (lg >= w) ? prow[c] * dc : 0.0;
        }
        const int x = lg - w;
        const int k0 = d1.y;
        for (int y0 = 0; y0 < nb; y0 += 4) {
            int kk[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int y = y0 + j;
                kk[j] = (y < nb && x >= y && x < nb) ? S.upd_dst[k0 + y * nb - (y * (y - 1)) / 2 + (x - y)] : -1;
            }
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int y = y0 + j;
                if (y < nb) {
                    double u = 0.0;
#pragma unroll
                    for (int c = 0; c < SN_WMAX; c++) {
                        if (c < w) {
                            const double pv = __shfl_sync(gm, prow[c], w + y, GS);
                            u = fma(ldv[c], pv, u);
                        }
                    }
                    if (kk[j] >= 0) atomicAdd(&P[kk[j]], -u);
                }
            }
        }
    }
}

// one thread, one single-column panel of at most SN_R1 rows
__device__ __forceinline__ void sn1_factor_panel(const SnProgram &S, const int4 d0, const int4 d1, double *P, double *invD,
                                                 int G, int sg, double tau, double rho, double bad_abs, int *bad)
{
    const int a = d0.x, R = d0.z, off = d0.w, nb = R - 1;
    double v[SN_R1];
#pragma unroll
    for (int r = 0; r < SN_R1; r++) v[r] = (r < R) ? __ldcg(&P[(size_t)off + r]) : 0.0;
    const double sgn = (double)S.sign[a];
    double d = v[0];
    if (!(sgn * d > tau)) {
        if (!(fabs(d) <= bad_abs)) *bad = 1;
        d = sgn * rho;
    }
    const double inv = 1.0 / d;
    invD[(size_t)a * G + sg] = inv;
    P[off] = d;
#pragma unroll
    for (int r = 1; r < SN_R1; r++)
        if (r < R) { v[r] *= inv; P[(size_t)off + r] = v[r]; }
    int k = d1.y;
#pragma unroll
    for (int y = 0; y < SN_R1 - 1; y++) {
        if (y < nb) {
            const double ly = v[1 + y] * d;
#pragma unroll
            for (int x = y; x < SN_R1 - 1; x++)
                if (x < nb) { atomicAdd(&P[S.upd_dst[k]], -(v[1 + x] * ly)); k++; }
        }
    }
}

// ---- forward substitution of one panel: x_S = L_SS^-1 x_S, then the rows below receive -L_below,S x_S ---------
// v: the substitution vector of this seed group in shared memory (entry i of seed sg at v[i*G + sg])
template <int GS>
__device__ __forceinline__ void sn_forward_panel(const SnProgram &S, const int4 d0, const int4 d1, const double *P,
                                                 double *v, int G, int sg)
{
    const unsigned gm = sn_group_mask<GS>();
    const int lg = (int)(threadIdx.x & (GS - 1));
    const int w = d0.y, R = d0.z, off = d0.w;
    const int ri = (lg < R) ? S.rows[d1.x + lg] : 0;
    double prow[SN_WMAX];
#pragma unroll
    for (int c = 0; c < SN_WMAX; c++)
        prow[c] = (c < w && lg < R && lg > c) ? __ldcg(&P[(size_t)off + lg + (size_t)R * c]) : 0.0;
    double x = (lg < w) ? v[(size_t)ri * G + sg] : 0.0;
#pragma unroll
    for (int c = 0; c < SN_WMAX; c++) {
        if (c < w) {
            const double xc = __shfl_sync(gm, x, c, GS);
            x = fma(-prow[c], xc, x);      // prow[c] is zero for lanes <= c
        }
    }
    if (lg < w) v[(size_t)ri * G + sg] = x;
    else if (lg < R) atomicAdd(&v[(size_t)ri * G + sg], x);   // siblings update common ancestors
}

__device__ __forceinline__ void sn1_forward_panel(const SnProgram &S, const int4 d0, const int4 d1, const double *P,
                                                  double *v, int G, int sg)
{
    const int R = d0.z, off = d0.w;
    int ri[SN_R1];
    double l[SN_R1];
#pragma unroll
    for (int r = 0; r < SN_R1; r++) {
        ri[r] = (r < R) ? S.rows[d1.x + r] : 0;
        l[r] = (r > 0 && r < R) ? __ldcg(&P[(size_t)off + r]) : 0.0;
    }
    const double xc = v[(size_t)ri[0] * G + sg];
#pragma unroll
    for (int r = 1; r < SN_R1; r++)
        if (r < R) atomicAdd(&v[(size_t)ri[r] * G + sg], -l[r] * xc);
}

// ---- backward substitution of one panel: x_S = L_SS^-T (x_S - L_below,S^T x_below) -----------------------------
template <int GS>
__device__ __forceinline__ void sn_backward_panel(const SnProgram &S, const int4 d0, const int4 d1, const double *P,
                                                  double *v, int G, int sg)
{
    const unsigned gm = sn_group_mask<GS>();
    const int lg = (int)(threadIdx.x & (GS - 1));
    const int w = d0.y, R = d0.z, off = d0.w;
    const int ri = (lg < R) ? S.rows[d1.x + lg] : 0;
    double prow[SN_WMAX];
#pragma unroll
    for (int c = 0; c < SN_WMAX; c++)
        prow[c] = (c < w && lg < R && lg > c) ? __ldcg(&P[(size_t)off + lg + (size_t)R * c]) : 0.0;
    double x = (lg < R) ? v[(size_t)ri * G + sg] : 0.0;
#pragma unroll
    for (int c = SN_WMAX - 1; c >= 0; c--) {
        if (c < w) {
            double t = prow[c] * x;
#pragma unroll
            for (int o = GS / 2; o > 0; o >>= 1) t += __shfl_xor_sync(gm, t, o, GS);
            if (lg == c) x -= t;
        }
    }
    if (lg < w) v[(size_t)ri * G + sg] = x;
}

__device__ __forceinline__ void sn1_backward_panel(const SnProgram &S, const int4 d0, const int4 d1, const double *P,
                                                   double *v, int G, int sg)
{
    const int R = d0.z, off = d0.w;
    int ri[SN_R1];
    double l[SN_R1];
#pragma unroll
    for (int r = 0; r < SN_R1; r++) {
        ri[r] = (r < R) ? S.rows[d1.x + r] : 0;
        l[r] = (r > 0 && r < R) ? __ldcg(&P[(size_t)off + r]) : 0.0;
    }
    double x = v[(size_t)ri[0] * G + sg];
#pragma unroll
    for (int r = 1; r < SN_R1; r++)
        if (r < R) x = fma(-l[r], v[(size_t)ri[r] * G + sg], x);
    v[(size_t)ri[0] * G + sg] = x;
}

// ---- hybrid program (cone_symbolic_build_hybrid): the same panel routines for the TOP supernodes, addressed in place
// on the scalar storage.  Column c of a supernode that starts at column a with R panel rows holds R-c-1 entries, and
// consecutive columns are consecutive in the CSC arrays of L, so entry (r, c), r > c, sits at L position
//   lb + c*(R-1) - c*(c-1)/2 + (r-c-1),   lb = L_cp[a];   its diagonal is target nnzL + a + c.
// Arrays are group-blocked: entry e of seed sg at [e*G + sg].  Descriptors: d0 = {a, w, R, lb}, d1 = {offset into rows,
// offset into upd_dst, pivot-sign bits, -}.

__device__ __forceinline__ int hy_pos(int lb, int R, int r, int c) { return lb + c * (R - 1) - ((c * (c - 1)) >> 1) + (r - c - 1); }

// numeric factorisation of one top panel; Y: unscaled targets (already holding every contribution of the low columns
// and of the top supernodes below), Ls: the scaled factor in column order, invD.  The Schur complement of the rows below
// goes to the ancestors' targets with fire-and-forget atomics.
template <int GS>
__device__ __forceinline__ void hy_factor_panel(const int4 d0, const int4 d1, const int *__restrict__ upd_dst, double *Y,
                                                double *Ls, double *invD, int nnzL, int G, int sg, double tau, double rho,
                                                double bad_abs, int *bad)
{
    const unsigned gm = sn_group_mask<GS>();
    const int lg = (int)(threadIdx.x & (GS - 1));
    const int a = d0.x, w = d0.y, R = d0.z, lb = d0.w;
    double prow[SN_WMAX];
#pragma unroll
    for (int c = 0; c < SN_WMAX; c++) {
        prow[c] = 0.0;
        if (c < w && lg < R && lg >= c) {
            const int t = (lg == c) ? nnzL + a + c : hy_pos(lb, R, lg, c);
            prow[c] = __ldcg(&Y[(size_t)t * G + sg]);
        }
    }
#pragma unroll
    for (int c = 0; c < SN_WMAX; c++) {
        if (c < w) {
            double d = __shfl_sync(gm, prow[c], c, GS);
            const double sgn = ((d1.z >> c) & 1) ? 1.0 : -1.0;
            if (!(sgn * d > tau)) {
                if (lg == 0 && !(fabs(d) <= bad_abs)) *bad = 1;
                d = sgn * rho;
            }
            const double inv = 1.0 / d;
            const double l = (lg > c && lg < R) ? prow[c] * inv : 0.0;
            const double ld = l * d;
#pragma unroll
            for (int c2 = c + 1; c2 < SN_WMAX; c2++) {
                if (c2 < w) {
                    const double lc2 = __shfl_sync(gm, l, c2, GS);
                    if (lg >= c2) prow[c2] = fma(-ld, lc2, prow[c2]);
                }
            }
            if (lg == c) { prow[c] = d; invD[(size_t)(a + c) * G + sg] = inv; }
            else if (lg > c) { prow[c] = l; if (lg < R) Ls[(size_t)hy_pos(lb, R, lg, c) * G + sg] = l; }
        }
    }
    const int nb = R - w;
    if (nb > 0) {
        double ldv[SN_WMAX];
#pragma unroll
        for (int c = 0; c < SN_WMAX; c++) {
            const double dc = (c < w) ? __shfl_sync(gm, prow[c], c, GS) : 0.0;
            ldv[c] = (lg >= w) ? prow[c] * dc : 0.0;
        }
        const int x = lg - w;
        const int k0 = d1.y;
        for (int y0 = 0; y0 < nb; y0 += 4) {
            int kk[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int y = y0 + j;
                kk[j] = (y < nb && x >= y && x < nb) ? upd_dst[k0 + y * nb - (y * (y - 1)) / 2 + (x - y)] : -1;
            }
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int y = y0 + j;
                if (y < nb) {
                    double u = 0.0;
#pragma unroll
                    for (int c = 0; c < SN_WMAX; c++) {
                        if (c < w) {
                            const double pv = __shfl_sync(gm, prow[c], w + y, GS);
                            u = fma(ldv[c], pv, u);
                        }
                    }
                    if (kk[j] >= 0) atomicAdd(&Y[(size_t)kk[j] * G + sg], -u);
                }
            }
        }
    }
}

// forward substitution of one top panel on the shared-memory vector v: x_S = L_SS^-1 x_S, rows below receive -L x_S
template <int GS>
__device__ __forceinline__ void hy_forward_panel(const int4 d0, const int4 d1, const int *__restrict__ rows,
                                                 const double *__restrict__ Ls, double *v, int G, int sg)
{
    const unsigned gm = sn_group_mask<GS>();
    const int lg = (int)(threadIdx.x & (GS - 1));
    const int w = d0.y, R = d0.z, lb = d0.w;
    const int ri = (lg < R) ? rows[d1.x + lg] : 0;
    double prow[SN_WMAX];
#pragma unroll
    for (int c = 0; c < SN_WMAX; c++)
        prow[c] = (c < w && lg < R && lg > c) ? Ls[(size_t)hy_pos(lb, R, lg, c) * G + sg] : 0.0;
    double x = (lg < w) ? v[(size_t)ri * G + sg] : 0.0;
#pragma unroll
    for (int c = 0; c < SN_WMAX; c++) {
        if (c < w) {
            const double xc = __shfl_sync(gm, x, c, GS);
            x = fma(-prow[c], xc, x);      // prow[c] is zero for lanes <= c
        }
    }
    if (lg < w) v[(size_t)ri * G + sg] = x;
    else if (lg < R) atomicAdd(&v[(size_t)ri * G + sg], x);   // siblings update common ancestors
}

// backward substitution of one top panel: x_S = L_SS^-T (x_S - L_below,S^T x_below)
template <int GS>
__device__ __forceinline__ void hy_backward_panel(const int4 d0, const int4 d1, const int *__restrict__ rows,
                                                  const double *__restrict__ Ls, double *v, int G, int sg)
{
    const unsigned gm = sn_group_mask<GS>();
    const int lg = (int)(threadIdx.x & (GS - 1));
    const int w = d0.y, R = d0.z, lb = d0.w;
    const int ri = (lg < R) ? rows[d1.x + lg] : 0;
    double prow[SN_WMAX];
#pragma unroll
    for (int c = 0; c < SN_WMAX; c++)
        prow[c] = (c < w && lg < R && lg > c) ? Ls[(size_t)hy_pos(lb, R, lg, c) * G + sg] : 0.0;
    double x = (lg < R) ? v[(size_t)ri * G + sg] : 0.0;
#pragma unroll
    for (int c = SN_WMAX - 1; c >= 0; c--) {
        if (c < w) {
            double t = prow[c] * x;
#pragma unroll
            for (int o = GS / 2; o > 0; o >>= 1) t += __shfl_xor_sync(gm, t, o, GS);
            if (lg == c) x -= t;
        }
    }
    if (lg < w) v[(size_t)ri * G + sg] = x;
}
#endif  // __CUDACC__

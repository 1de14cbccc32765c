// constraints.cuh -- device packs for the nonconvex path constraints s(t,k,x,u,p) <= 0 and their Jacobians
// C = ds/dx, D = ds/du, G = ds/dp  (the device twins of traj.s/C/D/G, src/parser/problem.jl:560-587,
// consumed by add_nonconvex_constraints!, src/solvers/scp.jl:744-794).
// Outputs are dense, row-major: s[NS], C[NS][NX], D[NS][NU], G[NS][NG]; entries that are structurally zero
// for a pack are simply never referenced by the host template.  G is PACKED: it holds NG columns and gcol(k, j) names
// the parameter each one belongs to at node k (0-based).  For most packs NG = np and gcol(k, j) = j; the free-flyer
// has np = 1 + 6N parameters of which only the six room-SDF slacks of node k enter s at that node
// (freeflyer/definition.jl:398-424), so its G has six columns instead of 481.
#pragma once
#include <math_constants.h>
#include "models.cuh"

template <int ID>
struct Constr {
    static constexpr int NS = 0;
};

// starship_flip/definition.jl:704-810.  par: [9] rate_delay, [10] deltadot_max, [11] gamma_gs, [12] thetamax2,
// [8] tau_s (shared with the dynamics pack).
template <>
struct Constr<SCPB_MODEL_STARSHIP> {
    static constexpr int NS = 23, NX = 8, NU = 3, NP = 10, NG = 10;
    __device__ static constexpr int gcol(int, int j) { return j; }
    __device__ static void eval(const ModelPar &P, double t, int N, int, const double *x, const double *u, const double *p,
                                double *s, double *C, double *D, double *G)
    {
        const double taus = P.v[8], rd = P.v[9], ddmax = P.v[10], cgs = cos(P.v[11]), thmax = P.v[12];
        // phase_switch / phase2, definition.jl:707-721
        const double dt = 1.0 / (double)(N - 1), tol = 1e-3;
        const bool psw = ((taus - dt) + tol <= t) && (t <= taus + tol);
        const bool ph2 = psw || (t > taus);
        for (int i = 0; i < NS; i++) s[i] = 0.0;
        for (int i = 0; i < NS * NX; i++) C[i] = 0.0;
        for (int i = 0; i < NS * NU; i++) D[i] = 0.0;
        for (int i = 0; i < NS * NP; i++) G[i] = 0.0;
        const double r0 = x[0], r1 = x[1], th = x[4], dd = x[7], de = u[1], ddot = u[2];
        s[0] = (de - dd) - ddot * rd;
        s[1] = ddot * rd - (de - dd);
        s[2] = ddot - ddmax;
        s[3] = -ddmax - ddot;
        const double nr = sqrt(r0 * r0 + r1 * r1);
        s[4] = nr * cgs - r1;
        C[0 * NX + 7] = -1.0;
        C[1 * NX + 7] = 1.0;
        const bool tiny = nr < 1.4901161193847656e-08;  // sqrt(eps)
        C[4 * NX + 0] = (tiny ? 0.0 : r0 / nr) * cgs;
        C[4 * NX + 1] = (tiny ? 0.0 : r1 / nr) * cgs - 1.0;
        D[0 * NU + 1] = 1.0;  D[0 * NU + 2] = -rd;
        D[1 * NU + 1] = -1.0; D[1 * NU + 2] = rd;
        D[2 * NU + 2] = 1.0;
        D[3 * NU + 2] = -1.0;
        if (psw) {
            for (int i = 0; i < NX; i++) {
                s[5 + i] = p[2 + i] - x[i];
                s[13 + i] = x[i] - p[2 + i];
                C[(5 + i) * NX + i] = -1.0;
                C[(13 + i) * NX + i] = 1.0;
                G[(5 + i) * NP + 2 + i] = 1.0;
                G[(13 + i) * NP + 2 + i] = -1.0;
            }
        }
        if (ph2) {
            s[21] = th - thmax;
            s[22] = -thmax - th;
            C[21 * NX + 4] = 1.0;
            C[22 * NX + 4] = -1.0;
        }
    }
};

// shared helper: ellipsoidal keep-out zones  s_i = 1 - |H_i (r - c_i)|,  ds_i/dr = -(H_i' H_i)(r - c_i) / |H_i (r - c_i)|
// (src/utils/ellipsoid.jl:99-118; quadrotor/definition.jl:237-262, freeflyer/definition.jl:380-404).
// par block per obstacle: H (3x3 column-major), c (3).
__device__ __forceinline__ void ellipsoid_rows(const double *ob, int nobs, const double *r, double *s, double *C, int NX)
{
    for (int i = 0; i < nobs; i++) {
        const double *H = ob + 12 * i, *c = H + 9;
        const double d0 = r[0] - c[0], d1 = r[1] - c[1], d2 = r[2] - c[2];
        double y[3];
        for (int a = 0; a < 3; a++) y[a] = H[a + 3 * 0] * d0 + H[a + 3 * 1] * d1 + H[a + 3 * 2] * d2;
        const double E = sqrt(y[0] * y[0] + y[1] * y[1] + y[2] * y[2]);
        s[i] = 1.0 - E;
        for (int b = 0; b < 3; b++) {          // (H'H d)_b = sum_a H[a,b] y_a
            const double g = H[0 + 3 * b] * y[0] + H[1 + 3 * b] * y[1] + H[2 + 3 * b] * y[2];
            C[i * NX + b] = -g / E;
        }
    }
}

// quadrotor/definition.jl:237-262: two ellipsoidal obstacles.  par: g(3), then 2 x {H(9), c(3)} from [3].
template <>
struct Constr<SCPB_MODEL_QUADROTOR> {
    static constexpr int NS = 2, NX = 6, NU = 4, NP = 1, NG = 1;
    __device__ static constexpr int gcol(int, int j) { return j; }
    __device__ static void eval(const ModelPar &P, double, int, int, const double *x, const double *, const double *,
                                double *s, double *C, double *D, double *G)
    {
        for (int i = 0; i < NS * NX; i++) C[i] = 0.0;
        for (int i = 0; i < NS * NU; i++) D[i] = 0.0;
        for (int i = 0; i < NS * NG; i++) G[i] = 0.0;
        ellipsoid_rows(&P.v[3], NS, x, s, C, NX);
    }
};

// freeflyer/definition.jl:376-442: three ellipsoidal obstacles and the space-station flight-space SDF
//   s_4 = -logsumexp(delta[:, k]; t = hom)  (src/utils/helper.jl:623-662), dG = -softmax weights on the six slacks of node k.
// par: mass [0], J [1..9], Jinv [10..18] (dynamics pack), hom [19], then 3 x {H(9), c(3)} from [20].
template <>
struct Constr<SCPB_MODEL_FREEFLYER> {
    static constexpr int NS = 4, NX = 13, NU = 6, NG = 6, NISS = 6, NOBS = 3;
    __device__ static constexpr int gcol(int k, int j) { return 1 + NISS * k + j; }
    __device__ static void eval(const ModelPar &P, double, int, int k, const double *x, const double *, const double *p,
                                double *s, double *C, double *D, double *G)
    {
        for (int i = 0; i < NS * NX; i++) C[i] = 0.0;
        for (int i = 0; i < NS * NU; i++) D[i] = 0.0;
        for (int i = 0; i < NS * NG; i++) G[i] = 0.0;
        ellipsoid_rows(&P.v[20], NOBS, x, s, C, NX);
        const double hom = P.v[19];
        double dl[NISS], a = -CUDART_INF;
        for (int j = 0; j < NISS; j++) { dl[j] = p[gcol(k, j)]; a = fmax(a, hom * dl[j]); }
        double E = 0.0;
        for (int j = 0; j < NISS; j++) E += exp(hom * dl[j] - a);
        s[3] = -((a + log(E)) / hom);
        for (int j = 0; j < NISS; j++) G[3 * NG + j] = -(exp(hom * dl[j] - a) / E);
    }
};

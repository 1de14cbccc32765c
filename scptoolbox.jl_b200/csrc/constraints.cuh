// constraints.cuh -- device packs for the nonconvex path constraints s(t,k,x,u,p) <= 0 and their Jacobians
// C = ds/dx, D = ds/du, G = ds/dp  (the device twins of traj.s/C/D/G, src/parser/problem.jl:560-587,
// consumed by add_nonconvex_constraints!, src/solvers/scp.jl:744-794).
// Outputs are dense, row-major: s[NS], C[NS][NX], D[NS][NU], G[NS][NP]; entries that are structurally zero
// for a pack are simply never referenced by the host template.
#pragma once
#include "models.cuh"

template <int ID>
struct Constr {
    static constexpr int NS = 0;
};

// starship_flip/definition.jl:704-810.  par: [9] rate_delay, [10] deltadot_max, [11] gamma_gs, [12] thetamax2,
// [8] tau_s (shared with the dynamics pack).
template <>
struct Constr<SCPB_MODEL_STARSHIP> {
    static constexpr int NS = 23, NX = 8, NU = 3, NP = 10;
    __device__ static void eval(const ModelPar &P, double t, int N, const double *x, const double *u, const double *p,
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

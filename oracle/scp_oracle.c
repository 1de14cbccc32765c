/* scp_oracle.c -- CPU fp64 ORACLE (test infrastructure, see scp_oracle.h).
 *
 * Restates, function by function, the reference's discretization path:
 *   discretize!          src/solvers/discretization.jl:160-217
 *   derivs_foh           src/solvers/discretization.jl:235-286
 *   set_update_matrices  src/solvers/discretization.jl:354-406
 *   rk4_generic/core     src/utils/helper.jl:451-501, 411-424
 *   linterp/get_interval src/utils/helper.jl:84-118
 *   propagate (FOH)      src/solvers/discretization.jl:515-562
 * and the user dynamics of the example problems (file:line at each model).
 *
 * Build with -ffp-contract=off so that the arithmetic is the plain IEEE
 * sequence Julia executes (no FMA contraction), in particular the LinRange
 * time arithmetic that decides the starship phase switch (`t <= tau_s`).
 *
 * Parity status: unpinned against the reference binary (no Julia here);
 * pinned by the known-answer checks in tests/test_oracle_discretize.py.
 */
#include "scp_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define IDX(i, j, ld) ((i) + (size_t)(j) * (ld))

/* ------------------------------------------------------------------ */
/* Julia LinRange element: lerpi(j, d, a, b) = (1 - j/d)*a + (j/d)*b   */
/* (Base range.jl, used by scp.jl:147 and discretization.jl:197)       */
static double linrange_at(double a, double b, int j, int d)
{
    double t = (double)j / (double)d;
    return (1.0 - t) * a + t * b;
}

/* ------------------------------------------------------------------ */
/* Models                                                              */

/* --- double integrator with friction, free final time (new definition on
 * the dynamics of double_integrator/parameters.jl:64: f = [x2; u - g]) --- */
static void dblint_f(const orc_model *m, const double *x, const double *u, const double *p, double *f)
{
    double g = m->par[0];
    f[0] = p[0] * x[1];
    f[1] = p[0] * (u[0] - g);
}
static void dblint_A(const orc_model *m, const double *p, double *A)
{
    (void)m;
    memset(A, 0, sizeof(double) * 4);
    A[IDX(0, 1, 2)] = p[0];
}
static void dblint_B(const orc_model *m, const double *p, double *B)
{
    (void)m;
    B[0] = 0.0;
    B[1] = p[0];
}

/* --- rocket: rocket_landing/parameters.jl:110-121 (A_c, B_c, p_c), scaled by
 * the time-dilation parameter p[0] (free final time PTR variant) --- */
static void rocket_mats(const orc_model *m, double *Ac, double *Bc, double *pc)
{
    const double *g = &m->par[0], *w = &m->par[3];
    double alpha = m->par[6];
    double S[9], S2[9];
    /* skew(w), helper.jl:65-70 */
    memset(S, 0, sizeof S);
    S[IDX(0, 1, 3)] = -w[2]; S[IDX(0, 2, 3)] = w[1]; S[IDX(1, 2, 3)] = -w[0];
    S[IDX(1, 0, 3)] = w[2];  S[IDX(2, 0, 3)] = -w[1]; S[IDX(2, 1, 3)] = w[0];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double s = 0.0;
            for (int k = 0; k < 3; k++) s += S[IDX(i, k, 3)] * S[IDX(k, j, 3)];
            S2[IDX(i, j, 3)] = s;
        }
    memset(Ac, 0, sizeof(double) * 49);
    memset(Bc, 0, sizeof(double) * 28);
    memset(pc, 0, sizeof(double) * 7);
    for (int i = 0; i < 3; i++) {
        Ac[IDX(i, 3 + i, 7)] = 1.0;
        for (int j = 0; j < 3; j++) {
            Ac[IDX(3 + i, j, 7)] = -S2[IDX(i, j, 3)];
            Ac[IDX(3 + i, 3 + j, 7)] = -2.0 * S[IDX(i, j, 3)];
        }
        Bc[IDX(3 + i, i, 7)] = 1.0;
        pc[3 + i] = g[i];
    }
    Bc[IDX(6, 3, 7)] = -alpha;
}
static void rocket_f(const orc_model *m, const double *x, const double *u, const double *p, double *f)
{
    double Ac[49], Bc[28], pc[7];
    rocket_mats(m, Ac, Bc, pc);
    for (int i = 0; i < 7; i++) {
        double s = pc[i];
        for (int j = 0; j < 7; j++) s += Ac[IDX(i, j, 7)] * x[j];
        for (int j = 0; j < 4; j++) s += Bc[IDX(i, j, 7)] * u[j];
        f[i] = p[0] * s;
    }
}

/* --- starship: starship_flip/definition.jl:498-550 (dynamics) --- */
enum { SS_m = 0, SS_J, SS_lcg, SS_lcp, SS_CD, SS_ae, SS_rd, SS_g0, SS_taus };

static double starship_tdil(const orc_model *m, double t, const double *p)
{
    double taus = m->par[SS_taus];
    /* definition.jl:521 */
    return (t <= taus) ? p[0] / taus : p[1] / (1.0 - taus);
}
static void starship_f(const orc_model *m, double t, const double *x, const double *u, const double *p, double *f)
{
    const double *P = m->par;
    double v0 = x[2], v1 = x[3], th = x[4], om = x[5], dd = x[7];
    double T = u[0], de = u[1];
    double tdil = starship_tdil(m, t, p);
    double leng = -P[SS_lcg];
    double lcp = P[SS_lcp] - P[SS_lcg];
    double ei0 = cos(th), ei1 = sin(th);   /* parameters.jl:124 */
    double ej0 = -sin(th), ej1 = cos(th);  /* parameters.jl:125 */
    double Tv0 = T * (-sin(de) * ei0 + cos(de) * ej0);
    double Tv1 = T * (-sin(de) * ei1 + cos(de) * ej1);
    double MT = leng * T * sin(de);
    double nv = sqrt(v0 * v0 + v1 * v1);
    double D0 = -P[SS_CD] * nv * v0, D1 = -P[SS_CD] * nv * v1;
    double MD = -lcp * (D0 * ei0 + D1 * ei1);
    f[0] = v0;
    f[1] = v1;
    f[2] = (Tv0 + D0) / P[SS_m] + 0.0;
    f[3] = (Tv1 + D1) / P[SS_m] + (-P[SS_g0]);
    f[4] = om;
    f[5] = (MT + MD) / P[SS_J];
    f[6] = P[SS_ae] * T;
    f[7] = (de - dd) / P[SS_rd];
    for (int i = 0; i < 8; i++) f[i] *= tdil;
}
/* definition.jl:562-593 */
static void starship_A(const orc_model *m, double t, const double *x, const double *u, const double *p, double *A)
{
    const double *P = m->par;
    double v0 = x[2], v1 = x[3], th = x[4];
    double T = u[0], de = u[1];
    double tdil = starship_tdil(m, t, p);
    double lcp = P[SS_lcp] - P[SS_lcg];
    double ei0 = cos(th), ei1 = sin(th), ej0 = -sin(th), ej1 = cos(th);
    double nv = sqrt(v0 * v0 + v1 * v1);
    double D0 = -P[SS_CD] * nv * v0, D1 = -P[SS_CD] * nv * v1;
    double dTv0 = T * (-sin(de) * ej0 + cos(de) * -ei0);
    double dTv1 = T * (-sin(de) * ej1 + cos(de) * -ei1);
    /* grad_v D = -CD (|v| I + v v'/|v|), symmetric */
    double G00 = -P[SS_CD] * (nv + v0 * v0 / nv);
    double G01 = -P[SS_CD] * (v0 * v1 / nv);
    double G11 = -P[SS_CD] * (nv + v1 * v1 / nv);
    /* grad_v MD = -lcp * G' * ei */
    double gMD0 = -lcp * (G00 * ei0 + G01 * ei1);
    double gMD1 = -lcp * (G01 * ei0 + G11 * ei1);
    double dthMD = -lcp * (D0 * ej0 + D1 * ej1);
    memset(A, 0, sizeof(double) * 64);
    A[IDX(0, 2, 8)] = 1.0;
    A[IDX(1, 3, 8)] = 1.0;
    A[IDX(2, 2, 8)] = G00 / P[SS_m];
    A[IDX(2, 3, 8)] = G01 / P[SS_m];
    A[IDX(3, 2, 8)] = G01 / P[SS_m];
    A[IDX(3, 3, 8)] = G11 / P[SS_m];
    A[IDX(2, 4, 8)] = dTv0 / P[SS_m];
    A[IDX(3, 4, 8)] = dTv1 / P[SS_m];
    A[IDX(4, 5, 8)] = 1.0;
    A[IDX(5, 2, 8)] = gMD0 / P[SS_J];
    A[IDX(5, 3, 8)] = gMD1 / P[SS_J];
    A[IDX(5, 4, 8)] = dthMD / P[SS_J];
    A[IDX(7, 7, 8)] = -1.0 / P[SS_rd];
    for (int i = 0; i < 64; i++) A[i] *= tdil;
}
/* definition.jl:595-624 */
static void starship_B(const orc_model *m, double t, const double *x, const double *u, const double *p, double *B)
{
    const double *P = m->par;
    double th = x[4];
    double T = u[0], de = u[1];
    double tdil = starship_tdil(m, t, p);
    double leng = -P[SS_lcg];
    double ei0 = cos(th), ei1 = sin(th), ej0 = -sin(th), ej1 = cos(th);
    double dT0 = -sin(de) * ei0 + cos(de) * ej0;
    double dT1 = -sin(de) * ei1 + cos(de) * ej1;
    double dd0 = T * (-cos(de) * ei0 - sin(de) * ej0);
    double dd1 = T * (-cos(de) * ei1 - sin(de) * ej1);
    memset(B, 0, sizeof(double) * 24);
    B[IDX(2, 0, 8)] = dT0 / P[SS_m];
    B[IDX(3, 0, 8)] = dT1 / P[SS_m];
    B[IDX(2, 1, 8)] = dd0 / P[SS_m];
    B[IDX(3, 1, 8)] = dd1 / P[SS_m];
    B[IDX(5, 0, 8)] = (leng * sin(de)) / P[SS_J];
    B[IDX(5, 1, 8)] = (leng * T * cos(de)) / P[SS_J];
    B[IDX(6, 0, 8)] = P[SS_ae];
    B[IDX(7, 1, 8)] = 1.0 / P[SS_rd];
    for (int i = 0; i < 24; i++) B[i] *= tdil;
}

/* --- quadrotor: quadrotor/definition.jl:140-186 --- */
static void quad_f(const orc_model *m, const double *x, const double *u, const double *p, double *f)
{
    const double *g = &m->par[0];
    for (int i = 0; i < 3; i++) {
        f[i] = x[3 + i];
        f[3 + i] = u[i] + g[i];
    }
    for (int i = 0; i < 6; i++) f[i] *= p[0];
}

/* --- freeflyer: freeflyer/definition.jl:224-284, quaternion.jl:190-214 --- */
enum { FF_m = 0, FF_J = 1, FF_Jinv = 10 };
static void skew3(const double *v, double *S)
{
    memset(S, 0, sizeof(double) * 9);
    S[IDX(0, 1, 3)] = -v[2]; S[IDX(0, 2, 3)] = v[1]; S[IDX(1, 2, 3)] = -v[0];
    S[IDX(1, 0, 3)] = v[2];  S[IDX(2, 0, 3)] = -v[1]; S[IDX(2, 1, 3)] = v[0];
}
/* skew(q, side) 4x4 (quaternion.jl:190-197), q = [v; w] scalar last */
static void qskew(const double *v, double w, int left, double *S)
{
    double Sv[9];
    skew3(v, Sv);
    double sg = left ? 1.0 : -1.0;
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++) S[IDX(i, j, 4)] = ((i == j) ? w : 0.0) + sg * Sv[IDX(i, j, 3)];
        S[IDX(i, 3, 4)] = v[i];
        S[IDX(3, i, 4)] = -v[i];
    }
    S[IDX(3, 3, 4)] = w;
}
static void ff_f(const orc_model *m, const double *x, const double *u, const double *p, double *f)
{
    const double *J = &m->par[FF_J], *Ji = &m->par[FF_Jinv];
    double mass = m->par[FF_m], tdil = p[0];
    const double *q = &x[6], *om = &x[10];
    double Sq[16];
    qskew(q, q[3], 1, Sq);
    /* f_q = 0.5 * vec(q * omega) = 0.5 * skew(q) * [omega; 0] */
    for (int i = 0; i < 4; i++) {
        double s = 0.0;
        for (int j = 0; j < 3; j++) s += Sq[IDX(i, j, 4)] * om[j];
        f[6 + i] = 0.5 * s;
    }
    double Jw[3], c[3], rhs[3];
    for (int i = 0; i < 3; i++) {
        Jw[i] = 0.0;
        for (int j = 0; j < 3; j++) Jw[i] += J[IDX(i, j, 3)] * om[j];
    }
    c[0] = om[1] * Jw[2] - om[2] * Jw[1];
    c[1] = om[2] * Jw[0] - om[0] * Jw[2];
    c[2] = om[0] * Jw[1] - om[1] * Jw[0];
    for (int i = 0; i < 3; i++) rhs[i] = u[3 + i] - c[i];
    for (int i = 0; i < 3; i++) {
        f[i] = x[3 + i];
        f[3 + i] = u[i] / mass;
        double s = 0.0;
        for (int j = 0; j < 3; j++) s += Ji[IDX(i, j, 3)] * rhs[j];
        f[10 + i] = s;
    }
    for (int i = 0; i < 13; i++) f[i] *= tdil;
}
static void ff_A(const orc_model *m, const double *x, const double *p, double *A)
{
    const double *J = &m->par[FF_J], *Ji = &m->par[FF_Jinv];
    double tdil = p[0];
    const double *q = &x[6], *om = &x[10];
    double SR[16], SL[16], So[9], SJw[9], Jw[3], M[9];
    qskew(om, 0.0, 0, SR); /* skew(Quaternion(omega), :R) */
    qskew(q, q[3], 1, SL); /* skew(q) */
    skew3(om, So);
    for (int i = 0; i < 3; i++) {
        Jw[i] = 0.0;
        for (int j = 0; j < 3; j++) Jw[i] += J[IDX(i, j, 3)] * om[j];
    }
    skew3(Jw, SJw);
    /* M = skew(om)*J - skew(J*om) */
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double s = 0.0;
            for (int k = 0; k < 3; k++) s += So[IDX(i, k, 3)] * J[IDX(k, j, 3)];
            M[IDX(i, j, 3)] = s - SJw[IDX(i, j, 3)];
        }
    memset(A, 0, sizeof(double) * 169);
    for (int i = 0; i < 3; i++) A[IDX(i, 3 + i, 13)] = 1.0;
    for (int i = 0; i < 4; i++) {
        for (int j = 0; j < 4; j++) A[IDX(6 + i, 6 + j, 13)] = 0.5 * SR[IDX(i, j, 4)];
        for (int j = 0; j < 3; j++) A[IDX(6 + i, 10 + j, 13)] = 0.5 * SL[IDX(i, j, 4)];
    }
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double s = 0.0;
            for (int k = 0; k < 3; k++) s += Ji[IDX(i, k, 3)] * M[IDX(k, j, 3)];
            A[IDX(10 + i, 10 + j, 13)] = -s;
        }
    for (int i = 0; i < 169; i++) A[i] *= tdil;
}
static void ff_B(const orc_model *m, const double *p, double *B)
{
    const double *Ji = &m->par[FF_Jinv];
    double mass = m->par[FF_m], tdil = p[0];
    memset(B, 0, sizeof(double) * 13 * 6);
    for (int i = 0; i < 3; i++) {
        B[IDX(3 + i, i, 13)] = 1.0 / mass;
        for (int j = 0; j < 3; j++) B[IDX(10 + i, 3 + j, 13)] = Ji[IDX(i, j, 3)];
    }
    for (int i = 0; i < 13 * 6; i++) B[i] *= tdil;
}

/* ------------------------------------------------------------------ */
/* ---------------------------------------------------------------- planar rendezvous (impulsive RCS)
 * rendezvous_planar/definition.jl:147-243, parameters.jl:87-111: x = [r(2) v(2) theta omega], u[0..2] = (f-, f+, f0)
 * (the other 9 inputs are references / absolute values and do not enter the dynamics), p = [tdil];
 * par: m, J, lu, lv, n.  xh = (1,0), yh = (0,1); uh = (-cos th, sin th), vh = (-sin th, -cos th).
 * k < 0 selects the impulse: f returns the jump of (v, omega) only and is NOT scaled by tdil. */
static void rdv_f(const orc_model *m, int k, const double *x, const double *u, const double *p, double *f)
{
    const double ms = m->par[0], J = m->par[1], lu = m->par[2], lv = m->par[3], n = m->par[4];
    const double th = x[4], fm = u[0], fp = u[1], f0 = u[2], tdil = p[0];
    const double uh0 = -cos(th), uh1 = sin(th), vh0 = -sin(th), vh1 = -cos(th);
    for (int i = 0; i < 6; i++) f[i] = 0.0;
    f[2] = ((fm + fp) * uh0 + f0 * vh0) / ms;
    f[3] = ((fm + fp) * uh1 + f0 * vh1) / ms;
    f[5] = ((fp - fm) * lv - f0 * lu) / J;
    if (k >= 0) {
        f[0] = x[2]; f[1] = x[3];
        f[2] += 2.0 * n * x[3];                        /* (2 n yh.v) xh */
        f[3] += 3.0 * n * n * x[1] - 2.0 * n * x[2];   /* (3 n^2 yh.r - 2 n xh.v) yh */
        f[4] = x[5];
        for (int i = 0; i < 6; i++) f[i] *= tdil;
    }
}
static void rdv_A(const orc_model *m, const double *x, const double *u, const double *p, double *A)
{
    const double ms = m->par[0], n = m->par[4];
    const double th = x[4], fm = u[0], fp = u[1], f0 = u[2], tdil = p[0];
    /* d uh / d th = -vh = (sin th, cos th); d vh / d th = uh = (-cos th, sin th) */
    const double duh0 = sin(th), duh1 = cos(th), dvh0 = -cos(th), dvh1 = sin(th);
    for (int i = 0; i < 36; i++) A[i] = 0.0;
    A[IDX(0, 2, 6)] = 1.0; A[IDX(1, 3, 6)] = 1.0;
    A[IDX(3, 1, 6)] = 3.0 * n * n;                         /* 3 n^2 yh yh' */
    A[IDX(2, 3, 6)] = 2.0 * n; A[IDX(3, 2, 6)] = -2.0 * n; /* 2 n (xh yh' - yh xh') */
    A[IDX(2, 4, 6)] = ((fm + fp) * duh0 + f0 * dvh0) / ms;
    A[IDX(3, 4, 6)] = ((fm + fp) * duh1 + f0 * dvh1) / ms;
    A[IDX(4, 5, 6)] = 1.0;
    for (int i = 0; i < 36; i++) A[i] *= tdil;
}
static void rdv_B(const orc_model *m, int k, const double *x, const double *p, double *B)
{
    const double ms = m->par[0], J = m->par[1], lu = m->par[2], lv = m->par[3];
    const double th = x[4], tdil = p[0];
    const double uh0 = -cos(th), uh1 = sin(th), vh0 = -sin(th), vh1 = -cos(th);
    for (int i = 0; i < 6 * m->nu; i++) B[i] = 0.0;
    B[IDX(2, 0, 6)] = uh0 / ms; B[IDX(3, 0, 6)] = uh1 / ms; B[IDX(5, 0, 6)] = -lv / J;
    B[IDX(2, 1, 6)] = uh0 / ms; B[IDX(3, 1, 6)] = uh1 / ms; B[IDX(5, 1, 6)] = lv / J;
    B[IDX(2, 2, 6)] = vh0 / ms; B[IDX(3, 2, 6)] = vh1 / ms; B[IDX(5, 2, 6)] = -lu / J;
    if (k >= 0)
        for (int i = 0; i < 6 * m->nu; i++) B[i] *= tdil;
}

void orc_f(const orc_model *m, double t, int k, const double *x, const double *u, const double *p, double *f)
{
    switch (m->model_id) {
    case ORC_MODEL_RENDEZVOUS2D: rdv_f(m, k, x, u, p, f); break;
    case ORC_MODEL_DBLINT: dblint_f(m, x, u, p, f); break;
    case ORC_MODEL_ROCKET: rocket_f(m, x, u, p, f); break;
    case ORC_MODEL_STARSHIP: starship_f(m, t, x, u, p, f); break;
    case ORC_MODEL_QUADROTOR: quad_f(m, x, u, p, f); break;
    case ORC_MODEL_FREEFLYER: ff_f(m, x, u, p, f); break;
    default: break;
    }
}
void orc_A(const orc_model *m, double t, int k, const double *x, const double *u, const double *p, double *A)
{
    (void)k;
    switch (m->model_id) {
    case ORC_MODEL_RENDEZVOUS2D: rdv_A(m, x, u, p, A); break;
    case ORC_MODEL_DBLINT: dblint_A(m, p, A); break;
    case ORC_MODEL_ROCKET: {
        double Bc[28], pc[7];
        rocket_mats(m, A, Bc, pc);
        for (int i = 0; i < 49; i++) A[i] *= p[0];
        break;
    }
    case ORC_MODEL_STARSHIP: starship_A(m, t, x, u, p, A); break;
    case ORC_MODEL_QUADROTOR:
        memset(A, 0, sizeof(double) * 36);
        for (int i = 0; i < 3; i++) A[IDX(i, 3 + i, 6)] = p[0];
        break;
    case ORC_MODEL_FREEFLYER: ff_A(m, x, p, A); break;
    default: break;
    }
}
void orc_B(const orc_model *m, double t, int k, const double *x, const double *u, const double *p, double *B)
{
    switch (m->model_id) {
    case ORC_MODEL_RENDEZVOUS2D: rdv_B(m, k, x, p, B); break;
    case ORC_MODEL_DBLINT: dblint_B(m, p, B); break;
    case ORC_MODEL_ROCKET: {
        double Ac[49], pc[7];
        rocket_mats(m, Ac, B, pc);
        for (int i = 0; i < 28; i++) B[i] *= p[0];
        break;
    }
    case ORC_MODEL_STARSHIP: starship_B(m, t, x, u, p, B); break;
    case ORC_MODEL_QUADROTOR:
        memset(B, 0, sizeof(double) * 24);
        for (int i = 0; i < 3; i++) B[IDX(3 + i, i, 6)] = p[0];
        break;
    case ORC_MODEL_FREEFLYER: ff_B(m, p, B); break;
    default: break;
    }
}
/* F[:, id_t] = f / p[id_t]  (starship definition.jl:626-633, quadrotor :177-183,
 * freeflyer :273-281); the LTI models use the same time-dilation convention. */
void orc_F(const orc_model *m, double t, int k, const double *x, const double *u, const double *p, double *F)
{
    int nx = m->nx, np = m->np;
    int id_t = 0;
    double f[32];
    memset(F, 0, sizeof(double) * (size_t)nx * np);
    if (m->model_id == ORC_MODEL_STARSHIP) id_t = (t <= m->par[SS_taus]) ? 0 : 1;
    orc_f(m, t, k, x, u, p, f);
    for (int i = 0; i < nx; i++) F[IDX(i, id_t, nx)] = f[i] / p[id_t];
}

/* integration actions (helper.jl:492-496): freeflyer renormalises the quaternion
 * (freeflyer/definition.jl:69-82) */
static void integ_actions(const orc_model *m, double *X)
{
    if (m->model_id == ORC_MODEL_FREEFLYER) {
        double *q = &X[6];
        double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
        for (int i = 0; i < 4; i++) q[i] = q[i] / n;
    }
}

/* ------------------------------------------------------------------ */
/* dense LU with partial pivoting (what Julia's `Phi \ I(nx)` does via LAPACK
 * getrf/getrs): returns inv(Phi) in iPhi; 0 on success */
static int dense_inverse(int n, const double *Phi, double *iPhi, double *work)
{
    double *L = work; /* n*n */
    int piv[64];
    memcpy(L, Phi, sizeof(double) * n * n);
    for (int c = 0; c < n; c++) {
        int pr = c;
        double pm = fabs(L[IDX(c, c, n)]);
        for (int r2 = c + 1; r2 < n; r2++)
            if (fabs(L[IDX(r2, c, n)]) > pm) { pm = fabs(L[IDX(r2, c, n)]); pr = r2; }
        piv[c] = pr;
        if (pm == 0.0) return -1;
        if (pr != c)
            for (int j = 0; j < n; j++) {
                double tmp = L[IDX(c, j, n)];
                L[IDX(c, j, n)] = L[IDX(pr, j, n)];
                L[IDX(pr, j, n)] = tmp;
            }
        for (int r2 = c + 1; r2 < n; r2++) {
            double l = L[IDX(r2, c, n)] / L[IDX(c, c, n)];
            L[IDX(r2, c, n)] = l;
            for (int j = c + 1; j < n; j++) L[IDX(r2, j, n)] -= l * L[IDX(c, j, n)];
        }
    }
    for (int j = 0; j < n; j++) {
        double *y = &iPhi[(size_t)j * n];
        for (int i = 0; i < n; i++) y[i] = (i == j) ? 1.0 : 0.0;
        for (int c = 0; c < n; c++)
            if (piv[c] != c) { double tmp = y[c]; y[c] = y[piv[c]]; y[piv[c]] = tmp; }
        for (int i = 0; i < n; i++)
            for (int k2 = 0; k2 < i; k2++) y[i] -= L[IDX(i, k2, n)] * y[k2];
        for (int i = n - 1; i >= 0; i--) {
            for (int k2 = i + 1; k2 < n; k2++) y[i] -= L[IDX(i, k2, n)] * y[k2];
            y[i] /= L[IDX(i, i, n)];
        }
    }
    return 0;
}

static void matmul(int m_, int k_, int n_, const double *A, const double *B, double *C)
{
    for (int j = 0; j < n_; j++)
        for (int i = 0; i < m_; i++) {
            double s = 0.0;
            for (int k2 = 0; k2 < k_; k2++) s += A[IDX(i, k2, m_)] * B[IDX(k2, j, k_)];
            C[IDX(i, j, m_)] = s;
        }
}

typedef struct {
    int nx, nu, np, len;
    int ox, oA, oBm, oBp, oF, or_, oE; /* DiscretizationIndices, discretization.jl:113-144 */
} disc_idx;

static disc_idx make_idx(int nx, int nu, int np)
{
    disc_idx id;
    id.nx = nx; id.nu = nu; id.np = np;
    id.ox = 0;
    id.oA = nx;
    id.oBm = id.oA + nx * nx;
    id.oBp = id.oBm + nx * nu;
    id.oF = id.oBp + nx * nu;
    id.or_ = id.oF + nx * np;
    id.oE = id.or_ + nx;
    id.len = id.oE + nx * nx;
    return id;
}

typedef struct {
    const orc_model *m;
    disc_idx id;
    int k;               /* 1-based segment */
    double t1, t2;       /* t_span */
    const double *uk, *ukp1, *p;
    double *wk;          /* scratch */
    int err;
} foh_ctx;

/* derivs_foh, discretization.jl:235-286 */
static void derivs_foh(foh_ctx *c, double t, const double *V, double *dV)
{
    const orc_model *m = c->m;
    int nx = c->id.nx, nu = c->id.nu, np = c->id.np;
    double *w = c->wk;
    double *u = w;             w += nu;
    double *f = w;             w += nx;
    double *A = w;             w += nx * nx;
    double *B = w;             w += nx * nu;
    double *F = w;             w += nx * np;
    double *r = w;             w += nx;
    double *iPhi = w;          w += nx * nx;
    double *lu = w;            w += nx * nx;
    const double *x = &V[c->id.ox];
    const double *Phi = &V[c->id.oA];

    /* linterp on the 2-point span (helper.jl:107-118) */
    double ts = fmax(c->t1, fmin(c->t2, t));
    double cc = (c->t2 - ts) / (c->t2 - c->t1);
    for (int i = 0; i < nu; i++) u[i] = cc * c->uk[i] + (1.0 - cc) * c->ukp1[i];
    double sm = (c->t2 - t) / (c->t2 - c->t1);
    double sp = (t - c->t1) / (c->t2 - c->t1);

    orc_f(m, t, c->k, x, u, c->p, f);
    orc_A(m, t, c->k, x, u, c->p, A);
    orc_B(m, t, c->k, x, u, c->p, B);
    orc_F(m, t, c->k, x, u, c->p, F);
    /* r = f - A x - B u - F p */
    for (int i = 0; i < nx; i++) {
        double ax = 0.0, bu = 0.0, fp = 0.0;
        for (int j = 0; j < nx; j++) ax += A[IDX(i, j, nx)] * x[j];
        for (int j = 0; j < nu; j++) bu += B[IDX(i, j, nx)] * u[j];
        for (int j = 0; j < np; j++) fp += F[IDX(i, j, nx)] * c->p[j];
        r[i] = f[i] - ax - bu - fp;
    }
    if (dense_inverse(nx, Phi, iPhi, lu) != 0) c->err = -1;

    memcpy(&dV[c->id.ox], f, sizeof(double) * nx);
    matmul(nx, nx, nx, A, Phi, &dV[c->id.oA]);
    /* reference order: B_m = sm*B, then iPhi*B_m (discretization.jl:260-270) */
    for (int j = 0; j < nu; j++)
        for (int i = 0; i < nx; i++) {
            double s1 = 0.0, s2 = 0.0;
            for (int k2 = 0; k2 < nx; k2++) {
                s1 += iPhi[IDX(i, k2, nx)] * (sm * B[IDX(k2, j, nx)]);
                s2 += iPhi[IDX(i, k2, nx)] * (sp * B[IDX(k2, j, nx)]);
            }
            dV[c->id.oBm + IDX(i, j, nx)] = s1;
            dV[c->id.oBp + IDX(i, j, nx)] = s2;
        }
    matmul(nx, nx, np, iPhi, F, &dV[c->id.oF]);
    matmul(nx, nx, 1, iPhi, r, &dV[c->id.or_]);
    /* E = I(nx) (scp.jl:149) => dE = iPhi */
    memcpy(&dV[c->id.oE], iPhi, sizeof(double) * nx * nx);
}

/* rk4_core_step, helper.jl:411-424 */
static void rk4_step_foh(foh_ctx *c, double *X, double t, double tp, double *k1, double *k2, double *k3,
                         double *k4, double *xt)
{
    int n = c->id.len;
    double h = tp - t;
    derivs_foh(c, t, X, k1);
    for (int i = 0; i < n; i++) xt[i] = X[i] + h / 2 * k1[i];
    derivs_foh(c, t + h / 2, xt, k2);
    for (int i = 0; i < n; i++) xt[i] = X[i] + h / 2 * k2[i];
    derivs_foh(c, t + h / 2, xt, k3);
    for (int i = 0; i < n; i++) xt[i] = X[i] + h * k3[i];
    derivs_foh(c, t + h, xt, k4);
    for (int i = 0; i < n; i++) X[i] = X[i] + h / 6 * (k1[i] + 2 * k2[i] + 2 * k3[i] + k4[i]);
}

int orc_discretize_foh(const orc_model *m, int N, int Nsub, const double *t_grid,
                       const double *xd, const double *ud, const double *p,
                       const double *iSx_diag, double feas_tol,
                       double *A, double *Bm, double *Bp, double *F, double *r, double *E,
                       double *defect, int *feas)
{
    int nx = m->nx, nu = m->nu, np = m->np;
    disc_idx id = make_idx(nx, nu, np);
    int n = id.len;
    size_t wk_sz = (size_t)nu + nx + nx * nx + nx * nu + (size_t)nx * np + nx + nx * nx +
                   (size_t)nx * (nx + nu + np) + nx * nx + 16;
    double *buf = (double *)malloc(sizeof(double) * (6 * (size_t)n + wk_sz));
    if (!buf) return -2;
    double *V = buf, *k1 = V + n, *k2 = k1 + n, *k3 = k2 + n, *k4 = k3 + n, *xt = k4 + n;
    foh_ctx c;
    c.m = m; c.id = id; c.p = p; c.wk = xt + n; c.err = 0;
    *feas = 1;

    for (int k = 1; k <= N - 1; k++) {
        /* V0 = [x_k; vec(I); 0...] (discretization.jl:177-185) */
        memset(V, 0, sizeof(double) * n);
        for (int i = 0; i < nx; i++) V[id.oA + IDX(i, i, nx)] = 1.0;
        memcpy(&V[id.ox], &xd[(size_t)(k - 1) * nx], sizeof(double) * nx);
        c.k = k;
        c.t1 = t_grid[k - 1];
        c.t2 = t_grid[k];
        c.uk = &ud[(size_t)(k - 1) * nu];
        c.ukp1 = &ud[(size_t)k * nu];
        /* rk4_generic over t_subgrid = LinRange(t[k], t[k+1], Nsub) */
        for (int j = 1; j < Nsub; j++) {
            double t = linrange_at(c.t1, c.t2, j - 1, Nsub - 1);
            double tp = linrange_at(c.t1, c.t2, j, Nsub - 1);
            rk4_step_foh(&c, V, t, tp, k1, k2, k3, k4, xt);
            integ_actions(m, V);
        }
        /* set_update_matrices, discretization.jl:354-406 */
        const double *Ak = &V[id.oA];
        size_t s = (size_t)(k - 1);
        memcpy(&A[s * nx * nx], Ak, sizeof(double) * nx * nx);
        matmul(nx, nx, nu, Ak, &V[id.oBm], &Bm[s * nx * nu]);
        matmul(nx, nx, nu, Ak, &V[id.oBp], &Bp[s * nx * nu]);
        matmul(nx, nx, np, Ak, &V[id.oF], &F[s * nx * np]);
        matmul(nx, nx, 1, Ak, &V[id.or_], &r[s * nx]);
        matmul(nx, nx, nx, Ak, &V[id.oE], &E[s * nx * nx]);
        /* defect, discretization.jl:205-210 */
        double nrm = 0.0;
        for (int i = 0; i < nx; i++) {
            double d = xd[(size_t)k * nx + i] - V[id.ox + i];
            defect[s * nx + i] = d;
            double a = fabs(iSx_diag[i] * d);
            if (a > nrm || a != a) nrm = a;
        }
        if (nrm > feas_tol) *feas = 0;
    }
    int err = c.err;
    free(buf);
    return err;
}

int orc_discretize_foh_batch(const orc_model *m, int nb, int N, int Nsub, const double *t_grid,
                             const double *xd, const double *ud, const double *p,
                             const double *iSx_diag, double feas_tol,
                             double *A, double *Bm, double *Bp, double *F, double *r, double *E,
                             double *defect, int *feas, int nthreads)
{
    int nx = m->nx, nu = m->nu, np = m->np;
    int err = 0;
    (void)nthreads;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(dynamic) reduction(| : err)
#endif
    for (int b = 0; b < nb; b++) {
        size_t s = (size_t)b, M = (size_t)(N - 1);
        int e = orc_discretize_foh(m, N, Nsub, t_grid, &xd[s * nx * N], &ud[s * nu * N], &p[s * np],
                                   iSx_diag, feas_tol, &A[s * M * nx * nx], &Bm[s * M * nx * nu],
                                   &Bp[s * M * nx * nu], &F[s * M * nx * np], &r[s * M * nx],
                                   &E[s * M * nx * nx], &defect[s * M * nx], &feas[b]);
        err |= (e != 0);
    }
    return err ? -1 : 0;
}

/* ------------------------------------------------------------------ */
/* IMPULSE discretization.  V = [x; vec(Phi); vec(PF); Pr; vec(PE)]  (DiscretizationIndices without B blocks,
 * discretization.jl:113-144); derivs_impulse (discretization.jl:304-340): u = 0 during the coast. */
typedef struct {
    const orc_model *m;
    int k, ox, oA, oF, or_, oE, len;
    const double *p;
    double *wk;
    int err;
} imp_ctx;

static void derivs_impulse(imp_ctx *c, double t, const double *V, double *dV)
{
    const orc_model *m = c->m;
    int nx = m->nx, nu = m->nu, np = m->np;
    double *w = c->wk;
    double *u0 = w;            w += nu;
    double *f = w;             w += nx;
    double *A = w;             w += nx * nx;
    double *F = w;             w += nx * np;
    double *r = w;             w += nx;
    double *iPhi = w;          w += nx * nx;
    double *lu = w;            w += nx * nx;
    const double *x = &V[c->ox];
    const double *Phi = &V[c->oA];
    for (int i = 0; i < nu; i++) u0[i] = 0.0;
    orc_f(m, t, c->k, x, u0, c->p, f);
    orc_A(m, t, c->k, x, u0, c->p, A);
    orc_F(m, t, c->k, x, u0, c->p, F);
    for (int i = 0; i < nx; i++) {   /* r = f - A x - F p */
        double ax = 0.0, fp = 0.0;
        for (int j = 0; j < nx; j++) ax += A[IDX(i, j, nx)] * x[j];
        for (int j = 0; j < np; j++) fp += F[IDX(i, j, nx)] * c->p[j];
        r[i] = f[i] - ax - fp;
    }
    if (dense_inverse(nx, Phi, iPhi, lu) != 0) c->err = -1;
    memcpy(&dV[c->ox], f, sizeof(double) * nx);
    matmul(nx, nx, nx, A, Phi, &dV[c->oA]);
    matmul(nx, nx, np, iPhi, F, &dV[c->oF]);
    matmul(nx, nx, 1, iPhi, r, &dV[c->or_]);
    memcpy(&dV[c->oE], iPhi, sizeof(double) * nx * nx);
}

static void rk4_step_imp(imp_ctx *c, double *X, double t, double tp, double *k1, double *k2, double *k3,
                         double *k4, double *xt)
{
    int n = c->len;
    double h = tp - t;
    derivs_impulse(c, t, X, k1);
    for (int i = 0; i < n; i++) xt[i] = X[i] + h / 2 * k1[i];
    derivs_impulse(c, t + h / 2, xt, k2);
    for (int i = 0; i < n; i++) xt[i] = X[i] + h / 2 * k2[i];
    derivs_impulse(c, t + h / 2, xt, k3);
    for (int i = 0; i < n; i++) xt[i] = X[i] + h * k3[i];
    derivs_impulse(c, t + h, xt, k4);
    for (int i = 0; i < n; i++) X[i] = X[i] + h / 6 * (k1[i] + 2 * k2[i] + 2 * k3[i] + k4[i]);
}

int orc_discretize_impulse(const orc_model *m, int N, int Nsub, const double *t_grid,
                           const double *xd, const double *ud, const double *p,
                           const double *iSx_diag, double feas_tol,
                           double *A, double *Bm, double *F, double *r, double *E,
                           double *defect, int *feas)
{
    int nx = m->nx, nu = m->nu, np = m->np;
    imp_ctx c;
    c.m = m; c.p = p; c.err = 0;
    c.ox = 0; c.oA = nx; c.oF = c.oA + nx * nx; c.or_ = c.oF + nx * np; c.oE = c.or_ + nx; c.len = c.oE + nx * nx;
    int n = c.len;
    size_t wk_sz = (size_t)nu + nx + nx * nx + (size_t)nx * np + nx + 2 * nx * nx + 16;
    double *buf = (double *)malloc(sizeof(double) * (6 * (size_t)n + wk_sz + nx + (size_t)nx * nu));
    if (!buf) return -2;
    double *V = buf, *k1 = V + n, *k2 = k1 + n, *k3 = k2 + n, *k4 = k3 + n, *xt = k4 + n;
    c.wk = xt + n;
    double *jump = c.wk + wk_sz, *Btk = jump + nx;
    *feas = 1;
    for (int k = 1; k <= N - 1; k++) {
        const double *xk = &xd[(size_t)(k - 1) * nx], *uk = &ud[(size_t)(k - 1) * nu];
        memset(V, 0, sizeof(double) * n);
        for (int i = 0; i < nx; i++) V[c.oA + IDX(i, i, nx)] = 1.0;
        /* impulse-update the state: xk_plus = xk + f(tk, -k, xk, uk, p)  (discretization.jl:186-193) */
        orc_f(m, t_grid[k - 1], -k, xk, uk, p, jump);
        for (int i = 0; i < nx; i++) V[c.ox + i] = xk[i] + jump[i];
        c.k = k;
        for (int j = 1; j < Nsub; j++) {
            double t = linrange_at(t_grid[k - 1], t_grid[k], j - 1, Nsub - 1);
            double tp = linrange_at(t_grid[k - 1], t_grid[k], j, Nsub - 1);
            rk4_step_imp(&c, V, t, tp, k1, k2, k3, k4, xt);
            integ_actions(m, V);
        }
        const double *Ak = &V[c.oA];
        size_t s = (size_t)(k - 1);
        memcpy(&A[s * nx * nx], Ak, sizeof(double) * nx * nx);
        orc_B(m, t_grid[k - 1], -k, xk, uk, p, Btk);      /* B_k = A_k * B(tk, -k, xk, uk, p)  (:384-390) */
        matmul(nx, nx, nu, Ak, Btk, &Bm[s * nx * nu]);
        matmul(nx, nx, np, Ak, &V[c.oF], &F[s * nx * np]);
        matmul(nx, nx, 1, Ak, &V[c.or_], &r[s * nx]);
        matmul(nx, nx, nx, Ak, &V[c.oE], &E[s * nx * nx]);
        double nrm = 0.0;
        for (int i = 0; i < nx; i++) {
            double d = xd[(size_t)k * nx + i] - V[c.ox + i];
            defect[s * nx + i] = d;
            double a = fabs(iSx_diag[i] * d);
            if (a > nrm || a != a) nrm = a;
        }
        if (nrm > feas_tol) *feas = 0;
    }
    int err = c.err;
    free(buf);
    return err;
}

int orc_propagate_impulse(const orc_model *m, int N, int res, const double *t_grid,
                          const double *xd, const double *ud, const double *p, double *xc)
{
    int nx = m->nx, nu = m->nu;
    int subres = (res + (N - 1) - 1) / (N - 1);      /* ceil(Int, res / (N - 1)) */
    double X[32], k1[32], k2[32], k3[32], k4[32], xt[32], u0[64], jump[32];
    for (int i = 0; i < nu; i++) u0[i] = 0.0;
    memcpy(xc, xd, sizeof(double) * nx);             /* xc_intvl[1] = xd[:, 1] */
    size_t col = 1;
    for (int k = 1; k <= N - 1; k++) {
        const double *xk = &xd[(size_t)(k - 1) * nx], *uk = &ud[(size_t)(k - 1) * nu];
        orc_f(m, t_grid[k - 1], -k, xk, uk, p, jump);
        for (int i = 0; i < nx; i++) X[i] = xk[i] + jump[i];
        memcpy(&xc[col * nx], X, sizeof(double) * nx); col++;
        for (int j = 1; j < subres; j++) {
            double t = linrange_at(t_grid[k - 1], t_grid[k], j - 1, subres - 1);
            double tp = linrange_at(t_grid[k - 1], t_grid[k], j, subres - 1);
            double h = tp - t;
            orc_f(m, t, N, X, u0, p, k1);
            for (int i = 0; i < nx; i++) xt[i] = X[i] + h / 2 * k1[i];
            orc_f(m, t + h / 2, N, xt, u0, p, k2);
            for (int i = 0; i < nx; i++) xt[i] = X[i] + h / 2 * k2[i];
            orc_f(m, t + h / 2, N, xt, u0, p, k3);
            for (int i = 0; i < nx; i++) xt[i] = X[i] + h * k3[i];
            orc_f(m, t + h, N, xt, u0, p, k4);
            for (int i = 0; i < nx; i++) X[i] = X[i] + h / 6 * (k1[i] + 2 * k2[i] + 2 * k3[i] + k4[i]);
            integ_actions(m, X);
            memcpy(&xc[col * nx], X, sizeof(double) * nx); col++;
        }
    }
    return 0;
}

/* propagate, FOH branch (discretization.jl:533-538): RK4 of f over LinRange(0,1,res)
 * with u = linear interpolation over the whole grid; the reference's k(t) helper
 * always evaluates to N (discretization.jl:530). xc is nx*res. */
static void lin_u(int N, int nu, const double *t_grid, const double *ud, double t, double *u)
{
    double ts = fmax(t_grid[0], fmin(t_grid[N - 1], t));
    int k = 0;
    for (int i = 0; i < N; i++) k += (ts > t_grid[i]);
    if (k == 0) k = 1;
    if (k > N - 1) k = N - 1;
    double c = (t_grid[k] - ts) / (t_grid[k] - t_grid[k - 1]);
    for (int i = 0; i < nu; i++) u[i] = c * ud[(size_t)(k - 1) * nu + i] + (1.0 - c) * ud[(size_t)k * nu + i];
}

int orc_propagate_foh(const orc_model *m, int N, int res, const double *t_grid,
                      const double *xd, const double *ud, const double *p, double *xc)
{
    int nx = m->nx, nu = m->nu;
    double X[32], k1[32], k2[32], k3[32], k4[32], xt[32], u[32];
    memcpy(X, xd, sizeof(double) * nx);
    memcpy(xc, X, sizeof(double) * nx);
    for (int j = 1; j < res; j++) {
        double t = linrange_at(0.0, 1.0, j - 1, res - 1);
        double tp = linrange_at(0.0, 1.0, j, res - 1);
        double h = tp - t;
        lin_u(N, nu, t_grid, ud, t, u);
        orc_f(m, t, N, X, u, p, k1);
        for (int i = 0; i < nx; i++) xt[i] = X[i] + h / 2 * k1[i];
        lin_u(N, nu, t_grid, ud, t + h / 2, u);
        orc_f(m, t + h / 2, N, xt, u, p, k2);
        for (int i = 0; i < nx; i++) xt[i] = X[i] + h / 2 * k2[i];
        orc_f(m, t + h / 2, N, xt, u, p, k3);
        for (int i = 0; i < nx; i++) xt[i] = X[i] + h * k3[i];
        lin_u(N, nu, t_grid, ud, t + h, u);
        orc_f(m, t + h, N, xt, u, p, k4);
        for (int i = 0; i < nx; i++) X[i] = X[i] + h / 6 * (k1[i] + 2 * k2[i] + 2 * k3[i] + k4[i]);
        integ_actions(m, X);
        memcpy(&xc[(size_t)j * nx], X, sizeof(double) * nx);
    }
    return 0;
}

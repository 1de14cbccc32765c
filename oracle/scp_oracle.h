/* scp_oracle.h -- CPU fp64 ORACLE for the SCP hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is a plain-C restatement of the reference algorithm
 * (UW-ACL/SCPToolbox.jl @ d5b6691); only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * path (scptoolbox.jl_b200/) never links or imports anything in oracle/.
 *
 * PARITY STATUS: "parity unpinned" -- the reference ships no golden vectors
 * for this path (status-only asserts) and cannot be
 * executed in this image (no Julia, no ECOS).  The oracle is pinned instead by
 * internal known-answer checks (tests/test_oracle_*.py): Phi(t_k)=I, LTI
 * A_k == expm(A dt), finite-difference Jacobians, RK4 roll-out => zero defect.
 *
 * All matrices are column-major (Julia memory order).
 */
#ifndef SCP_ORACLE_H
#define SCP_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

enum {
    ORC_MODEL_DBLINT = 1,   /* double integrator, free final time (new PTR definition)            */
    ORC_MODEL_ROCKET = 2,   /* Mars powered-descent LTI rocket  (rocket_landing/parameters.jl:110) */
    ORC_MODEL_STARSHIP = 3, /* starship_flip/definition.jl:498-637                                 */
    ORC_MODEL_QUADROTOR = 4,/* quadrotor/definition.jl:140-186                                     */
    ORC_MODEL_FREEFLYER = 5,/* freeflyer/definition.jl:224-284                                     */
    ORC_MODEL_RENDEZVOUS2D = 6 /* rendezvous_planar/definition.jl:147-243: impulsive RCS thrust (k < 0 = the jump) */
};

#define ORC_MAX_PAR 64

typedef struct {
    int model_id;
    int nx, nu, np;
    double par[ORC_MAX_PAR]; /* model constants, layout documented in include/scpb.h */
} orc_model;

/* user dynamics and Jacobians at (t, k, x, u, p); k is the 1-based segment index */
void orc_f(const orc_model *m, double t, int k, const double *x, const double *u, const double *p, double *f);
void orc_A(const orc_model *m, double t, int k, const double *x, const double *u, const double *p, double *A);
void orc_B(const orc_model *m, double t, int k, const double *x, const double *u, const double *p, double *B);
void orc_F(const orc_model *m, double t, int k, const double *x, const double *u, const double *p, double *F);

/* discretize! with FOH (discretization.jl:160-217).
 * xd[nx*N], ud[nu*N], p[np], t_grid[N], iSx_diag[nx];
 * outputs A[nx*nx*(N-1)], Bm/Bp[nx*nu*(N-1)], F[nx*np*(N-1)], r[nx*(N-1)],
 * E[nx*nx*(N-1)], defect[nx*(N-1)], *feas (1 = dynamically feasible).
 * returns 0, or -1 on a singular Phi. */
int orc_discretize_foh(const orc_model *m, int N, int Nsub, const double *t_grid,
                       const double *xd, const double *ud, const double *p,
                       const double *iSx_diag, double feas_tol,
                       double *A, double *Bm, double *Bp, double *F, double *r, double *E,
                       double *defect, int *feas);

/* batch driver used for the CPU baseline: seeds are independent; OpenMP over seeds. */
int orc_discretize_foh_batch(const orc_model *m, int nb, int N, int Nsub, const double *t_grid,
                             const double *xd, const double *ud, const double *p,
                             const double *iSx_diag, double feas_tol,
                             double *A, double *Bm, double *Bp, double *F, double *r, double *E,
                             double *defect, int *feas, int nthreads);

/* propagate (discretization.jl:515-562, FOH branch): full RK4 roll-out on res points */
int orc_propagate_foh(const orc_model *m, int N, int res, const double *t_grid,
                      const double *xd, const double *ud, const double *p, double *xc);

/* discretize! with IMPULSE (discretization.jl:186-193 jump, :304-340 derivs_impulse, :384-390 B_k = A_k B(t_k,-k)):
 * same arrays as the FOH variant; Bp is not written (dyn.B has one block for IMPULSE). */
int orc_discretize_impulse(const orc_model *m, int N, int Nsub, const double *t_grid,
                           const double *xd, const double *ud, const double *p,
                           const double *iSx_diag, double feas_tol,
                           double *A, double *Bm, double *F, double *r, double *E,
                           double *defect, int *feas);

/* propagate, IMPULSE branch (discretization.jl:539-558): per interval RK4 of the coasting dynamics from the
 * impulse-updated state on LinRange(t_k, t_k+1, subres), subres = ceil(res/(N-1)); xc is nx*(1+(N-1)*subres). */
int orc_propagate_impulse(const orc_model *m, int N, int res, const double *t_grid,
                          const double *xd, const double *ud, const double *p, double *xc);

#ifdef __cplusplus
}
#endif
#endif

// conic_sn.cuh -- supernodal LDL' : the per-panel routines of the next kernel generation.
//
// STATUS: executed and checked on the CPU (lane emulation, scpb_debug_kkt_solve_sn with mode 1,
// tests/test_conic_symbolic.py); wired into k_ipm_solve behind SCPB_SUPERNODAL=1 but NOT yet run on a GPU -- the
// default path is the scalar level-scheduled one.
//
// One warp owns one (supernode, seed) item of a supernodal level (conic_symbolic.h, "supernodal program"): a dense
// column-major R x w panel, D on its diagonal, unit-lower L below.  The routines are written as a sequence of lane
// phases -- SN_LANES(l) { ... } SN_SYNC() -- whose bodies only read what earlier phases wrote: on the device a phase is
// the warp's 32 lanes followed by __syncwarp(), in the CPU build it is a plain loop over 32 lanes.  Per-lane state that
// crosses a phase lives in the warp's scratch (shared memory on the device), never in registers, which is what makes
// the two builds the same program.
#pragma once

#ifdef SN_EMULATE
#define SN_FN static inline
#define SN_LANES(l) for (int l = 0; l < 32; l++)
#define SN_SYNC() ((void)0)
#define SN_ATOMIC_SUB(p, v) (*(p) -= (v))
#define SN_LDCG(p) (*(p))
#else
#define SN_FN __device__ __forceinline__
#define SN_LANES(l) for (int l = (int)(threadIdx.x & 31), l##_go = 1; l##_go; l##_go = 0)
#define SN_SYNC() __syncwarp()
#define SN_ATOMIC_SUB(p, v) atomicAdd((p), -(v))
#define SN_LDCG(p) __ldcg(p)
#endif

#define SN_MAXROWS 64      // panel rows a warp can hold in its sweep scratch
#define SN_SCRATCH 384     // doubles of factor scratch per warp (panel copy): R*w must fit

struct SnProgram {          // device view of the supernodal program (all arrays shared by the batch)
    const int *first, *width, *nrows, *rows_ptr, *rows, *lvl_ptr, *lvl_nodes, *upd_xy, *sign;
    const int *panel_off, *upd_ptr, *upd_dst;
    int nlevels;
};

// ---- numeric factorisation of one panel + Schur update of the ancestors ------------------------------------
// P: this seed group's panels (entry e of seed sg at P[e*G + sg]); scr: >= R*w doubles of warp scratch
SN_FN void sn_factor_item(const SnProgram &S, int s, double *P, double *invD, int G, int sg, double delta_dyn,
                          double *scr)
{
    const int a = S.first[s], w = S.width[s], R = S.nrows[s];
    const int off = S.panel_off[s];
    SN_LANES(l) { for (int e = l; e < R * w; e += 32) scr[e] = SN_LDCG(&P[(size_t)(off + e) * G + sg]); }
    SN_SYNC();
    for (int c = 0; c < w; c++) {
        double d = scr[c + R * c];
        const double sgn = (double)S.sign[a + c];
        if (!(sgn * d > delta_dyn)) d = sgn * delta_dyn;   // dynamic regularisation keeps the expected inertia
        SN_LANES(l) {
            for (int r = c + 1 + l; r < R; r += 32) scr[r + R * c] /= d;
        }
        SN_SYNC();
        SN_LANES(l) {
            if (l == 0) { scr[c + R * c] = d; invD[(size_t)(a + c) * G + sg] = 1.0 / d; }
            const int span = (w - c - 1) * R;
            for (int idx = l; idx < span; idx += 32) {
                const int c2 = c + 1 + idx / R, r = idx % R;
                if (r >= c2) scr[r + R * c2] -= scr[r + R * c] * d * scr[c2 + R * c];
            }
        }
        SN_SYNC();
    }
    SN_LANES(l) { for (int e = l; e < R * w; e += 32) P[(size_t)(off + e) * G + sg] = scr[e]; }
    const int k0 = S.upd_ptr[s], k1 = S.upd_ptr[s + 1];
    SN_LANES(l) {
        for (int k = k0 + l; k < k1; k += 32) {
            const int xy = S.upd_xy[k], x = xy & 0xffff, y = xy >> 16;
            double u = 0.0;
            for (int c = 0; c < w; c++) u += scr[w + x + R * c] * scr[c + R * c] * scr[w + y + R * c];
            SN_ATOMIC_SUB(&P[(size_t)S.upd_dst[k] * G + sg], u);
        }
    }
    SN_SYNC();
}

// ---- forward substitution of one panel: x_S = L_SS^-1 x_S, then the rows below receive -L_below,S x_S ---------
// v: the substitution vector of this seed group (entry i of seed sg at v[i*G + sg]); xs: >= R doubles of warp scratch
SN_FN void sn_forward_item(const SnProgram &S, int s, const double *P, double *v, int G, int sg, double *xs)
{
    const int w = S.width[s], R = S.nrows[s];
    const int off = S.panel_off[s];
    const int *rows = S.rows + S.rows_ptr[s];
    SN_LANES(l) { for (int r = l; r < R; r += 32) xs[r] = (r < w) ? v[(size_t)rows[r] * G + sg] : 0.0; }
    SN_SYNC();
    for (int c = 0; c < w; c++) {
        SN_LANES(l) {
            const double xc = xs[c];
            for (int r = c + 1 + l; r < R; r += 32) xs[r] -= P[(size_t)(off + r + R * c) * G + sg] * xc;
        }
        SN_SYNC();
    }
    SN_LANES(l) {
        for (int r = l; r < R; r += 32) {
            double *t = &v[(size_t)rows[r] * G + sg];
            if (r < w) *t = xs[r]; else SN_ATOMIC_SUB(t, -xs[r]);   // siblings update common ancestors: atomic
        }
    }
    SN_SYNC();
}

// ---- backward substitution of one panel: x_S = L_SS^-T (x_S - L_below,S^T x_below) -----------------------------
// xs: >= R + 32 doubles of warp scratch
SN_FN void sn_backward_item(const SnProgram &S, int s, const double *P, double *v, int G, int sg, double *xs)
{
    const int w = S.width[s], R = S.nrows[s];
    const int off = S.panel_off[s];
    const int *rows = S.rows + S.rows_ptr[s];
    double *ps = xs + SN_MAXROWS;
    SN_LANES(l) { for (int r = l; r < R; r += 32) xs[r] = v[(size_t)rows[r] * G + sg]; }
    SN_SYNC();
    for (int c = w - 1; c >= 0; c--) {
        SN_LANES(l) {
            double acc = 0.0;
            for (int r = c + 1 + l; r < R; r += 32) acc += P[(size_t)(off + r + R * c) * G + sg] * xs[r];
            ps[l] = acc;
        }
        SN_SYNC();
        SN_LANES(l) {
            if (l == 0) {
                double t = 0.0;
                for (int j = 0; j < 32; j++) t += ps[j];
                xs[c] -= t;
            }
        }
        SN_SYNC();
    }
    SN_LANES(l) { for (int r = l; r < w; r += 32) v[(size_t)rows[r] * G + sg] = xs[r]; }
    SN_SYNC();
}

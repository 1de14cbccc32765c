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

// A panel is worked on by a GROUP of GS = 8, 16 or 32 lanes (small leaf panels would waste a whole warp): SN_LANES
// runs l over the lanes of the group, SN_SYNC synchronises the group only (groups of one warp may take different trip
// counts, so the mask matters).
#ifdef SN_EMULATE
#define SN_FN template <int GS> static inline
#define SN_LANES(l) for (int l = 0; l < GS; l++)
#define SN_SYNC() ((void)0)
#define SN_ATOMIC_SUB(p, v) (*(p) -= (v))
#define SN_LDCG(p) (*(p))
#else
#define SN_FN template <int GS> __device__ __forceinline__
#define SN_LANES(l) for (int l = (int)(threadIdx.x & (GS - 1)), l##_go = 1; l##_go; l##_go = 0)
#define SN_SYNC() __syncwarp(GS == 32 ? 0xffffffffu : (((1u << GS) - 1u) << ((threadIdx.x & 31) & ~(GS - 1))))
#define SN_ATOMIC_SUB(p, v) atomicAdd((p), -(v))
#define SN_LDCG(p) __ldcg(p)
#endif

#define SN_MAXROWS 64      // panel rows a 32-lane group can hold in its sweep scratch
#define SN_SCRATCH 384     // doubles of factor scratch per warp (panel copies of its groups): see sn_class()

// lane-group size of a panel: the smallest of 8 / 16 / 32 lanes that covers its rows and whose share of the warp's
// scratch (SN_SCRATCH * GS / 32 for the panel copy, (SN_MAXROWS + 32) * GS / 32 for the sweeps) holds it; 0 = too large
static inline int sn_class(int R, int w)
{
    for (int gs = 8; gs <= 32; gs *= 2)
        if (R <= gs && R * w <= SN_SCRATCH * gs / 32 && R + gs <= (SN_MAXROWS + 32) * gs / 32) return gs;
    if (R <= SN_MAXROWS && R * w <= SN_SCRATCH) return 32;
    return 0;
}

struct SnProgram {          // device view of the supernodal program (all arrays shared by the batch)
    const int *first, *width, *nrows, *rows_ptr, *rows, *lvl_ptr, *lvl_nodes, *upd_xy, *sign;
    const int *cls_ptr;     // [nlevels][4]: inside a level the supernodes are sorted by lane-group size 8 | 16 | 32
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
    SN_LANES(l) { for (int e = l; e < R * w; e += GS) scr[e] = SN_LDCG(&P[(size_t)(off + e) * G + sg]); }
    SN_SYNC();
    for (int c = 0; c < w; c++) {
        double d = scr[c + R * c];
        const double sgn = (double)S.sign[a + c];
        if (!(sgn * d > delta_dyn)) d = sgn * delta_dyn;   // dynamic regularisation keeps the expected inertia
        SN_LANES(l) {
            for (int r = c + 1 + l; r < R; r += GS) scr[r + R * c] /= d;
        }
        SN_SYNC();
        SN_LANES(l) {
            if (l == 0) { scr[c + R * c] = d; invD[(size_t)(a + c) * G + sg] = 1.0 / d; }
            const int span = (w - c - 1) * R;
            for (int idx = l; idx < span; idx += GS) {
                const int c2 = c + 1 + idx / R, r = idx % R;
                if (r >= c2) scr[r + R * c2] -= scr[r + R * c] * d * scr[c2 + R * c];
            }
        }
        SN_SYNC();
    }
    SN_LANES(l) { for (int e = l; e < R * w; e += GS) P[(size_t)(off + e) * G + sg] = scr[e]; }
    const int k0 = S.upd_ptr[s], k1 = S.upd_ptr[s + 1];
    SN_LANES(l) {
        for (int k = k0 + l; k < k1; k += GS) {
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
    SN_LANES(l) { for (int r = l; r < R; r += GS) xs[r] = (r < w) ? v[(size_t)rows[r] * G + sg] : 0.0; }
    SN_SYNC();
    for (int c = 0; c < w; c++) {
        SN_LANES(l) {
            const double xc = xs[c];
            for (int r = c + 1 + l; r < R; r += GS) xs[r] -= P[(size_t)(off + r + R * c) * G + sg] * xc;
        }
        SN_SYNC();
    }
    SN_LANES(l) {
        for (int r = l; r < R; r += GS) {
            double *t = &v[(size_t)rows[r] * G + sg];
            if (r < w) *t = xs[r]; else SN_ATOMIC_SUB(t, -xs[r]);   // siblings update common ancestors: atomic
        }
    }
    SN_SYNC();
}

// ---- backward substitution of one panel: x_S = L_SS^-T (x_S - L_below,S^T x_below) -----------------------------
// xs: >= SN_MAXROWS + GS doubles of group scratch (rows, then the partial sums)
SN_FN void sn_backward_item(const SnProgram &S, int s, const double *P, double *v, int G, int sg, double *xs)
{
    const int w = S.width[s], R = S.nrows[s];
    const int off = S.panel_off[s];
    const int *rows = S.rows + S.rows_ptr[s];
    double *ps = xs + (GS == 32 ? SN_MAXROWS : GS);
    SN_LANES(l) { for (int r = l; r < R; r += GS) xs[r] = v[(size_t)rows[r] * G + sg]; }
    SN_SYNC();
    for (int c = w - 1; c >= 0; c--) {
        SN_LANES(l) {
            double acc = 0.0;
            for (int r = c + 1 + l; r < R; r += GS) acc += P[(size_t)(off + r + R * c) * G + sg] * xs[r];
            ps[l] = acc;
        }
        SN_SYNC();
        SN_LANES(l) {
            if (l == 0) {
                double t = 0.0;
                for (int j = 0; j < GS; j++) t += ps[j];
                xs[c] -= t;
            }
        }
        SN_SYNC();
    }
    SN_LANES(l) { for (int r = l; r < w; r += GS) v[(size_t)rows[r] * G + sg] = xs[r]; }
    SN_SYNC();
}

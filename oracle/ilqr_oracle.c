/*
 * ilqr_oracle.c -- CPU ORACLE in plain C (test infrastructure; never linked into or
 * called by the product library).
 *
 * A sequential, one-problem-at-a-time restatement of IterativeLQR.jl's solve path,
 * array-of-arrays per problem like the reference, threaded over the batch with OpenMP
 * (one independent solver state per problem).  Every function cites the reference
 * file:line it restates (paths relative to /root/reference/).
 *
 * PARITY STATUS: "parity unpinned" against the Julia package (no Julia in this image;
 * the reference's tests hold no golden histories -- SURVEY.md 8c).  It is pinned to
 * oracle/ilqr_oracle.py (the literal numpy restatement, itself checked against the
 * reference's restated test files) by tests/test_oracle.py::test_c_oracle_matches_py_oracle at the north-star tolerance
 * (same iteration counts, histories to 1e-9 relative, trajectories to 1e-7) and, over 256 acrobot + 256 car problems with
 * libm-based model functions as well, by oracle/flip_rate_study.py (profiles/r2_flip_rate.json: 0 mismatches).
 *
 * ARITHMETIC CONTRACT: this file and the CUDA engine implement the same
 * floating-point specification -- IEEE binary64, no implicit contraction
 * (-ffp-contract=off), dot products accumulated in ascending index order as
 *     acc = a0*b0; acc = fma(a_k, b_k, acc)
 * "C += A*B" forms computed as C = C + dot(...), and the model functions taken from the
 * same generated header -- so the engine is expected to reproduce this oracle BIT FOR
 * BIT.  Two documented deviations from the literal Julia statement order (both at the
 * 1-ulp level, both unknowable for the real package anyway because it calls OpenBLAS):
 * (1) the AL cost is J = (sum_t g_t) + (sum_t AL terms) with the AL terms accumulated
 * separately (src/augmented_lagrangian.jl:41-63 adds them onto the running J);
 * (2) delta_grad_product = (sum_t Lx_t.dx_t) + (sum_t Lu_t.du_t) (src/forward_pass.jl:20
 * is one flat BLAS dot).
 *
 * Build: see oracle/build_oracle.py (gcc -O2 -ffp-contract=off -mfma -fopenmp,
 * -include <generated model header>).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "ilqr_cuda.h" /* ilqr_options only */
#ifndef ILQR_MODEL_GEN_H
#error "compile with -include <generated model header>"
#endif

enum { N = ILQR_N, M = ILQR_M, NP = ILQR_P, CS = ILQR_CS, CT = ILQR_CT };
#define DIM(x) ((x) > 0 ? (x) : 1)

typedef struct {
    int T, cap;
    /* ProblemData: src/data/problem.jl:3-23 */
    double *x, *u, *xb, *ub, *w, *z;
    /* ModelData / ObjectiveData: src/data/model.jl:5-10, src/data/objective.jl:3-10 */
    double *fx, *fu, *gx, *gu, *gxx, *guu, *gux;
    /* PolicyData: src/data/policy.jl:23-42 */
    double *K, *k, *P, *p, *Qx, *Qu, *Qxx, *Quu, *Qux;
    /* SolverData: src/data/solver.jl:4-18 (gradient split into its x and u blocks) */
    double *Lx, *Lu;
    double objective, max_violation, step_size;
    int status, iterations;
    /* AugmentedLagrangianCosts + ConstraintsData: src/augmented_lagrangian.jl:1-11 */
    double *c, *lam, *rho;
    int* a;
    /* instrumentation */
    double *h_cost, *h_gnorm, *h_viol, *h_alpha;
    int32_t* h_outer;
    uint8_t* h_status;
    uint32_t flags;
    int outer;
} prob_t;

typedef struct {
    int T, B, cap;
    ilqr_options opt;
    prob_t* probs;
} oracle_t;

/* -------------------------------------------------------------------------- helpers */
static inline double dotf(const double* a, int sa, const double* b, int sb, int len) {
    if (len <= 0) return 0.0;
    double acc = a[0] * b[0];
    for (int k = 1; k < len; ++k) acc = ilqr_fma(a[k * sa], b[k * sb], acc);
    return acc;
}
static inline int rows_at(const prob_t* s, int t) { return t < s->T - 1 ? CS : CT; }
static inline double* crow(double* base, const prob_t* s, int t) { return base + (size_t)t * CS; }
static inline int is_ineq(const prob_t* s, int t, int i) { return t < s->T - 1 ? ilqr_ineq_s(i) : ilqr_ineq_T(i); }

static double* zalloc(size_t n) { return (double*)calloc(n > 0 ? n : 1, sizeof(double)); }

/* -------------------------------------------------------------------------- costs */
/* cost(costs, states, actions, parameters): src/costs.jl:48-55 */
static double plain_cost(const prob_t* s, const double* X, const double* U) {
    const int T = s->T;
    double J = 0.0, g;
    double dummy[1] = {0};
    for (int t = 0; t < T - 1; ++t) {
        ilqr_cost_s(&g, X + t * N, U + t * M, s->w + (size_t)t * NP);
        J += g;
    }
    ilqr_cost_T(&g, X + (T - 1) * N, dummy, s->w + (size_t)(T - 1) * NP);
    J += g;
    return J;
}

/* constraint!: src/constraints.jl:66-73 via src/data/constraints.jl:19-21 */
static void eval_constraints(prob_t* s, const double* X, const double* U) {
    const int T = s->T;
    double dummy[1] = {0};
    (void)dummy;
#if ILQR_CS > 0
    for (int t = 0; t < T - 1; ++t) ilqr_con_s(crow(s->c, s, t), X + t * N, U + t * M, s->w + (size_t)t * NP);
#endif
#if ILQR_CT > 0
    ilqr_con_T(crow(s->c, s, T - 1), X + (T - 1) * N, dummy, s->w + (size_t)(T - 1) * NP);
#endif
    (void)X; (void)U; (void)T;
}

/* active_set!: src/augmented_lagrangian.jl:68-85 */
static void active_set(prob_t* s) {
    for (int t = 0; t < s->T; ++t) {
        const int r = rows_at(s, t);
        double* c = crow(s->c, s, t);
        double* lam = crow(s->lam, s, t);
        int* a = s->a + (size_t)t * CS;
        for (int i = 0; i < r; ++i) {
            a[i] = 1;
            if (is_ineq(s, t, i) && c[i] < 0.0 && lam[i] == 0.0) a[i] = 0;
        }
    }
}

/* cost(::AugmentedLagrangianCosts, ...): src/augmented_lagrangian.jl:39-66 */
static double al_cost(prob_t* s, const double* X, const double* U) {
    const double Jc = plain_cost(s, X, U);
    eval_constraints(s, X, U);
    active_set(s);
    double Jal = 0.0;
    for (int t = 0; t < s->T; ++t) {
        const int r = rows_at(s, t);
        if (r == 0) continue;
        const double* c = crow(s->c, s, t);
        const double* lam = crow(s->lam, s, t);
        const double* rho = crow(s->rho, s, t);
        const int* a = s->a + (size_t)t * CS;
        Jal += dotf(lam, 1, c, 1, r);
        for (int i = 0; i < r; ++i)
            if (a[i] == 1) Jal += (0.5 * rho[i]) * (c[i] * c[i]);
    }
    return Jc + Jal;
}

/* constraint_violation: src/data/constraints.jl:23-39 (inf-norm always) */
static double violation_of_buffer(const prob_t* s) {
    double mv = 0.0;
    for (int t = 0; t < s->T; ++t) {
        const int r = rows_at(s, t);
        const double* c = s->c + (size_t)t * CS;
        for (int i = 0; i < r; ++i) {
            const double ci = c[i];
            const double v = is_ineq(s, t, i) ? (ci > 0.0 ? ci : 0.0) : fabs(ci);
            if (v > mv) mv = v;
        }
    }
    return mv;
}

/* cost!: src/data/methods.jl:13-30.  mode 0 = nominal, 1 = current */
static void cost_bang(prob_t* s, int mode) {
    const double* X = mode ? s->x : s->xb;
    const double* U = mode ? s->u : s->ub;
    if (CS + CT > 0) {
        s->objective = al_cost(s, X, U);
        /* Q2: max_violation is ALWAYS taken on the current trajectory and overwrites c */
        eval_constraints(s, s->x, s->u);
        s->max_violation = violation_of_buffer(s);
    } else {
        s->objective = plain_cost(s, X, U);
    }
}

/* -------------------------------------------------------------------------- gradients! */
/* src/gradients.jl:92-98 -> :1-8 (dynamics.jl:41-50), :10-21 (costs.jl:57-84), :83-90
 * (constraints.jl:75-87), :54-80 (AL terms).  Always on the nominal trajectory. */
static void gradients(prob_t* s) {
    const int T = s->T;
    double dummy[1] = {0};
    for (int t = 0; t < T; ++t) {
        const double* x = s->xb + t * N;
        const double* u = t < T - 1 ? s->ub + t * M : dummy;
        const double* w = s->w + (size_t)t * NP;
        double* gx = s->gx + t * N;
        double* gxx = s->gxx + t * N * N;
        double hxx[N * N];
        if (t < T - 1) {
            double* gu = s->gu + t * M;
            double* guu = s->guu + t * M * M;
            double* gux = s->gux + t * M * N;
            double huu[M * M], hux[M * N];
            ilqr_dyn_jac(s->fx + t * N * N, s->fu + t * N * M, x, u, w);              /* overwrite */
            ilqr_cost_s_grad(gx, gu, hxx, huu, hux, x, u, w);                        /* gx, gu overwrite */
            for (int i = 0; i < N * N; ++i) gxx[i] = gxx[i] + hxx[i];                /* Q1: accumulate */
            for (int i = 0; i < M * M; ++i) guu[i] = guu[i] + huu[i];
            for (int i = 0; i < M * N; ++i) gux[i] = gux[i] + hux[i];
#if ILQR_CS > 0
            {
                double cx[CS * N], cu[CS * M], cxt[CS * N], cut[CS * M], d[CS], v[CS];
                const double* c = crow(s->c, s, t);
                const double* lam = crow(s->lam, s, t);
                const double* rho = crow(s->rho, s, t);
                const int* a = s->a + (size_t)t * CS;
                ilqr_con_s_jac(cx, cu, x, u, w);
                for (int i = 0; i < CS; ++i) {
                    d[i] = rho[i] * (double)a[i];      /* :56-58 */
                    v[i] = lam[i] + d[i] * c[i];       /* :59-62 */
                }
                for (int j = 0; j < N; ++j) gx[j] = gx[j] + dotf(cx + j * CS, 1, v, 1, CS);       /* :63 */
                for (int j = 0; j < N; ++j)
                    for (int i = 0; i < CS; ++i) cxt[i + j * CS] = d[i] * cx[i + j * CS];          /* :66 */
                for (int l = 0; l < N; ++l)
                    for (int j = 0; j < N; ++j)
                        gxx[j + l * N] = gxx[j + l * N] + dotf(cx + j * CS, 1, cxt + l * CS, 1, CS); /* :67 */
                for (int b = 0; b < M; ++b) gu[b] = gu[b] + dotf(cu + b * CS, 1, v, 1, CS);       /* :72 */
                for (int b = 0; b < M; ++b)
                    for (int i = 0; i < CS; ++i) cut[i + b * CS] = d[i] * cu[i + b * CS];          /* :75 */
                for (int e = 0; e < M; ++e)
                    for (int b = 0; b < M; ++b)
                        guu[b + e * M] = guu[b + e * M] + dotf(cu + b * CS, 1, cut + e * CS, 1, CS); /* :76 */
                for (int j = 0; j < N; ++j)
                    for (int b = 0; b < M; ++b)
                        gux[b + j * M] = gux[b + j * M] + dotf(cu + b * CS, 1, cxt + j * CS, 1, CS); /* :79 */
            }
#endif
        } else {
            ilqr_cost_T_grad(gx, hxx, x, u, w);
            for (int i = 0; i < N * N; ++i) gxx[i] = gxx[i] + hxx[i];
#if ILQR_CT > 0
            {
                double cx[CT * N], cxt[CT * N], d[CT], v[CT];
                const double* c = crow(s->c, s, t);
                const double* lam = crow(s->lam, s, t);
                const double* rho = crow(s->rho, s, t);
                const int* a = s->a + (size_t)t * CS;
                ilqr_con_T_jac(cx, x, u, w);
                for (int i = 0; i < CT; ++i) {
                    d[i] = rho[i] * (double)a[i];
                    v[i] = lam[i] + d[i] * c[i];
                }
                for (int j = 0; j < N; ++j) gx[j] = gx[j] + dotf(cx + j * CT, 1, v, 1, CT);
                for (int j = 0; j < N; ++j)
                    for (int i = 0; i < CT; ++i) cxt[i + j * CT] = d[i] * cx[i + j * CT];
                for (int l = 0; l < N; ++l)
                    for (int j = 0; j < N; ++j)
                        gxx[j + l * N] = gxx[j + l * N] + dotf(cx + j * CT, 1, cxt + l * CT, 1, CT);
            }
#endif
        }
    }
}

/* -------------------------------------------------------------------------- backward_pass! */
/* upper Cholesky in place, unblocked, stops at the first non-positive pivot like
 * LAPACK potrf (info ignored by the reference: src/backward_pass.jl:68-69, Q3).
 * rinv[j] = 1 / A[j,j] of whatever the diagonal holds afterwards (the factor's diagonal, or
 * the unfactored / failed entries -- "garbage in" exactly as potrs would consume them). */
static int chol_upper(double* A, double* rinv) {
    int info = 0;
    for (int j = 0; j < M && info == 0; ++j) {
        double ajj = A[j + j * M];
        for (int k = 0; k < j; ++k) ajj = ilqr_fma(-A[k + j * M], A[k + j * M], ajj);
        if (!(ajj > 0.0)) {
            A[j + j * M] = ajj;
            info = j + 1;
            break;
        }
        const double ujj = sqrt(ajj);
        A[j + j * M] = ujj;
        const double r = 1.0 / ujj;
        for (int i = j + 1; i < M; ++i) {
            double sum = A[j + i * M];
            for (int k = 0; k < j; ++k) sum = ilqr_fma(-A[k + j * M], A[k + i * M], sum);
            A[j + i * M] = sum * r;
        }
    }
    for (int j = 0; j < M; ++j) rinv[j] = 1.0 / A[j + j * M];
    return info;
}
/* potrs 'U' for one right-hand side: solve U'U x = b in place (src/backward_pass.jl:72-73).
 * Contract: the triangular solves multiply by the reciprocal diagonal (as optimised BLAS
 * trsm kernels do) instead of dividing. */
static void chol_solve(const double* U, const double* rinv, double* b) {
    for (int i = 0; i < M; ++i) {
        double sum = b[i];
        for (int k = 0; k < i; ++k) sum = ilqr_fma(-U[k + i * M], b[k], sum);
        b[i] = sum * rinv[i];
    }
    for (int i = M - 1; i >= 0; --i) {
        double sum = b[i];
        for (int k = i + 1; k < M; ++k) sum = ilqr_fma(-U[i + k * M], b[k], sum);
        b[i] = sum * rinv[i];
    }
}

/* src/backward_pass.jl:39-90 + lagrangian_gradient! src/solve.jl:67-83 */
static void backward_pass(prob_t* s) {
    const int T = s->T;
    memcpy(s->P + (T - 1) * N * N, s->gxx + (T - 1) * N * N, sizeof(double) * N * N); /* :39 */
    memcpy(s->p + (T - 1) * N, s->gx + (T - 1) * N, sizeof(double) * N);             /* :40 */
    for (int t = T - 2; t >= 0; --t) {
        const double* fx = s->fx + t * N * N;
        const double* fu = s->fu + t * N * M;
        const double* Pn = s->P + (t + 1) * N * N;
        const double* pn = s->p + (t + 1) * N;
        double* Qx = s->Qx + t * N;
        double* Qu = s->Qu + t * M;
        double* Qxx = s->Qxx + t * N * N;
        double* Quu = s->Quu + t * M * M;
        double* Qux = s->Qux + t * M * N;
        double* K = s->K + t * M * N;
        double* k = s->k + t * M;
        double xxh[N * N], uxh[M * N], uu[M * M], uxt[M * N], rinv[M];
        for (int i = 0; i < N; ++i) Qx[i] = dotf(fx + i * N, 1, pn, 1, N) + s->gx[t * N + i];   /* :44-45 */
        for (int a = 0; a < M; ++a) Qu[a] = dotf(fu + a * N, 1, pn, 1, N) + s->gu[t * M + a];   /* :48-49 */
        for (int l = 0; l < N; ++l)
            for (int i = 0; i < N; ++i) xxh[i + l * N] = dotf(fx + i * N, 1, Pn + l * N, 1, N);  /* :52 */
        for (int j = 0; j < N; ++j)
            for (int i = 0; i < N; ++i)
                Qxx[i + j * N] = dotf(xxh + i, N, fx + j * N, 1, N) + s->gxx[t * N * N + i + j * N]; /* :53-54 */
        for (int l = 0; l < N; ++l)
            for (int a = 0; a < M; ++a) uxh[a + l * M] = dotf(fu + a * N, 1, Pn + l * N, 1, N);  /* :57 (= :62) */
        for (int b = 0; b < M; ++b)
            for (int a = 0; a < M; ++a)
                Quu[a + b * M] = dotf(uxh + a, M, fu + b * N, 1, N) + s->guu[t * M * M + a + b * M]; /* :58-59 */
        for (int j = 0; j < N; ++j)
            for (int a = 0; a < M; ++a)
                Qux[a + j * M] = dotf(uxh + a, M, fx + j * N, 1, N) + s->gux[t * M * N + a + j * M]; /* :63-64 */
        memcpy(uu, Quu, sizeof(uu));                                                             /* :68 */
        if (chol_upper(uu, rinv) != 0) s->flags |= ILQR_FLAG_CHOL_FAIL;                                /* :69 */
        for (int j = 0; j < N; ++j) {                                                            /* :70,72,74 */
            double col[M];
            for (int a = 0; a < M; ++a) col[a] = Qux[a + j * M];
            chol_solve(uu, rinv, col);
            for (int a = 0; a < M; ++a) K[a + j * M] = -col[a];
        }
        {                                                                                        /* :71,73,75 */
            double col[M];
            for (int a = 0; a < M; ++a) col[a] = Qu[a];
            chol_solve(uu, rinv, col);
            for (int a = 0; a < M; ++a) k[a] = -col[a];
        }
        for (int j = 0; j < N; ++j)
            for (int a = 0; a < M; ++a) uxt[a + j * M] = dotf(Quu + a, M, K + j * M, 1, M);      /* :79 */
        double* P = s->P + t * N * N;
        double* p = s->p + t * N;
        for (int j = 0; j < N; ++j)
            for (int i = 0; i < N; ++i) {
                double v = dotf(K + i * M, 1, uxt + j * M, 1, M);                                /* :81 */
                v = v + dotf(K + i * M, 1, Qux + j * M, 1, M);                                   /* :82 */
                v = v + dotf(Qux + i * M, 1, K + j * M, 1, M);                                   /* :83 */
                P[i + j * N] = v + Qxx[i + j * N];                                               /* :84 */
            }
        for (int i = 0; i < N; ++i) {
            double v = dotf(uxt + i * M, 1, k, 1, M);                                            /* :86 */
            v = v + dotf(K + i * M, 1, Qu, 1, M);                                                /* :87 */
            v = v + dotf(Qux + i * M, 1, k, 1, M);                                               /* :88 */
            p[i] = v + Qx[i];                                                                    /* :89 */
        }
    }
    /* lagrangian_gradient!: src/solve.jl:73-81 (the x_T block is never written, Q10) */
    for (int t = 0; t < T - 1; ++t) {
        for (int i = 0; i < N; ++i) s->Lx[t * N + i] = s->Qx[t * N + i] - s->p[t * N + i];
        for (int a = 0; a < M; ++a) s->Lu[t * M + a] = s->Qu[t * M + a];
    }
}

static double gradient_norm(const prob_t* s) { /* norm(data.gradient, Inf): src/solve.jl:36 */
    double g = 0.0;
    for (int i = 0; i < (s->T - 1) * N; ++i) { const double a = fabs(s->Lx[i]); if (a > g || a != a) g = a; }
    for (int i = 0; i < (s->T - 1) * M; ++i) { const double a = fabs(s->Lu[i]); if (a > g || a != a) g = a; }
    return g;
}

/* -------------------------------------------------------------------------- forward_pass! */
/* trajectory_sensitivities + gradient' * trajectory: src/data/methods.jl:42-54, src/forward_pass.jl:19-20 */
static double delta_grad_product(prob_t* s) {
    const int T = s->T;
    double zx[N], zy[N], zu[M];
    double sx = 0.0, su = 0.0;
    for (int i = 0; i < N; ++i) zx[i] = 0.0;
    for (int t = 0; t < T - 1; ++t) {
        const double* K = s->K + t * M * N;
        const double* fx = s->fx + t * N * N;
        const double* fu = s->fu + t * N * M;
        for (int a = 0; a < M; ++a) zu[a] = s->k[t * M + a] + dotf(K + a, M, zx, 1, N);      /* :49-50 */
        for (int i = 0; i < N; ++i) {
            const double v = dotf(fu + i, N, zu, 1, M);                                      /* :51 */
            zy[i] = v + dotf(fx + i, N, zx, 1, N);                                           /* :52 */
        }
        for (int i = 0; i < N; ++i) sx = ilqr_fma(s->Lx[t * N + i], zx[i], sx);
        for (int a = 0; a < M; ++a) su = ilqr_fma(s->Lu[t * M + a], zu[a], su);
        for (int i = 0; i < N; ++i) zx[i] = zy[i];
    }
    return sx + su;
}

/* rollout!: src/rollout.jl:1-31 */
static void rollout_bang(prob_t* s, double alpha) {
    const int T = s->T;
    for (int i = 0; i < N; ++i) s->x[i] = s->xb[i];
    for (int t = 0; t < T - 1; ++t) {
        const double* K = s->K + t * M * N;
        double* u = s->u + t * M;
        const double* x = s->x + t * N;
        for (int a = 0; a < M; ++a) {
            double v = s->k[t * M + a] * alpha;              /* :24-25 */
            v = v + s->ub[t * M + a];                        /* :26 */
            v = v + dotf(K + a, M, x, 1, N);                 /* :27 */
            v = v - dotf(K + a, M, s->xb + t * N, 1, N);     /* :28 */
            u[a] = v;
        }
        ilqr_dyn(s->x + (t + 1) * N, x, u, s->w + (size_t)t * NP); /* :29 */
    }
}

static int count_trials(const ilqr_options* o) { /* src/forward_pass.jl:28-29: alpha >= min, iteration <= 25 */
    int n = 0;
    double a = 1.0;
    while (a >= o->min_step_size && n < 25) { ++n; a *= 0.5; }
    return n;
}

/* src/forward_pass.jl:1-56 */
static void forward_pass(prob_t* s, const ilqr_options* o) {
    s->status = 0;
    const double J_prev = s->objective;
    const double dgp = (o->line_search == ILQR_LINE_SEARCH_ARMIJO) ? delta_grad_product(s) : 0.0;
    const int trials = count_trials(o);
    double alpha = 1.0;
    for (int it = 0; it < trials; ++it) {
        rollout_bang(s, alpha);
        cost_bang(s, 1);
        const double J = s->objective;
        if (!(J - J == 0.0)) s->flags |= ILQR_FLAG_NONFINITE;
        if (J <= J_prev + (1.0e-4 * alpha) * dgp) {                        /* :44 */
            memcpy(s->xb, s->x, sizeof(double) * (size_t)s->T * N);       /* :46 */
            memcpy(s->ub, s->u, sizeof(double) * (size_t)(s->T - 1) * M);
            s->status = 1;
            break;
        }
        alpha *= 0.5;
    }
    s->step_size = alpha;
}

/* -------------------------------------------------------------------------- solve.jl */
static void record(prob_t* s, double gnorm) {
    const int r = s->iterations - 1;
    if (r >= 0 && r < s->cap) {
        s->h_cost[r] = s->objective; s->h_gnorm[r] = gnorm; s->h_viol[r] = s->max_violation;
        s->h_alpha[r] = s->step_size; s->h_outer[r] = s->outer; s->h_status[r] = (uint8_t)s->status;
    }
}

static void reset_data(prob_t* s) { /* reset!(data): src/data/solver.jl:49-59 */
    s->objective = 0.0; s->max_violation = 0.0; s->status = 0; s->iterations = 0;
    memset(s->Lx, 0, sizeof(double) * DIM((s->T - 1) * N));
    memset(s->Lu, 0, sizeof(double) * DIM((s->T - 1) * M));
}

/* ilqr_solve!: src/solve.jl:1-54 */
static void ilqr_solve_inner(prob_t* s, const ilqr_options* o) {
    const int T = s->T;
    /* reset!(problem.model); reset!(problem.objective): :9-10 */
    memset(s->fx, 0, sizeof(double) * (T - 1) * N * N);
    memset(s->fu, 0, sizeof(double) * DIM((T - 1) * N * M));
    memset(s->gx, 0, sizeof(double) * T * N);
    memset(s->gu, 0, sizeof(double) * DIM((T - 1) * M));
    memset(s->gxx, 0, sizeof(double) * T * N * N);
    memset(s->guu, 0, sizeof(double) * DIM((T - 1) * M * M));
    memset(s->gux, 0, sizeof(double) * DIM((T - 1) * M * N));
    if (o->reset_cache) reset_data(s);                                    /* :12 */
    cost_bang(s, 0);                                                      /* :14 */
    gradients(s);                                                         /* :16 */
    backward_pass(s);                                                     /* :18 */
    double gnorm = gradient_norm(s);
    double obj_prev = s->objective;                                       /* :21 */
    for (int i = 1; i <= o->max_iterations; ++i) {
        forward_pass(s, o);                                               /* :23 */
        if (o->line_search != ILQR_LINE_SEARCH_NONE) {                    /* :27-33 */
            gradients(s);
            backward_pass(s);
            gnorm = gradient_norm(s);
        }
        s->iterations += 1;                                               /* :39 */
        record(s, gnorm);                                                 /* :40-45 */
        if (gnorm < o->lagrangian_gradient_tolerance) break;              /* :48 */
        if (fabs(s->objective - obj_prev) < o->objective_tolerance) break; /* :49 */
        obj_prev = s->objective;
        if (!s->status) break;                                            /* :50 */
    }
}

/* augmented_lagrangian_update!: src/augmented_lagrangian.jl:87-110 */
static void al_update(prob_t* s, const ilqr_options* o) {
    for (int t = 0; t < s->T; ++t) {
        const int r = rows_at(s, t);
        double* c = crow(s->c, s, t);
        double* lam = crow(s->lam, s, t);
        double* rho = crow(s->rho, s, t);
        for (int i = 0; i < r; ++i) {
            lam[i] = lam[i] + rho[i] * c[i];
            if (is_ineq(s, t, i)) lam[i] = lam[i] > 0.0 ? lam[i] : 0.0;  /* max(0.0, lam) */
            const double sc = o->scaling_penalty * rho[i];
            rho[i] = sc < o->max_penalty ? sc : o->max_penalty;           /* min(scaling*rho, max_penalty) */
        }
    }
}

/* solve!: src/solve.jl:137-143; constrained_ilqr_solve!: :88-129 */
static void solve_one(prob_t* s, const ilqr_options* o) {
    s->flags = 0;
    if (CS + CT == 0) {
        s->outer = 0;
        ilqr_solve_inner(s, o);
        return;
    }
    const int rows = (s->T - 1) * CS + CT;
    reset_data(s);                                                        /* :93 */
    for (int i = 0; i < rows; ++i) { s->lam[i] = 0.0; s->rho[i] = o->initial_constraint_penalty; } /* :96-103 */
    for (int i = 1; i <= o->max_dual_updates; ++i) {
        s->outer = i;
        ilqr_solve_inner(s, o);                                           /* :109 */
        cost_bang(s, 0);                                                  /* :113 */
        if (s->max_violation <= o->constraint_tolerance) break;           /* :117 */
        al_update(s, o);                                                  /* :120-122 */
    }
}

/* -------------------------------------------------------------------------- C entry points (ctypes) */
void oracle_dims(int32_t* n, int32_t* m, int32_t* p, int32_t* cs, int32_t* ct) { *n = N; *m = M; *p = NP; *cs = CS; *ct = CT; }
const char* oracle_model_hash(void) { return ILQR_MODEL_HASH; }
int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

oracle_t* oracle_create(int T, int B, int cap, const ilqr_options* opt) {
    oracle_t* h = (oracle_t*)calloc(1, sizeof(oracle_t));
    h->T = T; h->B = B; h->cap = cap > 0 ? cap : 1000; h->opt = *opt;
    h->probs = (prob_t*)calloc((size_t)B, sizeof(prob_t));
    const size_t rows = (size_t)(T - 1) * CS + DIM(CT) + CS; /* slack so crow(T-1) + CT stays in range */
    for (int b = 0; b < B; ++b) {
        prob_t* s = &h->probs[b];
        s->T = T; s->cap = h->cap;
        s->x = zalloc((size_t)T * N); s->xb = zalloc((size_t)T * N);
        s->u = zalloc((size_t)(T - 1) * M); s->ub = zalloc((size_t)(T - 1) * M);
        s->w = zalloc((size_t)T * NP); s->z = zalloc(1);
        s->fx = zalloc((size_t)(T - 1) * N * N); s->fu = zalloc((size_t)(T - 1) * N * M);
        s->gx = zalloc((size_t)T * N); s->gu = zalloc((size_t)(T - 1) * M);
        s->gxx = zalloc((size_t)T * N * N); s->guu = zalloc((size_t)(T - 1) * M * M); s->gux = zalloc((size_t)(T - 1) * M * N);
        s->K = zalloc((size_t)(T - 1) * M * N); s->k = zalloc((size_t)(T - 1) * M);
        s->P = zalloc((size_t)T * N * N); s->p = zalloc((size_t)T * N);
        s->Qx = zalloc((size_t)(T - 1) * N); s->Qu = zalloc((size_t)(T - 1) * M);
        s->Qxx = zalloc((size_t)(T - 1) * N * N); s->Quu = zalloc((size_t)(T - 1) * M * M); s->Qux = zalloc((size_t)(T - 1) * M * N);
        s->Lx = zalloc((size_t)(T - 1) * N); s->Lu = zalloc((size_t)(T - 1) * M);
        s->c = zalloc(rows); s->lam = zalloc(rows); s->rho = zalloc(rows);
        s->a = (int*)calloc(rows, sizeof(int));
        for (size_t i = 0; i < rows; ++i) { s->rho[i] = 1.0; s->a[i] = 1; }   /* src/augmented_lagrangian.jl:17-22 */
        s->h_cost = zalloc(h->cap); s->h_gnorm = zalloc(h->cap); s->h_viol = zalloc(h->cap); s->h_alpha = zalloc(h->cap);
        s->h_outer = (int32_t*)calloc(h->cap, sizeof(int32_t)); s->h_status = (uint8_t*)calloc(h->cap, 1);
        s->objective = INFINITY; s->step_size = 1.0;                           /* src/data/solver.jl:37-39 */
    }
    return h;
}

void oracle_destroy(oracle_t* h) {
    if (!h) return;
    for (int b = 0; b < h->B; ++b) {
        prob_t* s = &h->probs[b];
        double* ptrs[] = {s->x, s->xb, s->u, s->ub, s->w, s->z, s->fx, s->fu, s->gx, s->gu, s->gxx, s->guu, s->gux, s->K, s->k,
                          s->P, s->p, s->Qx, s->Qu, s->Qxx, s->Quu, s->Qux, s->Lx, s->Lu, s->c, s->lam, s->rho,
                          s->h_cost, s->h_gnorm, s->h_viol, s->h_alpha};
        for (size_t i = 0; i < sizeof(ptrs) / sizeof(ptrs[0]); ++i) free(ptrs[i]);
        free(s->a); free(s->h_outer); free(s->h_status);
    }
    free(h->probs);
    free(h);
}

void oracle_set_options(oracle_t* h, const ilqr_options* opt) { h->opt = *opt; }

void oracle_initialize_controls(oracle_t* h, const double* u) { /* src/solver.jl:56-60 */
    const size_t len = (size_t)(h->T - 1) * M;
    for (int b = 0; b < h->B; ++b) memcpy(h->probs[b].ub, u + b * len, sizeof(double) * len);
}
void oracle_initialize_states(oracle_t* h, const double* x) { /* src/solver.jl:62-66 */
    const size_t len = (size_t)h->T * N;
    for (int b = 0; b < h->B; ++b) memcpy(h->probs[b].xb, x + b * len, sizeof(double) * len);
}
void oracle_set_parameters(oracle_t* h, const double* w) {
    const size_t len = (size_t)h->T * NP;
    for (int b = 0; b < h->B; ++b) memcpy(h->probs[b].w, w + b * len, sizeof(double) * len);
}

/* rollout: src/rollout.jl:33-42 */
void oracle_rollout(oracle_t* h, const double* x1, const double* u, double* x_out) {
    const int T = h->T;
#pragma omp parallel for schedule(static)
    for (int b = 0; b < h->B; ++b) {
        double* X = x_out + (size_t)b * T * N;
        const double* U = u + (size_t)b * (T - 1) * M;
        for (int i = 0; i < N; ++i) X[i] = x1[(size_t)b * N + i];
        for (int t = 0; t < T - 1; ++t) ilqr_dyn(X + (t + 1) * N, X + t * N, U + t * M, h->probs[b].w + (size_t)t * NP);
    }
}

void oracle_solve(oracle_t* h, int nthreads) {
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    (void)nthreads;
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < h->B; ++b) solve_one(&h->probs[b], &h->opt);
}

/* receding-horizon step, same definition as ilqr_mpc_step in include/ilqr_cuda.h */
void oracle_mpc_step(oracle_t* h, int nthreads, double* applied_u, double* x_next) {
    const int T = h->T;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    (void)nthreads;
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < h->B; ++b) {
        prob_t* s = &h->probs[b];
        double xn[N], u0[DIM(M)];
        for (int a = 0; a < M; ++a) u0[a] = s->ub[a];
        ilqr_dyn(xn, s->xb, u0, s->w);
        if (applied_u) for (int a = 0; a < M; ++a) applied_u[(size_t)b * M + a] = u0[a];
        if (x_next) for (int i = 0; i < N; ++i) x_next[(size_t)b * N + i] = xn[i];
        for (int t = 0; t < T - 2; ++t)
            for (int a = 0; a < M; ++a) s->ub[t * M + a] = s->ub[(t + 1) * M + a];
        for (int i = 0; i < N; ++i) s->xb[i] = xn[i];
        for (int t = 0; t < T - 1; ++t) ilqr_dyn(s->xb + (t + 1) * N, s->xb + t * N, s->ub + t * M, s->w + (size_t)t * NP);
        solve_one(s, &h->opt);
    }
}

void oracle_get_trajectory(oracle_t* h, double* x, double* u, int current) {
    const size_t lx = (size_t)h->T * N, lu = (size_t)(h->T - 1) * M;
    for (int b = 0; b < h->B; ++b) {
        if (x) memcpy(x + b * lx, current ? h->probs[b].x : h->probs[b].xb, sizeof(double) * lx);
        if (u) memcpy(u + b * lu, current ? h->probs[b].u : h->probs[b].ub, sizeof(double) * lu);
    }
}

void oracle_get_stats(oracle_t* h, int32_t* iterations, uint8_t* status, double* objective, double* max_violation,
                      double* step_size, uint32_t* flags) {
    for (int b = 0; b < h->B; ++b) {
        const prob_t* s = &h->probs[b];
        if (iterations) iterations[b] = s->iterations;
        if (status) status[b] = (uint8_t)s->status;
        if (objective) objective[b] = s->objective;
        if (max_violation) max_violation[b] = s->max_violation;
        if (step_size) step_size[b] = s->step_size;
        if (flags) flags[b] = s->flags;
    }
}

void oracle_get_history(oracle_t* h, int cap, double* cost, double* gnorm, double* viol, double* alpha, int32_t* outer,
                        uint8_t* status) {
    for (int b = 0; b < h->B; ++b) {
        const prob_t* s = &h->probs[b];
        for (int r = 0; r < cap; ++r) {
            const int ok = r < s->cap;
            if (cost) cost[(size_t)b * cap + r] = ok ? s->h_cost[r] : 0.0;
            if (gnorm) gnorm[(size_t)b * cap + r] = ok ? s->h_gnorm[r] : 0.0;
            if (viol) viol[(size_t)b * cap + r] = ok ? s->h_viol[r] : 0.0;
            if (alpha) alpha[(size_t)b * cap + r] = ok ? s->h_alpha[r] : 0.0;
            if (outer) outer[(size_t)b * cap + r] = ok ? s->h_outer[r] : 0;
            if (status) status[(size_t)b * cap + r] = ok ? s->h_status[r] : 0;
        }
    }
}

void oracle_get_duals(oracle_t* h, double* dual, double* penalty, double* violations, int32_t* active) {
    const size_t rows = (size_t)(h->T - 1) * CS + CT;
    for (int b = 0; b < h->B; ++b) {
        const prob_t* s = &h->probs[b];
        for (size_t i = 0; i < rows; ++i) {
            if (dual) dual[b * rows + i] = s->lam[i];
            if (penalty) penalty[b * rows + i] = s->rho[i];
            if (violations) violations[b * rows + i] = s->c[i];
            if (active) active[b * rows + i] = s->a[i];
        }
    }
}

void oracle_get_policy(oracle_t* h, double* K, double* k) {
    const size_t lK = (size_t)(h->T - 1) * M * N, lk = (size_t)(h->T - 1) * M;
    for (int b = 0; b < h->B; ++b) {
        if (K) memcpy(K + b * lK, h->probs[b].K, sizeof(double) * lK);
        if (k) memcpy(k + b * lk, h->probs[b].k, sizeof(double) * lk);
    }
}

/* single calls of the generated model functions, for codegen tests */
void oracle_eval_dyn(const double* x, const double* u, const double* w, double* y, double* fx, double* fu) {
    ilqr_dyn(y, x, u, w);
    ilqr_dyn_jac(fx, fu, x, u, w);
}
void oracle_eval_cost(int terminal, const double* x, const double* u, const double* w, double* g, double* gx, double* gu,
                      double* gxx, double* guu, double* gux) {
    if (terminal) { ilqr_cost_T(g, x, u, w); ilqr_cost_T_grad(gx, gxx, x, u, w); }
    else { ilqr_cost_s(g, x, u, w); ilqr_cost_s_grad(gx, gu, gxx, guu, gux, x, u, w); }
}
void oracle_eval_con(int terminal, const double* x, const double* u, const double* w, double* c, double* cx, double* cu) {
    (void)x; (void)u; (void)w; (void)c; (void)cx; (void)cu;
    if (terminal) {
#if ILQR_CT > 0
        ilqr_con_T(c, x, u, w); ilqr_con_T_jac(cx, x, u, w);
#endif
    } else {
#if ILQR_CS > 0
        ilqr_con_s(c, x, u, w); ilqr_con_s_jac(cx, cu, x, u, w);
#endif
    }
}

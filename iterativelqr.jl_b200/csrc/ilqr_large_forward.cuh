/*
 * ilqr_large_forward.cuh -- forward-pass pieces for models whose per-step matrices do not fit a thread's
 * registers (BASELINE config 4: n = 64, m = 16).  Same names and the same arithmetic contract as the
 * register-resident versions in ilqr_kernels.cuh; vectors live in local memory, matrices are streamed
 * from HBM inside the dot-product loops (structure-of-arrays: every load is coalesced across the warp's
 * 32 problems), and the generated model functions are called out of line.
 */
#pragma once

/* rollout! + cost!(mode=:current) for one line-search trial: src/rollout.jl:19-29, src/data/methods.jl:13-30 */
__device__ __noinline__ void rollout_eval(const Params& P, const TrialOut& o, int b, double alpha, double& J_out,
                                          double& viol_out, double* /*ring_lane*/) {
    const Dev& d = P.d;
    const size_t Bp = P.Bp;
    const int T = P.T;
    double x[N], u[d1(M)], xn[N], wv[d1(NP)];
    double Jc = 0.0, Jal = 0.0, mv = 0.0;
    for (int i = 0; i < N; ++i) x[i] = d.xb[(size_t)i * Bp + b];
    for (int t = 0; t < T - 1; ++t) {
        const double* Kt = d.K + (size_t)t * M * N * Bp + b;
        const double* xbt = d.xb + (size_t)t * N * Bp + b;
        {
            /* K x and K xbar, traversed column by column: the M loads of a column are independent and in flight
             * together, every output's chain still runs over j ascending (the contract's order) */
            double acc1[d1(M)], acc2[d1(M)];
#pragma unroll 2
            for (int j = 0; j < N; ++j) {
                const double xj = x[j], xbj = xbt[(size_t)j * Bp];
                double kc[d1(M)];
#pragma unroll
                for (int a = 0; a < M; ++a) kc[a] = Kt[((size_t)a + (size_t)j * M) * Bp];
#pragma unroll
                for (int a = 0; a < M; ++a) {
                    acc1[a] = (j == 0) ? kc[a] * xj : ilqr_fma(kc[a], xj, acc1[a]);
                    acc2[a] = (j == 0) ? kc[a] * xbj : ilqr_fma(kc[a], xbj, acc2[a]);
                }
            }
#pragma unroll
            for (int a = 0; a < M; ++a) {
                double v = d.k[((size_t)t * M + a) * Bp + b] * alpha;     /* src/rollout.jl:24-25 */
                v = v + d.ub[((size_t)t * M + a) * Bp + b];               /* :26 */
                v = v + acc1[a];                                          /* :27 */
                v = v - acc2[a];                                          /* :28 */
                u[a] = v;
            }
        }
        if (o.x) { /* a speculative trial only reports its cost (forward_finish) */
            for (int i = 0; i < N; ++i) o.x[((size_t)t * N + i) * Bp + b] = x[i];
            for (int a = 0; a < M; ++a) o.u[((size_t)t * M + a) * Bp + b] = u[a];
        }
        for (int i = 0; i < NP; ++i) wv[i] = d.w[((size_t)t * NP + i) * Bp + b];
        double g;
        ilqr_cost_s(&g, x, u, wv);
        Jc += g;
        if (CS > 0) {
            double c[d1(CS)], lam[d1(CS)], rho[d1(CS)];
            uint8_t a[d1(CS)];
#if ILQR_CS > 0
            ilqr_con_s(c, x, u, wv);
#endif
            ld_rows<CS>(lam, d.lam, (size_t)t * CS, (int)Bp, b);
            ld_rows<CS>(rho, d.rho, (size_t)t * CS, (int)Bp, b);
            al_stage_cost<CS, false>(c, lam, rho, a, Jal);
            for (int i = 0; i < CS; ++i) viol_update(mv, c[i], ilqr_ineq_s(i));
            if (o.c) {
                st_rows<CS>(c, o.c, (size_t)t * CS, (int)Bp, b);
                for (int i = 0; i < CS; ++i) o.a[((size_t)t * CS + i) * Bp + b] = a[i];
            }
        }
        ilqr_dyn(xn, x, u, wv);                                       /* :29 */
        for (int i = 0; i < N; ++i) x[i] = xn[i];
    }
    {
        const int t = T - 1;
        for (int i = 0; i < NP; ++i) wv[i] = d.w[((size_t)t * NP + i) * Bp + b];
        if (o.x) for (int i = 0; i < N; ++i) o.x[((size_t)t * N + i) * Bp + b] = x[i];
        double g;
        ilqr_cost_T(&g, x, u, wv);
        Jc += g;
        if (CT > 0) {
            double c[d1(CT)], lam[d1(CT)], rho[d1(CT)];
            uint8_t a[d1(CT)];
#if ILQR_CT > 0
            ilqr_con_T(c, x, u, wv);
#endif
            ld_rows<CT>(lam, d.lam, (size_t)t * CS, (int)Bp, b);
            ld_rows<CT>(rho, d.rho, (size_t)t * CS, (int)Bp, b);
            al_stage_cost<CT, true>(c, lam, rho, a, Jal);
            for (int i = 0; i < CT; ++i) viol_update(mv, c[i], ilqr_ineq_T(i));
            if (o.c) {
                st_rows<CT>(c, o.c, (size_t)t * CS, (int)Bp, b);
                for (int i = 0; i < CT; ++i) o.a[((size_t)t * CS + i) * Bp + b] = a[i];
            }
        }
    }
    J_out = CONSTRAINED ? (Jc + Jal) : Jc;
    viol_out = mv;
}

/* trajectory_sensitivities + gradient' * trajectory: src/data/methods.jl:42-54, src/forward_pass.jl:19-20 */
constexpr int DG_SMEM_BYTES = 0, PR_WARP_DOUBLES = 0, FWD_SMEM_BYTES = 0; /* no rings: the wide-model version streams straight from HBM */
__device__ __noinline__ double delta_grad_product(const Params& P, int b, double* /*ring*/, int /*lane*/) {
    const Dev& d = P.d;
    const size_t Bp = P.Bp;
    const int T = P.T;
    double zx[N], zy[N], zu[d1(M)];
    double sx = 0.0, su = 0.0;
    for (int i = 0; i < N; ++i) zx[i] = 0.0;
    for (int t = 0; t < T - 1; ++t) {
        const double* Kt = d.K + (size_t)t * M * N * Bp + b;
        const double* fx = d.fx + (size_t)t * N * N * Bp + b;
        const double* fu = d.fu + (size_t)t * N * M * Bp + b;
        {
            double acc[d1(M)];
#pragma unroll 2
            for (int j = 0; j < N; ++j) {
                const double zj = zx[j];
                double kc[d1(M)];
#pragma unroll
                for (int a = 0; a < M; ++a) kc[a] = Kt[((size_t)a + (size_t)j * M) * Bp];
#pragma unroll
                for (int a = 0; a < M; ++a) acc[a] = (j == 0) ? kc[a] * zj : ilqr_fma(kc[a], zj, acc[a]);
            }
#pragma unroll
            for (int a = 0; a < M; ++a) zu[a] = d.k[((size_t)t * M + a) * Bp + b] + acc[a];           /* :49-50 */
        }
        /* zy = fu zu + fx zx in blocks of DG_ROWS outputs; within a block the column's loads are independent */
        constexpr int DG_ROWS = 16;
        for (int i0 = 0; i0 < N; i0 += DG_ROWS) {
            double av[DG_ROWS], ax[DG_ROWS];
            for (int a = 0; a < M; ++a) {
                const double za = zu[a];
#pragma unroll
                for (int ii = 0; ii < DG_ROWS; ++ii)
                    if (i0 + ii < N) {
                        const double f = fu[((size_t)(i0 + ii) + (size_t)a * N) * Bp];
                        av[ii] = (a == 0) ? f * za : ilqr_fma(f, za, av[ii]);                         /* :51 */
                    }
            }
#pragma unroll 2
            for (int j = 0; j < N; ++j) {
                const double zj = zx[j];
#pragma unroll
                for (int ii = 0; ii < DG_ROWS; ++ii)
                    if (i0 + ii < N) {
                        const double f = fx[((size_t)(i0 + ii) + (size_t)j * N) * Bp];
                        ax[ii] = (j == 0) ? f * zj : ilqr_fma(f, zj, ax[ii]);
                    }
            }
#pragma unroll
            for (int ii = 0; ii < DG_ROWS; ++ii)
                if (i0 + ii < N) zy[i0 + ii] = av[ii] + ax[ii];                                       /* :52 */
        }
        for (int i = 0; i < N; ++i) sx = ilqr_fma(d.Lx[((size_t)t * N + i) * Bp + b], zx[i], sx);
        for (int a = 0; a < M; ++a) su = ilqr_fma(d.Lu[((size_t)t * M + a) * Bp + b], zu[a], su);
        for (int i = 0; i < N; ++i) zx[i] = zy[i];
    }
    return sx + su;
}

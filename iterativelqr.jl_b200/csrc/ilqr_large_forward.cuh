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
        const double* Kt = K_block(d, T, b, t);
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
                for (int a = 0; a < M; ++a) kc[a] = Kt[a + j * M];
#pragma unroll
                for (int a = 0; a < M; ++a) {
                    acc1[a] = (j == 0) ? kc[a] * xj : ilqr_fma(kc[a], xj, acc1[a]);
                    acc2[a] = (j == 0) ? kc[a] * xbj : ilqr_fma(kc[a], xbj, acc2[a]);
                }
            }
#pragma unroll
            for (int a = 0; a < M; ++a) {
                double v = k_block(d, T, b, t)[a] * alpha;                /* src/rollout.jl:24-25 */
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
        const double* Kt = K_block(d, T, b, t);
        const double* fx = jac_block(d, T, b, t);                           /* staged block: fx(k, i) = fx[k LDF + i] */
        const double* fu = fx + JAC_FU;                                     /* fu(k, a) = fu[k LDU + a] */
        {
            double acc[d1(M)];
#pragma unroll 2
            for (int j = 0; j < N; ++j) {
                const double zj = zx[j];
                double kc[d1(M)];
#pragma unroll
                for (int a = 0; a < M; ++a) kc[a] = Kt[a + j * M];
#pragma unroll
                for (int a = 0; a < M; ++a) acc[a] = (j == 0) ? kc[a] * zj : ilqr_fma(kc[a], zj, acc[a]);
            }
#pragma unroll
            for (int a = 0; a < M; ++a) zu[a] = k_block(d, T, b, t)[a] + acc[a];                       /* :49-50 */
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
                        const double f = fu[(i0 + ii) * LDU + a];
                        av[ii] = (a == 0) ? f * za : ilqr_fma(f, za, av[ii]);                         /* :51 */
                    }
            }
#pragma unroll 2
            for (int j = 0; j < N; ++j) {
                const double zj = zx[j];
#pragma unroll
                for (int ii = 0; ii < DG_ROWS; ++ii)
                    if (i0 + ii < N) {
                        const double f = fx[(i0 + ii) * LDF + j];
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

/* ==================================================================================================================
 * k_forward_wp ("warp per problem"): the forward half of a tick for wide unconstrained models whose dynamics come as a
 * dense table block (codegen emits ilqr_dyn_part: a slice of the rows per caller).  The thread-per-problem rollout above
 * walks a 64-state step as ONE in-order stream (~10 k dependent operations, vectors in local memory): 71-126 ms per
 * launch at batch 1024, more than the Riccati kernel.  Here a CTA owns one problem: warps 0 / 1 roll out the two step
 * sizes of the round, warp 2 carries the expected-decrease sweep, lane 0 of warp 3 does the between-solves bookkeeping.
 * Inside a warp the OUTPUTS of every matrix-vector product are spread over the lanes -- each output is still one
 * ascending fma chain (the contract), so the bits are those of rollout_eval / delta_grad_product -- and the vectors live
 * in shared memory.  Scalar pieces (stage cost, the Armijo sums) run on lane 0.
 * ================================================================================================================== */
#if defined(ILQR_HAVE_ILQR_DYN_PART) && (ILQR_CS == 0) && (ILQR_CT == 0)
#define ILQR_FWD_WP 1
/* Everything a step reads that does not depend on the rollout itself -- K_t, k_t, the nominal x_t, u_t, the parameters w_t
 * (and Lx_t, Lu_t for the expected-decrease sweep) -- is fetched ONE STEP AHEAD by asynchronous copies into a double buffer
 * of the warp (the gain as 16-byte copies out of its problem-major block, the structure-of-arrays vectors as 8-byte ones):
 * a step's critical path is its arithmetic, not a chain of global-memory round trips. */
constexpr int WP_KSZ = (M * N + 1) & ~1; /* doubles per gain buffer */
constexpr bool WP_K16 = (M * N) % 2 == 0; /* blocks start on 16-byte boundaries then */
struct WpTrial { double x[N], xn[N], u[d1(M)], xb[2][N], ub[2][d1(M)], kf[2][d1(M)], w[2][d1(NP)]; };
struct WpDg { double zx[N], zy[N], zu[d1(M)], Lx[2][N], Lu[2][d1(M)], kf[2][d1(M)]; };
/* dynamic shared memory: two gain buffers per working warp, then the expected-decrease warp's copy of the step's staged Jacobians */
constexpr size_t WP_DYN_SMEM = ((size_t)(FWD_TRIAL_WARPS + 1) * 2 * WP_KSZ + (size_t)JAC_BLOCK) * sizeof(double);

__device__ __forceinline__ void wp_cp8(double* smem_dst, const double* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void wp_cp16(double* smem_dst, const double* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void wp_fetch_gain(const Params& P, int b, int t, double* Kbuf, int lane) {
    const double* src = K_block(P.d, P.T, b, t);
    if (WP_K16) { for (int c = lane; c < (M * N) / 2; c += 32) wp_cp16(Kbuf + 2 * c, src + 2 * c); }
    else        { for (int c = lane; c < M * N; c += 32) wp_cp8(Kbuf + c, src + c); }
}

/* rollout! + cost!(mode=:current), one trial, one warp (src/rollout.jl:19-29, src/costs.jl:48-55) */
__device__ __forceinline__ double rollout_wp(const Params& P, const TrialOut& o, int b, double alpha, WpTrial& s, double* Kbuf, int lane) {
    const Dev& d = P.d;
    const size_t Bp = P.Bp;
    const int T = P.T;
    double Jc = 0.0; /* lane 0's */
    auto fetch = [&](int t, int q) { /* step t's inputs into buffer q (t = T-1: the terminal parameters only) */
        for (int i = lane; i < NP; i += 32) wp_cp8(&s.w[q][i], &d.w[((size_t)t * NP + i) * Bp + b]);
        if (t < T - 1) {
            for (int i = lane; i < N; i += 32) wp_cp8(&s.xb[q][i], &d.xb[((size_t)t * N + i) * Bp + b]);
            for (int a = lane; a < M; a += 32) {
                wp_cp8(&s.ub[q][a], &d.ub[((size_t)t * M + a) * Bp + b]);
                wp_cp8(&s.kf[q][a], k_block(d, T, b, t) + a);
            }
            wp_fetch_gain(P, b, t, Kbuf + q * WP_KSZ, lane);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    fetch(0, 0);
    for (int i = lane; i < N; i += 32) s.x[i] = d.xb[(size_t)i * Bp + b];
    for (int t = 0; t < T - 1; ++t) {
        const int q = t & 1;
        fetch(t + 1, q ^ 1);
        asm volatile("cp.async.wait_group 1;" ::: "memory"); /* this lane's copies for step t; the next step's may still fly */
        __syncwarp();
        const double* Kt = Kbuf + q * WP_KSZ;
        for (int a = lane; a < M; a += 32) {                                /* one action component per lane */
            double acc1 = 0.0, acc2 = 0.0;
#pragma unroll 8
            for (int j = 0; j < N; ++j) {
                const double kc = Kt[a + j * M];
                acc1 = (j == 0) ? kc * s.x[j] : ilqr_fma(kc, s.x[j], acc1);
                acc2 = (j == 0) ? kc * s.xb[q][j] : ilqr_fma(kc, s.xb[q][j], acc2);
            }
            double v = s.kf[q][a] * alpha;                                  /* src/rollout.jl:24-25 */
            v = v + s.ub[q][a];                                             /* :26 */
            v = v + acc1;                                                   /* :27 */
            v = v - acc2;                                                   /* :28 */
            s.u[a] = v;
        }
        __syncwarp();
        if (o.x) { /* a speculative trial only reports its cost */
            for (int i = lane; i < N; i += 32) o.x[((size_t)t * N + i) * Bp + b] = s.x[i];
            for (int a = lane; a < M; a += 32) o.u[((size_t)t * M + a) * Bp + b] = s.u[a];
        }
        if (lane == 0) {
            double g;
            ilqr_cost_s(&g, s.x, s.u, s.w[q]);
            Jc += g;
        }
        ilqr_dyn_part(s.xn, s.x, s.u, s.w[q], lane, 32);                    /* :29, rows lane, lane + 32, ... */
        __syncwarp();
        for (int i = lane; i < N; i += 32) s.x[i] = s.xn[i];
        __syncwarp();
    }
    {
        const int t = T - 1, q = t & 1;
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        if (o.x) for (int i = lane; i < N; i += 32) o.x[((size_t)t * N + i) * Bp + b] = s.x[i];
        __syncwarp();
        if (lane == 0) {
            double g;
            ilqr_cost_T(&g, s.x, s.u, s.w[q]);
            Jc += g;
        }
    }
    return Jc;
}

/* trajectory_sensitivities + gradient' * trajectory, one warp (src/data/methods.jl:42-54, src/forward_pass.jl:19-20) */
__device__ __forceinline__ double dgp_wp(const Params& P, int b, WpDg& s, double* Kbuf, double* Jbuf, uint64_t* jbar, int lane) {
    const Dev& d = P.d;
    const size_t Bp = P.Bp;
    const int T = P.T;
    double sx = 0.0, su = 0.0; /* lane 0's */
    auto fetch = [&](int t, int q) {
        if (t < T - 1) {
            for (int i = lane; i < N; i += 32) wp_cp8(&s.Lx[q][i], &d.Lx[((size_t)t * N + i) * Bp + b]);
            for (int a = lane; a < M; a += 32) {
                wp_cp8(&s.Lu[q][a], &d.Lu[((size_t)t * M + a) * Bp + b]);
                wp_cp8(&s.kf[q][a], k_block(d, T, b, t) + a);
            }
            wp_fetch_gain(P, b, t, Kbuf + q * WP_KSZ, lane);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    /* the step's staged Jacobian block (fx | fu rows, 48 KB for n = 64) comes as bulk copies into ONE buffer: issued as soon as
     * the previous step's rows have been consumed, it lands under that step's closing sums and the next step's K zx product.
     * (Read straight from global memory the two 80-term chains of a lane stalled on every cache line: the sweep was the
     * critical warp of the CTA, 43 % of the kernel's stall samples.) */
    auto jac_issue = [&](int t) {
        if (lane == 0) {
            constexpr unsigned BYTES = (unsigned)JAC_BLOCK * 8u, CHUNK = 16384u;
            const char* src = (const char*)jac_block(d, T, b, t);
            mbar_arrive_expect_tx(jbar, BYTES);
            for (unsigned o = 0; o < BYTES; o += CHUNK)
                bulk_copy_g2s((double*)((char*)Jbuf + o), (const double*)(src + o), BYTES - o < CHUNK ? BYTES - o : CHUNK, jbar);
        }
    };
    if (lane == 0) {
        mbar_init(jbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    fetch(0, 0);
    jac_issue(0);
    for (int i = lane; i < N; i += 32) s.zx[i] = 0.0;
    for (int t = 0; t < T - 1; ++t) {
        const int q = t & 1;
        fetch(t + 1, q ^ 1);
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncwarp();
        const double* Kt = Kbuf + q * WP_KSZ;
        for (int a = lane; a < M; a += 32) {
            double acc = 0.0;
#pragma unroll 8
            for (int j = 0; j < N; ++j) {
                const double kc = Kt[a + j * M];
                acc = (j == 0) ? kc * s.zx[j] : ilqr_fma(kc, s.zx[j], acc);
            }
            s.zu[a] = s.kf[q][a] + acc;                                     /* :49-50 */
        }
        __syncwarp();
        mbar_wait(jbar, (unsigned)t & 1u); /* the step's Jacobian block has landed */
        for (int i = lane; i < N; i += 32) {                                /* one next-state component per lane */
            const double* fx = Jbuf + (size_t)i * LDF;                      /* row i of the staged fx */
            const double* fu = Jbuf + JAC_FU + (size_t)i * LDU;
            double av = 0.0, ax = 0.0;
#pragma unroll 8
            for (int a = 0; a < M; ++a) {
                const double f = fu[a];
                av = (a == 0) ? f * s.zu[a] : ilqr_fma(f, s.zu[a], av);     /* :51 */
            }
#pragma unroll 8
            for (int j = 0; j < N; ++j) {
                const double f = fx[j];
                ax = (j == 0) ? f * s.zx[j] : ilqr_fma(f, s.zx[j], ax);
            }
            s.zy[i] = av + ax;                                              /* :52 */
        }
        __syncwarp(); /* every lane is done with the block */
        if (t + 1 < T - 1) jac_issue(t + 1);
        if (lane == 0) {
            for (int i = 0; i < N; ++i) sx = ilqr_fma(s.Lx[q][i], s.zx[i], sx);
            for (int a = 0; a < M; ++a) su = ilqr_fma(s.Lu[q][a], s.zu[a], su);
        }
        __syncwarp();
        for (int i = lane; i < N; i += 32) s.zx[i] = s.zy[i];
        __syncwarp();
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    return sx + su;
}

__global__ void __launch_bounds__(32 * (FWD_TRIAL_WARPS + 2)) k_forward_wp(const __grid_constant__ Params P) {
    extern __shared__ __align__(16) double wp_gain[]; /* per warp: two gain buffers */
    __shared__ WpTrial tr[FWD_TRIAL_WARPS];
    __shared__ WpDg dg;
    __shared__ double sJ[FWD_TRIAL_WARPS], sDgp;
    __shared__ uint64_t s_jbar;
    const Dev& d = P.d;
    const int lane = threadIdx.x, wid = threadIdx.y;
    constexpr int NWc = FWD_TRIAL_WARPS, NT = 32 * (FWD_TRIAL_WARPS + 2);
    const int n_alpha = P.n_alpha;
    const int b = blockIdx.x; /* one CTA per problem */
    const size_t Bp = P.Bp;
    int phase = d.phase[b];
    const bool start_now = P.mode == MODE_STREAM && d.pending[b] == 1 + (P.tick & 7); /* see k_forward */
    if (start_now) phase = PH_START;
    const bool iter = phase == PH_ITER;
    if (b == 0 && wid == 0 && lane == 0) {
        d.active[(P.tick + 4) & 7] = 0;
        if (P.mode == MODE_STREAM) d.done_count[(P.tick + 2) & 3] = 0;
    }
    const int base = iter ? d.ls_base[b] : 0;
    const bool open_ls = iter && base < n_alpha;
    if (wid < NWc) {
        const int c_mine = base + wid;
        if (open_ls && c_mine < n_alpha) {
            TrialOut o;
            if (wid == 0) { o.x = d.xc; o.u = d.uc; o.c = d.c; o.a = d.act; }
            else { o.x = nullptr; o.u = nullptr; o.c = nullptr; o.a = nullptr; }
            const double J = rollout_wp(P, o, b, pow2neg(c_mine), tr[wid], wp_gain + (size_t)wid * 2 * WP_KSZ, lane);
            if (lane == 0) sJ[wid] = J;
        }
    } else if (wid == NWc) {
        if (iter) {
            double v;
            if (base == 0) {
                v = (P.o.line_search == ILQR_LINE_SEARCH_ARMIJO) ? dgp_wp(P, b, dg, wp_gain + (size_t)NWc * 2 * WP_KSZ, wp_gain + (size_t)(NWc + 1) * 2 * WP_KSZ, &s_jbar, lane) : 0.0;
                if (lane == 0) d.dgp[b] = v;
            } else {
                v = d.dgp[b];
            }
            if (lane == 0) sDgp = v;
        }
    } else if (lane == 0) { /* between two inner solves / two receding-horizon steps */
        if (phase == PH_START) {
            if (start_now) { d.pending[b] = 0; d.refilling[b] = 0; d.phase[b] = PH_START; }
            start_bookkeeping(P, b);
        } else if (phase == PH_SHIFT) {
            const Job& J = *P.job;
            const size_t s = (size_t)d.mpc_step[b] * P.B + b;
            mpc_shift_slot(P, b, J.mpc_u ? J.mpc_u + s * M : nullptr, J.mpc_x ? J.mpc_x + s * N : nullptr);
        } else if (!iter) {
            d.kind[b] = KIND_NONE;
        }
    }
    __syncthreads();
    /* selection and finalisation as forward_finish (one problem per CTA: every thread sees the same values) */
    int win = -1, wwin = -1;
    bool accepted = false, nonfinite = false;
    double Jwin = 0.0;
    const double Jp = iter ? d.J[b] : 0.0;
    if (open_ls) {
        const double dgpv = sDgp;
        for (int w = 0; w < NWc && base + w < n_alpha; ++w) {
            const int c = base + w;
            const double Jc = sJ[w];
            if (!(Jc - Jc == 0.0)) nonfinite = true;
            win = c; wwin = w; Jwin = Jc;
            if (Jc <= Jp + (1.0e-4 * pow2neg(c)) * dgpv) { accepted = true; break; }
        }
    }
    const bool more_rounds = open_ls && !accepted && base + NWc < n_alpha;
    const bool redo = open_ls && !more_rounds && wwin > 0;
    const int tid = wid * 32 + lane;
    if (more_rounds || redo) {
        if (tid == 0) {
            d.ls_base[b] = redo ? base + wwin : base + NWc;
            if (nonfinite) d.flags[b] |= ILQR_FLAG_NONFINITE;
            d.kind[b] = KIND_NONE;
        }
        return;
    }
    if (iter && accepted) { /* update_nominal_trajectory! (src/data/methods.jl:32-39) */
        for (int r = tid; r < P.T * N; r += NT) d.xb[(size_t)r * Bp + b] = d.xc[(size_t)r * Bp + b];
        for (int r = tid; r < (P.T - 1) * M; r += NT) d.ub[(size_t)r * Bp + b] = d.uc[(size_t)r * Bp + b];
    }
    if (tid == 0 && iter) {
        if (n_alpha > 0) d.J[b] = Jwin;
        d.alpha[b] = accepted ? pow2neg(win) : pow2neg(n_alpha);
        d.status[b] = accepted ? 1 : 0;
        if (nonfinite) d.flags[b] |= ILQR_FLAG_NONFINITE;
        d.ls_base[b] = 0;
        d.kind[b] = KIND_ITER;
    }
}
#else
#define ILQR_FWD_WP 0
#endif

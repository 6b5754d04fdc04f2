/*
 * ilqr_large_backward.cuh -- gradients! and backward_pass! for models too wide for registers
 * (BASELINE config 4: n = 64, m = 16, T = 256).  Same arithmetic contract as the register-resident
 * kernels (every output element is one ascending-k fma chain), different decomposition:
 *
 *   k_linearize : one thread per (problem, time step); the generated Jacobian / Hessian functions are
 *                 called out of line into local memory and copied to the structure-of-arrays buffers
 *                 (coalesced across the warp's problems).
 *   k_backward  : ONE CTA PER PROBLEM.  The value function P (n x n), the step's fx and the products
 *                 fx'P, Qxx live in shared memory; the five dense contractions of
 *                 /root/reference/src/backward_pass.jl:52-84 are register-tiled FP64 FMA loops over shared
 *                 memory (4 x 4 outputs per thread, strided by 16 so that a half-warp reads consecutive words),
 *                 the m x m Cholesky and the triangular solves of :68-75
 *                 run on one warp / one thread per right-hand side.  FP64 on B200 has the same peak on the
 *                 vector pipe as on the tensor pipe and tcgen05 has no FP64 kind, so these are DFMA loops;
 *                 a chain order identical to the oracle's is what keeps the result bit-exact.
 */
#pragma once

constexpr int BK_STAGES = 0;
constexpr int BK_STAGE_BYTES = 0;
constexpr bool BK_FUSED = false;

/* -------------------------------------------------------------------------------------------- k_linearize */
/* The Jacobians go to the problem-major staged blocks (ilqr_kernels.cuh): the warp's 32 problems hand their values over
 * through a 32 x 32 shared-memory tile so that every store instruction writes one full 256-byte run of ONE problem's block.
 * Called by whole warps (lanes whose problem sits this tick out pass active = false and only help with the tile). */
__device__ __noinline__ void lin_dynamics(const Params& P, int b, int t, bool active, const double* x, const double* u, const double* wv,
                                          double (*tile)[33]) {
    const Dev& d = P.d;
    const int lane = threadIdx.x & 31;
    double jl[N * N + d1(N * M)];
    if (active) ilqr_dyn_jac(jl, jl + N * N, x, u, wv);                     /* src/dynamics.jl:41-50 */
    const unsigned live = __ballot_sync(0xffffffffu, active);
    if (live == 0u) return;
    const size_t pstride = (size_t)(P.T - 1) * JAC_BLOCK;                   /* from one problem's block of step t to the next problem's */
    double* const blk0 = d.fx + ((size_t)(b - lane) * (P.T - 1) + t) * JAC_BLOCK;
    for (int q0 = 0; q0 < JAC_BLOCK; q0 += 32) {
        if (active) {
#pragma unroll 8
            for (int j = 0; j < 32; ++j) {
                const int q = q0 + j;
                double v = 0.0;                                             /* the padding is written too: a block is copied as a whole */
                if (q < JAC_FU) {
                    const int k = q / LDF, i = q - k * LDF;
                    if (i < N) v = jl[k + i * N];
                } else if (q < JAC_FU + N * LDU) {
                    const int r = q - JAC_FU, k = r / LDU, a = r - k * LDU;
                    if (a < M) v = jl[N * N + k + a * N];
                }
                tile[lane][j] = v;
            }
        }
        __syncwarp();
        if (q0 + lane < JAC_BLOCK) {
            for (int pl = 0; pl < 32; ++pl)
                if ((live >> pl) & 1u) blk0[(size_t)pl * pstride + q0 + lane] = tile[pl][lane];
        }
        __syncwarp();
    }
}

/* AL terms of src/gradients.jl:54-80 on top of (gx, gxx[, gu, guu, gux]) held in global memory rows */
template <int R, bool TERM>
__device__ __noinline__ void lin_al_terms(const Params& P, int b, int t, const double* cx, const double* cu, double* gx, double* gu) {
    const Dev& d = P.d;
    const size_t Bp = P.Bp;
    double dd[d1(R)], v[d1(R)];
    for (int i = 0; i < R; ++i) {
        const size_t idx = ((size_t)t * CS + i) * Bp + b;
        dd[i] = d.rho[idx] * (double)d.act[idx];                            /* :56-58 */
        v[i] = d.lam[idx] + dd[i] * d.c[idx];                               /* :59-62 */
    }
    for (int j = 0; j < N; ++j) {                                           /* :63 */
        double acc = cx[j * R] * v[0];
        for (int i = 1; i < R; ++i) acc = ilqr_fma(cx[i + j * R], v[i], acc);
        gx[j] = gx[j] + acc;
    }
    for (int l = 0; l < N; ++l)                                             /* :66-67 */
        for (int j = 0; j < N; ++j) {
            double acc = cx[j * R] * (dd[0] * cx[l * R]);
            for (int i = 1; i < R; ++i) acc = ilqr_fma(cx[i + j * R], dd[i] * cx[i + l * R], acc);
            const size_t g = ((size_t)t * N * N + j + (size_t)l * N) * Bp + b;
            d.gxx[g] = d.gxx[g] + acc;
        }
    if (!TERM) {
        for (int e = 0; e < M; ++e) {                                       /* :72 */
            double acc = cu[e * R] * v[0];
            for (int i = 1; i < R; ++i) acc = ilqr_fma(cu[i + e * R], v[i], acc);
            gu[e] = gu[e] + acc;
        }
        for (int e = 0; e < M; ++e)                                         /* :75-76 */
            for (int a = 0; a < M; ++a) {
                double acc = cu[a * R] * (dd[0] * cu[e * R]);
                for (int i = 1; i < R; ++i) acc = ilqr_fma(cu[i + a * R], dd[i] * cu[i + e * R], acc);
                const size_t g = ((size_t)t * M * M + a + (size_t)e * M) * Bp + b;
                d.guu[g] = d.guu[g] + acc;
            }
        for (int j = 0; j < N; ++j)                                         /* :79 */
            for (int a = 0; a < M; ++a) {
                double acc = cu[a * R] * (dd[0] * cx[j * R]);
                for (int i = 1; i < R; ++i) acc = ilqr_fma(cu[i + a * R], dd[i] * cx[i + j * R], acc);
                const size_t g = ((size_t)t * M * N + a + (size_t)j * M) * Bp + b;
                d.gux[g] = d.gux[g] + acc;
            }
    }
}

__device__ __noinline__ void lin_cost_stage(const Params& P, int b, int t, bool fresh, const double* x, const double* u, const double* wv) {
    const Dev& d = P.d;
    const size_t Bp = P.Bp;
    double gx[N], gu[d1(M)];
    if (HACC_L) {
        ilqr_cost_s_grad1(gx, gu, x, u, wv);                                /* the Hessians: one accumulator per problem (lin_hacc_advance) */
    } else {
        double hxx[N * N], huu[d1(M * M)], hux[d1(M * N)];
        ilqr_cost_s_grad(gx, gu, hxx, huu, hux, x, u, wv);                  /* src/costs.jl:57-84 */
        for (int r = 0; r < N * N; ++r) {                                   /* Q1: accumulate */
            const size_t g = ((size_t)t * N * N + r) * Bp + b;
            d.gxx[g] = (fresh ? 0.0 : d.gxx[g]) + hxx[r];
        }
        for (int r = 0; r < M * M; ++r) {
            const size_t g = ((size_t)t * M * M + r) * Bp + b;
            d.guu[g] = (fresh ? 0.0 : d.guu[g]) + huu[r];
        }
        for (int r = 0; r < M * N; ++r) {
            const size_t g = ((size_t)t * M * N + r) * Bp + b;
            d.gux[g] = (fresh ? 0.0 : d.gux[g]) + hux[r];
        }
    }
#if ILQR_CS > 0
    {
        double cx[CS * N], cu[CS * M];
        ilqr_con_s_jac(cx, cu, x, u, wv);                                   /* src/constraints.jl:75-87 */
        lin_al_terms<CS, false>(P, b, t, cx, cu, gx, gu);
    }
#endif
    for (int i = 0; i < N; ++i) d.gx[((size_t)t * N + i) * Bp + b] = gx[i];
    for (int a = 0; a < M; ++a) d.gu[((size_t)t * M + a) * Bp + b] = gu[a];
}

/* HACC_L: the stage Hessians this tick's gradients! call leaves behind, accumulator = (fresh ? 0 : accumulator) + constants
 * (src/costs.jl:74,79,80; the same additions every step's accumulator would receive) */
__device__ __noinline__ void lin_hacc_advance(const Params& P, int b, bool fresh) {
    const Dev& d = P.d;
    const size_t Bp = P.Bp;
    double zx[N], zu[d1(M)], zw[d1(NP)], gx[N], gu[d1(M)], hc[NH];
    for (int i = 0; i < N; ++i) zx[i] = 0.0;
    for (int i = 0; i < d1(M); ++i) zu[i] = 0.0;
    for (int i = 0; i < d1(NP); ++i) zw[i] = 0.0;
    ilqr_cost_s_grad(gx, gu, hc, hc + N * N, hc + N * N + M * M, zx, zu, zw); /* constants: the arguments do not enter */
    for (int r = 0; r < NH; ++r) {
        const size_t g = (size_t)r * Bp + b;
        d.hacc[g] = (fresh ? 0.0 : d.hacc[g]) + hc[r];
    }
}

__device__ __noinline__ void lin_cost_terminal(const Params& P, int b, bool fresh, const double* x, const double* u, const double* wv) {
    const Dev& d = P.d;
    const size_t Bp = P.Bp;
    const int t = P.T - 1;
    double gx[N];
    {
        double hxx[N * N];
        ilqr_cost_T_grad(gx, hxx, x, u, wv);
        for (int r = 0; r < N * N; ++r) {
            const size_t g = ((size_t)t * N * N + r) * Bp + b;
            d.gxx[g] = (fresh ? 0.0 : d.gxx[g]) + hxx[r];
        }
    }
#if ILQR_CT > 0
    {
        double cx[CT * N];
        ilqr_con_T_jac(cx, x, u, wv);
        lin_al_terms<CT, true>(P, b, t, cx, nullptr, gx, nullptr);
    }
#endif
    for (int i = 0; i < N; ++i) d.gx[((size_t)t * N + i) * Bp + b] = gx[i];
}

/* LIN_COOP: models whose generated Jacobians are table blocks come with ilqr_dyn_jac_part (an entry slice per caller, see
 * codegen.py): ONE CTA per (problem, time step) evaluates them together into shared memory and writes the staged block
 * with coalesced stores -- no 40 KB of per-thread local memory, no transposing tile.  k_linearize keeps the cost part. */
#if defined(ILQR_HAVE_ILQR_DYN_JAC_PART) && !defined(ILQR_NO_LIN_COOP)
constexpr bool LIN_COOP = !JAC_CONST;
#else
constexpr bool LIN_COOP = false;
#endif
/* The terminal cost derivatives (src/gradients.jl via cost_gradient!/cost_hessian! at t = T) and the per-problem Hessian
 * accumulator's advance are done by the Riccati CTA of the problem in its prologue (256 threads on the 4096 + 5376 row
 * updates) instead of by ONE thread per problem of k_linearize, whose 32 warps were the tail of that kernel. */
constexpr bool RL_PRO_FUSED = N >= M;                   /* (the prologue's scratch area holds the NH constants then) */
constexpr bool RL_PRO_TERM = RL_PRO_FUSED && CT == 0;   /* with terminal constraints the AL terms are added by k_linearize, as before */
constexpr bool RL_PRO_HACC = RL_PRO_FUSED && HACC_L;
constexpr int LC_THREADS = 256;
constexpr int LC_CTAS_PER_SM = 3; /* persistent CTAs: 3 x 43 KB of shared memory leave the L1 room for the generated tables (~90 KB for n = 64) */
constexpr int LC_LD = N + 1; /* odd column stride: the transposing read-out below is conflict-free */
__global__ void __launch_bounds__(LC_THREADS) k_linearize_jac(const __grid_constant__ Params P) {
#if defined(ILQR_HAVE_ILQR_DYN_JAC_PART) && !defined(ILQR_NO_LIN_COOP)
    constexpr int NIN = N + M + NP;                                  /* x | u | w of an item */
    constexpr int NV = (NIN + LC_THREADS - 1) / LC_THREADS;          /* ... words per thread */
    __shared__ double s_in[NIN + 1], s_t[ILQR_ILQR_DYN_JAC_PART_NT];
    __shared__ double s_fx[LC_LD * N], s_fu[LC_LD * d1(M)];
    const double *s_x = s_in, *s_u = s_in + N, *s_w = s_in + N + M;
    const Dev& d = P.d;
    const size_t Bp = P.Bp;
    const int tid = threadIdx.x;
    const size_t items = (size_t)(P.T - 1) * Bp;
    /* an item's inputs are one useful word per line (the batch-interleaved rows): each thread fetches its word(s) of the NEXT
     * item while the CTA works on the current one (L2 only: the L1 is wanted for the generated tables) */
    auto fetch = [&](size_t item, double (&v)[NV]) {
        if (item >= items) return;
        const int b = (int)(item % Bp), t = (int)(item / Bp);
#pragma unroll
        for (int r = 0; r < NV; ++r) {
            const int i = tid + r * LC_THREADS;
            if (i < N) v[r] = __ldcg(&d.xb[((size_t)t * N + i) * Bp + b]);
            else if (i < N + M) v[r] = __ldcg(&d.ub[((size_t)t * M + (i - N)) * Bp + b]);
            else if (i < NIN) v[r] = __ldcg(&d.w[((size_t)t * NP + (i - N - M)) * Bp + b]);
        }
    };
    double vin[NV];
    fetch(blockIdx.x, vin);
    for (size_t item = blockIdx.x; item < items; item += gridDim.x) { /* neighbouring CTAs: neighbouring problems of one step (shared sectors) */
        const int b = (int)(item % Bp), t = (int)(item / Bp);
        const int kind = d.kind[b];
        const bool active = !(kind == KIND_NONE || (kind == KIND_ITER && P.o.line_search == ILQR_LINE_SEARCH_NONE)); /* src/solve.jl:27 */
        if (active) {
#pragma unroll
            for (int r = 0; r < NV; ++r)
                if (tid + r * LC_THREADS < NIN) s_in[tid + r * LC_THREADS] = vin[r];
        }
        fetch(item + gridDim.x, vin);
        if (!active) continue;
        __syncthreads();
        ilqr_dyn_jac_part_t(s_t, s_x, s_u, s_w, tid, LC_THREADS);
        __syncthreads();
        ilqr_dyn_jac_part(s_fx, s_fu, s_t, tid, LC_THREADS, LC_LD - N);          /* src/dynamics.jl:41-50 */
        __syncthreads();
        double* const blk = jac_block(d, P.T, b, t);
        for (int q = tid; q < JAC_BLOCK; q += LC_THREADS) {
            double v = 0.0;                                                     /* the padding is written too: a block is copied as a whole */
            if (q < JAC_FU) {
                const int k = q / LDF, i = q - k * LDF;
                if (i < N) v = s_fx[k + i * LC_LD];
            } else if (q < JAC_FU + N * LDU) {
                const int r = q - JAC_FU, k = r / LDU, a = r - k * LDU;
                if (a < M) v = s_fu[k + a * LC_LD];
            }
            __stcs(&blk[q], v); /* streaming: read next by another SM's Riccati CTA */
        }
        __syncthreads(); /* the next item overwrites the staging arrays */
    }
#endif
}

/* JAC_CONST: the one staged block of the workspace (any arguments: the generated Jacobians do not read them) */
__global__ void k_jac_const_init(const __grid_constant__ Params P) {
    double x[N], u[d1(M)], wv[d1(NP)], jl[N * N + d1(N * M)];
    for (int i = 0; i < N; ++i) x[i] = 0.0;
    for (int i = 0; i < d1(M); ++i) u[i] = 0.0;
    for (int i = 0; i < d1(NP); ++i) wv[i] = 0.0;
    ilqr_dyn_jac(jl, jl + N * N, x, u, wv);
    double* blk = P.d.fx;
    for (int q = 0; q < JAC_BLOCK; ++q) blk[q] = 0.0;
    for (int k = 0; k < N; ++k) {
        for (int i = 0; i < N; ++i) blk[k * LDF + i] = jl[k + i * N];
        for (int a = 0; a < M; ++a) blk[JAC_FU + k * LDU + a] = jl[N * N + k + a * N];
    }
}

__global__ void __launch_bounds__(64) k_linearize(const __grid_constant__ Params P) {
    __shared__ double s_tile[2][32][33];
    const Dev& d = P.d;
    const size_t Bp = P.Bp;
    const int T = P.T;
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int b = (int)(g % Bp);
    const int t = (int)(g / Bp);            /* the same for a warp's 32 lanes (Bp is a multiple of 32) */
    if (t >= T) return;
    const int kind = d.kind[b];
    const bool active = kind != KIND_NONE && !(kind == KIND_ITER && P.o.line_search == ILQR_LINE_SEARCH_NONE); /* src/solve.jl:27 */
    if (t == T - 1 && !active) return;
    const bool fresh = kind == KIND_PRELOOP;
    double x[N], u[d1(M)], wv[d1(NP)];
    if (active) {
        for (int i = 0; i < N; ++i) x[i] = d.xb[((size_t)t * N + i) * Bp + b];
        for (int i = 0; i < NP; ++i) wv[i] = d.w[((size_t)t * NP + i) * Bp + b];
    }
    if (t < T - 1) {
        if (active)
            for (int a = 0; a < M; ++a) u[a] = d.ub[((size_t)t * M + a) * Bp + b];
        if (!JAC_CONST && !LIN_COOP) lin_dynamics(P, b, t, active, x, u, wv, s_tile[threadIdx.x >> 5]);
        if (active) lin_cost_stage(P, b, t, fresh, x, u, wv);
    } else {
        for (int a = 0; a < M; ++a) u[a] = 0.0;
        if (!RL_PRO_TERM) lin_cost_terminal(P, b, fresh, x, u, wv);
        if (HACC_L && !RL_PRO_HACC) lin_hacc_advance(P, b, fresh);
    }
}

/* -------------------------------------------------------------------------------------------- k_backward */
constexpr int RL_THREADS = 256;
static_assert(N + M <= RL_THREADS && N <= 16 * 16 && M <= 16 * 16, "k_backward (large) thread mapping");
constexpr int TI = (N + 15) / 16;  /* outputs per thread along one index of an N x N result on the 16 x 16 thread grid */
constexpr int TA = (M + 15) / 16;  /* ... for the M-row results */
/* DMMA variant of the dense contractions (phases B, C, G): FP64 tensor-core tiles, mma.sync.aligned.m8n8k4.f64 (tcgen05 has
 * no FP64 kind).  One warp owns a block of 8 x 8 output tiles; an A / B fragment is ONE double per lane per k-step of 4, so
 * a k-step costs TM + TN shared-memory loads for TM x TN DMMAs of 256 fused multiply-adds each -- against 8 loads per 16
 * DFMAs of the register-tiled loops, whose shared-memory traffic was the busiest unit of the kernel (58 % in the r1 capture).
 * Every output is still accumulated over k ascending starting from an exact zero, i.e. the contract's fma chain (tests
 * compare bit for bit).  Needs n, m multiples of 8; -DILQR_RL_DMMA=0 (build variant "nodmma") keeps the DFMA loops. */
/* (RL_DMMA, LDF, LDU: ilqr_kernels.cuh, with the layout of the staged Jacobian blocks) */
static_assert(!RL_DMMA || RL_THREADS == 256, "DMMA warp -> tile map");
constexpr int LDP = RL_DMMA ? N + 4 : N + 1;
constexpr int LDK = RL_DMMA ? M + 4 : M + 1;
constexpr bool RL_HSMEM = HACC_L && !RL_DMMA; /* shared-memory copy of the constant Hessians (the DMMA variant spends that room on padding and reads them from L2) */

struct RlSmem { /* carve-up of the dynamic shared memory, all doubles */
    double *P, *p, *fxT, *fuT, *xxhT, *uxhT, *Qxx, *Qux, *Quu, *uu, *K, *uxt, *Qx, *Qu, *kk, *rinv, *gxs, *gus;
    double *Hxx, *Hux, *Huu; /* HACC_L: the problem's stage Hessians (the same for every step), loaded once */
};
constexpr size_t RL_SMEM_DOUBLES = (size_t)LDP * N /*P*/ + N /*p*/ + 1 /*alignment*/ + (size_t)JAC_BLOCK /*fxT | fuT*/ + (size_t)LDF * N /*xxhT*/ +
                                   (size_t)LDU * N /*uxhT*/ + (size_t)N * N /*Qxx*/ + (size_t)LDK * N /*Qux*/ + (size_t)M * M /*Quu*/ +
                                   (size_t)M * M /*uu*/ + (size_t)LDK * N /*K*/ + (size_t)LDK * N /*uxt*/ + N + M + M + M + 2 * N + 2 * M +
                                   (RL_HSMEM ? (size_t)N * N + (size_t)LDK * N + (size_t)M * M : 0);
static_assert(RL_SMEM_DOUBLES * 8 + 64 <= 227 * 1024 || !ILQR_LARGE, "wide-model Riccati kernel: shared memory budget");
constexpr size_t RL_SMEM_BYTES = RL_SMEM_DOUBLES * 8 + 64;

__device__ __forceinline__ void rl_cp8(double* smem_dst, const double* gsrc) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gsrc) : "memory");
}
/* named barriers (producer: arrive, consumers: sync; PTX bar.arrive / bar.sync order the participants' memory accesses) */
__device__ __forceinline__ void rl_bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void rl_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
/* RL_OVERLAP: the Cholesky factorisation (warp 0) and the triangular solves (warps 1..) of a step run UNDER the Qxx contraction
 * of the other warps instead of after it -- they only need Quu and Qux, which are computed first.  -DILQR_RL_OVERLAP=0 keeps the
 * phases in sequence. */
#ifndef ILQR_RL_OVERLAP
#define ILQR_RL_OVERLAP 1
#endif
#ifndef ILQR_RL_QVEC_DMMA   /* Qx, Qu as one more column tile of the fx'P, fu'P products instead of 64-term chains on three warps: measured 0.2 ms SLOWER per launch (the chains run under the other warps' tiles), off */
#define ILQR_RL_QVEC_DMMA 0
#endif
#ifndef ILQR_RL_FG_DMMA     /* uxt = Quu K on the tensor cores, the three sums of p spread over neighbouring lanes */
#define ILQR_RL_FG_DMMA 1
#endif
#ifndef ILQR_RL_FWARP       /* measured (bench c4, ms per launch): 22.1 with, 21.9 without -- off, see RL_FWARP below */
#define ILQR_RL_FWARP 0
#endif
#ifndef ILQR_RL_CHOL_RIGHT  /* register Cholesky right-looking (1) or left-looking (0): same bits; measured 0.5 ms per launch SLOWER, off */
#define ILQR_RL_CHOL_RIGHT 0
#endif
constexpr int RL_E_THREADS = ((N + 1 + 31) / 32) * 32; /* whole warps on the right-hand sides: bar.sync counts are multiples of 32 */
constexpr bool RL_OVERLAP = (ILQR_RL_OVERLAP != 0) && RL_DMMA && M <= 32 && 32 + RL_E_THREADS <= RL_THREADS;
/* RL_FWARP: warp 0 does NOTHING but the factorisation.  Its tiles of every product go to warp 4, the other warp of its SM
 * sub-partition -- a DMMA.8x8x4 occupies the sub-partition's FP64 tensor pipe for 16 cycles, ONE warp saturates it, so two
 * blocks on warp 4 take the pipe exactly as long as one block on each of two warps does in the other sub-partitions.  The
 * step is reordered so that Quu exists as early as possible -- fu'P, then Quu, and warp 0 starts -- and the big products
 * (fx'P, Qux, Qxx) of warps 1-7 run under the factorisation; warps 1-3 then do the triangular solves.  Hand-offs by named
 * barriers: 2 = the seven tile warps among themselves, 3 = Quu ready (warps 1-4 -> warp 0), 1 = factor ready (warp 0 ->
 * warps 1-3).  -DILQR_RL_FWARP=0: the RL_OVERLAP schedule (factorisation under Qxx only).
 * Measured: no gain (22.1 against 21.9 ms per launch).  ONE warp does not keep the tensor pipe full here -- its fragment
 * loads from shared memory are exposed between k-steps, which a second warp on the sub-partition hides -- so warp 4 with two
 * blocks becomes the straggler of every product.  Kept as a build variant ("rl_fwarp"); the schedule that would work gives
 * the factorisation and the solves warps of their OWN (12 warps, setmaxnreg) and leaves the eight tile warps symmetric. */
constexpr bool RL_FWARP = (ILQR_RL_FWARP != 0) && RL_OVERLAP && N == 64 && RL_THREADS == 256;
/* RL_HELPERS: the factorisation and the solves get warps of their OWN -- warp 8 factorises, warps 9-11 do the 65 triangular
 * solves, the eight tile warps keep one block of every product each -- and the step is ordered for the serial spine
 * fu'P -> Quu -> Cholesky -> solves: the tile warps compute fu'P and Quu first, hand Quu over (named barrier 3) and go on
 * with fx'P, Qux (barrier 5 tells the solve warps) and Qxx UNDER the factorisation; barrier 1 passes the factor from warp 8
 * to the solve warps; the only CTA-wide barrier of a step is the one that publishes K.  The tile warps synchronise among
 * themselves on barrier 4.  -DILQR_RL_HELPERS=0: 8 warps, RL_OVERLAP schedule.
 * Measured: 4 % SLOWER than the 8-warp RL_OVERLAP schedule, like RL_FWARP within a few percent of it.  The factorisation's
 * dependent DFMA / MUFU chain and the tiles' DMMAs share the sub-partition's FP64 pipe (a DMMA.8x8x4 holds it for 16 cycles),
 * so running the spine UNDER more tile work stretches the spine instead of hiding it; every ordering tried lands at
 * 22-24 ms per launch (0.39-0.43 of the DMMA peak).  Kept as build variants ("rl_helpers", "rl_fwarp"), parity-tested. */
#ifndef ILQR_RL_HELPERS      /* measured (bench c4): 23.7 ms per launch with, 22.7 without -- off */
#define ILQR_RL_HELPERS 0
#endif
constexpr bool RL_HELPERS = (ILQR_RL_HELPERS != 0) && RL_OVERLAP && !RL_FWARP && N == 64 && RL_THREADS == 256;
constexpr int RL_CTA_THREADS = RL_HELPERS ? RL_THREADS + 32 + RL_E_THREADS : RL_THREADS;
__device__ __forceinline__ void rl_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void rl_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void rl_wait_but_one() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

/* C[i + 16 ii][l + 16 ll] (TR x TC per thread, strided by 16) = sum_k A[k*lda + row] * B(k, col); the k loop is unrolled
 * by 4 so that the shared-memory loads of the next k run under the FMAs of the current one.  Every output is one
 * ascending-k fma chain starting from a0*b0 (the contract). */
template <int TR, int TC, bool B_KMAJOR>
__device__ __forceinline__ void rl_gemm(double (&acc)[TR][TC], const double* __restrict__ A, int lda, int nrow, int ti,
                                        const double* __restrict__ Bm, int ldb, int ncol, int tl, int K) {
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
        double av[TR], bv[TC];
#pragma unroll
        for (int ii = 0; ii < TR; ++ii) av[ii] = (ti + 16 * ii < nrow) ? A[k * lda + ti + 16 * ii] : 0.0;
#pragma unroll
        for (int ll = 0; ll < TC; ++ll) {
            const int c = tl + 16 * ll;
            bv[ll] = (c < ncol) ? (B_KMAJOR ? Bm[k * ldb + c] : Bm[k + c * ldb]) : 0.0;
        }
#pragma unroll
        for (int ii = 0; ii < TR; ++ii)
#pragma unroll
            for (int ll = 0; ll < TC; ++ll) acc[ii][ll] = (k == 0) ? av[ii] * bv[ll] : ilqr_fma(av[ii], bv[ll], acc[ii][ll]);
    }
}

__device__ __forceinline__ void rl_dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
/* One warp: acc[tm][tn] = the 8 x 8 tile at rows i0 + 8 tm, columns j0 + 8 tn of C = A B, C(i, j) = sum_k A(i, k) B(k, j), k =
 * 0 .. K-1 ascending.  Operand storage: *_KMAJOR: X(idx, k) = Xp[k * ld + idx]; otherwise X(idx, k) = Xp[k + idx * ld].
 * Fragment layout of mma.m8n8k4.f64 (PTX ISA): lane = 4 g + q; A holds A(g, q), B holds B(q, g), C holds C(g, 2 q + {0, 1}). */
template <int TM, int TN, bool A_KMAJOR, bool B_KMAJOR>
__device__ __forceinline__ void rl_dmma(double (&acc)[TM][TN][2], const double* __restrict__ Ap, int lda, int i0,
                                        const double* __restrict__ Bp, int ldb, int j0, int K, int lane) {
    const int g = lane >> 2, q = lane & 3;
#pragma unroll
    for (int tm = 0; tm < TM; ++tm)
#pragma unroll
        for (int tn = 0; tn < TN; ++tn) { acc[tm][tn][0] = 0.0; acc[tm][tn][1] = 0.0; }
#pragma unroll 4
    for (int k0 = 0; k0 < K; k0 += 4) {
        double a[TM], bv[TN];
#pragma unroll
        for (int tm = 0; tm < TM; ++tm) {
            const int i = i0 + 8 * tm + g;
            a[tm] = A_KMAJOR ? Ap[(k0 + q) * lda + i] : Ap[(k0 + q) + i * lda];
        }
#pragma unroll
        for (int tn = 0; tn < TN; ++tn) {
            const int j = j0 + 8 * tn + g;
            bv[tn] = B_KMAJOR ? Bp[(k0 + q) * ldb + j] : Bp[(k0 + q) + j * ldb];
        }
#pragma unroll
        for (int tm = 0; tm < TM; ++tm)
#pragma unroll
            for (int tn = 0; tn < TN; ++tn) rl_dmma884(acc[tm][tn][0], acc[tm][tn][1], a[tm], bv[tn]);
    }
}

/* One right-hand side of K = -Quu \ Qux, k = -Quu \ Qu (src/backward_pass.jl:70-75) given the factor U (upper, column-major
 * M x M in shared memory) and 1 / diag(U): forward substitution with U', back substitution with U, every sum one ascending-k
 * fma chain.  Out of line ON PURPOSE: for small m the loops are fully unrolled and the right-hand side lives in registers
 * (the rolled version walked it in local memory: 6.4 k cycles per step for 272 fused multiply-adds), and as a separate
 * function that unrolled code does not raise the register pressure inside the caller's DMMA loops. */
__device__ __noinline__ void rl_trisolve(const double* __restrict__ U, const double* __restrict__ rinv, const double* __restrict__ rhs,
                                         double* __restrict__ out_s, double* __restrict__ out_g, size_t gstride) {
    constexpr int EU = M <= 24 ? M : 4;
    double bv[d1(M)];
#pragma unroll EU
    for (int a = 0; a < M; ++a) bv[a] = rhs[a];
#pragma unroll EU
    for (int i = 0; i < M; ++i) {
        double sum = bv[i];
#pragma unroll EU
        for (int k = 0; k < i; ++k) sum = ilqr_fma(-U[k + i * M], bv[k], sum);
        bv[i] = sum * rinv[i];
        asm volatile("" ::: "memory"); /* one row's loads in flight at a time (all 136 at once do not fit the registers) */
    }
#pragma unroll EU
    for (int i = M - 1; i >= 0; --i) {
        double sum = bv[i];
#pragma unroll EU
        for (int k = i + 1; k < M; ++k) sum = ilqr_fma(-U[i + k * M], bv[k], sum);
        bv[i] = sum * rinv[i];
        asm volatile("" ::: "memory");
    }
#pragma unroll EU
    for (int a = 0; a < M; ++a) {
        out_s[a] = -bv[a];
        out_g[(size_t)a * gstride] = -bv[a];
    }
}

/* Layouts in shared memory (chosen so that a half-warp reads consecutive words):
 *   P[k + l*LDP]      column-major like the reference (padded leading dimension)
 *   fxT[k*N + i]  = fx[k, i]      ("k-major": the i's of one k are contiguous)
 *   fuT[k*M + a]  = fu[k, a]
 *   xxhT[l*N + i] = (fx' P)[i, l]
 *   uxhT[l*M + a] = (fu' P)[a, l]
 *   Qxx[i + j*N], Quu[a + e*M] column-major; Qux, K, uxt column-major with padded leading dimension LDK
 * Pipeline across time steps: fx/fu of step t-1 are copied (cp.async) into fxT/fuT as soon as phase C of step t has
 * released them; gxx/gux/guu of step t-1 are copied into the Qxx/Qux/Quu buffers after phase G of step t, where
 * phase C of step t-1 adds the contraction onto them in place. */
__global__ void __launch_bounds__(RL_CTA_THREADS, 1) k_backward(const __grid_constant__ Params P) {
    extern __shared__ __align__(16) double rl_smem[];
    const Dev& d = P.d;
    const size_t Bp = P.Bp;
    const int T = P.T;
    const int b = blockIdx.x;            /* one CTA per problem */
    const int tid = threadIdx.x;
    const int kind = d.kind[b];
    const bool skip_ls_none = (kind == KIND_ITER && P.o.line_search == ILQR_LINE_SEARCH_NONE);
    __shared__ double s_gn[RL_THREADS / 32];
    __shared__ int s_cholfail;
    __shared__ uint64_t s_jbar; /* completion of the step's Jacobian block (bulk copy) */
    RlSmem s;
    {
        double* q = rl_smem;
        s.P = q; q += LDP * N; s.p = q; q += N; q += (q - rl_smem) & 1; /* 16-byte aligned: the bulk copy's destination */
        s.fxT = q; s.fuT = q + JAC_FU; q += JAC_BLOCK; s.xxhT = q; q += LDF * N;
        s.uxhT = q; q += LDU * N; s.Qxx = q; q += N * N; s.Qux = q; q += LDK * N; s.Quu = q; q += M * M; s.uu = q; q += M * M;
        s.K = q; q += LDK * N; s.uxt = q; q += LDK * N; s.Qx = q; q += N; s.Qu = q; q += M; s.kk = q; q += M; s.rinv = q; q += M;
        s.gxs = q; q += 2 * N; s.gus = q; q += 2 * M; /* double-buffered by step parity */
        s.Hxx = q; q += RL_HSMEM ? N * N : 0; s.Hux = q; q += RL_HSMEM ? LDK * N : 0; s.Huu = q;
    }
    double gn = 0.0;
    const int ln = tid & 31;
#define RL_SYNC() do { if (RL_HELPERS) rl_bar_sync(4, RL_THREADS); else __syncthreads(); } while (0)
    /* Cholesky of Quu in registers (m <= 32), one warp: lane j holds column j of the factor; U(k, jj) travels by shuffle.  Same
     * operations in the same order as the shared-memory version (and the oracle): every sum is one ascending-k fma chain. */
    auto chol_regs = [&]() {
                    constexpr int MC = M <= 32 ? M : 1;
                    const int lj = ln < MC ? ln : MC - 1;
                    double col[MC];
#pragma unroll
                    for (int k = 0; k < MC; ++k) col[k] = ln < MC ? s.uu[k + lj * M] : 1.0; /* (lanes beyond m: idle columns, no reads) */
                    bool ok = true;
#if ILQR_RL_CHOL_RIGHT
                    /* right-looking: as soon as row k of the factor exists, every remaining entry (r, i) takes its k-th term --
                     * the same terms in the same ascending order as the left-looking sums, but off the critical path: a pivot
                     * waits for ONE fused multiply-add after the previous row instead of for a chain of k */
#pragma unroll
                    for (int k = 0; k < MC; ++k) {
                        const double akk = __shfl_sync(0xffffffffu, col[k], k);
                        if (ok && !(akk > 0.0)) { /* stop (Q3): the failed pivot keeps its value, what lies behind it stays the ORIGINAL matrix */
                            ok = false;
#pragma unroll
                            for (int r = k; r < MC; ++r)
                                if (r > k || ln != k) col[r] = s.uu[r + lj * M];
                            if (ln == k) s_cholfail = 1;
                        }
                        if (ok) {
                            const double ukk = sqrt(akk);
                            const double rr = 1.0 / ukk;
                            if (ln == k) col[k] = ukk;
                            else if (ln > k) col[k] = col[k] * rr;
#pragma unroll
                            for (int r = k + 1; r < MC; ++r) {
                                const double ukr = __shfl_sync(0xffffffffu, col[k], r);
                                if (ln >= r) col[r] = ilqr_fma(-ukr, col[k], col[r]);
                            }
                        }
                    }
#else
#pragma unroll
                    for (int jj = 0; jj < MC; ++jj) {
                        double ajj = col[jj];
#pragma unroll
                        for (int k = 0; k < jj; ++k) ajj = ilqr_fma(-col[k], col[k], ajj);
                        const double ajj_b = __shfl_sync(0xffffffffu, ajj, jj);
                        if (ok && !(ajj_b > 0.0)) {
                            ok = false;
                            if (ln == jj) { col[jj] = ajj_b; s_cholfail = 1; }
                        }
                        double sum = col[jj]; /* lane i: A(jj, i) */
#pragma unroll
                        for (int k = 0; k < jj; ++k) {
                            const double ukj = __shfl_sync(0xffffffffu, col[k], jj);
                            sum = ilqr_fma(-ukj, col[k], sum);
                        }
                        if (ok) {
                            const double ujj = sqrt(ajj_b);
                            const double r = 1.0 / ujj;
                            if (ln == jj) col[jj] = ujj;
                            else if (ln > jj) col[jj] = sum * r;
                        }
                    }
#endif
                    if (ln < MC) {
#pragma unroll
                        for (int k = 0; k < MC; ++k)
                            if (k <= ln) s.uu[k + ln * M] = col[k];
                        s.rinv[ln] = 1.0 / col[ln];
                    }
    };
    if (RL_HELPERS && tid >= RL_THREADS) {
        if (kind != KIND_NONE && !skip_ls_none) {
            const int hw = (tid - RL_THREADS) >> 5; /* 0: the factorisation warp, 1 ..: the solve warps */
            for (int t = T - 2; t >= 0; --t) {
                if (hw == 0) {
                    rl_bar_sync(3, 160);                 /* Quu (and its copy uu) from tile warps 0-3 */
                    chol_regs();                         /* src/backward_pass.jl:68-69 */
                    rl_bar_arrive(1, 32 + RL_E_THREADS); /* the factor is in shared memory */
                } else {
                    rl_bar_sync(5, RL_THREADS + RL_E_THREADS); /* Qux (and Qu) from the tile warps */
                    rl_bar_sync(1, 32 + RL_E_THREADS);
                    const int col = tid - RL_THREADS - 32;     /* :70-75, one right-hand side per thread */
                    if (col < N) rl_trisolve(s.uu, s.rinv, s.Qux + col * LDK, s.K + col * LDK, K_block(d, T, b, t) + (size_t)col * M, 1);
                    else if (col == N) rl_trisolve(s.uu, s.rinv, s.Qu, s.kk, k_block(d, T, b, t), 1);
                    __syncwarp(); /* converged again (the out-of-line call returns per thread) before the aligned barrier */
                }
                __syncthreads();
            }
        }
        return;
    }
    if (kind != KIND_NONE && !skip_ls_none) {
        if (tid == 0) {
            s_cholfail = 0;
            mbar_init(&s_jbar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        RL_SYNC();
        if (RL_PRO_TERM || RL_PRO_HACC) { /* this tick's gradients! work of the terminal step (see RL_PRO_FUSED) */
            const bool fresh = kind == KIND_PRELOOP;
            double* const scr = s.fxT; /* fxT | fuT | xxhT: at least NH doubles, idle until the first Jacobian block arrives */
            static_assert(!RL_PRO_FUSED || (size_t)JAC_BLOCK + (size_t)LDF * N >= (size_t)NH, "prologue scratch");
            if (RL_PRO_HACC && tid == 0) { /* the constant stage Hessians (the arguments do not enter) */
                double zx[N], zu[d1(M)], zw[d1(NP)], gx_[N], gu_[d1(M)];
                for (int i = 0; i < N; ++i) zx[i] = 0.0;
                for (int i = 0; i < d1(M); ++i) zu[i] = 0.0;
                for (int i = 0; i < d1(NP); ++i) zw[i] = 0.0;
                ilqr_cost_s_grad(gx_, gu_, scr, scr + N * N, scr + N * N + M * M, zx, zu, zw);
            }
            if (RL_PRO_TERM && tid == 32) { /* terminal cost gradient and Hessian at the nominal x_T (src/costs.jl:57-84) */
                double xT[N], u0[d1(M)], wT[d1(NP)];
                for (int i = 0; i < N; ++i) xT[i] = d.xb[((size_t)(T - 1) * N + i) * Bp + b];
                for (int i = 0; i < d1(M); ++i) u0[i] = 0.0;
                for (int i = 0; i < NP; ++i) wT[i] = d.w[((size_t)(T - 1) * NP + i) * Bp + b];
                ilqr_cost_T_grad(s.p, s.Qxx, xT, u0, wT);
            }
            RL_SYNC();
            if (RL_PRO_HACC)
                for (int r = tid; r < NH; r += RL_THREADS) {
                    const size_t g = (size_t)r * Bp + b;
                    d.hacc[g] = (fresh ? 0.0 : d.hacc[g]) + scr[r];                                /* Q1: accumulate */
                }
            if (RL_PRO_TERM) {
                for (int r = tid; r < N * N; r += RL_THREADS) {
                    const size_t g = ((size_t)(T - 1) * N * N + r) * Bp + b;
                    const double v = (fresh ? 0.0 : d.gxx[g]) + s.Qxx[r];                          /* Q1: accumulate */
                    d.gxx[g] = v;
                    s.P[(r % N) + (r / N) * LDP] = v;                                              /* src/backward_pass.jl:39 */
                }
                for (int r = tid; r < N; r += RL_THREADS) d.gx[((size_t)(T - 1) * N + r) * Bp + b] = s.p[r]; /* s.p: :40 */
            }
            RL_SYNC();
        }
        /* terminal value function: src/backward_pass.jl:39-40 */
        if (!RL_PRO_TERM) {
            for (int r = tid; r < N * N; r += RL_THREADS) s.P[(r % N) + (r / N) * LDP] = d.gxx[((size_t)(T - 1) * N * N + r) * Bp + b];
            for (int r = tid; r < N; r += RL_THREADS) s.p[r] = d.gx[((size_t)(T - 1) * N + r) * Bp + b];
        }
        /* asynchronous copies of one step's inputs */
        /* threads first, first + 1, ... < RL_THREADS copy (everybody commits a group: the wait counts are per thread) */
        auto issue_jac = [&](int t, int first) {
            const int nth = RL_THREADS - first, me = tid - first;
            if (me == 0) { /* the staged block IS the image of fxT | fuT: one thread, a few bulk copies */
                constexpr unsigned BYTES = (unsigned)JAC_BLOCK * 8u, CHUNK = 16384u;
                const char* src = (const char*)jac_block(d, T, b, t);
                mbar_arrive_expect_tx(&s_jbar, BYTES);
                for (unsigned o = 0; o < BYTES; o += CHUNK)
                    bulk_copy_g2s((double*)((char*)s.fxT + o), (const double*)(src + o), BYTES - o < CHUNK ? BYTES - o : CHUNK, &s_jbar);
            }
            if (me >= 0) {
                const int par = t & 1;
                for (int r = me; r < N; r += nth) rl_cp8(&s.gxs[par * N + r], &d.gx[((size_t)t * N + r) * Bp + b]);
                for (int r = me; r < M; r += nth) rl_cp8(&s.gus[par * M + r], &d.gu[((size_t)t * M + r) * Bp + b]);
            }
            rl_commit();
        };
        auto issue_hess = [&](int t) {
            for (int r = tid; r < N * N; r += RL_THREADS) rl_cp8(&s.Qxx[r], &d.gxx[((size_t)t * N * N + r) * Bp + b]);
            for (int r = tid; r < M * N; r += RL_THREADS) rl_cp8(&s.Qux[(r % M) + (r / M) * LDK], &d.gux[((size_t)t * M * N + r) * Bp + b]);
            for (int r = tid; r < M * M; r += RL_THREADS) rl_cp8(&s.Quu[r], &d.guu[((size_t)t * M * M + r) * Bp + b]);
            rl_commit();
        };
        issue_jac(T - 2, 0);
        if (HACC_L) { /* advanced by k_linearize's terminal thread of this tick */
            if (RL_HSMEM) {
                for (int r = tid; r < N * N; r += RL_THREADS) s.Hxx[r] = d.hacc[(size_t)r * Bp + b];
                for (int r = tid; r < M * M; r += RL_THREADS) s.Huu[r] = d.hacc[((size_t)N * N + r) * Bp + b];
                for (int r = tid; r < M * N; r += RL_THREADS) s.Hux[(r % M) + (r / M) * LDK] = d.hacc[((size_t)N * N + M * M + r) * Bp + b];
            }
            rl_commit(); /* keeps the group count of the two-group wait scheme */
        } else {
            issue_hess(T - 2);
        }
        const int ti = tid & 15, tl = tid >> 4; /* 16 x 16 thread grid */
        const int wq = tid >> 5, fg = ln >> 2, fq = ln & 3; /* DMMA: warp, lane, fragment row group / column pair */
        const int wi0 = 16 * (wq >> 1), wj0 = 32 * (wq & 1);               /* ... and the warp's 16 x 32 block of an n x n result */
        /* DMMA + HACC_L: the constant Hessians at this thread's fragment positions, read once (20 doubles) */
        constexpr int MTH = RL_DMMA ? M / 8 : 1;
        double hxx_r[2][4][2], hux_r[MTH][2], huu_r[2];
        if (RL_DMMA && HACC_L) {
            const double* Hg = d.hacc + b;
#pragma unroll
            for (int tm = 0; tm < 2; ++tm)
#pragma unroll
                for (int tn = 0; tn < 4; ++tn)
#pragma unroll
                    for (int e = 0; e < 2; ++e) hxx_r[tm][tn][e] = Hg[(size_t)((wi0 + 8 * tm + fg) + (wj0 + 8 * tn + 2 * fq + e) * N) * Bp];
#pragma unroll
            for (int tm = 0; tm < MTH; ++tm)
#pragma unroll
                for (int e = 0; e < 2; ++e) hux_r[tm][e] = Hg[(size_t)(N * N + M * M + (8 * tm + fg) + (8 * wq + 2 * fq + e) * M) * Bp];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int qt = RL_FWARP ? wq - 1 : wq; /* the Quu tile of this warp */
                huu_r[e] = (qt >= 0 && qt < MTH * MTH) ? Hg[(size_t)(N * N + (8 * (qt / MTH) + fg) + (8 * (qt % MTH) + 2 * fq + e) * M) * Bp] : 0.0;
            }
        }
#ifdef ILQR_RL_PHASE_TIMERS
        long long ph[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tprev = clock64();
#define RL_TICK(i) do { if (tid == 0) { const long long now_ = clock64(); ph[i] += now_ - tprev; tprev = now_; } } while (0)
#else
#define RL_TICK(i) do { } while (0)
#endif
        for (int t = T - 2; t >= 0; --t) {
            const double* gxs = s.gxs + (t & 1) * N;
            const double* gus = s.gus + (t & 1) * M;
            if ((RL_FWARP || RL_HELPERS) && !HACC_L) rl_wait_all(); /* (the reordered schedule reads the per-step Hessian blocks right after the first barrier) */
            else rl_wait_but_one(); /* this thread's gradient copies for step t have landed (the Hessian group may still fly) */
            mbar_wait(&s_jbar, (unsigned)(T - 2 - t) & 1u); /* ... and the step's Jacobian block */
            RL_SYNC();
            RL_TICK(0);
            if (RL_HELPERS) { /* the eight tile warps (the helper warps run their own loop, see above) */
                constexpr int MT = RL_DMMA ? M / 8 : 1;
                {   /* uxh = fu' P (:57): column tile wq */
                    double acc[MT][1][2];
                    rl_dmma<MT, 1, true, false>(acc, s.fuT, LDU, 0, s.P, LDP, 8 * wq, N, ln);
#pragma unroll
                    for (int tm = 0; tm < MT; ++tm)
#pragma unroll
                        for (int e = 0; e < 2; ++e) s.uxhT[(8 * wq + 2 * fq + e) * LDU + 8 * tm + fg] = acc[tm][0][e];
                }
                RL_SYNC();
                if (wq < MT * MT) { /* Quu = uxh fu + guu (:58-59), tile wq; warps 0-3 report to the factorisation warp */
                    double acc[1][1][2];
                    const int a0 = 8 * (wq / MT), e0 = 8 * (wq % MT);
                    rl_dmma<1, 1, true, true>(acc, s.uxhT, LDU, a0, s.fuT, LDU, e0, N, ln);
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int o = (a0 + fg) + (e0 + 2 * fq + e) * M;
                        const double q = acc[0][0][e] + (HACC_L ? huu_r[e] : s.Quu[o]);
                        s.Quu[o] = q;
                        s.uu[o] = q;                                                            /* :68 */
                    }
                }
                if (wq < 4) rl_bar_arrive(3, 160);
                {   /* xxh = fx' P (:52): the warp's 16 x 32 block */
                    double acc[2][4][2];
                    rl_dmma<2, 4, true, false>(acc, s.fxT, LDF, wi0, s.P, LDP, wj0, N, ln);
#pragma unroll
                    for (int tm = 0; tm < 2; ++tm)
#pragma unroll
                        for (int tn = 0; tn < 4; ++tn)
#pragma unroll
                            for (int e = 0; e < 2; ++e) s.xxhT[(wj0 + 8 * tn + 2 * fq + e) * LDF + wi0 + 8 * tm + fg] = acc[tm][tn][e];
                }
                if (wq == 4 || wq == 5) { /* Qx (:44-45) behind the blocks of warps 4, 5; Qu (:48-49) on warp 6 */
                    const int i = tid - 128;
                    double acc = s.fxT[i] * s.p[0];
                    for (int k = 1; k < N; ++k) acc = ilqr_fma(s.fxT[k * LDF + i], s.p[k], acc);
                    s.Qx[i] = acc + gxs[i];
                } else if (wq == 6 && ln < M) {
                    double acc = s.fuT[ln] * s.p[0];
                    for (int k = 1; k < N; ++k) acc = ilqr_fma(s.fuT[k * LDU + ln], s.p[k], acc);
                    s.Qu[ln] = acc + gus[ln];
                }
                RL_SYNC();
                {   /* Qux = uxh fx + gux (:63-64): column tile wq */
                    double acc[MT][1][2];
                    rl_dmma<MT, 1, true, true>(acc, s.uxhT, LDU, 0, s.fxT, LDF, 8 * wq, N, ln);
#pragma unroll
                    for (int tm = 0; tm < MT; ++tm)
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int a = 8 * tm + fg, j = 8 * wq + 2 * fq + e;
                            s.Qux[a + j * LDK] = acc[tm][0][e] + (HACC_L ? hux_r[tm][e] : s.Qux[a + j * LDK]);
                        }
                }
                rl_bar_arrive(5, RL_THREADS + RL_E_THREADS); /* Qux and Qu are in shared memory: the solve warps may read them */
                {   /* Qxx = xxh fx + gxx (:53-54): the warp's 16 x 32 block */
                    double acc[2][4][2];
                    rl_dmma<2, 4, true, true>(acc, s.xxhT, LDF, wi0, s.fxT, LDF, wj0, N, ln);
#pragma unroll
                    for (int tm = 0; tm < 2; ++tm)
#pragma unroll
                        for (int tn = 0; tn < 4; ++tn)
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                const int i = wi0 + 8 * tm + fg, j = wj0 + 8 * tn + 2 * fq + e;
                                s.Qxx[i + j * N] = acc[tm][tn][e] + (HACC_L ? hxx_r[tm][tn][e] : s.Qxx[i + j * N]);
                            }
                }
            } else {
            if (RL_FWARP) {
                constexpr int MT = RL_DMMA ? M / 8 : 1;
                const double* Hg = d.hacc + b;
                if (wq == 0) {
                    rl_bar_sync(3, 160); /* Quu (and its copy uu) written by warps 1-4 */
                    RL_TICK(1);
                    chol_regs();                                                                /* :68-69 */
                    rl_bar_arrive(1, 32 + RL_E_THREADS); /* the factor is in shared memory: releases the solves */
                    RL_TICK(4);
                } else {
                    const int nblk = wq == 4 ? 2 : 1; /* warp 4 also takes warp 0's block / tile of every product */
                    /* uxh = fu' P (:57): column tile per warp */
                    for (int r = 0; r < nblk; ++r) {
                        const int wb = r == 0 ? wq : 0;
                        double acc[MT][1][2];
                        rl_dmma<MT, 1, true, false>(acc, s.fuT, LDU, 0, s.P, LDP, 8 * wb, N, ln);
#pragma unroll
                        for (int tm = 0; tm < MT; ++tm)
#pragma unroll
                            for (int e = 0; e < 2; ++e) s.uxhT[(8 * wb + 2 * fq + e) * LDU + 8 * tm + fg] = acc[tm][0][e];
                    }
                    rl_bar_sync(2, 224);
                    /* Quu = uxh fu + guu (:58-59), tile wq - 1 on warps 1 .. MT^2; warps 1-4 report to warp 0 */
                    if (wq <= MT * MT) {
                        double acc[1][1][2];
                        const int qt = wq - 1, a0 = 8 * (qt / MT), e0 = 8 * (qt % MT);
                        rl_dmma<1, 1, true, true>(acc, s.uxhT, LDU, a0, s.fuT, LDU, e0, N, ln);
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int o = (a0 + fg) + (e0 + 2 * fq + e) * M;
                            const double q = acc[0][0][e] + (HACC_L ? huu_r[e] : s.Quu[o]);
                            s.Quu[o] = q;
                            s.uu[o] = q;                                                        /* :68 */
                        }
                    }
                    if (wq <= 4) rl_bar_arrive(3, 160);
                    /* xxh = fx' P (:52): 16 x 32 block per warp; Qx (:44-45) on warps 5, 6 and Qu (:48-49) on warp 7 behind their blocks */
                    for (int r = 0; r < nblk; ++r) {
                        const int wb = r == 0 ? wq : 0, bi0 = 16 * (wb >> 1), bj0 = 32 * (wb & 1);
                        double acc[2][4][2];
                        rl_dmma<2, 4, true, false>(acc, s.fxT, LDF, bi0, s.P, LDP, bj0, N, ln);
#pragma unroll
                        for (int tm = 0; tm < 2; ++tm)
#pragma unroll
                            for (int tn = 0; tn < 4; ++tn)
#pragma unroll
                                for (int e = 0; e < 2; ++e) s.xxhT[(bj0 + 8 * tn + 2 * fq + e) * LDF + bi0 + 8 * tm + fg] = acc[tm][tn][e];
                    }
                    if (wq == 5 || wq == 6) {
                        const int i = tid - 160;
                        double acc = s.fxT[i] * s.p[0];
                        for (int k = 1; k < N; ++k) acc = ilqr_fma(s.fxT[k * LDF + i], s.p[k], acc);
                        s.Qx[i] = acc + gxs[i];
                    } else if (wq == 7 && ln < M) {
                        double acc = s.fuT[ln] * s.p[0];
                        for (int k = 1; k < N; ++k) acc = ilqr_fma(s.fuT[k * LDU + ln], s.p[k], acc);
                        s.Qu[ln] = acc + gus[ln];
                    }
                    rl_bar_sync(2, 224);
                    /* Qux = uxh fx + gux (:63-64): column tile per warp */
                    for (int r = 0; r < nblk; ++r) {
                        const int wb = r == 0 ? wq : 0;
                        double acc[MT][1][2];
                        rl_dmma<MT, 1, true, true>(acc, s.uxhT, LDU, 0, s.fxT, LDF, 8 * wb, N, ln);
#pragma unroll
                        for (int tm = 0; tm < MT; ++tm)
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                const int a = 8 * tm + fg, j = 8 * wb + 2 * fq + e;
                                const double h = !HACC_L ? s.Qux[a + j * LDK] : r == 0 ? hux_r[tm][e] : Hg[(size_t)(N * N + M * M + a + j * M) * Bp];
                                s.Qux[a + j * LDK] = acc[tm][0][e] + h;
                            }
                    }
                    rl_bar_sync(2, 224);
                    /* Qxx = xxh fx + gxx (:53-54): 16 x 32 block per warp */
                    for (int r = 0; r < nblk; ++r) {
                        const int wb = r == 0 ? wq : 0, bi0 = 16 * (wb >> 1), bj0 = 32 * (wb & 1);
                        double hx[2][4][2];
                        if (HACC_L && r > 0) { /* the constants of the foreign block: from the accumulator (L2), in flight under the tiles */
#pragma unroll
                            for (int tm = 0; tm < 2; ++tm)
#pragma unroll
                                for (int tn = 0; tn < 4; ++tn)
#pragma unroll
                                    for (int e = 0; e < 2; ++e) hx[tm][tn][e] = Hg[(size_t)((bi0 + 8 * tm + fg) + (bj0 + 8 * tn + 2 * fq + e) * N) * Bp];
                        }
                        double acc[2][4][2];
                        rl_dmma<2, 4, true, true>(acc, s.xxhT, LDF, bi0, s.fxT, LDF, bj0, N, ln);
#pragma unroll
                        for (int tm = 0; tm < 2; ++tm)
#pragma unroll
                            for (int tn = 0; tn < 4; ++tn)
#pragma unroll
                                for (int e = 0; e < 2; ++e) {
                                    const int i = bi0 + 8 * tm + fg, j = bj0 + 8 * tn + 2 * fq + e;
                                    const double h = !HACC_L ? s.Qxx[i + j * N] : r == 0 ? hxx_r[tm][tn][e] : hx[tm][tn][e];
                                    s.Qxx[i + j * N] = acc[tm][tn][e] + h;
                                }
                    }
                    /* K = -Quu \ Qux, k = -Quu \ Qu (:70-75) on warps 1-3, one right-hand side per thread */
                    if (tid < 32 + RL_E_THREADS) {
                        rl_bar_sync(1, 32 + RL_E_THREADS);
                        const int col = tid - 32;
                        if (col < N) rl_trisolve(s.uu, s.rinv, s.Qux + col * LDK, s.K + col * LDK, K_block(d, T, b, t) + (size_t)col * M, 1);
                        else if (col == N) rl_trisolve(s.uu, s.rinv, s.Qu, s.kk, k_block(d, T, b, t), 1);
                        __syncwarp(); /* converged again (the out-of-line call returns per thread) before the aligned barrier */
                    }
                }
            } else {
            /* ---- B: xxh = fx' P (:52), uxh = fu' P (:57), Qx (:44-45), Qu (:48-49) */
            if (RL_DMMA) { /* warp wq: the 16 x 32 block (rows 16 (wq / 2), columns 32 (wq % 2)) of xxh; column tile wq of uxh */
                constexpr int MT = RL_DMMA ? M / 8 : 1;
                {
                    double acc[2][4][2];
                    rl_dmma<2, 4, true, false>(acc, s.fxT, LDF, wi0, s.P, LDP, wj0, N, ln);
#pragma unroll
                    for (int tm = 0; tm < 2; ++tm)
#pragma unroll
                        for (int tn = 0; tn < 4; ++tn)
#pragma unroll
                            for (int e = 0; e < 2; ++e) s.xxhT[(wj0 + 8 * tn + 2 * fq + e) * LDF + wi0 + 8 * tm + fg] = acc[tm][tn][e];
                }
                {
                    double acc[MT][1][2];
                    rl_dmma<MT, 1, true, false>(acc, s.fuT, LDU, 0, s.P, LDP, 8 * wq, N, ln);
#pragma unroll
                    for (int tm = 0; tm < MT; ++tm)
#pragma unroll
                        for (int e = 0; e < 2; ++e) s.uxhT[(8 * wq + 2 * fq + e) * LDU + 8 * tm + fg] = acc[tm][0][e];
                }
            } else {
                {
                    double acc[TI][TI];
                    rl_gemm<TI, TI, false>(acc, s.fxT, LDF, N, ti, s.P, LDP, N, tl, N);
#pragma unroll
                    for (int ii = 0; ii < TI; ++ii)
#pragma unroll
                        for (int ll = 0; ll < TI; ++ll)
                            if (ti + 16 * ii < N && tl + 16 * ll < N) s.xxhT[(tl + 16 * ll) * LDF + ti + 16 * ii] = acc[ii][ll];
                }
                {
                    double acc[TA][TI];
                    rl_gemm<TA, TI, false>(acc, s.fuT, LDU, M, ti, s.P, LDP, N, tl, N);
#pragma unroll
                    for (int aa = 0; aa < TA; ++aa)
#pragma unroll
                        for (int ll = 0; ll < TI; ++ll)
                            if (ti + 16 * aa < M && tl + 16 * ll < N) s.uxhT[(tl + 16 * ll) * LDU + ti + 16 * aa] = acc[aa][ll];
                }
            }
            if (RL_DMMA && ILQR_RL_QVEC_DMMA) {
                /* Qx = fx'p + gx (:44-45), Qu = fu'p + gu (:48-49): p sits right behind P's 64 columns in shared memory, i.e. it IS
                 * column 64 of P -- one more 8-column tile of the two products above (columns 65.. of it read whatever follows
                 * and are dropped; a column of a product depends on no other).  Row tile wq per warp instead of 64-long serial
                 * chains on warps 0, 1 and 7 while the others wait. */
                constexpr int MT = RL_DMMA ? M / 8 : 1;
                double acc[1][1][2];
                rl_dmma<1, 1, true, false>(acc, s.fxT, LDF, 8 * wq, s.P, LDP, N, N, ln);
                if (fq == 0) s.Qx[8 * wq + fg] = acc[0][0][0] + gxs[8 * wq + fg];
                if (wq < MT) {
                    rl_dmma<1, 1, true, false>(acc, s.fuT, LDU, 8 * wq, s.P, LDP, N, N, ln);
                    if (fq == 0) s.Qu[8 * wq + fg] = acc[0][0][0] + gus[8 * wq + fg];
                }
            } else if (tid < N) {
                double acc = s.fxT[tid] * s.p[0];
                for (int k = 1; k < N; ++k) acc = ilqr_fma(s.fxT[k * LDF + tid], s.p[k], acc);
                s.Qx[tid] = acc + gxs[tid];
            } else if (tid >= RL_THREADS - M) {
                const int a = tid - (RL_THREADS - M);
                double acc = s.fuT[a] * s.p[0];
                for (int k = 1; k < N; ++k) acc = ilqr_fma(s.fuT[k * LDU + a], s.p[k], acc);
                s.Qu[a] = acc + gus[a];
            }
            RL_TICK(1);
            rl_wait_all(); /* ... and the Hessian blocks that phase C adds onto */
            RL_SYNC();
            RL_TICK(2);
            /* ---- C: Qxx = xxh fx + gxx (:53-54), Quu = uxh fu + guu (:58-59), Qux = uxh fx + gux (:63-64);
             *         gxx, gux, guu are already sitting in the Qxx, Qux, Quu buffers */
            /* the two halves of phase C on the tensor cores; Quu and Qux first when the factorisation is overlapped with Qxx */
            auto c_qxx = [&]() {
                {
                    double acc[2][4][2];
                    rl_dmma<2, 4, true, true>(acc, s.xxhT, LDF, wi0, s.fxT, LDF, wj0, N, ln);
#pragma unroll
                    for (int tm = 0; tm < 2; ++tm)
#pragma unroll
                        for (int tn = 0; tn < 4; ++tn)
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                const int i = wi0 + 8 * tm + fg, j = wj0 + 8 * tn + 2 * fq + e;
                                s.Qxx[i + j * N] = acc[tm][tn][e] + (HACC_L ? hxx_r[tm][tn][e] : s.Qxx[i + j * N]);
                            }
                }
            };
            auto c_qux_quu = [&]() {
                constexpr int MT = RL_DMMA ? M / 8 : 1;
                {
                    double acc[MT][1][2];
                    rl_dmma<MT, 1, true, true>(acc, s.uxhT, LDU, 0, s.fxT, LDF, 8 * wq, N, ln);
#pragma unroll
                    for (int tm = 0; tm < MT; ++tm)
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int a = 8 * tm + fg, j = 8 * wq + 2 * fq + e;
                            s.Qux[a + j * LDK] = acc[tm][0][e] + (HACC_L ? hux_r[tm][e] : s.Qux[a + j * LDK]);
                        }
                }
                if (wq < MT * MT) {
                    double acc[1][1][2];
                    const int a0 = 8 * (wq / MT), e0 = 8 * (wq % MT);
                    rl_dmma<1, 1, true, true>(acc, s.uxhT, LDU, a0, s.fuT, LDU, e0, N, ln);
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int o = (a0 + fg) + (e0 + 2 * fq + e) * M;
                        const double q = acc[0][0][e] + (HACC_L ? huu_r[e] : s.Quu[o]);
                        s.Quu[o] = q;
                        s.uu[o] = q;                                                          /* :68 */
                    }
                }
            };
            if (RL_DMMA && RL_OVERLAP) {
                c_qux_quu();
            } else if (RL_DMMA) { /* the constant Hessians (HACC_L) come from the registers, the per-step ones sit in the buffers */
                c_qxx();
                c_qux_quu();
            } else {
                {
                    double acc[TI][TI];
                    rl_gemm<TI, TI, true>(acc, s.xxhT, LDF, N, ti, s.fxT, LDF, N, tl, N);
#pragma unroll
                    for (int ii = 0; ii < TI; ++ii)
#pragma unroll
                        for (int jj = 0; jj < TI; ++jj) {
                            const int i = ti + 16 * ii, j = tl + 16 * jj;
                            if (i < N && j < N) s.Qxx[i + j * N] = acc[ii][jj] + (HACC_L ? s.Hxx[i + j * N] : s.Qxx[i + j * N]);
                        }
                }
                {
                    double acc[TA][TI];
                    rl_gemm<TA, TI, true>(acc, s.uxhT, LDU, M, ti, s.fxT, LDF, N, tl, N);
#pragma unroll
                    for (int aa = 0; aa < TA; ++aa)
#pragma unroll
                        for (int jj = 0; jj < TI; ++jj) {
                            const int a = ti + 16 * aa, j = tl + 16 * jj;
                            if (a < M && j < N) s.Qux[a + j * LDK] = acc[aa][jj] + (HACC_L ? s.Hux[a + j * LDK] : s.Qux[a + j * LDK]);
                        }
                }
                for (int o = tid; o < M * M; o += RL_THREADS) {
                    const int a = o % M, e = o / M;
                    double acc = s.uxhT[a] * s.fuT[e];
                    for (int l = 1; l < N; ++l) acc = ilqr_fma(s.uxhT[l * LDU + a], s.fuT[l * LDU + e], acc);
                    const double q = acc + (HACC_L ? s.Huu[o] : s.Quu[o]);
                    s.Quu[o] = q;
                    s.uu[o] = q;                                                              /* :68 */
                }
            }
            RL_SYNC();
            RL_TICK(3);
            /* fxT, fuT are free from here to the next step's phase B: warps 1.. fetch the next step's Jacobians while warp 0
             * factorises (with M <= 32; otherwise everybody copies first) */
            if (!RL_OVERLAP && t > 0) issue_jac(t - 1, M <= 32 ? 32 : 0);
            /* ---- D: Cholesky of Quu on warp 0, unblocked upper, stop at the first bad pivot (:69, Q3) */
            if (M <= 32) {
                /* in registers: lane j holds column j of the factor; U(k, jj) travels by shuffle.  Same operations in the same
                 * order as the shared-memory version below (and the oracle): every sum is one ascending-k fma chain. */
                if (tid < 32) {
                    chol_regs();
                    if (RL_OVERLAP) rl_bar_arrive(1, 32 + RL_E_THREADS); /* the factor is in shared memory: releases the solves */
                    RL_TICK(4);
                }
                if (RL_OVERLAP) {
                    /* Qxx (every warp; warp 0 after its factorisation) under the factorisation and the triangular solves, which
                     * only need Quu and Qux: the solves run on warps 1.. as soon as those have done their Qxx blocks AND warp 0
                     * has signalled the factor (named barrier 1) */
                    c_qxx();
                    if (tid >= 32 && tid < 32 + RL_E_THREADS) {
                        rl_bar_sync(1, 32 + RL_E_THREADS);
                        const int col = tid - 32;
                        if (col < N) rl_trisolve(s.uu, s.rinv, s.Qux + col * LDK, s.K + col * LDK, K_block(d, T, b, t) + (size_t)col * M, 1);
                        else if (col == N) rl_trisolve(s.uu, s.rinv, s.Qu, s.kk, k_block(d, T, b, t), 1);
                        __syncwarp(); /* converged again (the out-of-line call returns per thread) before the aligned barrier */
                    }
                }
            } else
            if (tid < 32) {
                bool ok = true;
                for (int j = 0; j < M; ++j) {
                    if (ok) {
                        double ajj = s.uu[j + j * M];
#pragma unroll 4
                        for (int k = 0; k < j; ++k) ajj = ilqr_fma(-s.uu[k + j * M], s.uu[k + j * M], ajj);
                        if (!(ajj > 0.0)) {
                            ok = false;
                            __syncwarp();
                            if (tid == 0) { s.uu[j + j * M] = ajj; s_cholfail = 1; }
                        } else {
                            const double ujj = sqrt(ajj);
                            const double r = 1.0 / ujj;
                            __syncwarp();
                            for (int i = j + 1 + tid; i < M; i += 32) {
                                double sum = s.uu[j + i * M];
#pragma unroll 4
                                for (int k = 0; k < j; ++k) sum = ilqr_fma(-s.uu[k + j * M], s.uu[k + i * M], sum);
                                s.uu[j + i * M] = sum * r;
                            }
                            if (tid == 0) s.uu[j + j * M] = ujj;
                        }
                    }
                    __syncwarp();
                }
                for (int j = tid; j < M; j += 32) s.rinv[j] = 1.0 / s.uu[j + j * M];
            }
            if (!RL_OVERLAP) {
                RL_SYNC();
                RL_TICK(4);
            }
            /* ---- E: K = -Quu \ Qux, k = -Quu \ Qu (:70-75): one thread per right-hand side */
            if (!RL_OVERLAP)
            for (int col = tid; col < N + 1; col += RL_THREADS) {
                if (col < N) rl_trisolve(s.uu, s.rinv, s.Qux + col * LDK, s.K + col * LDK, K_block(d, T, b, t) + (size_t)col * M, 1);
                else rl_trisolve(s.uu, s.rinv, s.Qu, s.kk, k_block(d, T, b, t), 1);
            }
            __syncwarp();
            } /* !RL_FWARP */
            } /* !RL_HELPERS */
            __syncthreads(); /* K, k are complete (CTA-wide also with helper warps) */
            RL_TICK(5);
            if (RL_OVERLAP && t > 0) issue_jac(t - 1, 0); /* fxT, fuT are free from here (Qxx read them) to the next step's phase B */
            /* ---- F: uxt = Quu K (:79) */
            if (RL_DMMA && ILQR_RL_FG_DMMA) { /* column tile wq of the m x n product per warp */
                constexpr int MT = RL_DMMA ? M / 8 : 1;
                double acc[MT][1][2];
                rl_dmma<MT, 1, true, false>(acc, s.Quu, M, 0, s.K, LDK, 8 * wq, M, ln);
#pragma unroll
                for (int tm = 0; tm < MT; ++tm)
#pragma unroll
                    for (int e = 0; e < 2; ++e) s.uxt[(8 * tm + fg) + (8 * wq + 2 * fq + e) * LDK] = acc[tm][0][e];
            } else
            for (int o = tid; o < M * N; o += RL_THREADS) {
                const int a = o % M, j = o / M;
                double acc = s.Quu[a] * s.K[j * LDK];
                for (int e = 1; e < M; ++e) acc = ilqr_fma(s.Quu[a + e * M], s.K[e + j * LDK], acc);
                s.uxt[a + j * LDK] = acc;
            }
            RL_SYNC();
            /* ---- G: P = K'uxt + K'Qux + Qux'K + Qxx (:81-84), p (:86-89), Lagrangian gradient (src/solve.jl:75-78).
             *         The old P and p are dead since phase B, so they are overwritten in place. */
            if (RL_DMMA) {
                double a1[2][4][2], a2[2][4][2], a3[2][4][2];
                rl_dmma<2, 4, false, false>(a1, s.K, LDK, wi0, s.uxt, LDK, wj0, M, ln);   /* :81  K' (Quu K) */
                rl_dmma<2, 4, false, false>(a2, s.K, LDK, wi0, s.Qux, LDK, wj0, M, ln);   /* :82  K' Qux */
                rl_dmma<2, 4, false, false>(a3, s.Qux, LDK, wi0, s.K, LDK, wj0, M, ln);   /* :83  Qux' K */
#pragma unroll
                for (int tm = 0; tm < 2; ++tm)
#pragma unroll
                    for (int tn = 0; tn < 4; ++tn)
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int i = wi0 + 8 * tm + fg, j = wj0 + 8 * tn + 2 * fq + e;
                            double v = a1[tm][tn][e];
                            v = v + a2[tm][tn][e];
                            v = v + a3[tm][tn][e];
                            s.P[i + j * LDP] = v + s.Qxx[i + j * N];                                           /* :84 */
                        }
            } else
                {
                    double a1[TI][TI], a2[TI][TI], a3[TI][TI];
#pragma unroll 2
                    for (int a = 0; a < M; ++a) {
                        double ki[TI], kj[TI], uj[TI], qj[TI], qi[TI];
#pragma unroll
                        for (int ii = 0; ii < TI; ++ii) {
                            const int i = ti + 16 * ii;
                            ki[ii] = i < N ? s.K[a + i * LDK] : 0.0;
                            qi[ii] = i < N ? s.Qux[a + i * LDK] : 0.0;
                        }
#pragma unroll
                        for (int jj = 0; jj < TI; ++jj) {
                            const int j = tl + 16 * jj;
                            kj[jj] = j < N ? s.K[a + j * LDK] : 0.0;
                            uj[jj] = j < N ? s.uxt[a + j * LDK] : 0.0;
                            qj[jj] = j < N ? s.Qux[a + j * LDK] : 0.0;
                        }
#pragma unroll
                        for (int ii = 0; ii < TI; ++ii)
#pragma unroll
                            for (int jj = 0; jj < TI; ++jj) {
                                a1[ii][jj] = (a == 0) ? ki[ii] * uj[jj] : ilqr_fma(ki[ii], uj[jj], a1[ii][jj]);   /* :81 */
                                a2[ii][jj] = (a == 0) ? ki[ii] * qj[jj] : ilqr_fma(ki[ii], qj[jj], a2[ii][jj]);   /* :82 */
                                a3[ii][jj] = (a == 0) ? qi[ii] * kj[jj] : ilqr_fma(qi[ii], kj[jj], a3[ii][jj]);   /* :83 */
                            }
                    }
#pragma unroll
                    for (int ii = 0; ii < TI; ++ii)
#pragma unroll
                        for (int jj = 0; jj < TI; ++jj) {
                            const int i = ti + 16 * ii, j = tl + 16 * jj;
                            if (i < N && j < N) {
                                double v = a1[ii][jj];
                                v = v + a2[ii][jj];
                                v = v + a3[ii][jj];
                                s.P[i + j * LDP] = v + s.Qxx[i + j * N];                                           /* :84 */
                            }
                        }
                }
            if (RL_DMMA && ILQR_RL_FG_DMMA && 4 * N <= RL_THREADS) {
                /* p (:86-89) and Lx: the three 16-term sums of an entry on three neighbouring lanes (the fourth idles), combined in
                 * the reference's order; every warp takes its share after its tiles of P */
                const int i = tid >> 2, c = tid & 3;
                const double* A_ = c == 0 ? s.uxt + i * LDK : c == 1 ? s.K + i * LDK : s.Qux + i * LDK;
                const double* v_ = c == 1 ? s.Qu : s.kk;
                double ac = 0.0;
                if (c < 3 && i < N) {
                    ac = A_[0] * v_[0];
#pragma unroll 4
                    for (int a = 1; a < M; ++a) ac = ilqr_fma(A_[a], v_[a], ac);
                }
                const double a2 = __shfl_down_sync(0xffffffffu, ac, 1), a3 = __shfl_down_sync(0xffffffffu, ac, 2);
                if (c == 0 && i < N) {
                    double v = ac;
                    v = v + a2;
                    v = v + a3;
                    const double newp = v + s.Qx[i];
                    const double lx = s.Qx[i] - newp;
                    s.p[i] = newp;
                    d.Lx[((size_t)t * N + i) * Bp + b] = lx;
                    const double av = fabs(lx);
                    if (av > gn || av != av) gn = av;
                }
                if (tid < M) {
                    const double qu = s.Qu[tid];
                    d.Lu[((size_t)t * M + tid) * Bp + b] = qu;
                    const double av = fabs(qu);
                    if (av > gn || av != av) gn = av;
                }
            } else
            if (tid < N) {
                const int i = tid;
                double a1 = s.uxt[i * LDK] * s.kk[0];
                for (int a = 1; a < M; ++a) a1 = ilqr_fma(s.uxt[a + i * LDK], s.kk[a], a1);
                double a2 = s.K[i * LDK] * s.Qu[0];
                for (int a = 1; a < M; ++a) a2 = ilqr_fma(s.K[a + i * LDK], s.Qu[a], a2);
                double a3 = s.Qux[i * LDK] * s.kk[0];
                for (int a = 1; a < M; ++a) a3 = ilqr_fma(s.Qux[a + i * LDK], s.kk[a], a3);
                double v = a1;
                v = v + a2;
                v = v + a3;
                const double newp = v + s.Qx[i];
                const double lx = s.Qx[i] - newp;
                s.p[i] = newp;
                d.Lx[((size_t)t * N + i) * Bp + b] = lx;
                const double av = fabs(lx);
                if (av > gn || av != av) gn = av;
            } else if (tid >= RL_THREADS - M) {
                const int a = tid - (RL_THREADS - M);
                const double qu = s.Qu[a];
                d.Lu[((size_t)t * M + a) * Bp + b] = qu;
                const double av = fabs(qu);
                if (av > gn || av != av) gn = av;
            }
            RL_SYNC(); /* Qxx, Qux, Quu have been consumed: next step's Hessian blocks may land in them */
            RL_TICK(6);
            if (t > 0) { if (HACC_L) rl_commit(); else issue_hess(t - 1); }
            RL_TICK(7);
        }
#ifdef ILQR_RL_PHASE_TIMERS
        if (tid == 0 && b == 0) {
            printf("phase cycles per step: wait_jac %lld  B %lld  wait_hess %lld  C %lld  D(chol+issue) %lld  E %lld  F+G %lld  issue_hess %lld\n",
                   ph[0] / (T - 1), ph[1] / (T - 1), ph[2] / (T - 1), ph[3] / (T - 1), ph[4] / (T - 1), ph[5] / (T - 1), ph[6] / (T - 1), ph[7] / (T - 1));
        }
#endif
        rl_wait_all();
        /* gradient norm: max over the CTA, NaN-propagating like norm(., Inf) */
        for (int off = 16; off > 0; off >>= 1) {
            const double o = __shfl_xor_sync(0xffffffffu, gn, off);
            if (o > gn || o != o) gn = o;
        }
        if ((tid & 31) == 0) s_gn[tid >> 5] = gn;
        RL_SYNC();
        if (tid == 0) {
            double g = 0.0;
            for (int w = 0; w < RL_THREADS / 32; ++w) { const double o = s_gn[w]; if (o > g || o != o) g = o; }
            gn = g;
            if (s_cholfail) d.flags[b] |= ILQR_FLAG_CHOL_FAIL;
            d.gnorm[b] = gn;
        }
    } else if (skip_ls_none) {
        gn = d.gnorm[b];
    }
    if (tid == 0 && b < P.B) {
        const bool running = tick_epilogue(P, b, kind, gn);
        if (running) atomicAdd(&d.active[P.tick & 7], 1);
    }
}

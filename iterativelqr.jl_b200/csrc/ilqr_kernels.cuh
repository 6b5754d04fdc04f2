/*
 * ilqr_kernels.cuh -- sm_100a kernels of the batched iLQR engine, specialised at compile
 * time on one generated model (force-included header: ILQR_N, ILQR_M, ... and the
 * ilqr_dyn / ilqr_cost_* / ilqr_con_* device functions).
 *
 * One lock-step batch iteration ("tick") is three launches:
 *   k_forward   -- forward_pass! (/root/reference/src/forward_pass.jl:1-56): every line-search
 *                  step size of a problem is rolled out at once, one warp per step size,
 *                  32 problems per warp; plus the per-problem solver bookkeeping that sits
 *                  between inner solves (src/solve.jl:93-126, :9-21).
 *   k_linearize -- gradients! (src/gradients.jl:1-98): one thread per (problem, time step).
 *   k_backward  -- backward_pass! + lagrangian_gradient! (src/backward_pass.jl:39-90,
 *                  src/solve.jl:67-83) and the convergence tests of src/solve.jl:36-50:
 *                  one thread per problem, value function in registers.
 *
 * HBM layout: structure of arrays, [field][time][component][problem] with the problem
 * index fastest (padded to a multiple of 32), so a warp touching one component of 32
 * consecutive problems makes one 256-byte transaction.
 *
 * Arithmetic follows the contract stated in oracle/ilqr_oracle.c and DESIGN.md; the file is
 * compiled with -fmad=false so that only the explicit ilqr_fma calls fuse.
 */
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace ilqr {

constexpr int N = ILQR_N, M = ILQR_M, NP = ILQR_P, CS = ILQR_CS, CT = ILQR_CT;
constexpr bool CONSTRAINED = (CS + CT) > 0;
/* models whose per-step matrices do not fit a thread's registers take the loop-based kernels of
 * ilqr_large_*.cuh (same names, same arithmetic, different decomposition) */
#if (ILQR_N * ILQR_N > 100) || defined(ILQR_FORCE_LARGE)
#define ILQR_LARGE 1
#else
#define ILQR_LARGE 0
#endif
/* HACC: ONE stage-Hessian accumulator per problem instead of one per time step.  cost_hessian!
 * (/root/reference/src/costs.jl:70-84) ADDS the stage Hessians onto gxx_t, guu_t, gux_t at every gradients! call of an
 * inner solve (Q1).  When the generated Hessians are constants (ILQR_HESS_CONST: no x, u, w in them) and there are no
 * stage constraints (whose AL terms, src/gradients.jl:66-79, depend on t), every step's accumulator receives the same
 * sequence of additions and holds the same bits -- so one copy per problem, advanced once per tick, replaces
 * 2 x (n^2 + m^2 + mn) rows of HBM traffic per (problem, time step) and the same number of ring doubles. */
#ifndef ILQR_HESS_CONST
#define ILQR_HESS_CONST 0
#endif
#if ILQR_HESS_CONST && (ILQR_CS == 0) && !ILQR_LARGE && !defined(ILQR_NO_HACC)
#define ILQR_HACC 1
#else
#define ILQR_HACC 0
#endif
constexpr bool HACC = ILQR_HACC != 0;
/* the same shortcut on the wide-model path (csrc/ilqr_large_*.cuh): there it removes 2 x 5376 rows per (problem, step)
 * from k_linearize's traffic and the per-step Hessian copies from the Riccati kernel */
#if ILQR_HESS_CONST && (ILQR_CS == 0) && ILQR_LARGE && !defined(ILQR_NO_HACC)
#define ILQR_HACC_L 1
#else
#define ILQR_HACC_L 0
#endif
constexpr bool HACC_L = ILQR_HACC_L != 0;
#if ILQR_LARGE
/* Wide models: the dynamics Jacobians are kept PROBLEM-MAJOR, one contiguous block per (problem, time step) that is already
 * the shared-memory image the CTA-per-problem Riccati kernel wants (fx and fu row by row, rows padded to LDF / LDU doubles
 * for conflict-free DMMA fragment loads): block (b, t) = Dev::fx + (b (T-1) + t) JAC_BLOCK,
 *     fx(k, i) = block[k LDF + i],   fu(k, a) = block[JAC_FU + k LDU + a].
 * The Riccati kernel fetches a step with one bulk copy (TMA) instead of 5120 scattered 8-byte copies whose addresses were
 * a batch row apart; k_forward_wp's sensitivity sweep reads rows contiguously.  Dev::fu is unused on this path. */
#ifndef ILQR_RL_DMMA
#define ILQR_RL_DMMA 1
#endif
constexpr bool RL_DMMA = (ILQR_RL_DMMA != 0) && (N == 64) && (M % 8 == 0) && (M <= 16); /* the warp -> tile map of k_backward is written for n = 64 */
constexpr int LDF = RL_DMMA ? N + 8 : N;   /* fxT, xxhT: [k][i], i contiguous */
constexpr int LDU = RL_DMMA ? M + 8 : M;   /* fuT, uxhT: [k][a], a contiguous */
constexpr int JAC_FU = N * LDF;
constexpr int JAC_BLOCK = (N * LDF + N * LDU + 1) & ~1; /* doubles; even, so that a block is a whole number of 16-byte units */
/* JAC_CONST: the generated Jacobians contain no x, u, w (a linear time-invariant plant, ILQR_JAC_CONST from the code
 * generator): ONE block, written when the workspace is created, serves every problem and time step -- gradients! has no
 * dynamics part left and the Riccati kernel's Jacobian reads come out of L2.  -DILQR_NO_JAC_CONST keeps per-step blocks. */
#if defined(ILQR_JAC_CONST) && ILQR_JAC_CONST && !defined(ILQR_NO_JAC_CONST)
constexpr bool JAC_CONST = true;
#else
constexpr bool JAC_CONST = false;
#endif
#endif
constexpr int NH = ILQR_N * ILQR_N + ILQR_M * ILQR_M + ILQR_M * ILQR_N; /* gxx | guu | gux, column-major each */
__host__ __device__ constexpr int d1(int v) { return v > 0 ? v : 1; }

enum : int { PH_DONE = 0, PH_START = 1, PH_ITER = 2, PH_SHIFT = 3,
              PH_PAUSED = 4 /* ilqr_solve_outer: waiting for the host's augmented_lagrangian_callback! after a dual update */ };
enum : int { MODE_BATCH = 0, MODE_STREAM = 1, MODE_MPC = 2 };
enum : int { KIND_NONE = 0, KIND_PRELOOP = 1, KIND_ITER = 2 };

struct Dev {
    /* ProblemData (src/data/problem.jl:3-23): nominal and current trajectories, parameters */
    double *xb, *ub, *xc, *uc, *w;
    /* ModelData / ObjectiveData (src/data/model.jl:5-10, src/data/objective.jl:3-10) */
    double *fx, *fu, *gx, *gu, *gxx, *guu, *gux;
    double* hacc; /* HACC: [NH][Bp] per-problem stage-Hessian accumulator (gxx | guu | gux); only the terminal block of gxx is used then */
    /* PolicyData gains (src/data/policy.jl:25-26) and the Lagrangian gradient blocks (src/data/solver.jl:6) */
    double *K, *k, *Lx, *Lu;
    /* AugmentedLagrangianCosts (src/augmented_lagrangian.jl:1-11): rows = (T-1)*CS + CT */
    double *c, *lam, *rho;
    uint8_t* act;
    /* SolverData scalars (src/data/solver.jl:4-18), one per problem */
    double *J, *obj_prev, *viol, *alpha, *gnorm;
    double* dgp;      /* expected-decrease term of the line search in progress (src/forward_pass.jl:19-20) */
    int32_t* ls_base; /* first step-size index the next k_forward evaluates for this problem (0 = new search) */
    int32_t *status, *iters, *outer, *it, *phase, *kind, *inner_done;
    int32_t* iters0; /* data.iterations when the solve in progress began (unconstrained solves never reset it: src/solve.jl:137-139) */
    uint32_t* flags;
    /* iteration records (src/solve.jl:40-45) [record][problem] */
    double *h_cost, *h_gnorm, *h_viol, *h_alpha;
    int32_t* h_outer;
    uint8_t* h_status;
    int32_t* active; /* ring of 8 counters: problems still running after a tick */
    /* continuous batching (ilqr_solve_stream): slot -> problem id, slots that finished this tick */
    int32_t* pid;
    int32_t* done_list;  /* [4][Bp]: ring indexed by tick & 3 */
    int32_t* done_count; /* [4] */
    int32_t* pending;    /* slot refilled by k_refill: 1 + (tick & 7) of the k_forward that is to start it, 0 = none */
    int32_t* refilling;  /* ticks for which a just-finished slot still counts as running (its refill is in flight) */
    int32_t *mpc_step, *mpc_iters; /* MODE_MPC: re-solves completed, iterations summed over them */
    /* drain compaction (MODE_STREAM): the per-slot arrays that are live between two ticks, and the move plan */
    const struct MoveEntry* mv;
    int32_t n_mv;
    int32_t *cmp_src, *cmp_dst; /* [Bp] each: slot cmp_src[i] moves to slot cmp_dst[i] */
    int32_t* cmp_n;             /* [2]: sources found, destinations found */
};
#if ILQR_LARGE
__device__ __forceinline__ double* jac_block(const Dev& d, int T, int b, int t) {
    return JAC_CONST ? d.fx : d.fx + ((size_t)b * (T - 1) + t) * JAC_BLOCK;
}
/* Wide models keep the policy PROBLEM-MAJOR as well: K_t (m x n, column-major) of problem b at Dev::K + (b (T-1) + t) m n,
 * k_t at Dev::k + (b (T-1) + t) m -- what the CTA-per-problem kernels on both sides of it want (the Riccati kernel writes a
 * gain as one contiguous run, k_forward_wp fetches it with 16-byte asynchronous copies); it is also the host layout of
 * ilqr_get_policy. */
__device__ __forceinline__ double* K_block(const Dev& d, int T, int b, int t) { return d.K + ((size_t)b * (T - 1) + t) * (M * N); }
__device__ __forceinline__ double* k_block(const Dev& d, int T, int b, int t) { return d.k + ((size_t)b * (T - 1) + t) * M; }
#endif
struct MoveEntry { char* base; int32_t rows; int32_t elsize; }; /* array [rows][Bp] of elsize-byte elements */

/* A streamed job: n_total independent problems flow through the handle's `batch` slots; a slot whose
 * problem has terminated is retired (results written out) and refilled from the queue between ticks.
 * All pointers are device memory in the ABI's [problem][time][component] layout; outputs may be NULL. */
struct Job {
    int32_t n_total;
    int32_t* next; /* queue head: next problem id to hand out */
    const double *in_x, *in_u, *in_w;
    double *out_x, *out_u;
    int32_t* out_iters;
    uint8_t* out_status;
    double *out_J, *out_viol, *out_alpha;
    uint32_t* out_flags;
    /* MODE_MPC: every slot re-solves mpc_steps times (plant step with the first action, shift, roll out, solve)
     * at its own pace; logs are [step][problem][component] or NULL */
    int32_t mpc_steps;
    double *mpc_u, *mpc_x;
};

struct Params {
    Dev d;
    int T, B, Bp, cap;
    int n_alpha; /* line-search trials: src/forward_pass.jl:28-29 */
    int tick;
    int mode;           /* MODE_BATCH (ilqr_solve), MODE_STREAM (ilqr_solve_stream), MODE_MPC (ilqr_mpc_run) */
    int pause_outer;    /* MODE_BATCH only: park every problem after each dual update (ilqr_solve_outer) */
    const Job* job;     /* device copy of the current Job */
    ilqr_options o;
};

/* ---- small helpers ---------------------------------------------------------------- */
template <int R>
__device__ __forceinline__ void ld_rows(double* dst, const double* __restrict__ base, size_t row0, int Bp, int b) {
#pragma unroll
    for (int i = 0; i < R; ++i) dst[i] = base[(row0 + i) * (size_t)Bp + b];
}
template <int R>
__device__ __forceinline__ void st_rows(const double* src, double* __restrict__ base, size_t row0, int Bp, int b) {
#pragma unroll
    for (int i = 0; i < R; ++i) base[(row0 + i) * (size_t)Bp + b] = src[i];
}

/* Rows whose content is DEAD for the lanes that sit this tick out (fx, fu, Lx, Lu of a problem that is in the middle of a
 * line search, between two inner solves or finished: the next gradients! call rewrites them before anything reads them)
 * are stored by ALL lanes of the warp, the idle ones writing zeros.  A warp store that skips lanes leaves partially
 * written 32-byte sectors behind, and the memory system has to READ those back from DRAM to merge them: ncu showed
 * 0.75 sectors of DRAM read per sector written by k_linback with 13 % of the lanes idle (profiles/README.md). */
template <int R>
__device__ __forceinline__ void store_rows_full(const double* src, bool live, double* __restrict__ base, size_t row0, int Bp, int b) {
#pragma unroll
    for (int i = 0; i < R; ++i) base[(row0 + i) * (size_t)Bp + b] = live ? src[i] : 0.0;
}

/* acc = a0*b0; acc = fma(a_k, b_k, acc) -- the contract's dot product */
template <int LEN, int SA, int SB>
__device__ __forceinline__ double dotf(const double* a, const double* b) {
    if (LEN <= 0) return 0.0;
    double acc = a[0] * b[0];
#pragma unroll
    for (int k = 1; k < LEN; ++k) acc = ilqr_fma(a[k * SA], b[k * SB], acc);
    return acc;
}

__device__ __forceinline__ double pow2neg(int c) { return __longlong_as_double((long long)(1023 - c) << 52); }

__device__ __forceinline__ void viol_update(double& mv, double ci, bool ineq) {
    const double v = ineq ? (ci > 0.0 ? ci : 0.0) : fabs(ci);
    if (v > mv) mv = v;
}

/* AL terms of one stage for the cost (src/augmented_lagrangian.jl:55-63) given c, lam, rho:
 * returns the active set in a[] and adds onto Jal in the contract's order */
template <int R, bool TERM>
__device__ __forceinline__ void al_stage_cost(const double* c, const double* lam, const double* rho, uint8_t* a, double& Jal) {
    if (R <= 0) return;
#pragma unroll
    for (int i = 0; i < R; ++i) {
        const bool ineq = TERM ? ilqr_ineq_T(i) : ilqr_ineq_s(i);
        a[i] = (ineq && c[i] < 0.0 && lam[i] == 0.0) ? 0 : 1; /* src/augmented_lagrangian.jl:77-83 */
    }
    Jal += dotf<R, 1, 1>(lam, c);
#pragma unroll
    for (int i = 0; i < R; ++i)
        if (a[i] == 1) Jal += (0.5 * rho[i]) * (c[i] * c[i]);
}

/* ---- closed-loop rollout + cost of one line-search trial ---------------------------
 * rollout! (src/rollout.jl:19-29) fused with cost!(mode=:current) (src/data/methods.jl:13-30
 * -> src/augmented_lagrangian.jl:39-66, src/data/constraints.jl:23-39).  Every trial writes
 * its trajectory / constraint values / active set into the slot buffers it is given (slot 0 is
 * the canonical "current" state of the problem), so that the accepted trial never has to be
 * rolled out a second time. */
struct TrialOut { double *x, *u, *c; uint8_t* a; };

/* ---- cp.async (LDGSTS) helpers: each lane copies / reads back its own 8-byte column of a [row][32] stage ---- */
__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int PENDING>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(PENDING) : "memory"); }
template <int R>
__device__ __forceinline__ void cp_rows(double*& dst, const double* __restrict__ base, size_t row0, int Bp, int b) {
#pragma unroll
    for (int i = 0; i < R; ++i) { cp_async8(dst, base + (row0 + i) * (size_t)Bp + b); dst += 32; }
}
template <int R>
__device__ __forceinline__ void lds_rows(double* dst, const double*& src) {
#pragma unroll
    for (int i = 0; i < R; ++i) { dst[i] = *src; src += 32; }
}

/* ---- mbarrier helpers (shared::cta) ---- */
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.shared::cta.b64 st, [%0];\n}" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n.reg .pred p;\nWAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(a), "r"(parity) : "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(double* smem_dst, const double* gsrc, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}

/* producers are ahead of the Riccati warp most of the time: wait politely (the spin of the plain loop was 30 % of
 * the instructions the kernel issued and competed with the Riccati warp for its sub-partition's issue slots) */
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, unsigned parity) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    unsigned done = 0;
    while (true) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(a), "r"(parity) : "memory");
        if (done) break;
        __nanosleep(256);
    }
}



#ifndef ILQR_FWD_TRIALS
#define ILQR_FWD_TRIALS 2 /* step sizes per k_forward launch; see k_forward */
#endif
constexpr int FWD_TRIAL_WARPS = ILQR_FWD_TRIALS;

#if !ILQR_LARGE
struct PolicyRow { /* what one rollout step reads besides the running state; fetched ahead of the step that uses it */
    double Kt[d1(M * N)], kt[d1(M)], ubt[d1(M)], xbt[N], lam[d1(CS)], rho[d1(CS)];
};
__device__ __forceinline__ void load_policy_row(PolicyRow& r, const Dev& d, int t, int Bp, int b) {
    ld_rows<M * N>(r.Kt, d.K, (size_t)t * M * N, Bp, b);
    ld_rows<M>(r.kt, d.k, (size_t)t * M, Bp, b);
    ld_rows<M>(r.ubt, d.ub, (size_t)t * M, Bp, b);
    ld_rows<N>(r.xbt, d.xb, (size_t)t * N, Bp, b);
    ld_rows<CS>(r.lam, d.lam, (size_t)t * CS, Bp, b);
    ld_rows<CS>(r.rho, d.rho, (size_t)t * CS, Bp, b);
}

/* Ring version of the same fetch.  A load issued one step ahead (register prefetch) still arrives late: ncu
 * shows the rollout warps spending a quarter of their time on the long scoreboard of the `cur = nxt` move,
 * i.e. ~1500 cycles of load latency against a ~1000-cycle step.  So every trial warp streams its rows through
 * its own PR_STAGES-deep shared-memory ring, PR_STAGES-1 steps ahead; lanes only ever touch their own column,
 * so no barrier is involved, only cp.async.wait_group.  Falls back to the register prefetch when the ring
 * does not fit (wide rows). */
constexpr int DG_ROWS_PER_STEP = M * N + M + N * N + N * M + N + M;
constexpr int DG_STAGES_FIT = (96 * 1024) / (DG_ROWS_PER_STEP * 32 * 8);
/* ring depths: 4 stages hide the load latency as well as 8 did (measured), and 50 KB of rings per CTA instead of
 * 90 KB lets three k_forward CTAs share an SM */
#ifndef ILQR_DG_MAX_STAGES
#define ILQR_DG_MAX_STAGES 4
#endif
#ifndef ILQR_PR_MAX_STAGES
#define ILQR_PR_MAX_STAGES 4
#endif
constexpr int DG_STAGES = DG_STAGES_FIT >= ILQR_DG_MAX_STAGES ? ILQR_DG_MAX_STAGES : (DG_STAGES_FIT >= 2 ? DG_STAGES_FIT : 2);
constexpr int DG_SMEM_BYTES = DG_STAGES * DG_ROWS_PER_STEP * 32 * 8;
constexpr int PR_ROWS_PER_STEP = M * N + M + M + N + CS + CS + NP;
constexpr int PR_STAGES_FIT = (96 * 1024) / (FWD_TRIAL_WARPS * PR_ROWS_PER_STEP * 32 * 8);
#ifdef ILQR_NO_POLICY_RING
constexpr int PR_STAGES = 0;
#else
constexpr int PR_STAGES = PR_STAGES_FIT >= ILQR_PR_MAX_STAGES ? ILQR_PR_MAX_STAGES : (PR_STAGES_FIT >= 3 ? PR_STAGES_FIT : 0);
#endif
constexpr int PR_WARP_DOUBLES = PR_STAGES * PR_ROWS_PER_STEP * 32;
constexpr int PR_SMEM_BYTES = FWD_TRIAL_WARPS * PR_WARP_DOUBLES * 8;
constexpr int FWD_SMEM_BYTES = DG_SMEM_BYTES + PR_SMEM_BYTES;

__device__ __forceinline__ void pr_issue(double* stage_lane, const Dev& d, int t, int Bp, int b) {
    double* p = stage_lane;
    cp_rows<M * N>(p, d.K, (size_t)t * M * N, Bp, b);
    cp_rows<M>(p, d.k, (size_t)t * M, Bp, b);
    cp_rows<M>(p, d.ub, (size_t)t * M, Bp, b);
    cp_rows<N>(p, d.xb, (size_t)t * N, Bp, b);
    cp_rows<CS>(p, d.lam, (size_t)t * CS, Bp, b);
    cp_rows<CS>(p, d.rho, (size_t)t * CS, Bp, b);
    cp_rows<NP>(p, d.w, (size_t)t * NP, Bp, b);
}

/* one step of rollout! + cost!(mode=:current) for one trial: control law (src/rollout.jl:24-28), the trial's trajectory /
 * constraint rows, stage cost and AL terms, dynamics (src/rollout.jl:29).  Shared by every forward kernel. */
struct RollAcc { double Jc = 0.0, Jal = 0.0, mv = 0.0; };
__device__ __forceinline__ void rollout_step(const TrialOut& o, int t, int Bp, int b, double alpha, const PolicyRow& cur,
                                             const double* wv, double* x, double* u, RollAcc& acc) {
    double xn[N];
#pragma unroll
    for (int a = 0; a < M; ++a) {
        double v = cur.kt[a] * alpha;                    /* src/rollout.jl:24-25 */
        v = v + cur.ubt[a];                              /* :26 */
        v = v + dotf<N, M, 1>(cur.Kt + a, x);            /* :27 */
        v = v - dotf<N, M, 1>(cur.Kt + a, cur.xbt);      /* :28 */
        u[a] = v;
    }
    if (o.x) { /* a speculative trial only reports its cost */
        st_rows<N>(x, o.x, (size_t)t * N, Bp, b);
        st_rows<M>(u, o.u, (size_t)t * M, Bp, b);
    }
    double g;
    ilqr_cost_s(&g, x, u, wv);
    acc.Jc += g;
    if (CS > 0) {
        double c[d1(CS)];
        uint8_t a[d1(CS)];
#if ILQR_CS > 0
        ilqr_con_s(c, x, u, wv);
#endif
        al_stage_cost<CS, false>(c, cur.lam, cur.rho, a, acc.Jal);
#pragma unroll
        for (int i = 0; i < CS; ++i) viol_update(acc.mv, c[i], ilqr_ineq_s(i));
        if (o.c) {
            st_rows<CS>(c, o.c, (size_t)t * CS, Bp, b);
#pragma unroll
            for (int i = 0; i < CS; ++i) o.a[((size_t)t * CS + i) * Bp + b] = a[i];
        }
    }
    ilqr_dyn(xn, x, u, wv);                              /* :29 */
#pragma unroll
    for (int i = 0; i < N; ++i) x[i] = xn[i];
}
/* the terminal stage of the same sweep: state, terminal cost and constraint rows */
__device__ __forceinline__ void rollout_terminal(const Params& P, const TrialOut& o, int b, const double* lamT, const double* rhoT,
                                                 double* x, double* u, RollAcc& acc) {
    const Dev& d = P.d;
    const int Bp = P.Bp, t = P.T - 1;
    double wv[d1(NP)];
    ld_rows<NP>(wv, d.w, (size_t)t * NP, Bp, b);
    if (o.x) st_rows<N>(x, o.x, (size_t)t * N, Bp, b);
    double g;
    ilqr_cost_T(&g, x, u, wv);
    acc.Jc += g;
    if (CT > 0) {
        double c[d1(CT)];
        uint8_t a[d1(CT)];
#if ILQR_CT > 0
        ilqr_con_T(c, x, u, wv);
#endif
        al_stage_cost<CT, true>(c, lamT, rhoT, a, acc.Jal);
#pragma unroll
        for (int i = 0; i < CT; ++i) viol_update(acc.mv, c[i], ilqr_ineq_T(i));
        if (o.c) {
            st_rows<CT>(c, o.c, (size_t)t * CS, Bp, b);
#pragma unroll
            for (int i = 0; i < CT; ++i) o.a[((size_t)t * CS + i) * Bp + b] = a[i];
        }
    }
}

/* ring_lane: this warp's ring + lane (PR_STAGES > 0), unused otherwise */
__device__ __forceinline__ void rollout_eval(const Params& P, const TrialOut& o, int b, double alpha, double& J_out,
                                             double& viol_out, double* ring_lane) {
    const Dev& d = P.d;
    const int Bp = P.Bp, T = P.T;
    constexpr bool RING = PR_STAGES > 0;
    constexpr int ST = RING ? PR_STAGES : 1;
    double x[N], u[d1(M)], wv[d1(NP)];
    RollAcc acc;
    ld_rows<N>(x, d.xb, 0, Bp, b);
    PolicyRow cur;
    if (RING) {
#pragma unroll 1
        for (int s0 = 0; s0 < ST - 1; ++s0) {
            if (s0 < T - 1) pr_issue(ring_lane + (size_t)s0 * PR_ROWS_PER_STEP * 32, d, s0, Bp, b);
            cp_async_commit();
        }
    } else {
        load_policy_row(cur, d, 0, Bp, b);
    }
    double lamT[d1(CT)], rhoT[d1(CT)];
    ld_rows<CT>(lamT, d.lam, (size_t)(T - 1) * CS, Bp, b);
    ld_rows<CT>(rhoT, d.rho, (size_t)(T - 1) * CS, Bp, b);
    int stage = 0;
#pragma unroll 1
    for (int t = 0; t < T - 1; ++t) {
        PolicyRow nxt; /* register prefetch (no ring): the next step's rows do not depend on this step's result */
        if (RING) {
            const int tp = t + ST - 1;
            int ps = stage + ST - 1;
            if (ps >= ST) ps -= ST;
            if (tp < T - 1) pr_issue(ring_lane + (size_t)ps * PR_ROWS_PER_STEP * 32, d, tp, Bp, b);
            cp_async_commit();
            cp_async_wait<ST - 1>(); /* this lane's copies for step t have landed */
            const double* q = ring_lane + (size_t)stage * PR_ROWS_PER_STEP * 32;
            lds_rows<M * N>(cur.Kt, q); lds_rows<M>(cur.kt, q); lds_rows<M>(cur.ubt, q); lds_rows<N>(cur.xbt, q);
            lds_rows<CS>(cur.lam, q); lds_rows<CS>(cur.rho, q); lds_rows<NP>(wv, q);
            if (++stage == ST) stage = 0;
        } else {
            if (t + 1 < T - 1) load_policy_row(nxt, d, t + 1, Bp, b);
            ld_rows<NP>(wv, d.w, (size_t)t * NP, Bp, b);
        }
        rollout_step(o, t, Bp, b, alpha, cur, wv, x, u, acc);
        if (!RING && t + 1 < T - 1) cur = nxt;
    }
    if (RING) cp_async_wait<0>();
    rollout_terminal(P, o, b, lamT, rhoT, x, u, acc);
    J_out = CONSTRAINED ? (acc.Jc + acc.Jal) : acc.Jc;
    viol_out = acc.mv;
}

#endif /* !ILQR_LARGE */

/* MODE_STREAM: a slot's problem has terminated at this tick.  It goes onto the tick's done list for k_refill, which
 * runs CONCURRENTLY with the next tick (the refilled slot restarts two ticks later, k_refill is off the critical
 * path); while its refill is in flight the slot keeps counting as running so that the host does not stop early. */
__device__ __forceinline__ void stream_hand_to_refill(const Params& P, int b) {
    const Dev& d = P.d;
    const int ring = P.tick & 3;
    const int idx = atomicAdd(&d.done_count[ring], 1);
    d.done_list[(size_t)ring * P.Bp + idx] = b;
    d.refilling[b] = (*P.job->next < P.job->n_total) ? 2 : 0;
}

/* a solve has just terminated: in MODE_MPC the slot goes on to its next receding-horizon step */
__device__ __forceinline__ int mpc_next_phase(const Params& P, int b) {
    if (P.mode != MODE_MPC) return PH_DONE;
    const Dev& d = P.d;
    const int s = d.mpc_step[b] + 1;
    d.mpc_step[b] = s;
    d.mpc_iters[b] += d.iters[b] - d.iters0[b]; /* this re-solve's own iterations */
    return s < P.job->mpc_steps ? PH_SHIFT : PH_DONE;
}

/* prologue of solve!/constrained_ilqr_solve! for one problem: reset!(data), duals, penalties (src/solve.jl:93-103) */
__device__ __forceinline__ void solve_begin_slot(const Params& P, int b) {
    const Dev& d = P.d;
    d.flags[b] = 0;
    d.inner_done[b] = 0;
    d.kind[b] = KIND_NONE;
    d.it[b] = 0;
    d.ls_base[b] = 0;
    if (CONSTRAINED) {
        d.J[b] = 0.0; d.viol[b] = 0.0; d.status[b] = 0; d.iters[b] = 0; d.gnorm[b] = 0.0;
        d.outer[b] = 1;
        const int rows = (P.T - 1) * CS + CT;
        for (int r = 0; r < rows; ++r) {
            d.lam[(size_t)r * P.Bp + b] = 0.0;
            d.rho[(size_t)r * P.Bp + b] = P.o.initial_constraint_penalty;
        }
        d.iters0[b] = 0;
        d.phase[b] = P.o.max_dual_updates > 0 ? PH_START : PH_DONE;
    } else {
        d.outer[b] = 0;
        d.iters0[b] = d.iters[b];
        d.phase[b] = PH_START;
    }
}

/* receding-horizon shift of one problem (ilqr_mpc_step / ilqr_mpc_run): plant step with the first nominal
 * action, shift the actions left (repeat the last), roll the nominal states out again, then the solve!
 * prologue.  The current trajectory and constraint buffers persist (warm start, src/solve.jl:131-135). */
__device__ __noinline__ void mpc_shift_slot(const Params& P, int b, double* log_u, double* log_x) {
    const Dev& d = P.d;
    const int Bp = P.Bp, T = P.T;
    double xv[N], xn[N], uv[d1(M)], wv[d1(NP)];
    ld_rows<N>(xv, d.xb, 0, Bp, b);
    ld_rows<M>(uv, d.ub, 0, Bp, b);
    ld_rows<NP>(wv, d.w, 0, Bp, b);
    ilqr_dyn(xn, xv, uv, wv);
    if (log_u) {
#pragma unroll
        for (int a = 0; a < M; ++a) log_u[a] = uv[a];
    }
#pragma unroll
    for (int i = 0; i < N; ++i) { if (log_x) log_x[i] = xn[i]; xv[i] = xn[i]; }
    st_rows<N>(xv, d.xb, 0, Bp, b);
    double unext[d1(M)];
    ld_rows<M>(unext, d.ub, (size_t)(T > 2 ? 1 : 0) * M, Bp, b);
    for (int t = 0; t < T - 1; ++t) {
#pragma unroll
        for (int a = 0; a < M; ++a) uv[a] = unext[a];
        if (t < T - 2) {
            st_rows<M>(uv, d.ub, (size_t)t * M, Bp, b);
            if (t + 2 < T - 1) ld_rows<M>(unext, d.ub, (size_t)(t + 2) * M, Bp, b); /* one step ahead of the chain */
        }
        ld_rows<NP>(wv, d.w, (size_t)t * NP, Bp, b);
        ilqr_dyn(xn, xv, uv, wv);
#pragma unroll
        for (int i = 0; i < N; ++i) xv[i] = xn[i];
        st_rows<N>(xv, d.xb, (size_t)(t + 1) * N, Bp, b);
    }
    solve_begin_slot(P, b);
}

/* cost!(mode=:nominal) (src/data/methods.jl:13-30): J and the active set from the NOMINAL
 * trajectory, then c and max_violation from the CURRENT one (Q2).  The sweep over t reads its
 * inputs in chunks of CB_CHUNK steps so that the DRAM latency is paid once per chunk. */
constexpr int CB_CHUNK = 4;
struct CostRow { double xb[N], ub[d1(M)], xc[N], uc[d1(M)], w[d1(NP)], lam[d1(CS)], rho[d1(CS)]; };

__device__ __noinline__ void cost_bang_nominal(const Params& P, int b, double& J_out, double& viol_out) {
    const Dev& d = P.d;
    const int Bp = P.Bp, T = P.T;
    double Jc = 0.0, Jal = 0.0, mv = 0.0;
#pragma unroll 1
    for (int t0 = 0; t0 < T - 1; t0 += CB_CHUNK) {
        CostRow rows[CB_CHUNK];
#pragma unroll
        for (int j = 0; j < CB_CHUNK; ++j) {
            const int t = t0 + j;
            if (t < T - 1) {
                ld_rows<N>(rows[j].xb, d.xb, (size_t)t * N, Bp, b);
                ld_rows<M>(rows[j].ub, d.ub, (size_t)t * M, Bp, b);
                ld_rows<NP>(rows[j].w, d.w, (size_t)t * NP, Bp, b);
                if (CS > 0) {
                    ld_rows<N>(rows[j].xc, d.xc, (size_t)t * N, Bp, b);
                    ld_rows<M>(rows[j].uc, d.uc, (size_t)t * M, Bp, b);
                    ld_rows<CS>(rows[j].lam, d.lam, (size_t)t * CS, Bp, b);
                    ld_rows<CS>(rows[j].rho, d.rho, (size_t)t * CS, Bp, b);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < CB_CHUNK; ++j) {
            const int t = t0 + j;
            if (t < T - 1) {
                double g;
                ilqr_cost_s(&g, rows[j].xb, rows[j].ub, rows[j].w);
                Jc += g;
                if (CS > 0) {
                    double c[d1(CS)];
                    uint8_t a[d1(CS)];
#if ILQR_CS > 0
                    ilqr_con_s(c, rows[j].xb, rows[j].ub, rows[j].w);
#endif
                    al_stage_cost<CS, false>(c, rows[j].lam, rows[j].rho, a, Jal);
#pragma unroll
                    for (int i = 0; i < CS; ++i) d.act[((size_t)t * CS + i) * Bp + b] = a[i];
#if ILQR_CS > 0
                    ilqr_con_s(c, rows[j].xc, rows[j].uc, rows[j].w);
#endif
                    st_rows<CS>(c, d.c, (size_t)t * CS, Bp, b);
#pragma unroll
                    for (int i = 0; i < CS; ++i) viol_update(mv, c[i], ilqr_ineq_s(i));
                }
            }
        }
    }
    {
        const int t = T - 1;
        double x[N], u[d1(M)], wv[d1(NP)];
        ld_rows<N>(x, d.xb, (size_t)t * N, Bp, b);
        ld_rows<NP>(wv, d.w, (size_t)t * NP, Bp, b);
        double g;
        ilqr_cost_T(&g, x, u, wv);
        Jc += g;
        if (CT > 0) {
            double c[d1(CT)], lam[d1(CT)], rho[d1(CT)];
            uint8_t a[d1(CT)];
#if ILQR_CT > 0
            ilqr_con_T(c, x, u, wv);
#endif
            ld_rows<CT>(lam, d.lam, (size_t)t * CS, Bp, b);
            ld_rows<CT>(rho, d.rho, (size_t)t * CS, Bp, b);
            al_stage_cost<CT, true>(c, lam, rho, a, Jal);
#pragma unroll
            for (int i = 0; i < CT; ++i) d.act[((size_t)t * CS + i) * Bp + b] = a[i];
            ld_rows<N>(x, d.xc, (size_t)t * N, Bp, b);
#if ILQR_CT > 0
            ilqr_con_T(c, x, u, wv);
#endif
            st_rows<CT>(c, d.c, (size_t)t * CS, Bp, b);
#pragma unroll
            for (int i = 0; i < CT; ++i) viol_update(mv, c[i], ilqr_ineq_T(i));
        }
    }
    J_out = CONSTRAINED ? (Jc + Jal) : Jc;
    viol_out = mv;
}

/* augmented_lagrangian_update! (src/augmented_lagrangian.jl:87-110), fused dual + penalty */
__device__ __forceinline__ void al_update(const Params& P, int b) {
    const Dev& d = P.d;
    const size_t Bp = P.Bp;
    const int rows_s = (P.T - 1) * CS;
#pragma unroll 4
    for (int r = 0; r < rows_s + CT; ++r) {
        const int i = r < rows_s ? (CS > 0 ? r % d1(CS) : 0) : r - rows_s;
        const bool ineq = r < rows_s ? ilqr_ineq_s(i) : ilqr_ineq_T(i);
        const size_t idx = (size_t)r * Bp + b;
        const double rho = d.rho[idx];
        double lam = d.lam[idx] + rho * d.c[idx];
        if (ineq) lam = lam > 0.0 ? lam : 0.0;
        d.lam[idx] = lam;
        const double sc = P.o.scaling_penalty * rho;
        d.rho[idx] = sc < P.o.max_penalty ? sc : P.o.max_penalty;
    }
}

/* What happens to one problem between two inner solves: the tail of the AL loop
 * (src/solve.jl:113-122) and the head of ilqr_solve! (src/solve.jl:9-14, :21). */
__device__ __noinline__ void start_bookkeeping(const Params& P, int b) {
    const Dev& d = P.d;
    double J, mv;
    if (CONSTRAINED && d.inner_done[b]) {
        cost_bang_nominal(P, b, J, mv);                   /* src/solve.jl:113 */
        d.J[b] = J;
        d.viol[b] = mv;
        bool done = mv <= P.o.constraint_tolerance;       /* :117 */
        if (!done) {
            al_update(P, b);                              /* :120-122 */
            const int outer = d.outer[b] + 1;
            d.outer[b] = outer;
            if (P.pause_outer) { /* :125: the host runs augmented_lagrangian_callback!(solver) before the loop goes on */
                d.inner_done[b] = 0;
                d.kind[b] = KIND_NONE;
                d.phase[b] = PH_PAUSED;
                return;
            }
            done = outer > P.o.max_dual_updates;          /* :105 */
        }
        if (done) {
            d.phase[b] = mpc_next_phase(P, b);
            d.kind[b] = KIND_NONE;
            if (P.mode == MODE_STREAM) stream_hand_to_refill(P, b);
            return;
        }
    }
    if (CONSTRAINED && P.pause_outer && d.outer[b] > P.o.max_dual_updates) { /* resumed after the last dual update: :105 ends the loop */
        d.phase[b] = mpc_next_phase(P, b);
        d.kind[b] = KIND_NONE;
        return;
    }
    if (P.o.reset_cache) {                                /* :12 */
        d.J[b] = 0.0; d.viol[b] = 0.0; d.status[b] = 0; d.iters[b] = 0; d.iters0[b] = 0;
    }
    cost_bang_nominal(P, b, J, mv);                       /* :14 */
    d.J[b] = J;
    if (CONSTRAINED) d.viol[b] = mv;
    d.obj_prev[b] = J;                                    /* :21 */
    d.inner_done[b] = 0;
    d.it[b] = 0;
    d.kind[b] = KIND_PRELOOP;
}

#if !ILQR_LARGE
/* trajectory_sensitivities + gradient' * trajectory (src/data/methods.jl:42-54, src/forward_pass.jl:19-20).
 * The recursion only carries dx, so the rows it reads (K, k, fx, fu, Lx, Lu of every step) are streamed
 * through a DG_STAGES-deep shared-memory ring with cp.async (LDGSTS): each lane copies its own column,
 * DG_STAGES-1 steps ahead of the step it is consuming, and never waits on DRAM latency. */
__device__ __forceinline__ void dg_issue(double* stage_lane, const Dev& d, int t, int Bp, int b) {
    double* p = stage_lane;
    cp_rows<M * N>(p, d.K, (size_t)t * M * N, Bp, b);
    cp_rows<M>(p, d.k, (size_t)t * M, Bp, b);
    cp_rows<N * N>(p, d.fx, (size_t)t * N * N, Bp, b);
    cp_rows<N * M>(p, d.fu, (size_t)t * N * M, Bp, b);
    cp_rows<N>(p, d.Lx, (size_t)t * N, Bp, b);
    cp_rows<M>(p, d.Lu, (size_t)t * M, Bp, b);
}

__device__ __noinline__ double delta_grad_product(const Params& P, int b, double* ring, int lane) {
    const Dev& d = P.d;
    const int Bp = P.Bp, T = P.T;
    double zx[N], zy[N], zu[d1(M)];
    double sx = 0.0, su = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) zx[i] = 0.0;
#pragma unroll 1
    for (int s0 = 0; s0 < DG_STAGES - 1; ++s0) {
        if (s0 < T - 1) dg_issue(ring + (size_t)s0 * DG_ROWS_PER_STEP * 32 + lane, d, s0, Bp, b);
        cp_async_commit();
    }
    int stage = 0;
#pragma unroll 1
    for (int t = 0; t < T - 1; ++t) {
        const int tp = t + DG_STAGES - 1;
        int ps = stage + DG_STAGES - 1;
        if (ps >= DG_STAGES) ps -= DG_STAGES;
        if (tp < T - 1) dg_issue(ring + (size_t)ps * DG_ROWS_PER_STEP * 32 + lane, d, tp, Bp, b);
        cp_async_commit();
        cp_async_wait<DG_STAGES - 1>(); /* this lane's copies for step t have landed */
        double Kt[d1(M * N)], kt[d1(M)], fx[N * N], fu[d1(N * M)], Lx[N], Lu[d1(M)];
        const double* q = ring + (size_t)stage * DG_ROWS_PER_STEP * 32 + lane;
        lds_rows<M * N>(Kt, q); lds_rows<M>(kt, q); lds_rows<N * N>(fx, q); lds_rows<N * M>(fu, q);
        lds_rows<N>(Lx, q); lds_rows<M>(Lu, q);
        if (++stage == DG_STAGES) stage = 0;
#pragma unroll
        for (int a = 0; a < M; ++a) zu[a] = kt[a] + dotf<N, M, 1>(Kt + a, zx);
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const double v = dotf<M, N, 1>(fu + i, zu);
            zy[i] = v + dotf<N, N, 1>(fx + i, zx);
        }
#pragma unroll
        for (int i = 0; i < N; ++i) sx = ilqr_fma(Lx[i], zx[i], sx);
#pragma unroll
        for (int a = 0; a < M; ++a) su = ilqr_fma(Lu[a], zu[a], su);
#pragma unroll
        for (int i = 0; i < N; ++i) zx[i] = zy[i];
    }
    cp_async_wait<0>();
    return sx + su;
}
#else
#include "ilqr_large_forward.cuh"
#endif

/* ==================================================================================== */
/* k_forward: grid = Bp/32 blocks, block = 32 problems x (FWD_TRIAL_WARPS trial warps + 2 aux warps).
 * One launch evaluates ONE round of every problem's line search: step sizes 2^-base ... 2^-(base+FWD_TRIAL_WARPS-1)
 * concurrently, one per trial warp, base being the problem's own position in its search; a problem whose round
 * fails continues at the next launch (and sits out the backward pass in between).  Measured on the acrobot
 * batch, 98.3% of the iterations accept the full step and the rest 1/2, 1/4 or 1/8, so every trial warp beyond
 * the first is speculation that is thrown away 98% of the time -- it costs FP64 issue slots and HBM traffic that
 * other CTAs on the SM could use.  Two trials per launch measured best (B200, 8192 slots: 71.7 k solves/s
 * against 61.4 k with four and 67.1 k with one, where a search that fails outright takes 17 launches).
 * Trial warp w writes its rollout into slot w (slot 0 = the problem's canonical current trajectory);
 * after the selection all warps copy the winning slot into the nominal (if accepted) and canonical
 * current buffers.  The aux warp computes the expected-decrease term of the Armijo test meanwhile (first
 * round only), or does the between-inner-solves bookkeeping for problems in that phase. */
constexpr int COPY_BATCH = 16;

/* rows first, first+stride, ... < count of column b: src -> dst1 and/or dst2 */
template <typename TV>
__device__ __forceinline__ void copy_rows(const TV* __restrict__ src, TV* __restrict__ dst1, TV* __restrict__ dst2, int count,
                                          int first, int stride, size_t Bp, int b) {
    for (int r0 = first; r0 < count; r0 += stride * COPY_BATCH) {
        TV v[COPY_BATCH];
#pragma unroll
        for (int j = 0; j < COPY_BATCH; ++j) {
            const int r = r0 + j * stride;
            if (r < count) v[j] = src[r * Bp + b];
        }
#pragma unroll
        for (int j = 0; j < COPY_BATCH; ++j) {
            const int r = r0 + j * stride;
            if (r < count) {
                if (dst1) dst1[r * Bp + b] = v[j];
                if (dst2) dst2[r * Bp + b] = v[j];
            }
        }
    }
}

/* Two instantiations: MINCTAS = 1 lets ptxas use the registers it wants (192 for the acrobot) -- best while the
 * grid is small and a launch is pure latency; MINCTAS = FWD_DENSE_CTAS = 3 caps it at 168 registers so that three
 * CTAs (12 warps) share an SM -- best once the grid exceeds two CTAs per SM (the acrobot fits without spilling, the
 * car spills 16 bytes).  Measured, acrobot: 14208 = 3 x 148 x 32 slots give 78.8 k solves/s against 73.0 k at 8192
 * slots; car at 2048 slots: 152 k solves/s uncapped against 122 k capped.  The engine picks per launch. */
#ifndef ILQR_FWD_MIN_CTAS
#if ILQR_LARGE
#define ILQR_FWD_MIN_CTAS 1
#else
#define ILQR_FWD_MIN_CTAS 3
#endif
#endif
/* What follows the rollouts of a k_forward / k_forward_tma launch: the first step size, in descending order, that passes
 * the Armijo test (src/forward_pass.jl:28-54), then update_nominal_trajectory! (src/data/methods.jl:32-39) and the
 * solver scalars.  Only trial warp 0 writes its trajectory (into the problem's canonical current buffers); the other
 * trial warps are SPECULATION that only reports its cost: when the deciding trial -- the one that is accepted, or the
 * last one of an exhausted search, whose trajectory becomes `current` (Q2) -- sits in a speculative slot (1.5 % of the
 * acrobot iterations), the problem keeps its state and evaluates that step size again as trial 0 at the next launch.
 * One tick more for those, no trajectory / constraint rows written and no slot copied for everybody else. */
template <int NWc, int NW>
__device__ __forceinline__ void forward_finish(const Params& P, int b, int lane, int wid, bool iter, int base,
                                               const double (*sJ)[32], const double (*sV)[32], const double* sDgp) {
    const Dev& d = P.d;
    const size_t Bp = P.Bp;
    const int n_alpha = P.n_alpha;
    int win = -1, wwin = -1;
    bool accepted = false, nonfinite = false;
    const bool open_ls = iter && base < n_alpha; /* this problem has trials to evaluate */
    double Jwin = 0.0, Vwin = 0.0;
    const double Jp = iter ? d.J[b] : 0.0;
    if (open_ls) {
        const double dgp = sDgp[lane];
        for (int w = 0; w < NWc && base + w < n_alpha; ++w) {
            const int c = base + w;
            const double Jc = sJ[w][lane];
            if (!(Jc - Jc == 0.0)) nonfinite = true;
            win = c;
            wwin = w;
            Jwin = Jc;
            Vwin = sV[w][lane];
            if (Jc <= Jp + (1.0e-4 * pow2neg(c)) * dgp) { accepted = true; break; }
        }
    }
    const bool more_rounds = open_ls && !accepted && base + NWc < n_alpha; /* next round at the next launch */
    const bool redo = open_ls && !more_rounds && wwin > 0;                  /* the deciding trial again, as trial 0 */
    if (more_rounds || redo) {
        if (wid == 0) {
            d.ls_base[b] = redo ? base + wwin : base + NWc;
            if (nonfinite) d.flags[b] |= ILQR_FLAG_NONFINITE;
            d.kind[b] = KIND_NONE;
        }
        return;
    }
    /* all warps cooperate in the nominal update, each thread moves rows of its own problem (coalesced across the warp),
     * COPY_BATCH independent loads in flight before the first store so the copy is not latency-serialised */
    if (iter && accepted) {
        copy_rows(d.xc, d.xb, (double*)nullptr, P.T * N, wid, NW, Bp, b);
        copy_rows(d.uc, d.ub, (double*)nullptr, (P.T - 1) * M, wid, NW, Bp, b);
    }
    if (wid == 0 && iter) {
        if (n_alpha > 0) {
            d.J[b] = Jwin;                                   /* data.objective[1]: src/data/methods.jl:19 */
            if (CONSTRAINED) d.viol[b] = Vwin;
        }
        d.alpha[b] = accepted ? pow2neg(win) : pow2neg(n_alpha); /* src/forward_pass.jl:26,51 */
        d.status[b] = accepted ? 1 : 0;
        if (nonfinite) d.flags[b] |= ILQR_FLAG_NONFINITE;
        d.ls_base[b] = 0;
        d.kind[b] = KIND_ITER;
    }
}

constexpr int FWD_DENSE_CTAS = ILQR_FWD_MIN_CTAS;
template <int MINCTAS>
__global__ void __launch_bounds__(32 * (FWD_TRIAL_WARPS + 2), MINCTAS) k_forward(const __grid_constant__ Params P) {
    extern __shared__ __align__(16) double dg_ring[]; /* the aux warp's cp.async ring (DG_SMEM_BYTES) */
    __shared__ double sJ[FWD_TRIAL_WARPS][32];
    __shared__ double sV[FWD_TRIAL_WARPS][32];
    __shared__ double sDgp[32];
    const Dev& d = P.d;
    const int lane = threadIdx.x, wid = threadIdx.y;
    constexpr int NWc = FWD_TRIAL_WARPS, NW = FWD_TRIAL_WARPS + 2;
    const int n_alpha = P.n_alpha;
    const int b = blockIdx.x * 32 + lane;
    int phase = d.phase[b];
    /* Refilled by the k_refill of two ticks ago: start it now.  k_refill(t) runs on the side branch CONCURRENTLY with
     * the kernels of tick t+1 and is joined before k_forward(t+2); it posts the tick it is meant for as the token, so
     * the k_forward of tick t+1, which may or may not see the store, ignores it either way (deterministic start, and
     * nothing of the refilled slot is read before the join has ordered it). */
    const bool start_now = P.mode == MODE_STREAM && d.pending[b] == 1 + (P.tick & 7);
    if (start_now) phase = PH_START;
    const bool iter = phase == PH_ITER;

    if (blockIdx.x == 0 && wid == 0 && lane == 0) {
        d.active[(P.tick + 4) & 7] = 0;
        if (P.mode == MODE_STREAM) d.done_count[(P.tick + 2) & 3] = 0; /* last read by the k_refill of tick-2, complete by now */
    }

    /* One ROUND of the line search per launch: step sizes 2^-base ... 2^-(base+NWc-1), one per trial warp, where
     * base is the problem's own position in its search (src/forward_pass.jl:28-54 walks them in this order).  A
     * problem none of whose trials passes and that has step sizes left keeps its state, skips this tick's backward
     * pass (KIND_NONE) and continues with the next round at the next launch; nothing waits for it. */
    const int base = iter ? d.ls_base[b] : 0;
    if (wid == NWc) { /* aux warp 1: the expected-decrease term of the Armijo test, once per search */
#ifdef ILQR_TIMING_SKIP_DGP /* timing experiments only: breaks the Armijo test */
        if (iter) sDgp[lane] = 0.0;
#else
        if (iter) {
            double v;
            if (base == 0) {
                v = (P.o.line_search == ILQR_LINE_SEARCH_ARMIJO) ? delta_grad_product(P, b, dg_ring, lane) : 0.0;
                d.dgp[b] = v;
            } else {
                v = d.dgp[b];
            }
            sDgp[lane] = v;
        }
#endif
    } else if (wid == NWc + 1) { /* aux warp 2: problems between two inner solves / two receding-horizon steps */
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

    const bool open_ls = iter && base < n_alpha; /* this problem has trials to evaluate */
    {
        const int c_mine = base + wid;
        if (wid < NWc && open_ls && c_mine < n_alpha) {
            TrialOut o; /* trial 0 writes the canonical current trajectory; the speculative trials write nothing */
            if (wid == 0) { o.x = d.xc; o.u = d.uc; o.c = d.c; o.a = d.act; }
            else { o.x = nullptr; o.u = nullptr; o.c = nullptr; o.a = nullptr; }
            double J, mv;
            rollout_eval(P, o, b, pow2neg(c_mine), J, mv, dg_ring + DG_SMEM_BYTES / 8 + (size_t)wid * PR_WARP_DOUBLES + lane);
            sJ[wid][lane] = J;
            sV[wid][lane] = mv;
        }
        __syncthreads();
    }
    forward_finish<NWc, NW>(P, b, lane, wid, iter, base, sJ, sV, sDgp);
}

#if !ILQR_LARGE
/* ==================================================================================== */
/* k_forward_tp ("thread per problem"): the forward half of a tick for DENSE grids.  One warp = 32 problems, no
 * warp specialisation.  A lane in the first round of its line search runs ONE loop over the horizon that carries both
 * the full-step linearised response of src/data/methods.jl:42-54 (for the Armijo term, src/forward_pass.jl:19-20) and
 * the rollout of the step size under test (src/rollout.jl:19-29): the gains K_t, k_t are fetched once for the two, and
 * one step size is evaluated per launch (98 % of the acrobot iterations accept the first) -- k_forward's second,
 * speculative trial warp and its separate expected-decrease warp re-read the same rows and write a trajectory that
 * is thrown away 98 % of the time, which at three CTAs per SM made it DRAM-bound on 4.5x its algorithmic bytes.
 * Rows moved per (problem, step), acrobot: 35 read + 5 written + 10 for the nominal update, against 70.
 * All of a step's rows stream through the warp's own FT_STAGES-deep cp.async ring (each lane copies and reads back its
 * own column: no barrier), so the loads of step t + FT_STAGES - 1 fly while step t computes. */
#ifndef ILQR_FT_WARPS_PER_SM
#define ILQR_FT_WARPS_PER_SM 8
#endif
constexpr int FT_WARPS_PER_SM = ILQR_FT_WARPS_PER_SM;
constexpr int FT_DG_ROWS = N * N + N * M + N + M;                /* fx, fu, Lx, Lu */
constexpr int FT_ROWS = PR_ROWS_PER_STEP + FT_DG_ROWS;           /* + K, k, ub, xb, lam, rho, w */
constexpr int FT_STAGE_BYTES = FT_ROWS * 32 * 8;
constexpr int FT_STAGES_FIT = (227 * 1024 / FT_WARPS_PER_SM - 1024) / FT_STAGE_BYTES;
constexpr int FT_STAGES = FT_STAGES_FIT >= 4 ? 4 : FT_STAGES_FIT;
constexpr bool FT_OK = FT_STAGES >= 2;
constexpr int FT_ST = FT_OK ? FT_STAGES : 2;
constexpr int FT_SMEM_BYTES = FT_ST * FT_STAGE_BYTES;

__device__ __forceinline__ void ft_issue(double* stage_lane, const Dev& d, int t, int Bp, int b, bool with_dg) {
    double* p = stage_lane;
    cp_rows<M * N>(p, d.K, (size_t)t * M * N, Bp, b);
    cp_rows<M>(p, d.k, (size_t)t * M, Bp, b);
    cp_rows<M>(p, d.ub, (size_t)t * M, Bp, b);
    cp_rows<N>(p, d.xb, (size_t)t * N, Bp, b);
    cp_rows<CS>(p, d.lam, (size_t)t * CS, Bp, b);
    cp_rows<CS>(p, d.rho, (size_t)t * CS, Bp, b);
    cp_rows<NP>(p, d.w, (size_t)t * NP, Bp, b);
    if (with_dg) {
        cp_rows<N * N>(p, d.fx, (size_t)t * N * N, Bp, b);
        cp_rows<N * M>(p, d.fu, (size_t)t * N * M, Bp, b);
        cp_rows<N>(p, d.Lx, (size_t)t * N, Bp, b);
        cp_rows<M>(p, d.Lu, (size_t)t * M, Bp, b);
    }
}

/* one step of trajectory_sensitivities + gradient' * trajectory (src/data/methods.jl:46-52, src/forward_pass.jl:20) */
__device__ __forceinline__ void dgp_step(const double* Kt, const double* kt, const double* fx, const double* fu, const double* Lx,
                                         const double* Lu, double* zx, double& sx, double& su) {
    double zy[N], zu[d1(M)];
#pragma unroll
    for (int a = 0; a < M; ++a) zu[a] = kt[a] + dotf<N, M, 1>(Kt + a, zx);
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const double v = dotf<M, N, 1>(fu + i, zu);
        zy[i] = v + dotf<N, N, 1>(fx + i, zx);
    }
#pragma unroll
    for (int i = 0; i < N; ++i) sx = ilqr_fma(Lx[i], zx[i], sx);
#pragma unroll
    for (int a = 0; a < M; ++a) su = ilqr_fma(Lu[a], zu[a], su);
#pragma unroll
    for (int i = 0; i < N; ++i) zx[i] = zy[i];
}

/* rollout_eval + delta_grad_product in one sweep; same statements, same order per quantity as the two functions */
__device__ __forceinline__ void rollout_dgp_eval(const Params& P, const TrialOut& o, int b, double alpha, bool with_dg,
                                                 double& J_out, double& viol_out, double& dgp_out, double* ring_lane) {
    const Dev& d = P.d;
    const int Bp = P.Bp, T = P.T;
    constexpr int ST = FT_ST;
    double x[N], u[d1(M)], wv[d1(NP)], zx[N];
    double sx = 0.0, su = 0.0;
    RollAcc acc;
    ld_rows<N>(x, d.xb, 0, Bp, b);
#pragma unroll
    for (int i = 0; i < N; ++i) zx[i] = 0.0;
#pragma unroll 1
    for (int s0 = 0; s0 < ST - 1; ++s0) {
        if (s0 < T - 1) ft_issue(ring_lane + (size_t)s0 * FT_ROWS * 32, d, s0, Bp, b, with_dg);
        cp_async_commit();
    }
    double lamT[d1(CT)], rhoT[d1(CT)];
    ld_rows<CT>(lamT, d.lam, (size_t)(T - 1) * CS, Bp, b);
    ld_rows<CT>(rhoT, d.rho, (size_t)(T - 1) * CS, Bp, b);
    int stage = 0;
#pragma unroll 1
    for (int t = 0; t < T - 1; ++t) {
        const int tp = t + ST - 1;
        int ps = stage + ST - 1;
        if (ps >= ST) ps -= ST;
        if (tp < T - 1) ft_issue(ring_lane + (size_t)ps * FT_ROWS * 32, d, tp, Bp, b, with_dg);
        cp_async_commit();
        cp_async_wait<ST - 1>(); /* this lane's copies for step t have landed */
        PolicyRow cur;
        const double* q = ring_lane + (size_t)stage * FT_ROWS * 32;
        lds_rows<M * N>(cur.Kt, q); lds_rows<M>(cur.kt, q); lds_rows<M>(cur.ubt, q); lds_rows<N>(cur.xbt, q);
        lds_rows<CS>(cur.lam, q); lds_rows<CS>(cur.rho, q); lds_rows<NP>(wv, q);
        if (++stage == ST) stage = 0;
        if (with_dg) {
            double fx[N * N], fu[d1(N * M)], Lx[N], Lu[d1(M)];
            lds_rows<N * N>(fx, q); lds_rows<N * M>(fu, q); lds_rows<N>(Lx, q); lds_rows<M>(Lu, q);
            dgp_step(cur.Kt, cur.kt, fx, fu, Lx, Lu, zx, sx, su);
        }
        rollout_step(o, t, Bp, b, alpha, cur, wv, x, u, acc);
    }
    cp_async_wait<0>();
    rollout_terminal(P, o, b, lamT, rhoT, x, u, acc);
    J_out = CONSTRAINED ? (acc.Jc + acc.Jal) : acc.Jc;
    viol_out = acc.mv;
    dgp_out = sx + su;
}

__global__ void __launch_bounds__(32, FT_WARPS_PER_SM) k_forward_tp(const __grid_constant__ Params P) {
    extern __shared__ __align__(16) double ft_ring[];
    const Dev& d = P.d;
    const int lane = threadIdx.x;
    const int b = blockIdx.x * 32 + lane;
    const int n_alpha = P.n_alpha;
    const size_t Bp = P.Bp;
    int phase = d.phase[b];
    const bool start_now = P.mode == MODE_STREAM && d.pending[b] == 1 + (P.tick & 7); /* see k_forward */
    if (start_now) phase = PH_START;
    const bool iter = phase == PH_ITER;
    if (blockIdx.x == 0 && lane == 0) {
        d.active[(P.tick + 4) & 7] = 0;
        if (P.mode == MODE_STREAM) d.done_count[(P.tick + 2) & 3] = 0;
    }
    /* problems between two inner solves / two receding-horizon steps (the lanes of a warp diverge here) */
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
    __syncwarp();
    if (!iter) return;
    /* one step size of the line search per launch: 2^-base, base = the problem's position in its search
     * (src/forward_pass.jl:28-54 walks them in this order) */
    const int base = d.ls_base[b];
    const bool open_ls = base < n_alpha;
    bool accepted = false, nonfinite = false;
    double Jt = 0.0, Vt = 0.0;
    if (open_ls) {
        const bool first = base == 0;
        const bool with_dg = first && P.o.line_search == ILQR_LINE_SEARCH_ARMIJO;
        TrialOut o;
        o.x = d.xc; o.u = d.uc; o.c = d.c; o.a = d.act;
        double dgp;
        rollout_dgp_eval(P, o, b, pow2neg(base), with_dg, Jt, Vt, dgp, ft_ring + lane);
        if (first) d.dgp[b] = dgp; /* 0 when the line search is off */
        else dgp = d.dgp[b];
        const double Jp = d.J[b];
        if (!(Jt - Jt == 0.0)) nonfinite = true;
        accepted = Jt <= Jp + (1.0e-4 * pow2neg(base)) * dgp;
        if (!accepted && base + 1 < n_alpha) { /* next step size at the next launch; no backward pass in between */
            d.ls_base[b] = base + 1;
            if (nonfinite) d.flags[b] |= ILQR_FLAG_NONFINITE;
            d.kind[b] = KIND_NONE;
            return;
        }
        if (accepted) { /* update_nominal_trajectory! (src/data/methods.jl:32-39) */
            copy_rows(d.xc, d.xb, (double*)nullptr, P.T * N, 0, 1, Bp, b);
            copy_rows(d.uc, d.ub, (double*)nullptr, (P.T - 1) * M, 0, 1, Bp, b);
        }
        d.J[b] = Jt;                                         /* data.objective[1]: src/data/methods.jl:19 */
        if (CONSTRAINED) d.viol[b] = Vt;
    }
    d.alpha[b] = accepted ? pow2neg(base) : pow2neg(n_alpha); /* src/forward_pass.jl:26,51 */
    d.status[b] = accepted ? 1 : 0;
    if (nonfinite) d.flags[b] |= ILQR_FLAG_NONFINITE;
    d.ls_base[b] = 0;
    d.kind[b] = KIND_ITER;
}
#endif /* !ILQR_LARGE */

#if !ILQR_LARGE
/* ==================================================================================== */
/* k_forward_tma: k_forward with ONE shared-memory ring per CTA, filled by TMA.
 * k_forward's three working warps (two line-search trials + the expected-decrease warp) each stream their own copy of
 * the step's gains K_t, k_t (and the trials of x̄_t, ū_t) through private cp.async rings: 50 rows of 256 bytes per
 * (32 problems, time step) for 35 distinct ones, 8-byte LDGSTS per lane -- on dense grids the kernel is DRAM-bound on
 * that traffic.  Here the rows of a step are fetched ONCE: lane 0 of the expected-decrease warp issues one bulk copy
 * (cp.async.bulk.shared.global, the TMA engine; SASS UBLKCP) per 256-byte row into the stage and arms its `full`
 * mbarrier with the byte count (expect_tx); the trial warps and the expected-decrease warp wait on `full`, read their
 * columns, and release the stage through its `empty` mbarrier (one arrival per consumer lane), after which the
 * producer refills it FWT_STAGES - 1 steps ahead.  Consumer warps that have nothing to do this launch (no lane with a
 * step size to try / no lane in the first round of its search) do not take part; the barrier counts are set from the
 * warps that do.  Everything after the rollouts is k_forward's. */
constexpr int FWT_STAGES_FIT = (64 * 1024) / FT_STAGE_BYTES;
constexpr int FWT_STAGES = FWT_STAGES_FIT >= 4 ? 4 : FWT_STAGES_FIT;
constexpr bool FWT_OK = FWT_STAGES >= 2 && FWD_TRIAL_WARPS == 2;
constexpr int FWT_ST = FWT_OK ? FWT_STAGES : 2;
constexpr int FWT_SMEM_BYTES = FWT_ST * FT_STAGE_BYTES;

template <int R>
__device__ __forceinline__ void bulk_rows(double*& dst, const double* __restrict__ base, size_t row0, size_t Bp, int b0, uint64_t* bar) {
#pragma unroll
    for (int i = 0; i < R; ++i) { bulk_copy_g2s(dst, base + (row0 + i) * Bp + b0, 256, bar); dst += 32; }
}
/* one elected lane: arm the stage's barrier and fetch step t's rows for the CTA's 32 problems (same row order as ft_issue) */
__device__ __forceinline__ void fwt_issue(double* stage, uint64_t* bar, const Dev& d, int t, size_t Bp, int b0, bool with_dg) {
    mbar_arrive_expect_tx(bar, (unsigned)((PR_ROWS_PER_STEP + (with_dg ? FT_DG_ROWS : 0)) * 256));
    double* p = stage;
    bulk_rows<M * N>(p, d.K, (size_t)t * M * N, Bp, b0, bar);
    bulk_rows<M>(p, d.k, (size_t)t * M, Bp, b0, bar);
    bulk_rows<M>(p, d.ub, (size_t)t * M, Bp, b0, bar);
    bulk_rows<N>(p, d.xb, (size_t)t * N, Bp, b0, bar);
    bulk_rows<CS>(p, d.lam, (size_t)t * CS, Bp, b0, bar);
    bulk_rows<CS>(p, d.rho, (size_t)t * CS, Bp, b0, bar);
    bulk_rows<NP>(p, d.w, (size_t)t * NP, Bp, b0, bar);
    if (with_dg) {
        bulk_rows<N * N>(p, d.fx, (size_t)t * N * N, Bp, b0, bar);
        bulk_rows<N * M>(p, d.fu, (size_t)t * N * M, Bp, b0, bar);
        bulk_rows<N>(p, d.Lx, (size_t)t * N, Bp, b0, bar);
        bulk_rows<M>(p, d.Lu, (size_t)t * M, Bp, b0, bar);
    }
}

/* cp.async flavour of the same ring: every lane of the producer warp copies its own 8-byte column of the stage's rows
 * (LDGSTS) and posts their completion on the stage's `full` mbarrier (cp.async.mbarrier.arrive.noinc, 32 arrivals). */
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}

/* TMA = true: lane 0 feeds the ring with bulk copies (UBLKCP); false: the producer warp's 32 lanes with cp.async (LDGSTS).
 * Measured (profiles/README.md): 35 bulk copies of 256 bytes per step and CTA saturate the SM's TMA unit, three CTAs
 * per SM then run slower than with private rings; the LDGSTS-fed shared ring keeps the traffic saving. */
template <int MINCTAS, bool TMA>
__global__ void __launch_bounds__(32 * (FWD_TRIAL_WARPS + 2), MINCTAS) k_forward_tma(const __grid_constant__ Params P) {
    extern __shared__ __align__(128) double fw_ring[]; /* [FWT_ST][FT_ROWS][32] */
    __shared__ uint64_t full_bar[FWT_ST], empty_bar[FWT_ST];
    __shared__ double sJ[FWD_TRIAL_WARPS][32];
    __shared__ double sV[FWD_TRIAL_WARPS][32];
    __shared__ double sDgp[32];
    __shared__ int s_cons[FWD_TRIAL_WARPS + 1];
    __shared__ unsigned s_need[FWD_TRIAL_WARPS + 1]; /* lanes (problems) whose rows some consumer warp reads */
    const Dev& d = P.d;
    const int lane = threadIdx.x, wid = threadIdx.y;
    constexpr int NWc = FWD_TRIAL_WARPS, NW = FWD_TRIAL_WARPS + 2, ST = FWT_ST;
    const int n_alpha = P.n_alpha;
    const int b0 = blockIdx.x * 32;
    const int b = b0 + lane;
    int phase = d.phase[b];
    const bool start_now = P.mode == MODE_STREAM && d.pending[b] == 1 + (P.tick & 7); /* see k_forward */
    if (start_now) phase = PH_START;
    const bool iter = phase == PH_ITER;
    const size_t Bp = P.Bp;
    const int T = P.T;
    if (blockIdx.x == 0 && wid == 0 && lane == 0) {
        d.active[(P.tick + 4) & 7] = 0;
        if (P.mode == MODE_STREAM) d.done_count[(P.tick + 2) & 3] = 0;
    }
    const int base = iter ? d.ls_base[b] : 0;
    const bool open_ls = iter && base < n_alpha;
    const int c_mine = base + wid;
    const bool armijo = P.o.line_search == ILQR_LINE_SEARCH_ARMIJO;
    /* what this lane has to do in its warp's role */
    const bool trial_lane = wid < NWc && open_ls && c_mine < n_alpha;
    const bool dg_lane = wid == NWc && iter && base == 0 && armijo;
    if (wid <= NWc) {
        const unsigned lanes = __ballot_sync(0xffffffffu, wid < NWc ? trial_lane : dg_lane);
        if (lane == 0) { s_cons[wid] = lanes ? 1 : 0; s_need[wid] = lanes; }
    }
    __syncthreads();
    int ncons = 0;
#pragma unroll
    for (int w = 0; w <= NWc; ++w) ncons += s_cons[w];
    const bool cta_dg = s_cons[NWc] != 0;
    unsigned need_mask = 0;
#pragma unroll
    for (int w = 0; w <= NWc; ++w) need_mask |= s_need[w];
    const bool need = (need_mask >> lane) & 1u; /* cp.async flavour: idle problems' columns are not fetched */
    if (wid == 0 && lane == 0 && ncons > 0) {
        for (int i = 0; i < ST; ++i) { mbar_init(&full_bar[i], TMA ? 1 : 32); mbar_init(&empty_bar[i], 32 * ncons); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (wid == NWc) {
        /* ---- producer + expected-decrease term (src/data/methods.jl:42-54, src/forward_pass.jl:19-20) ---- */
        double v = 0.0;
        if (ncons > 0) {
            double zx[N], sx = 0.0, su = 0.0;
#pragma unroll
            for (int i = 0; i < N; ++i) zx[i] = 0.0;
            if (TMA) {
                if (lane == 0) {
#pragma unroll 1
                    for (int s0 = 0; s0 < ST - 1 && s0 < T - 1; ++s0) fwt_issue(fw_ring + (size_t)s0 * FT_ROWS * 32, &full_bar[s0], d, s0, Bp, b0, cta_dg);
                }
            } else {
#pragma unroll 1
                for (int s0 = 0; s0 < ST - 1 && s0 < T - 1; ++s0) {
                    if (need) ft_issue(fw_ring + (size_t)s0 * FT_ROWS * 32 + lane, d, s0, (int)Bp, b, cta_dg);
                    cp_async_mbar_arrive_noinc(&full_bar[s0]);
                }
            }
#pragma unroll 1
            for (int t = 0; t < T - 1; ++t) {
                const int tp = t + ST - 1;
                if (tp < T - 1) {
                    const int ps = tp % ST;
                    const unsigned use = (unsigned)(tp / ST);
                    if (TMA) {
                        if (lane == 0) {
                            if (use > 0) mbar_wait(&empty_bar[ps], (use - 1) & 1); /* every consumer has read step tp - ST */
                            fwt_issue(fw_ring + (size_t)ps * FT_ROWS * 32, &full_bar[ps], d, tp, Bp, b0, cta_dg);
                        }
                        __syncwarp();
                    } else {
                        if (use > 0) mbar_wait(&empty_bar[ps], (use - 1) & 1);
                        if (need) ft_issue(fw_ring + (size_t)ps * FT_ROWS * 32 + lane, d, tp, (int)Bp, b, cta_dg);
                        cp_async_mbar_arrive_noinc(&full_bar[ps]);
                    }
                }
                if (cta_dg) {
                    const int stage = t % ST;
                    mbar_wait(&full_bar[stage], (unsigned)(t / ST) & 1);
                    if (dg_lane) {
                        double Kt[d1(M * N)], kt[d1(M)], fx[N * N], fu[d1(N * M)], Lx[N], Lu[d1(M)];
                        const double* q = fw_ring + (size_t)stage * FT_ROWS * 32 + lane;
                        lds_rows<M * N>(Kt, q); lds_rows<M>(kt, q);
                        q += (size_t)(PR_ROWS_PER_STEP - M * N - M) * 32;
                        lds_rows<N * N>(fx, q); lds_rows<N * M>(fu, q); lds_rows<N>(Lx, q); lds_rows<M>(Lu, q);
                        dgp_step(Kt, kt, fx, fu, Lx, Lu, zx, sx, su);
                    }
                    mbar_arrive(&empty_bar[stage]);
                }
            }
            v = sx + su;
        }
        if (iter) {
            if (base == 0) {
                if (!armijo) v = 0.0;
                d.dgp[b] = v;
            } else {
                v = d.dgp[b];
            }
            sDgp[lane] = v;
        }
    } else if (wid == NWc + 1) { /* aux warp: problems between two inner solves / two receding-horizon steps */
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
    } else if (s_cons[wid]) {
        /* ---- one line-search trial: rollout! + cost!(mode=:current), rows from the CTA's ring ---- */
        TrialOut o; /* trial 0 writes the canonical current trajectory; the speculative trial writes nothing */
        if (wid == 0) { o.x = d.xc; o.u = d.uc; o.c = d.c; o.a = d.act; }
        else { o.x = nullptr; o.u = nullptr; o.c = nullptr; o.a = nullptr; }
        const double alpha = pow2neg(c_mine);
        double x[N], u[d1(M)], wv[d1(NP)];
        RollAcc acc;
        double lamT[d1(CT)], rhoT[d1(CT)];
        if (trial_lane) {
            ld_rows<N>(x, d.xb, 0, (int)Bp, b);
            ld_rows<CT>(lamT, d.lam, (size_t)(T - 1) * CS, (int)Bp, b);
            ld_rows<CT>(rhoT, d.rho, (size_t)(T - 1) * CS, (int)Bp, b);
        }
#pragma unroll 1
        for (int t = 0; t < T - 1; ++t) {
            const int stage = t % ST;
            mbar_wait(&full_bar[stage], (unsigned)(t / ST) & 1);
            PolicyRow cur;
            if (trial_lane) {
                const double* q = fw_ring + (size_t)stage * FT_ROWS * 32 + lane;
                lds_rows<M * N>(cur.Kt, q); lds_rows<M>(cur.kt, q); lds_rows<M>(cur.ubt, q); lds_rows<N>(cur.xbt, q);
                lds_rows<CS>(cur.lam, q); lds_rows<CS>(cur.rho, q); lds_rows<NP>(wv, q);
            }
            mbar_arrive(&empty_bar[stage]);
            if (trial_lane) rollout_step(o, t, (int)Bp, b, alpha, cur, wv, x, u, acc);
        }
        if (trial_lane) {
            rollout_terminal(P, o, b, lamT, rhoT, x, u, acc);
            sJ[wid][lane] = CONSTRAINED ? (acc.Jc + acc.Jal) : acc.Jc;
            sV[wid][lane] = acc.mv;
        }
    }
    __syncthreads();

    forward_finish<NWc, NW>(P, b, lane, wid, iter, base, sJ, sV, sDgp);
}
#endif /* !ILQR_LARGE */


/* src/solve.jl:36-50 for one problem after its backward pass: iteration counter, the per-iteration
 * record, convergence tests, phase transition.  Returns whether the problem is still running. */
__device__ __forceinline__ bool tick_epilogue(const Params& P, int b, int kind, double gn) {
    const Dev& d = P.d;
    int phase = d.phase[b];
    bool inner_end = false;
    if (kind == KIND_PRELOOP) {
        if (P.o.max_iterations <= 0) inner_end = true;
        else phase = PH_ITER;
    } else if (kind == KIND_ITER) {
        const int it = d.it[b] + 1;
        const int iters = d.iters[b] + 1;                                   /* :39 */
        d.it[b] = it;
        d.iters[b] = iters;
        const double J = d.J[b];
        const int st = d.status[b];
        if (iters - 1 < P.cap) {                                            /* the printout of :40-45 */
            const size_t r = (size_t)(iters - 1) * P.Bp + b;
            d.h_cost[r] = J; d.h_gnorm[r] = gn; d.h_viol[r] = d.viol[b]; d.h_alpha[r] = d.alpha[b];
            d.h_outer[r] = d.outer[b]; d.h_status[r] = (uint8_t)st;
        }
        if (gn < P.o.lagrangian_gradient_tolerance) inner_end = true;       /* :48 */
        else if (fabs(J - d.obj_prev[b]) < P.o.objective_tolerance) inner_end = true; /* :49 */
        else {
            d.obj_prev[b] = J;
            if (!st) inner_end = true;                                      /* :50 */
        }
        if (!inner_end && it >= P.o.max_iterations) inner_end = true;       /* :22 */
    }
    if (inner_end) {
        if (CONSTRAINED) { phase = PH_START; d.inner_done[b] = 1; }
        else phase = mpc_next_phase(P, b);
    }
    if (kind != KIND_NONE) {
        d.phase[b] = phase;
        if (P.mode == MODE_STREAM && phase == PH_DONE) stream_hand_to_refill(P, b);
    }
    if (P.mode == MODE_STREAM && phase == PH_DONE) {
        const int r = d.refilling[b];
        if (r > 0) { d.refilling[b] = r - 1; return true; }
    }
    return phase != PH_DONE && phase != PH_PAUSED;
}

#if !ILQR_LARGE
/* ==================================================================================== */
/* gradients! (src/gradients.jl:92-98) for one (problem, time step).
 * fx, fu, gx, gu are overwritten; gxx, guu, gux are read-modify-written in HBM (Q1: they accumulate
 * over the iterations of one inner solve and restart from zero on its first call, `fresh`).
 * The step's linearisation is returned in registers (StepIn) for the fused Riccati path; STORE_ALL also
 * writes gx, gu to HBM (only the unfused k_backward reads them from there). */
struct StepIn { /* first-order linearisation of one time step */
    double fx[N * N], fu[d1(N * M)], gx[N], gu[d1(M)];
};
struct Hess { /* stage Hessians gxx | guu | gux, packed */
    double v[NH];
    __device__ __forceinline__ double* gxx() { return v; }
    __device__ __forceinline__ double* guu() { return v + N * N; }
    __device__ __forceinline__ double* gux() { return v + N * N + M * M; }
    __device__ __forceinline__ const double* gxx() const { return v; }
    __device__ __forceinline__ const double* guu() const { return v + N * N; }
    __device__ __forceinline__ const double* gux() const { return v + N * N + M * M; }
};
struct StepFull { StepIn s; Hess h; }; /* what a ring stage of the fused kernel carries when the Hessians are per step */
constexpr int SI_ROWS = N * N + N * M + N + M;
static_assert(sizeof(StepIn) == sizeof(double) * SI_ROWS, "StepIn must be packed");
static_assert(sizeof(StepFull) == sizeof(double) * (SI_ROWS + NH), "StepFull must be packed");
static_assert(N * M > 0 && M > 0, "models need at least one action");

struct LinIn { /* everything one (problem, time step) linearisation reads from HBM */
    double x[N], u[d1(M)], wv[d1(NP)];
    Hess h; /* the step's accumulators before this call (untouched with HACC) */
    double c[d1(CS)], lam[d1(CS)], rho[d1(CS)], act[d1(CS)];
};
__device__ __forceinline__ void linearize_load(const Params& P, int b, int t, bool fresh, LinIn& in) {
    const Dev& d = P.d;
    const int Bp = P.Bp;
    ld_rows<N>(in.x, d.xb, (size_t)t * N, Bp, b);
    ld_rows<M>(in.u, d.ub, (size_t)t * M, Bp, b);
    ld_rows<NP>(in.wv, d.w, (size_t)t * NP, Bp, b);
    if (!HACC) {
        if (fresh) {
#pragma unroll
            for (int i = 0; i < NH; ++i) in.h.v[i] = 0.0;
        } else {
            ld_rows<N * N>(in.h.gxx(), d.gxx, (size_t)t * N * N, Bp, b);
            ld_rows<M * M>(in.h.guu(), d.guu, (size_t)t * M * M, Bp, b);
            ld_rows<M * N>(in.h.gux(), d.gux, (size_t)t * M * N, Bp, b);
        }
    }
#if ILQR_CS > 0
    ld_rows<CS>(in.c, d.c, (size_t)t * CS, Bp, b);
    ld_rows<CS>(in.lam, d.lam, (size_t)t * CS, Bp, b);
    ld_rows<CS>(in.rho, d.rho, (size_t)t * CS, Bp, b);
#pragma unroll
    for (int i = 0; i < CS; ++i) in.act[i] = (double)d.act[((size_t)t * CS + i) * Bp + b];
#endif
}

/* HACC: the stage Hessians this tick's gradients! call leaves behind, H = (fresh ? 0 : accumulator) + constants
 * (src/costs.jl:74,79,80 on an accumulator that src/solve.jl:10 zeroed); `store` writes the accumulator back. */
__device__ __forceinline__ void hacc_advance(const Params& P, int b, bool fresh, Hess& H, bool store) {
    const Dev& d = P.d;
    const int Bp = P.Bp;
    double zx[N], zu[d1(M)], zw[d1(NP)], gx[N], gu[d1(M)];
#pragma unroll
    for (int i = 0; i < N; ++i) zx[i] = 0.0;
#pragma unroll
    for (int i = 0; i < d1(M); ++i) zu[i] = 0.0;
#pragma unroll
    for (int i = 0; i < d1(NP); ++i) zw[i] = 0.0;
    Hess c;
    ilqr_cost_s_grad(gx, gu, c.gxx(), c.guu(), c.gux(), zx, zu, zw); /* constants: the arguments do not enter */
    if (fresh) {
#pragma unroll
        for (int i = 0; i < NH; ++i) H.v[i] = 0.0;
    } else {
        ld_rows<NH>(H.v, d.hacc, 0, Bp, b);
    }
#pragma unroll
    for (int i = 0; i < NH; ++i) H.v[i] = H.v[i] + c.v[i];
    if (store) st_rows<NH>(H.v, d.hacc, 0, Bp, b);
}

/* One stage: s = (fx, fu, gx, gu); without HACC also the step's Hessian accumulators h (read-modify-written in HBM). */
template <bool STORE_ALL, bool STORE_F = true>
__device__ __forceinline__ void linearize_compute(const Params& P, int b, int t, const LinIn& in, StepIn& s, Hess& h) {
    const Dev& d = P.d;
    const int Bp = P.Bp;
    const double* x = in.x;
    const double* u = in.u;
    const double* wv = in.wv;
#if ILQR_CS > 0
    const double* c = in.c;
    const double* lam = in.lam;
    const double* rho = in.rho;
    const double* act = in.act;
#endif
    ilqr_dyn_jac(s.fx, s.fu, x, u, wv);                                     /* src/dynamics.jl:41-50 */
    if (STORE_F) { /* (the fused kernels store fx, fu themselves: store_rows_full) */
        st_rows<N * N>(s.fx, d.fx, (size_t)t * N * N, Bp, b);
        st_rows<N * M>(s.fu, d.fu, (size_t)t * N * M, Bp, b);
    }
    Hess hc;
    ilqr_cost_s_grad(s.gx, s.gu, hc.gxx(), hc.guu(), hc.gux(), x, u, wv);   /* src/costs.jl:57-84 */
    if (!HACC) {
#pragma unroll
        for (int i = 0; i < NH; ++i) h.v[i] = in.h.v[i] + hc.v[i];
    }
#if ILQR_CS > 0
    {
        double* gxx = h.gxx();
        double* guu = h.guu();
        double* gux = h.gux();
        double cx[CS * N], cu[CS * M], cxt[CS * N], cut[CS * M], dd[CS], v[CS];
        ilqr_con_s_jac(cx, cu, x, u, wv);                                   /* src/constraints.jl:75-87 */
#pragma unroll
        for (int i = 0; i < CS; ++i) {
            dd[i] = rho[i] * act[i];                                        /* src/gradients.jl:56-58 */
            v[i] = lam[i] + dd[i] * c[i];                                   /* :59-62 */
        }
#pragma unroll
        for (int j = 0; j < N; ++j) s.gx[j] = s.gx[j] + dotf<CS, 1, 1>(cx + j * CS, v);               /* :63 */
#pragma unroll
        for (int j = 0; j < N; ++j)
#pragma unroll
            for (int i = 0; i < CS; ++i) cxt[i + j * CS] = dd[i] * cx[i + j * CS];                    /* :66 */
#pragma unroll
        for (int l = 0; l < N; ++l)
#pragma unroll
            for (int j = 0; j < N; ++j)
                gxx[j + l * N] = gxx[j + l * N] + dotf<CS, 1, 1>(cx + j * CS, cxt + l * CS);          /* :67 */
#pragma unroll
        for (int e = 0; e < M; ++e) s.gu[e] = s.gu[e] + dotf<CS, 1, 1>(cu + e * CS, v);               /* :72 */
#pragma unroll
        for (int e = 0; e < M; ++e)
#pragma unroll
            for (int i = 0; i < CS; ++i) cut[i + e * CS] = dd[i] * cu[i + e * CS];                    /* :75 */
#pragma unroll
        for (int e = 0; e < M; ++e)
#pragma unroll
            for (int a = 0; a < M; ++a)
                guu[a + e * M] = guu[a + e * M] + dotf<CS, 1, 1>(cu + a * CS, cut + e * CS);          /* :76 */
#pragma unroll
        for (int j = 0; j < N; ++j)
#pragma unroll
            for (int a = 0; a < M; ++a)
                gux[a + j * M] = gux[a + j * M] + dotf<CS, 1, 1>(cu + a * CS, cxt + j * CS);          /* :79 */
    }
#endif
    if (!HACC) {
        st_rows<N * N>(h.gxx(), d.gxx, (size_t)t * N * N, Bp, b);
        st_rows<M * M>(h.guu(), d.guu, (size_t)t * M * M, Bp, b);
        st_rows<M * N>(h.gux(), d.gux, (size_t)t * M * N, Bp, b);
    }
    if (STORE_ALL) {
        st_rows<N>(s.gx, d.gx, (size_t)t * N, Bp, b);
        st_rows<M>(s.gu, d.gu, (size_t)t * M, Bp, b);
    }
}

template <bool STORE_ALL>
__device__ __forceinline__ void linearize_stage(const Params& P, int b, int t, bool fresh, StepIn& s, Hess& h) {
    LinIn in;
    linearize_load(P, b, t, fresh, in);
    linearize_compute<STORE_ALL>(P, b, t, in, s, h);
}

/* terminal stage (t = T-1): cost and constraint have no action part (Q13) */
template <bool STORE_ALL>
__device__ __forceinline__ void linearize_terminal(const Params& P, int b, bool fresh, double* gx, double* gxx) {
    const Dev& d = P.d;
    const int Bp = P.Bp, t = P.T - 1;
    double x[N], u[d1(M)], wv[d1(NP)], hxx[N * N];
    ld_rows<N>(x, d.xb, (size_t)t * N, Bp, b);
    ld_rows<NP>(wv, d.w, (size_t)t * NP, Bp, b);
    if (fresh) {
#pragma unroll
        for (int i = 0; i < N * N; ++i) gxx[i] = 0.0;
    } else {
        ld_rows<N * N>(gxx, d.gxx, (size_t)t * N * N, Bp, b);
    }
    ilqr_cost_T_grad(gx, hxx, x, u, wv);
#pragma unroll
    for (int i = 0; i < N * N; ++i) gxx[i] = gxx[i] + hxx[i];
#if ILQR_CT > 0
    {
        double cx[CT * N], cxt[CT * N], dd[CT], v[CT], c[CT], lam[CT], rho[CT];
        ld_rows<CT>(c, d.c, (size_t)t * CS, Bp, b);
        ld_rows<CT>(lam, d.lam, (size_t)t * CS, Bp, b);
        ld_rows<CT>(rho, d.rho, (size_t)t * CS, Bp, b);
        ilqr_con_T_jac(cx, x, u, wv);
#pragma unroll
        for (int i = 0; i < CT; ++i) {
            const double a = (double)d.act[((size_t)t * CS + i) * Bp + b];
            dd[i] = rho[i] * a;
            v[i] = lam[i] + dd[i] * c[i];
        }
#pragma unroll
        for (int j = 0; j < N; ++j) gx[j] = gx[j] + dotf<CT, 1, 1>(cx + j * CT, v);
#pragma unroll
        for (int j = 0; j < N; ++j)
#pragma unroll
            for (int i = 0; i < CT; ++i) cxt[i + j * CT] = dd[i] * cx[i + j * CT];
#pragma unroll
        for (int l = 0; l < N; ++l)
#pragma unroll
            for (int j = 0; j < N; ++j)
                gxx[j + l * N] = gxx[j + l * N] + dotf<CT, 1, 1>(cx + j * CT, cxt + l * CT);
    }
#endif
    st_rows<N * N>(gxx, d.gxx, (size_t)t * N * N, Bp, b);
    if (STORE_ALL) st_rows<N>(gx, d.gx, (size_t)t * N, Bp, b);
}

/* k_linearize (unfused path): one thread per (time step, problem) */
__global__ void __launch_bounds__(128) k_linearize(const __grid_constant__ Params P) {
    const Dev& d = P.d;
    const int Bp = P.Bp, T = P.T;
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int b = (int)(g % Bp);
    const int t = (int)(g / Bp);
    if (t >= T) return;
    const int kind = d.kind[b];
    if (kind == KIND_NONE) return;
    if (kind == KIND_ITER && P.o.line_search == ILQR_LINE_SEARCH_NONE) return; /* src/solve.jl:27 */
    const bool fresh = kind == KIND_PRELOOP; /* reset!(problem.objective): src/solve.jl:10 */
    if (t < T - 1) {
        StepIn s;
        Hess h;
        linearize_stage<true>(P, b, t, fresh, s, h);
    } else {
        double gx[N], gxx[N * N];
        linearize_terminal<true>(P, b, fresh, gx, gxx);
        if (HACC) { /* the terminal thread also advances the problem's stage-Hessian accumulator */
            Hess h;
            hacc_advance(P, b, fresh, h, true);
        }
    }
}

/* ==================================================================================== */
/* upper Cholesky / potrs with LAPACK's stop-at-first-bad-pivot behaviour (Q3); the solves
 * multiply by the reciprocal diagonal (contract, see oracle/ilqr_oracle.c) */
__device__ __forceinline__ bool chol_upper(double* A, double* rinv) {
    bool ok = true;
#pragma unroll
    for (int j = 0; j < M; ++j) {
        if (ok) {
            double ajj = A[j + j * M];
#pragma unroll
            for (int k = 0; k < j; ++k) ajj = ilqr_fma(-A[k + j * M], A[k + j * M], ajj);
            if (!(ajj > 0.0)) {
                A[j + j * M] = ajj;
                ok = false;
            } else {
                const double ujj = sqrt(ajj);
                A[j + j * M] = ujj;
                const double r = 1.0 / ujj;
#pragma unroll
                for (int i = j + 1; i < M; ++i) {
                    double sum = A[j + i * M];
#pragma unroll
                    for (int k = 0; k < j; ++k) sum = ilqr_fma(-A[k + j * M], A[k + i * M], sum);
                    A[j + i * M] = sum * r;
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < M; ++j) rinv[j] = 1.0 / A[j + j * M];
    return ok;
}
__device__ __forceinline__ void chol_solve(const double* U, const double* rinv, double* bv) {
#pragma unroll
    for (int i = 0; i < M; ++i) {
        double sum = bv[i];
#pragma unroll
        for (int k = 0; k < i; ++k) sum = ilqr_fma(-U[k + i * M], bv[k], sum);
        bv[i] = sum * rinv[i];
    }
#pragma unroll
    for (int i = M - 1; i >= 0; --i) {
        double sum = bv[i];
#pragma unroll
        for (int k = i + 1; k < M; ++k) sum = ilqr_fma(-U[i + k * M], bv[k], sum);
        bv[i] = sum * rinv[i];
    }
}

constexpr int BK_ROWS = SI_ROWS + (HACC ? 0 : NH); /* doubles per problem per step handed from linearisation to Riccati */
constexpr int BK_PAIRS = (BK_ROWS + 1) / 2;
constexpr int BK_STAGE_BYTES = BK_PAIRS * 32 * 16;
constexpr int BK_STAGES = (150 * 1024 / BK_STAGE_BYTES) >= 8 ? 8 : (150 * 1024 / BK_STAGE_BYTES);
constexpr bool BK_FUSED = BK_STAGES >= 4; /* models too large for the smem ring use k_linearize + k_backward */

__device__ __forceinline__ void load_step(StepIn& s, Hess& h, const Dev& d, int t, int Bp, int b) {
    ld_rows<N * N>(s.fx, d.fx, (size_t)t * N * N, Bp, b);
    ld_rows<N * M>(s.fu, d.fu, (size_t)t * N * M, Bp, b);
    ld_rows<N>(s.gx, d.gx, (size_t)t * N, Bp, b);
    ld_rows<M>(s.gu, d.gu, (size_t)t * M, Bp, b);
    if (!HACC) { /* with HACC the caller holds the problem's accumulator */
        ld_rows<N * N>(h.gxx(), d.gxx, (size_t)t * N * N, Bp, b);
        ld_rows<M * M>(h.guu(), d.guu, (size_t)t * M * M, Bp, b);
        ld_rows<M * N>(h.gux(), d.gux, (size_t)t * M * N, Bp, b);
    }
}

/* One Riccati step (src/backward_pass.jl:44-89 + src/solve.jl:75-78) in two halves, so that the fused kernel can
 * run them on two warps: the MATRIX half carries the value-function Hessian P (and produces the gain K and the
 * factor of Quu), the VECTOR half carries the value-function gradient p (and produces k, Lx, Lu).  P never
 * depends on p, so the vector half can trail one step behind. */
struct RicHand { /* what the matrix half hands to the vector half */
    double K[d1(M * N)], uxt[d1(M * N)], Qux[d1(M * N)], uu[d1(M * M)], rinv[d1(M)];
};

__device__ __forceinline__ void riccati_matrix_half(const StepIn& s, const Hess& H, double* Pm, RicHand& h, bool& chol_ok) {
    double Qxx[N * N], Quu[d1(M * M)], xxh[N * N], uxh[d1(M * N)];
#pragma unroll
    for (int l = 0; l < N; ++l)
#pragma unroll
        for (int i = 0; i < N; ++i) xxh[i + l * N] = dotf<N, 1, 1>(s.fx + i * N, Pm + l * N);       /* :52 */
#pragma unroll
    for (int j = 0; j < N; ++j)
#pragma unroll
        for (int i = 0; i < N; ++i)
            Qxx[i + j * N] = dotf<N, N, 1>(xxh + i, s.fx + j * N) + H.gxx()[i + j * N];               /* :53-54 */
#pragma unroll
    for (int l = 0; l < N; ++l)
#pragma unroll
        for (int a = 0; a < M; ++a) uxh[a + l * M] = dotf<N, 1, 1>(s.fu + a * N, Pm + l * N);       /* :57 (= :62) */
#pragma unroll
    for (int e = 0; e < M; ++e)
#pragma unroll
        for (int a = 0; a < M; ++a)
            Quu[a + e * M] = dotf<N, M, 1>(uxh + a, s.fu + e * N) + H.guu()[a + e * M];               /* :58-59 */
#pragma unroll
    for (int j = 0; j < N; ++j)
#pragma unroll
        for (int a = 0; a < M; ++a)
            h.Qux[a + j * M] = dotf<N, M, 1>(uxh + a, s.fx + j * N) + H.gux()[a + j * M];             /* :63-64 */
#pragma unroll
    for (int i = 0; i < M * M; ++i) h.uu[i] = Quu[i];                                               /* :68 */
    if (!chol_upper(h.uu, h.rinv)) chol_ok = false;                                                 /* :69 */
#pragma unroll
    for (int j = 0; j < N; ++j) {                                                                   /* :70,72,74 */
        double col[d1(M)];
#pragma unroll
        for (int a = 0; a < M; ++a) col[a] = h.Qux[a + j * M];
        chol_solve(h.uu, h.rinv, col);
#pragma unroll
        for (int a = 0; a < M; ++a) h.K[a + j * M] = -col[a];
    }
#pragma unroll
    for (int j = 0; j < N; ++j)
#pragma unroll
        for (int a = 0; a < M; ++a) h.uxt[a + j * M] = dotf<M, M, 1>(Quu + a, h.K + j * M);         /* :79 */
#pragma unroll
    for (int j = 0; j < N; ++j)
#pragma unroll
        for (int i = 0; i < N; ++i) {
            double v = dotf<M, 1, 1>(h.K + i * M, h.uxt + j * M);                                   /* :81 */
            v = v + dotf<M, 1, 1>(h.K + i * M, h.Qux + j * M);                                      /* :82 */
            v = v + dotf<M, 1, 1>(h.Qux + i * M, h.K + j * M);                                      /* :83 */
            Pm[i + j * N] = v + Qxx[i + j * N];                                                     /* :84 */
        }
}

__device__ __forceinline__ void riccati_vector_half(const StepIn& s, const RicHand& h, double* pv, double* kk, double* Lx,
                                                    double* Qu, double& gn) {
    double Qx[N];
#pragma unroll
    for (int i = 0; i < N; ++i) Qx[i] = dotf<N, 1, 1>(s.fx + i * N, pv) + s.gx[i];                  /* :44-45 */
#pragma unroll
    for (int a = 0; a < M; ++a) Qu[a] = dotf<N, 1, 1>(s.fu + a * N, pv) + s.gu[a];                  /* :48-49 */
    {                                                                                               /* :71,73,75 */
        double col[d1(M)];
#pragma unroll
        for (int a = 0; a < M; ++a) col[a] = Qu[a];
        chol_solve(h.uu, h.rinv, col);
#pragma unroll
        for (int a = 0; a < M; ++a) kk[a] = -col[a];
    }
#pragma unroll
    for (int i = 0; i < N; ++i) {
        double v = dotf<M, 1, 1>(h.uxt + i * M, kk);                                                /* :86 */
        v = v + dotf<M, 1, 1>(h.K + i * M, Qu);                                                     /* :87 */
        v = v + dotf<M, 1, 1>(h.Qux + i * M, kk);                                                   /* :88 */
        pv[i] = v + Qx[i];                                                                          /* :89 */
        Lx[i] = Qx[i] - pv[i];                                                                      /* src/solve.jl:75-76 */
        const double a = fabs(Lx[i]);
        if (a > gn || a != a) gn = a;
    }
#pragma unroll
    for (int a = 0; a < M; ++a) {
        const double v = fabs(Qu[a]);                                                               /* src/solve.jl:78 */
        if (v > gn || v != v) gn = v;
    }
}

__device__ __forceinline__ void riccati_step(const StepIn& s, const Hess& H, double* Pm, double* pv, double* K, double* kk,
                                             double* Lx, double* Qu, bool& chol_ok, double& gn) {
    RicHand h;
    riccati_matrix_half(s, H, Pm, h, chol_ok);
    riccati_vector_half(s, h, pv, kk, Lx, Qu, gn);
#pragma unroll
    for (int i = 0; i < M * N; ++i) K[i] = h.K[i];
}

/* k_backward (unfused path): one thread per problem reading k_linearize's output from HBM */
__global__ void __launch_bounds__(32) k_backward(const __grid_constant__ Params P) {
    const Dev& d = P.d;
    const int Bp = P.Bp, T = P.T;
    const int b = blockIdx.x * 32 + threadIdx.x;
    const int kind = d.kind[b];
    const bool skip_ls_none = (kind == KIND_ITER && P.o.line_search == ILQR_LINE_SEARCH_NONE);
    double gn = 0.0;
    if (kind != KIND_NONE && !skip_ls_none) {
        double Pm[N * N], pv[N];
        bool chol_ok = true;
        ld_rows<N * N>(Pm, d.gxx, (size_t)(T - 1) * N * N, Bp, b);            /* src/backward_pass.jl:39 */
        ld_rows<N>(pv, d.gx, (size_t)(T - 1) * N, Bp, b);                     /* :40 */
        Hess H;
        if (HACC) ld_rows<NH>(H.v, d.hacc, 0, Bp, b); /* advanced by k_linearize's terminal thread of this tick */
#pragma unroll 1
        for (int t = T - 2; t >= 0; --t) {
            StepIn s;
            load_step(s, H, d, t, Bp, b);
            double K[d1(M * N)], kk[d1(M)], Lx[N], Qu[d1(M)];
            riccati_step(s, H, Pm, pv, K, kk, Lx, Qu, chol_ok, gn);
            st_rows<M * N>(K, d.K, (size_t)t * M * N, Bp, b);
            st_rows<M>(kk, d.k, (size_t)t * M, Bp, b);
            st_rows<N>(Lx, d.Lx, (size_t)t * N, Bp, b);
            st_rows<M>(Qu, d.Lu, (size_t)t * M, Bp, b);
        }
        if (!chol_ok) d.flags[b] |= ILQR_FLAG_CHOL_FAIL;
        d.gnorm[b] = gn;
    } else if (skip_ls_none) {
        gn = d.gnorm[b];
    }
    const bool running = tick_epilogue(P, b, kind, gn);
    const unsigned mask = __ballot_sync(0xffffffffu, running);
    if (threadIdx.x == 0 && mask) atomicAdd(&d.active[P.tick & 7], __popc(mask));
}

/* k_linback (fused path): gradients! + backward_pass! + lagrangian_gradient! + the convergence tests in ONE
 * kernel.  CTA = 32 problems x 8 warps:
 *   warp 0        Riccati MATRIX warp: carries P in registers (xxh, Qxx, Quu, Qux, Cholesky, K, P), alone on its
 *                 SM sub-partition (warp 4 exits at once: measured -8 %)
 *   warp 1        Riccati VECTOR warp: carries p one step behind warp 0 (Qx, Qu, k, p, Lx, Lu, gradient norm, the
 *                 stores of K, k, Lx, Lu) and closes the tick with the bookkeeping of src/solve.jl:36-50
 *   warps 2,3,5-7 linearisation producers: time steps round-robin into a BK_STAGES-deep shared-memory ring
 * full/empty mbarriers guard the ring (both Riccati warps read a stage); a second, LB_HAND-deep ring hands
 * (K, Quu K, Qux, chol(Quu)) from warp 0 to warp 1.  The linearisation never makes the HBM round trip of the
 * unfused pair (only fx, fu -- needed by the next forward pass -- and the Hessian accumulators of Q1 are
 * written) and its latency hides under the sequential recursion. */
#ifdef ILQR_LB_KEEP_WARP4 /* experiment: no idle warp, the matrix warp shares its sub-partition with a producer */
constexpr bool LB_IDLE_WARPS = false;
#else
constexpr bool LB_IDLE_WARPS = true;
#endif
constexpr int HAND_DOUBLES = 3 * M * N + M * M + M; /* RicHand, packed */
constexpr int HAND_PAIRS = (HAND_DOUBLES + 1) / 2;
static_assert(sizeof(RicHand) == sizeof(double) * (3 * d1(M * N) + d1(M * M) + d1(M)), "RicHand must be packed");
/* Ring geometry for a linearisation ring of ST stages.  Hand ring depth: with ST + 2 slots (one of them reserved for
 * p_T) the matrix warp can never catch up with a slot the vector warp still reads: stage s of the linearisation ring is
 * only refilled after BOTH Riccati warps have released step s - ST, so when the matrix warp starts step s the vector
 * warp has finished step s - ST - 1 -- the slot step s overwrites.  No "empty" handshake is needed then, which takes
 * one ~100-cycle mbarrier wait out of every step of the critical warp.  Falls back to 4 slots + handshake if that does
 * not fit. */
template <int ST>
struct LbGeom {
    static constexpr int HAND_DEEP = ST + 2;
    static constexpr bool HAND_FREE = ST * BK_STAGE_BYTES + HAND_DEEP * HAND_PAIRS * 32 * 16 <= 220 * 1024;
    static constexpr int HAND = HAND_FREE ? HAND_DEEP : 4;
    static constexpr int SMEM_BYTES = ST * BK_STAGE_BYTES + HAND * HAND_PAIRS * 32 * 16;
};
/* Two instantiations of k_linback.  LATENCY: 8 warps (5 producers), one CTA per SM, the full ring -- best while the grid
 * has at most one CTA per SM and a tick is pure latency.  DENSE: 6 warps (3 producers), 168 registers, TWO CTAs per SM
 * (two Riccati matrix warps per SM instead of one; a shallower ring if two full ones do not fit) -- measured on B200,
 * acrobot, every slot busy: 10.2-11.3 ns per problem and tick from 9472 slots up against 13.2-14.1 for the 8-warp
 * kernel (profiles/README.md). */
constexpr int LB_WARPS = 8, LB_DENSE_WARPS = 6;
constexpr int LB_STAGES = BK_STAGES;
constexpr int LB_DENSE_STAGES = (2 * (LbGeom<d1(BK_STAGES)>::SMEM_BYTES + 2048) <= 227 * 1024) ? BK_STAGES : (BK_STAGES > 4 ? 4 : BK_STAGES);
/* only for the light steps of HACC models: capped at 168 registers the car (per-step Hessians + AL terms) spills 600 bytes
 * per thread and loses (measured, 9472 slots: 312 k solves/s against 366 k with the 8-warp kernel) */
constexpr bool LB_DENSE_OK = BK_FUSED && HACC && 2 * (LbGeom<d1(LB_DENSE_STAGES)>::SMEM_BYTES + 2048) <= 227 * 1024;

template <int PAIRS, int COUNT>
__device__ __forceinline__ void lane_write(double* base_lane, const double* v) {
    double2* p = reinterpret_cast<double2*>(base_lane);
#pragma unroll
    for (int q = 0; q < PAIRS; ++q) {
        double2 t;
        t.x = v[2 * q];
        t.y = (2 * q + 1 < COUNT) ? v[2 * q + 1] : 0.0;
        p[q * 32] = t;
    }
}
template <int PAIRS, int COUNT>
__device__ __forceinline__ void lane_read(double* v, const double* base_lane) {
    const double2* p = reinterpret_cast<const double2*>(base_lane);
#pragma unroll
    for (int q = 0; q < PAIRS; ++q) {
        const double2 t = p[q * 32];
        v[2 * q] = t.x;
        if (2 * q + 1 < COUNT) v[2 * q + 1] = t.y;
    }
}

template <int LBW, int MINCTAS, int ST>
__global__ void __launch_bounds__(32 * LBW, MINCTAS) k_linback(const __grid_constant__ Params P) {
    constexpr int NPROD = LBW - 2 - (LB_IDLE_WARPS ? (LBW - 1) / 4 : 0); /* warps 4, 8 (the matrix warp's sub-partition) stay idle */
    constexpr int LB_HAND = LbGeom<ST>::HAND;
    constexpr bool LB_HAND_FREE = LbGeom<ST>::HAND_FREE;
    extern __shared__ __align__(16) double ring[];
    __shared__ uint64_t full_bar[ST], empty_bar[ST];
    __shared__ uint64_t hfull_bar[LB_HAND], hempty_bar[LB_HAND];
    __shared__ int s_cholfail[32];
    const Dev& d = P.d;
    const int Bp = P.Bp, T = P.T;
    const int lane = threadIdx.x, wid = threadIdx.y;
    const int b = blockIdx.x * 32 + lane;
    const int kind = d.kind[b];
    const bool skip_ls_none = (kind == KIND_ITER && P.o.line_search == ILQR_LINE_SEARCH_NONE);
    const bool work = kind != KIND_NONE && !skip_ls_none;
    const bool fresh = kind == KIND_PRELOOP;
    double* hand = ring + (size_t)ST * (BK_STAGE_BYTES / 8);
    if (wid == 0 && lane == 0) {
        for (int i = 0; i < ST; ++i) { mbar_init(&full_bar[i], 32); mbar_init(&empty_bar[i], 64); }
        for (int i = 0; i < LB_HAND; ++i) { mbar_init(&hfull_bar[i], 32); mbar_init(&hempty_bar[i], 32); }
    }
    if (wid == 0) s_cholfail[lane] = 0;
    __syncthreads();
    if (!__syncthreads_or(work)) { /* nothing to linearise in this CTA: bookkeeping only */
        if (wid == 0) {
            const bool running = tick_epilogue(P, b, kind, skip_ls_none ? d.gnorm[b] : 0.0);
            const unsigned mask = __ballot_sync(0xffffffffu, running);
            if (lane == 0 && mask) atomicAdd(&d.active[P.tick & 7], __popc(mask));
        }
        return;
    }
    const int nsteps = T - 1;
    if (LB_IDLE_WARPS && wid >= 4 && (wid & 3) == 0) return; /* would share the matrix warp's sub-partition */
    const int p_idx = wid - 2 - (LB_IDLE_WARPS ? wid / 4 : 0);
    if (wid >= 2) {
        /* ---------------- producers ---------------- */
        const int p = p_idx; /* 0..NPROD-1 */
        LinIn cur;
        if (work && p < nsteps) linearize_load(P, b, T - 2 - p, fresh, cur);
        for (int s = p; s < nsteps; s += NPROD) {
            const int t = T - 2 - s;
            const int stage = s % ST;
            const unsigned use = (unsigned)(s / ST);
            LinIn nxt; /* this producer's next step: its loads fly while the current step is computed */
            const bool more = s + NPROD < nsteps;
            if (work && more) linearize_load(P, b, t - NPROD, fresh, nxt);
#ifdef ILQR_LB_LOAD_BARRIER
            asm volatile("" ::: "memory"); /* keep the prefetch loads ahead of the step's arithmetic */
#endif
            StepFull st;
            if (work) linearize_compute<false, false>(P, b, t, cur, st.s, st.h);
            store_rows_full<N * N>(st.s.fx, work, d.fx, (size_t)t * N * N, Bp, b);
            store_rows_full<N * M>(st.s.fu, work, d.fu, (size_t)t * N * M, Bp, b);
            if (use > 0) mbar_wait_backoff(&empty_bar[stage], (use - 1) & 1); /* both Riccati warps have drained this slot */
            if (work) lane_write<BK_PAIRS, BK_ROWS>(ring + (size_t)stage * (BK_STAGE_BYTES / 8) + lane * 2, st.s.fx);
            mbar_arrive(&full_bar[stage]);
            if (more) cur = nxt;
        }
    } else if (wid == 0) {
        /* ---------------- Riccati matrix warp ---------------- */
#ifdef ILQR_LB_TIMERS
        long long lb_t[4] = {0, 0, 0, 0};
#endif
        double Pm[N * N];
        bool chol_ok = true;
        Hess Hc; /* HACC: this tick's stage Hessians, the same for every step */
        if (HACC && work) hacc_advance(P, b, fresh, Hc, true);
        {
            double pv_dummy[N];
            if (work) linearize_terminal<false>(P, b, fresh, pv_dummy, Pm);         /* src/backward_pass.jl:39: P_T = gxx_T */
            /* hand p_T = gx_T to the vector warp through hand slot LB_HAND-1 (free at this point) */
            if (work) lane_write<(N + 1) / 2, N>(hand + (size_t)(LB_HAND - 1) * HAND_PAIRS * 64 + lane * 2, pv_dummy);
        }
        __syncwarp();
        mbar_arrive(&hfull_bar[LB_HAND - 1]); /* phase 0 of the last slot carries p_T */
#pragma unroll 1
        for (int s = 0; s < nsteps; ++s) {
            const int stage = s % ST;
            /* hand slots: step s uses slot s % (LB_HAND-1); the last slot is reserved for p_T */
            const int hs = s % (LB_HAND - 1);
            const unsigned huse = (unsigned)(s / (LB_HAND - 1));
#ifdef ILQR_LB_TIMERS
            const long long c0 = clock64();
#endif
            mbar_wait(&full_bar[stage], (unsigned)(s / ST) & 1);
#ifdef ILQR_LB_TIMERS
            const long long c1 = clock64();
#endif
            StepFull st;
            if (work) lane_read<BK_PAIRS, BK_ROWS>(st.s.fx, ring + (size_t)stage * (BK_STAGE_BYTES / 8) + lane * 2);
            mbar_arrive(&empty_bar[stage]);
#ifdef ILQR_LB_TIMERS
            const long long c2 = clock64();
#endif
            RicHand h;
            if (work) riccati_matrix_half(st.s, HACC ? Hc : st.h, Pm, h, chol_ok);
            if (work && !chol_ok) s_cholfail[lane] = 1; /* ordered before the hand-over below */
#ifdef ILQR_LB_TIMERS
            const long long c3 = clock64();
#endif
            if (!LB_HAND_FREE && huse > 0) mbar_wait(&hempty_bar[hs], (huse - 1) & 1);
            if (work) lane_write<HAND_PAIRS, HAND_DOUBLES>(hand + (size_t)hs * HAND_PAIRS * 64 + lane * 2, h.K);
            mbar_arrive(&hfull_bar[hs]);
#ifdef ILQR_LB_TIMERS
            const long long c4 = clock64();
            lb_t[0] += c1 - c0; lb_t[1] += c2 - c1; lb_t[2] += c3 - c2; lb_t[3] += c4 - c3;
#endif
        }
#ifdef ILQR_LB_TIMERS /* debug build (variant "lbtimers"): cycles per step of the critical warp.  Measured on B200, acrobot:
                       * wait_full 105 (a try_wait that succeeds at once), read 237, riccati 800, hand-over 107.  Tried and
                       * slower: phase tests issued before the arithmetic (the SYNCS unit serialises them anyway, 1430 in
                       * total) and ld.acquire/st.release flags instead of mbarriers (MEMBAR.ALL.CTA per post, 1600). */
        if (lane == 0 && (blockIdx.x == 0 || blockIdx.x == 200) && (P.tick & 7) == 4)
            printf("lb matrix warp block %d: wait_full %lld read %lld riccati %lld handover %lld cycles/step\n", blockIdx.x,
                   lb_t[0] / nsteps, lb_t[1] / nsteps, lb_t[2] / nsteps, lb_t[3] / nsteps);
#endif
    } else {
        /* ---------------- Riccati vector warp ---------------- */
        double pv[N], gn = 0.0;
        mbar_wait(&hfull_bar[LB_HAND - 1], 0);
        if (work) lane_read<(N + 1) / 2, N>(pv, hand + (size_t)(LB_HAND - 1) * HAND_PAIRS * 64 + lane * 2);   /* :40 p_T = gx_T */
#pragma unroll 1
        for (int s = 0; s < nsteps; ++s) {
            const int t = T - 2 - s;
            const int stage = s % ST;
            const int hs = s % (LB_HAND - 1);
            mbar_wait(&full_bar[stage], (unsigned)(s / ST) & 1);
            StepIn st; /* the vector half only needs the first-order part of the stage */
            if (work) lane_read<(SI_ROWS + 1) / 2, SI_ROWS>(st.fx, ring + (size_t)stage * (BK_STAGE_BYTES / 8) + lane * 2);
            mbar_arrive(&empty_bar[stage]);
            mbar_wait(&hfull_bar[hs], (unsigned)(s / (LB_HAND - 1)) & 1);
            RicHand h;
            if (work) lane_read<HAND_PAIRS, HAND_DOUBLES>(h.K, hand + (size_t)hs * HAND_PAIRS * 64 + lane * 2);
            if (!LB_HAND_FREE) mbar_arrive(&hempty_bar[hs]);
            double kk[d1(M)], Lx[N], Qu[d1(M)];
            if (work) {
                riccati_vector_half(st, h, pv, kk, Lx, Qu, gn);
                st_rows<M * N>(h.K, d.K, (size_t)t * M * N, Bp, b); /* the gains stay live for an idle lane: its line search goes on */
                st_rows<M>(kk, d.k, (size_t)t * M, Bp, b);
            }
            store_rows_full<N>(Lx, work, d.Lx, (size_t)t * N, Bp, b);
            store_rows_full<M>(Qu, work, d.Lu, (size_t)t * M, Bp, b);
        }
        /* the matrix warp raises its Cholesky flag before each hand-over, so it is visible here */
        if (work) {
            if (s_cholfail[lane]) d.flags[b] |= ILQR_FLAG_CHOL_FAIL;
            d.gnorm[b] = gn;
        } else if (skip_ls_none) {
            gn = d.gnorm[b];
        }
        const bool running = tick_epilogue(P, b, kind, gn);
        const unsigned mask = __ballot_sync(0xffffffffu, running);
        if (lane == 0 && mask) atomicAdd(&d.active[P.tick & 7], __popc(mask));
    }
}

/* k_linback_tp ("thread per problem"): the same tick -- gradients! + backward_pass! + lagrangian_gradient! + the
 * convergence tests -- with NO warp specialisation: every warp owns 32 problems and walks them through
 * linearise(t) -> Riccati(t) for t = T-2 ... 0, the next step's inputs prefetched into registers.  One step of one
 * warp is a long in-order stream (about 4 k cycles against 1.1 k for the specialised kernel's critical warp), but a
 * warp needs no shared memory and no partner, so TP_WARPS_PER_SM of them share an SM and keep all four FP64 pipes
 * busy: k_linback puts 32 problems on an SM at a time and leaves three of its four sub-partitions mostly idle.
 * The engine picks this kernel once the grid holds enough warps to fill the machine that way (dense streamed jobs);
 * small grids, where a tick is pure latency, keep the specialised kernel.  Same device functions, same bits. */
#ifndef ILQR_TP_WARPS_PER_SM
#define ILQR_TP_WARPS_PER_SM 8
#endif
constexpr int TP_WARPS_PER_SM = ILQR_TP_WARPS_PER_SM; /* 8 -> 255 registers per thread, 12 -> 168 */
__global__ void __launch_bounds__(32, TP_WARPS_PER_SM) k_linback_tp(const __grid_constant__ Params P) {
    const Dev& d = P.d;
    const int Bp = P.Bp, T = P.T;
    const int b = blockIdx.x * 32 + threadIdx.x;
    const int kind = d.kind[b];
    const bool skip_ls_none = (kind == KIND_ITER && P.o.line_search == ILQR_LINE_SEARCH_NONE);
    const bool work = kind != KIND_NONE && !skip_ls_none;
    const bool fresh = kind == KIND_PRELOOP;
    double gn = 0.0;
    if (__any_sync(0xffffffffu, work)) { /* the idle lanes of a working warp come along for the full-row stores */
        double Pm[N * N], pv[N];
        bool chol_ok = true;
        Hess H;
        LinIn cur;
        if (work) {
            linearize_terminal<false>(P, b, fresh, pv, Pm);                   /* src/backward_pass.jl:39-40 */
            if (HACC) hacc_advance(P, b, fresh, H, true);
            linearize_load(P, b, T - 2, fresh, cur);
        }
#pragma unroll 1
        for (int t = T - 2; t >= 0; --t) {
            LinIn nxt; /* the next step's loads fly while this step is computed */
            StepIn s;
            double K[d1(M * N)], kk[d1(M)], Lx[N], Qu[d1(M)];
            if (work) {
                if (t > 0) linearize_load(P, b, t - 1, fresh, nxt);
                Hess Hs;
                linearize_compute<false, false>(P, b, t, cur, s, Hs);
                riccati_step(s, HACC ? H : Hs, Pm, pv, K, kk, Lx, Qu, chol_ok, gn);
                st_rows<M * N>(K, d.K, (size_t)t * M * N, Bp, b);
                st_rows<M>(kk, d.k, (size_t)t * M, Bp, b);
            }
            store_rows_full<N * N>(s.fx, work, d.fx, (size_t)t * N * N, Bp, b);
            store_rows_full<N * M>(s.fu, work, d.fu, (size_t)t * N * M, Bp, b);
            store_rows_full<N>(Lx, work, d.Lx, (size_t)t * N, Bp, b);
            store_rows_full<M>(Qu, work, d.Lu, (size_t)t * M, Bp, b);
            if (work && t > 0) cur = nxt;
        }
        if (work) {
            if (!chol_ok) d.flags[b] |= ILQR_FLAG_CHOL_FAIL;
            d.gnorm[b] = gn;
        }
    }
    if (!work && skip_ls_none) gn = d.gnorm[b];
    const bool running = tick_epilogue(P, b, kind, gn);
    const unsigned mask = __ballot_sync(0xffffffffu, running);
    if (threadIdx.x == 0 && mask) atomicAdd(&d.active[P.tick & 7], __popc(mask));
}

#else
#include "ilqr_large_backward.cuh"
#endif /* ILQR_LARGE */

/* ==================================================================================== */
/* solve!/constrained_ilqr_solve! prologue: reset!(data), duals, penalties (src/solve.jl:93-103) */
__global__ void k_solve_begin(const __grid_constant__ Params P) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= P.B) return;
    solve_begin_slot(P, b);
}

/* ilqr_solve_outer(restart = 0): the problems parked after their dual update go on with the next inner solve; counts them */
__global__ void k_resume_paused(const __grid_constant__ Params P, int32_t* count) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= P.B) return;
    if (P.d.phase[b] == PH_PAUSED) { P.d.phase[b] = PH_START; atomicAdd(count, 1); }
}
__global__ void k_count_paused(const __grid_constant__ Params P, int32_t* count) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= P.B) return;
    if (P.d.phase[b] == PH_PAUSED) atomicAdd(count, 1);
}

/* ilqr_mpc_run prologue: every problem starts with a receding-horizon shift */
__global__ void k_mpc_begin(const __grid_constant__ Params P) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= P.B) return;
    P.d.mpc_step[b] = 0;
    P.d.mpc_iters[b] = 0;
    P.d.kind[b] = KIND_NONE;
    P.d.phase[b] = P.job->mpc_steps > 0 ? PH_SHIFT : PH_DONE;
}

/* Fresh-solver state of one slot + the prologue of solve! (src/solve.jl:93-103): what a new Solver holds
 * (src/data/problem.jl:32-38, src/data/solver.jl:37-39, src/augmented_lagrangian.jl:17-22) */
__device__ __forceinline__ void slot_reset_scalars(const Params& P, int b, bool set_phase) {
    const Dev& d = P.d;
    d.flags[b] = 0; d.inner_done[b] = 0; d.it[b] = 0; d.ls_base[b] = 0;
    d.iters[b] = 0; d.status[b] = 0; d.gnorm[b] = 0.0; d.viol[b] = 0.0; d.alpha[b] = 1.0;
    d.J[b] = CONSTRAINED ? 0.0 : __longlong_as_double(0x7ff0000000000000LL);
    d.outer[b] = CONSTRAINED ? 1 : 0;
    if (set_phase) { d.kind[b] = KIND_NONE; d.phase[b] = PH_START; }
}

/* k_refill: one CTA per finished slot (grid-stride over the done list of tick P.tick).  Retires the slot's problem
 * (nominal trajectory and solver scalars to the job's output arrays), takes the next problem id from the queue and
 * loads it as a fresh solver.  Runs on a side branch of the CUDA graph, concurrently with the kernels of the next
 * tick: the slot is DONE (inert) for them; `pending` tells the k_forward after next to start it.  Threads stride
 * over rows, so the job-side accesses are contiguous. */
__global__ void __launch_bounds__(128) k_refill(const __grid_constant__ Params P) {
    const Dev& d = P.d;
    const Job& J = *P.job;
    const int ring = P.tick & 3;
    const int count = d.done_count[ring];
    const size_t Bp = P.Bp;
    const int nxr = P.T * N, nur = (P.T - 1) * M, nwr = P.T * NP, ncr = (P.T - 1) * CS + CT;
    __shared__ int s_nid;
    for (int e = blockIdx.x; e < count; e += gridDim.x) {
        const int b = d.done_list[(size_t)ring * Bp + e];
        const int pid = d.pid[b];
        if (pid >= 0) { /* retire */
            if (J.out_x) for (int r = threadIdx.x; r < nxr; r += blockDim.x) J.out_x[(size_t)pid * nxr + r] = d.xb[r * Bp + b];
            if (J.out_u) for (int r = threadIdx.x; r < nur; r += blockDim.x) J.out_u[(size_t)pid * nur + r] = d.ub[r * Bp + b];
            if (threadIdx.x == 0) {
                if (J.out_iters) J.out_iters[pid] = d.iters[b];
                if (J.out_status) J.out_status[pid] = (uint8_t)d.status[b];
                if (J.out_J) J.out_J[pid] = d.J[b];
                if (J.out_viol) J.out_viol[pid] = d.viol[b];
                if (J.out_alpha) J.out_alpha[pid] = d.alpha[b];
                if (J.out_flags) J.out_flags[pid] = d.flags[b];
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) s_nid = atomicAdd(J.next, 1);
        __syncthreads();
        const int nid = s_nid;
        if (nid < J.n_total) { /* refill as a fresh solver */
            for (int r = threadIdx.x; r < nxr; r += blockDim.x) { d.xb[r * Bp + b] = J.in_x[(size_t)nid * nxr + r]; d.xc[r * Bp + b] = 0.0; }
            for (int r = threadIdx.x; r < nur; r += blockDim.x) { d.ub[r * Bp + b] = J.in_u[(size_t)nid * nur + r]; d.uc[r * Bp + b] = 0.0; }
            if (NP > 0 && J.in_w) for (int r = threadIdx.x; r < nwr; r += blockDim.x) d.w[r * Bp + b] = J.in_w[(size_t)nid * nwr + r];
            for (int r = threadIdx.x; r < ncr; r += blockDim.x) {
                d.lam[r * Bp + b] = 0.0; d.rho[r * Bp + b] = P.o.initial_constraint_penalty;
                d.c[r * Bp + b] = 0.0; d.act[r * Bp + b] = 1;
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                d.pid[b] = nid;
                slot_reset_scalars(P, b, false);
                __threadfence();
                d.pending[b] = 1 + ((P.tick + 2) & 7); /* token: the tick whose k_forward starts the slot */
            }
        } else if (threadIdx.x == 0) {
            d.pid[b] = -1;
        }
        __syncthreads();
    }
}

/* ---- drain compaction ---------------------------------------------------------------------------------------
 * Towards the end of a streamed job the queue is empty and the problems still running are scattered over the slots:
 * every warp keeps a few live lanes and a tick costs what it cost with all slots busy.  Between two tick graphs
 * (where every k_refill branch has been joined, so nothing else touches the slots) the engine therefore PACKS the
 * running problems into the lowest slots and launches smaller grids from then on -- a tick's cost follows the number
 * of running problems, and small grids fall back to the latency-optimised kernels.  Moving a problem is a copy of its
 * column in every array that is live between ticks (Dev::mv); its results do not depend on the slot it sits in.
 * k_compact_plan: A = problems running after the last tick.  Running slots at index >= A are sources, idle slots below
 * A are destinations; there are at least as many destinations as sources (A counts slots whose refill came back
 * empty as running for two ticks), so every source gets one. */
__device__ __forceinline__ bool slot_in_use(const Dev& d, int b) { return d.phase[b] != PH_DONE || d.pending[b] != 0; }

__global__ void k_compact_plan(const __grid_constant__ Params P, int slots_in_grid) {
    const Dev& d = P.d;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= slots_in_grid) return;
    const int A = d.active[P.tick & 7];
    const bool used = slot_in_use(d, b);
    if (used && b >= A) d.cmp_src[atomicAdd(&d.cmp_n[0], 1)] = b;
    if (!used && b < A) d.cmp_dst[atomicAdd(&d.cmp_n[1], 1)] = b;
}

__global__ void __launch_bounds__(128) k_compact_move(const __grid_constant__ Params P) {
    const Dev& d = P.d;
    const size_t Bp = P.Bp;
    const int n = d.cmp_n[0];
    for (int e = blockIdx.x; e < n; e += gridDim.x) {
        const int src = d.cmp_src[e], dst = d.cmp_dst[e];
        for (int k = 0; k < d.n_mv; ++k) {
            const MoveEntry m = d.mv[k];
            if (m.elsize == 8) {
                double* a = reinterpret_cast<double*>(m.base);
                for (int r = threadIdx.x; r < m.rows; r += blockDim.x) a[r * Bp + dst] = a[r * Bp + src];
            } else if (m.elsize == 4) {
                int32_t* a = reinterpret_cast<int32_t*>(m.base);
                for (int r = threadIdx.x; r < m.rows; r += blockDim.x) a[r * Bp + dst] = a[r * Bp + src];
            } else {
                uint8_t* a = reinterpret_cast<uint8_t*>(m.base);
                for (int r = threadIdx.x; r < m.rows; r += blockDim.x) a[r * Bp + dst] = a[r * Bp + src];
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) { /* the vacated slot idles */
            d.phase[src] = PH_DONE; d.kind[src] = KIND_NONE; d.pid[src] = -1; d.pending[src] = 0; d.refilling[src] = 0;
            d.ls_base[src] = 0; /* a problem moved in the middle of its line search must not leave its position behind */
        }
    }
}
__global__ void k_compact_reset(const __grid_constant__ Params P) { P.d.cmp_n[0] = 0; P.d.cmp_n[1] = 0; }

/* start of a streamed job: slots 0..min(B, n_total)-1 hold the first problems (their trajectories were
 * already transposed in by the host side); the other slots idle */
__global__ void k_stream_begin(const __grid_constant__ Params P) {
    const Dev& d = P.d;
    const Job& J = *P.job;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= P.Bp) return;
    const size_t Bp = P.Bp;
    const int ncr = (P.T - 1) * CS + CT;
    if (b < P.B && b < J.n_total) {
        for (int r = 0; r < ncr; ++r) {
            d.lam[r * Bp + b] = 0.0; d.rho[r * Bp + b] = P.o.initial_constraint_penalty;
            d.c[r * Bp + b] = 0.0; d.act[r * Bp + b] = 1;
        }
        d.pid[b] = b;
        slot_reset_scalars(P, b, true);
    } else {
        d.pid[b] = -1;
        d.phase[b] = PH_DONE;
        d.kind[b] = KIND_NONE;
    }
    d.pending[b] = 0;
    d.refilling[b] = 0;
    if (b == 0) { *J.next = P.B < J.n_total ? P.B : J.n_total; for (int i = 0; i < 4; ++i) d.done_count[i] = 0; }
}

/* rollout (src/rollout.jl:33-42), open loop: x [T][N][Bp] from x[0] and u [T-1][M][Bp] */
__global__ void k_rollout(const __grid_constant__ Params P, double* __restrict__ x, const double* __restrict__ u) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= P.B) return;
    const int Bp = P.Bp;
    double xv[N], xn[N], uv[d1(M)], wv[d1(NP)];
    ld_rows<N>(xv, x, 0, Bp, b);
    for (int t = 0; t < P.T - 1; ++t) {
        ld_rows<M>(uv, u, (size_t)t * M, Bp, b);
        ld_rows<NP>(wv, P.d.w, (size_t)t * NP, Bp, b);
        ilqr_dyn(xn, xv, uv, wv);
#pragma unroll
        for (int i = 0; i < N; ++i) xv[i] = xn[i];
        st_rows<N>(xv, x, (size_t)(t + 1) * N, Bp, b);
    }
}

/* receding-horizon shift for the whole batch (ilqr_mpc_step); logs are [component][Bp] scratch rows */
__global__ void k_mpc_shift(const __grid_constant__ Params P, double* __restrict__ applied_u, double* __restrict__ x_next) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= P.B) return;
    double lu[d1(M)], lx[N];
    mpc_shift_slot(P, b, lu, lx);
#pragma unroll
    for (int a = 0; a < M; ++a) applied_u[(size_t)a * P.Bp + b] = lu[a];
#pragma unroll
    for (int i = 0; i < N; ++i) x_next[(size_t)i * P.Bp + b] = lx[i];
}

/* layout changes at the ABI: host [problem][row] <-> device [row][problem(Bp)] */
template <typename TS, typename TD>
__global__ void k_to_soa(const TS* __restrict__ src, TD* __restrict__ dst, int B, int Bp, int rows) {
    __shared__ TD tile[32][33];
    const int r0 = blockIdx.x * 32, b0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int bb = b0 + j, r = r0 + threadIdx.x;
        if (bb < B && r < rows) tile[j][threadIdx.x] = (TD)src[(size_t)bb * rows + r];
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int r = r0 + j, bb = b0 + threadIdx.x;
        if (r < rows && bb < B) dst[(size_t)r * Bp + bb] = tile[threadIdx.x][j];
    }
}
template <typename TS, typename TD>
__global__ void k_from_soa(const TS* __restrict__ src, TD* __restrict__ dst, int B, int Bp, int rows) {
    __shared__ TD tile[32][33];
    const int r0 = blockIdx.x * 32, b0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int r = r0 + j, bb = b0 + threadIdx.x;
        if (r < rows && bb < B) tile[j][threadIdx.x] = (TD)src[(size_t)r * Bp + bb];
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int bb = b0 + j, r = r0 + threadIdx.x;
        if (bb < B && r < rows) dst[(size_t)bb * rows + r] = tile[threadIdx.x][j];
    }
}

} /* namespace ilqr */

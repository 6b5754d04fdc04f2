/*
 * ilqr_cuda.h -- C ABI of libilqr_cuda.so, the B200 batched iLQR engine.
 *
 * Drop-in boundary for the SOLVE PATH of thowell/IterativeLQR.jl.  The reference
 * has no FFI layer of its own; the entry points below are what a Julia host shim
 * binds with `ccall` in place of the reference's Julia methods.  Each one cites the
 * reference interface it replaces (paths relative to /root/reference/).
 *
 * Conventions
 *  - every function returns 0 on success or a negative ILQR_E* code; nothing throws
 *    across the boundary; ilqr_last_error() gives the message for the last failure
 *    on that handle (or, with a NULL handle, for the last failed ilqr_create);
 *  - host buffers are caller-owned, plain double arrays laid out
 *    [problem][time][component] (problem slowest); matrices column-major like Julia;
 *  - the library owns all device memory; a handle is bound to one CUDA device and is
 *    not thread-safe; distinct handles are independent;
 *  - there is NO CPU fallback: ilqr_create fails if no CUDA device is usable.
 *
 * A handle solves `batch` independent problems of one compiled MODEL (dynamics, costs,
 * constraints: a shared library emitted by the code generator and compiled with nvcc
 * for sm_100a).  All arithmetic is IEEE-754 binary64.
 */
#ifndef ILQR_CUDA_H
#define ILQR_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ILQR_ABI_VERSION 2

enum {
    ILQR_OK = 0,
    ILQR_EINVAL = -1,      /* bad argument / dimension mismatch */
    ILQR_ECUDA = -2,       /* CUDA runtime error (message has the cudaError string) */
    ILQR_EMODEL = -3,      /* model library missing, not loadable or ABI mismatch */
    ILQR_ENOMEM = -4,
    ILQR_ESTATE = -5       /* call sequence error */
};

enum { ILQR_LINE_SEARCH_ARMIJO = 0, ILQR_LINE_SEARCH_NONE = 1 };

/* per-problem flag bits returned by ilqr_get_stats (extra outputs; they do not alter
 * the control flow -- SURVEY.md Q3) */
enum {
    ILQR_FLAG_CHOL_FAIL = 1u,   /* some Quu was not positive definite (potrf info != 0) */
    ILQR_FLAG_NONFINITE = 2u    /* a line-search trial produced a non-finite cost */
};

/* Options -- src/options.jl:1-14, field for field (constraint_norm is accepted and
 * ignored, as in the reference: src/data/constraints.jl:23-39 always uses the inf-norm). */
typedef struct ilqr_options {
    int32_t line_search;           /* ILQR_LINE_SEARCH_*        (:armijo)  */
    int32_t max_iterations;        /* 100 */
    int32_t max_dual_updates;      /* 10  */
    int32_t reset_cache;           /* false */
    int32_t verbose;               /* ignored by the library; the host shim prints */
    int32_t reserved;
    double min_step_size;          /* 1e-5 */
    double objective_tolerance;    /* 1e-3 */
    double lagrangian_gradient_tolerance; /* 1e-3 */
    double constraint_tolerance;   /* 5e-3 */
    double constraint_norm;        /* Inf (unused) */
    double initial_constraint_penalty;    /* 1.0 */
    double scaling_penalty;        /* 10.0 */
    double max_penalty;            /* 1e8 */
} ilqr_options;

typedef struct ilqr_desc {
    int32_t abi_version;           /* ILQR_ABI_VERSION */
    int32_t T;                     /* knot points (states); T-1 actions  */
    int32_t n, m, p;               /* state / action / parameter dims: checked against the model */
    int32_t c_s, c_T;              /* stage / terminal constraint rows: checked against the model */
    int32_t batch;                 /* independent problems in this handle */
    int32_t device;                /* CUDA device ordinal */
    int32_t history_cap;           /* per-problem iteration records kept; 0 = 1000 (src/data/solver.jl:21) */
    const char* model_library;     /* path of the compiled model plug-in (.so) */
} ilqr_desc;

typedef struct ilqr_handle ilqr_handle;

/* Options{T}() defaults -- src/options.jl:1-14 */
void ilqr_options_default(ilqr_options* out);

/* Solver(dynamics, objective[, constraints]; options) -- src/solver.jl:11-46.
 * Allocates the batched workspace (ProblemData/PolicyData/SolverData/AL data:
 * src/data/problem.jl:25-45, src/data/policy.jl:44-77, src/data/solver.jl:20-47,
 * src/augmented_lagrangian.jl:13-37); current and nominal trajectories start at zero. */
int ilqr_create(const ilqr_desc* desc, const ilqr_options* options, ilqr_handle** out);
void ilqr_destroy(ilqr_handle* h);
const char* ilqr_last_error(const ilqr_handle* h);

/* Run all of this handle's device work on the caller's CUDA stream (a cudaStream_t passed as
 * void*; NULL restores the handle's own stream).  Calls stay synchronous for the host. */
int ilqr_set_stream(ilqr_handle* h, void* cuda_stream);

/* solver.options = ...  -- src/solver.jl:8 (Options is a mutable field) */
int ilqr_set_options(ilqr_handle* h, const ilqr_options* options);

/* initialize_controls!(solver, actions) -- src/solver.jl:56-60.   u: [batch][T-1][m] */
int ilqr_initialize_controls(ilqr_handle* h, const double* u);
/* initialize_states!(solver, states)   -- src/solver.jl:62-66.   x: [batch][T][n]   */
int ilqr_initialize_states(ilqr_handle* h, const double* x);
/* Solver(...; parameters) -- src/solver.jl:12, src/data/problem.jl:25-30.  w: [batch][T][p]
 * (the terminal entry w[T-1] feeds the terminal cost/constraint) */
int ilqr_set_parameters(ilqr_handle* h, const double* w);

/* The two initialisers above for inputs that are ALREADY in device memory of the handle's
 * device (same [batch][time][component] layout): no host round trip. */
int ilqr_initialize_controls_device(ilqr_handle* h, const double* d_u);
int ilqr_initialize_states_device(ilqr_handle* h, const double* d_x);

/* rollout(dynamics, initial_state, actions, parameters) -- src/rollout.jl:33-42.
 * x1: [batch][n], u: [batch][T-1][m] -> x_out: [batch][T][n].  Uses the handle's
 * parameters; does not touch the solver state. */
int ilqr_rollout(ilqr_handle* h, const double* x1, const double* u, double* x_out);

/* solve!(solver) -- src/solve.jl:137-143: constrained_ilqr_solve! (:88-129) when the model
 * has constraints, else ilqr_solve! (:1-54).  Runs every problem of the batch to the
 * reference's own termination. */
int ilqr_solve(ilqr_handle* h);
/* solve!(solver, states, actions) -- src/solve.jl:56-60, :131-135 (warm start) */
int ilqr_solve_warm(ilqr_handle* h, const double* x, const double* u);
/* constrained_ilqr_solve!(solver; augmented_lagrangian_callback!) -- src/solve.jl:87-129 with the user callback of :125.
 * The host drives the augmented-Lagrangian loop one outer iteration at a time:
 *     ilqr_solve_outer(h, 1, &n);  while (n > 0) { callback(solver);  ilqr_solve_outer(h, 0, &n); }
 * restart != 0 runs the prologue (:93-103) and then every problem up to and including its next dual update (:120-122) or
 * its termination; restart == 0 resumes the problems parked there.  *n_paused = problems waiting for the next call.
 * Every problem makes exactly the steps of ilqr_solve; unconstrained models finish in the first call. */
int ilqr_solve_outer(ilqr_handle* h, int32_t restart, int32_t* n_paused);

/* Continuous batching: solve n_problems FRESH problems (each exactly as `Solver(...); initialize_controls!;
 * initialize_states!; solve!` on a new solver would -- src/solver.jl:28-46, src/solve.jl:137-143) by streaming
 * them through the handle's `batch` slots: a slot whose problem has terminated is written out and refilled
 * with the next problem between lock-step iterations, so no slot idles while other problems are still
 * iterating.  Every pointer is DEVICE memory of the handle's device: inputs d_x [n][T][n_state],
 * d_u [n][T-1][m], d_w [n][T][p] (NULL if p = 0); outputs (any may be NULL) nominal trajectories and the
 * SolverData scalars per problem.  The handle's own trajectories / history are scratch afterwards. */
int ilqr_solve_stream(ilqr_handle* h, int32_t n_problems, const double* d_x, const double* d_u, const double* d_w,
                      double* d_x_out, double* d_u_out, int32_t* d_iterations, uint8_t* d_status, double* d_objective,
                      double* d_max_violation, double* d_step_size, uint32_t* d_flags);

/* ilqr_solve_stream with HOST buffers: the host->device copy of the inputs and the device->host copy of the
 * results happen inside the call. */
int ilqr_solve_stream_host(ilqr_handle* h, int32_t n_problems, const double* x, const double* u, const double* w,
                           double* x_out, double* u_out, int32_t* iterations, uint8_t* status, double* objective,
                           double* max_violation, double* step_size, uint32_t* flags);

/* get_trajectory(solver) -- src/solver.jl:48-50 (NOMINAL trajectory).
 * x: [batch][T][n], u: [batch][T-1][m]; either may be NULL. */
int ilqr_get_trajectory(ilqr_handle* h, double* x, double* u);
/* current_trajectory(solver) -- src/solver.jl:52-54 */
int ilqr_get_current_trajectory(ilqr_handle* h, double* x, double* u);
/* same as ilqr_get_trajectory but into DEVICE buffers of the handle's device (for a
 * host that gathers shards with NCCL); same layout. */
int ilqr_get_trajectory_device(ilqr_handle* h, double* d_x, double* d_u);

/* SolverData scalars -- src/data/solver.jl:4-18: iterations[1], status[1], objective[1],
 * max_violation[1], step_size[1]; plus flags (ILQR_FLAG_*).  Arrays of length batch; any may be NULL. */
int ilqr_get_stats(ilqr_handle* h, int32_t* iterations, uint8_t* status, double* objective,
                   double* max_violation, double* step_size, uint32_t* flags);

/* The per-iteration record the reference prints when verbose (src/solve.jl:40-45): cost,
 * gradient_norm, max_violation, step_size, plus the AL outer index (src/solve.jl:106) and
 * line-search status.  Arrays are [batch][cap]; record r of problem b is valid for
 * r < min(iterations[b], history_cap).  Any array may be NULL. */
int ilqr_get_history(ilqr_handle* h, int32_t cap, double* cost, double* gradient_norm,
                     double* max_violation, double* step_size, int32_t* outer, uint8_t* status);

/* AugmentedLagrangianCosts state -- src/augmented_lagrangian.jl:1-11.
 * Rows per problem: (T-1)*c_s stage rows then c_T terminal rows.  Any may be NULL. */
int ilqr_get_duals(ilqr_handle* h, double* dual, double* penalty, double* violations, int32_t* active_set);

/* PolicyData gains -- src/data/policy.jl:23-27.  K: [batch][T-1][m*n] (column-major m x n), k: [batch][T-1][m] */
int ilqr_get_policy(ilqr_handle* h, double* K, double* k);

/* Receding-horizon step (BASELINE config 5; the reference ships only the warm-start entry
 * src/solve.jl:131-135): apply the first nominal action to the plant (the model dynamics),
 * shift the nominal actions left repeating the last one, roll the nominal states out from
 * the new initial state and re-solve.  applied_u: [batch][m] or NULL, x_next: [batch][n] or NULL. */
int ilqr_mpc_step(ilqr_handle* h, double* applied_u, double* x_next);

/* n_steps receding-horizon steps for every problem of the handle, starting from its present nominal trajectory:
 * step = ilqr_mpc_step's shift followed by solve!(solver, x, u) (src/solve.jl:131-135).  Problems do not wait for
 * one another: each goes on to its next step as soon as its own solve terminates.  DEVICE pointers, any may be
 * NULL: d_applied_u [n_steps][batch][m] (action applied to the plant at each step), d_x_next [n_steps][batch][n]
 * (plant state after it), d_total_iterations [batch] (iLQR iterations summed over the steps). */
int ilqr_mpc_run(ilqr_handle* h, int32_t n_steps, double* d_applied_u, double* d_x_next, int32_t* d_total_iterations);

/* Instrumentation: number of lock-step batch iterations and kernel launches of the last
 * solve, and accumulated device time per kernel kind (CUDA events on the solve stream;
 * only collected while profiling is on).  kinds: 0 forward, 1 linearize, 2 backward. */
int ilqr_set_profiling(ilqr_handle* h, int32_t on);
int ilqr_get_counters(ilqr_handle* h, int64_t* ticks, int64_t* launches, double kernel_ms[3], int64_t kernel_launches[3]);
/* sum over the ticks of the last solves of the number of problems each tick worked on
 * (the unit count behind the roofline's algorithmic bytes); reset by ilqr_set_profiling(h, 1) */
int ilqr_get_problem_ticks(ilqr_handle* h, int64_t* problem_ticks);
/* how many times the streamed jobs of this handle packed their running problems into the lowest slots and shrank
 * the grid (drain compaction; ILQR_COMPACT_MIN_BLOCKS in the environment bounds it) */
int ilqr_get_compactions(ilqr_handle* h, int64_t* compactions);

/* ---- multi-GPU: the final gather (the reference has no counterpart -- it is single-threaded CPU code).
 * Problems are independent, so a batch is sharded over GPUs as one handle per device (one per process under
 * torchrun / MPI, or several in one process) with no data-path communication; the only collective is the gather
 * of trajectories / solver scalars at the end.  NCCL is loaded at run time (dlopen), never linked.
 *   ilqr_comm_unique_id  rank 0 creates the 128-byte rendezvous id (ncclGetUniqueId) and ships it to the other ranks
 *                        by whatever means the host has (torch.distributed / MPI broadcast, a file, ...);
 *   ilqr_comm_init       every rank joins with its handle (ncclCommInitRank on the handle's device);
 *   ilqr_gather          ncclAllGather of `bytes_per_rank` bytes from d_local into d_all ([n_ranks][bytes_per_rank],
 *                        DEVICE pointers) on the handle's stream; returns once it is enqueued. */
#define ILQR_COMM_ID_BYTES 128
int ilqr_comm_unique_id(char id[ILQR_COMM_ID_BYTES]);
int ilqr_comm_init(ilqr_handle* h, int32_t n_ranks, int32_t rank, const char id[ILQR_COMM_ID_BYTES]);
int ilqr_gather(ilqr_handle* h, const void* d_local, void* d_all, size_t bytes_per_rank);

/* model plug-in facts */
int ilqr_model_dims(const char* model_library, int32_t* n, int32_t* m, int32_t* p, int32_t* c_s, int32_t* c_T);

#ifdef __cplusplus
}
#endif
#endif /* ILQR_CUDA_H */

/*
 * ilqr_plugin.h -- private interface between libilqr_cuda.so (csrc/ilqr_front.cpp, the
 * C ABI of include/ilqr_cuda.h) and a compiled MODEL plug-in (csrc/ilqr_engine.cu built
 * against one generated model header).  The kernels are specialised at compile time on
 * the model's dimensions and inline its dynamics / cost / constraint functions, so each
 * model is its own shared library; the front library dlopen()s it and forwards calls
 * through this table.
 */
#ifndef ILQR_PLUGIN_H
#define ILQR_PLUGIN_H

#include <stddef.h>
#include <stdint.h>

#include "ilqr_cuda.h"

#define ILQR_PLUGIN_VERSION 7
#define ILQR_PLUGIN_SYMBOL "ilqr_plugin_table_v7"
#define ILQR_ERRLEN 512

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ilqr_plugin_table {
    int32_t plugin_version;
    int32_t n, m, p, c_s, c_T;
    const char* model_name;
    const char* model_hash;
    /* all functions: 0 or negative ILQR_E*, message written to err (ILQR_ERRLEN bytes) */
    int (*create)(const ilqr_desc*, const ilqr_options*, void** impl, char* err);
    void (*destroy)(void* impl);
    int (*set_options)(void* impl, const ilqr_options*, char* err);
    int (*initialize_controls)(void* impl, const double* u, int device_in, char* err);
    int (*initialize_states)(void* impl, const double* x, int device_in, char* err);
    int (*set_parameters)(void* impl, const double* w, char* err);
    int (*rollout)(void* impl, const double* x1, const double* u, double* x_out, char* err);
    int (*solve)(void* impl, char* err);
    int (*get_trajectory)(void* impl, double* x, double* u, int current, int device_out, char* err);
    int (*get_stats)(void* impl, int32_t*, uint8_t*, double*, double*, double*, uint32_t*, char* err);
    int (*get_history)(void* impl, int32_t cap, double*, double*, double*, double*, int32_t*, uint8_t*, char* err);
    int (*get_duals)(void* impl, double*, double*, double*, int32_t*, char* err);
    int (*get_policy)(void* impl, double* K, double* k, char* err);
    int (*mpc_step)(void* impl, double* applied_u, double* x_next, char* err);
    int (*set_profiling)(void* impl, int32_t on, char* err);
    int (*get_counters)(void* impl, int64_t* ticks, int64_t* launches, double* kernel_ms, int64_t* kernel_launches, char* err);
    int (*get_problem_ticks)(void* impl, int64_t* problem_ticks, char* err);
    int (*set_stream)(void* impl, void* cuda_stream, char* err);
    int (*solve_stream)(void* impl, int32_t n_problems, const double* d_x, const double* d_u, const double* d_w, double* d_x_out,
                        double* d_u_out, int32_t* d_iterations, uint8_t* d_status, double* d_objective, double* d_max_violation,
                        double* d_step_size, uint32_t* d_flags, char* err);
    int (*solve_stream_host)(void* impl, int32_t n_problems, const double* x, const double* u, const double* w, double* x_out,
                             double* u_out, int32_t* iterations, uint8_t* status, double* objective, double* max_violation,
                             double* step_size, uint32_t* flags, char* err);
    int (*mpc_run)(void* impl, int32_t n_steps, double* d_applied_u, double* d_x_next, int32_t* d_total_iterations, char* err);
    int (*get_compactions)(void* impl, int64_t* compactions, char* err);
    int (*comm_init)(void* impl, int32_t n_ranks, int32_t rank, const char* id, char* err);
    int (*gather)(void* impl, const void* d_local, void* d_all, size_t bytes_per_rank, char* err);
    int (*solve_outer)(void* impl, int32_t restart, int32_t* n_paused, char* err);
} ilqr_plugin_table;

#ifdef __cplusplus
}
#endif
#endif

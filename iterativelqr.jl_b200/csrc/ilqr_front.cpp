/*
 * ilqr_front.cpp -- libilqr_cuda.so: the C ABI of include/ilqr_cuda.h.
 *
 * Thin by design: it validates arguments, dlopen()s the compiled model plug-in named in
 * ilqr_desc::model_library (csrc/ilqr_engine.cu built against one generated model), and
 * forwards each entry point through the plug-in table.  All numerical work happens in
 * the plug-in's CUDA kernels; there is no CPU path here to fall back to.
 */
#include <dlfcn.h>

#include <cstdio>
#include <cstring>
#include <new>

#include "ilqr_cuda.h"
#include "ilqr_nccl_dyn.h"
#include "ilqr_plugin.h"

struct ilqr_handle {
    void* dl = nullptr;
    const ilqr_plugin_table* vt = nullptr;
    void* impl = nullptr;
    char err[ILQR_ERRLEN] = {0};
};

static thread_local char g_create_err[ILQR_ERRLEN] = {0};

static int load_table(const char* path, void** dl_out, const ilqr_plugin_table** vt_out, char* err) {
    if (!path || !*path) {
        snprintf(err, ILQR_ERRLEN, "model_library is empty");
        return ILQR_EMODEL;
    }
    void* dl = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!dl) {
        snprintf(err, ILQR_ERRLEN, "cannot load model library '%s': %s", path, dlerror());
        return ILQR_EMODEL;
    }
    const ilqr_plugin_table* vt = (const ilqr_plugin_table*)dlsym(dl, ILQR_PLUGIN_SYMBOL);
    if (!vt) {
        snprintf(err, ILQR_ERRLEN, "'%s' is not an ilqr model plug-in (symbol %s missing)", path, ILQR_PLUGIN_SYMBOL);
        dlclose(dl);
        return ILQR_EMODEL;
    }
    if (vt->plugin_version != ILQR_PLUGIN_VERSION) {
        snprintf(err, ILQR_ERRLEN, "model plug-in '%s' has version %d, library expects %d", path, vt->plugin_version, ILQR_PLUGIN_VERSION);
        dlclose(dl);
        return ILQR_EMODEL;
    }
    *dl_out = dl;
    *vt_out = vt;
    return 0;
}

extern "C" {

void ilqr_options_default(ilqr_options* o) { /* src/options.jl:1-14 */
    if (!o) return;
    memset(o, 0, sizeof(*o));
    o->line_search = ILQR_LINE_SEARCH_ARMIJO;
    o->max_iterations = 100;
    o->max_dual_updates = 10;
    o->reset_cache = 0;
    o->verbose = 1;
    o->min_step_size = 1.0e-5;
    o->objective_tolerance = 1.0e-3;
    o->lagrangian_gradient_tolerance = 1.0e-3;
    o->constraint_tolerance = 5.0e-3;
    o->constraint_norm = __builtin_inf();
    o->initial_constraint_penalty = 1.0;
    o->scaling_penalty = 10.0;
    o->max_penalty = 1.0e8;
}

int ilqr_create(const ilqr_desc* desc, const ilqr_options* options, ilqr_handle** out) {
    g_create_err[0] = 0;
    if (!desc || !out) {
        snprintf(g_create_err, ILQR_ERRLEN, "desc/out is NULL");
        return ILQR_EINVAL;
    }
    *out = nullptr;
    if (desc->abi_version != ILQR_ABI_VERSION) {
        snprintf(g_create_err, ILQR_ERRLEN, "abi_version %d != library ABI %d", desc->abi_version, ILQR_ABI_VERSION);
        return ILQR_EINVAL;
    }
    ilqr_options defaults;
    if (!options) {
        ilqr_options_default(&defaults);
        options = &defaults;
    }
    ilqr_handle* h = new (std::nothrow) ilqr_handle();
    if (!h) return ILQR_ENOMEM;
    int rc = load_table(desc->model_library, &h->dl, &h->vt, g_create_err);
    if (rc == 0) rc = h->vt->create(desc, options, &h->impl, g_create_err);
    if (rc != 0) {
        if (h->dl) dlclose(h->dl);
        delete h;
        return rc;
    }
    *out = h;
    return 0;
}

void ilqr_destroy(ilqr_handle* h) {
    if (!h) return;
    if (h->vt && h->impl) h->vt->destroy(h->impl);
    /* the plug-in stays mapped: unloading a library that registered CUDA kernels at exit time is unsafe */
    delete h;
}

const char* ilqr_last_error(const ilqr_handle* h) { return h ? h->err : g_create_err; }

#define CHECK_H(h)            \
    if (!(h) || !(h)->impl) return ILQR_EINVAL; \
    (h)->err[0] = 0

int ilqr_set_stream(ilqr_handle* h, void* cuda_stream) { CHECK_H(h); return h->vt->set_stream(h->impl, cuda_stream, h->err); }
int ilqr_set_options(ilqr_handle* h, const ilqr_options* o) { CHECK_H(h); return h->vt->set_options(h->impl, o, h->err); }
int ilqr_initialize_controls(ilqr_handle* h, const double* u) { CHECK_H(h); return h->vt->initialize_controls(h->impl, u, 0, h->err); }
int ilqr_initialize_states(ilqr_handle* h, const double* x) { CHECK_H(h); return h->vt->initialize_states(h->impl, x, 0, h->err); }
int ilqr_initialize_controls_device(ilqr_handle* h, const double* d_u) { CHECK_H(h); return h->vt->initialize_controls(h->impl, d_u, 1, h->err); }
int ilqr_initialize_states_device(ilqr_handle* h, const double* d_x) { CHECK_H(h); return h->vt->initialize_states(h->impl, d_x, 1, h->err); }
int ilqr_set_parameters(ilqr_handle* h, const double* w) { CHECK_H(h); return h->vt->set_parameters(h->impl, w, h->err); }
int ilqr_rollout(ilqr_handle* h, const double* x1, const double* u, double* x_out) { CHECK_H(h); return h->vt->rollout(h->impl, x1, u, x_out, h->err); }
int ilqr_solve(ilqr_handle* h) { CHECK_H(h); return h->vt->solve(h->impl, h->err); }

int ilqr_solve_warm(ilqr_handle* h, const double* x, const double* u) { /* src/solve.jl:56-60, :131-135 */
    CHECK_H(h);
    int rc = h->vt->initialize_controls(h->impl, u, 0, h->err);
    if (!rc) rc = h->vt->initialize_states(h->impl, x, 0, h->err);
    if (!rc) rc = h->vt->solve(h->impl, h->err);
    return rc;
}

int ilqr_solve_stream(ilqr_handle* h, int32_t n_problems, const double* d_x, const double* d_u, const double* d_w,
                      double* d_x_out, double* d_u_out, int32_t* d_iterations, uint8_t* d_status, double* d_objective,
                      double* d_max_violation, double* d_step_size, uint32_t* d_flags) {
    CHECK_H(h);
    return h->vt->solve_stream(h->impl, n_problems, d_x, d_u, d_w, d_x_out, d_u_out, d_iterations, d_status, d_objective,
                               d_max_violation, d_step_size, d_flags, h->err);
}

int ilqr_solve_stream_host(ilqr_handle* h, int32_t n_problems, const double* x, const double* u, const double* w, double* x_out,
                           double* u_out, int32_t* iterations, uint8_t* status, double* objective, double* max_violation,
                           double* step_size, uint32_t* flags) {
    CHECK_H(h);
    return h->vt->solve_stream_host(h->impl, n_problems, x, u, w, x_out, u_out, iterations, status, objective, max_violation,
                                    step_size, flags, h->err);
}

int ilqr_get_trajectory(ilqr_handle* h, double* x, double* u) { CHECK_H(h); return h->vt->get_trajectory(h->impl, x, u, 0, 0, h->err); }
int ilqr_get_current_trajectory(ilqr_handle* h, double* x, double* u) { CHECK_H(h); return h->vt->get_trajectory(h->impl, x, u, 1, 0, h->err); }
int ilqr_get_trajectory_device(ilqr_handle* h, double* d_x, double* d_u) { CHECK_H(h); return h->vt->get_trajectory(h->impl, d_x, d_u, 0, 1, h->err); }

int ilqr_get_stats(ilqr_handle* h, int32_t* iterations, uint8_t* status, double* objective, double* max_violation,
                   double* step_size, uint32_t* flags) {
    CHECK_H(h);
    return h->vt->get_stats(h->impl, iterations, status, objective, max_violation, step_size, flags, h->err);
}
int ilqr_get_history(ilqr_handle* h, int32_t cap, double* cost, double* gradient_norm, double* max_violation,
                     double* step_size, int32_t* outer, uint8_t* status) {
    CHECK_H(h);
    return h->vt->get_history(h->impl, cap, cost, gradient_norm, max_violation, step_size, outer, status, h->err);
}
int ilqr_get_duals(ilqr_handle* h, double* dual, double* penalty, double* violations, int32_t* active_set) {
    CHECK_H(h);
    return h->vt->get_duals(h->impl, dual, penalty, violations, active_set, h->err);
}
int ilqr_get_policy(ilqr_handle* h, double* K, double* k) { CHECK_H(h); return h->vt->get_policy(h->impl, K, k, h->err); }
int ilqr_mpc_step(ilqr_handle* h, double* applied_u, double* x_next) { CHECK_H(h); return h->vt->mpc_step(h->impl, applied_u, x_next, h->err); }
int ilqr_mpc_run(ilqr_handle* h, int32_t n_steps, double* d_applied_u, double* d_x_next, int32_t* d_total_iterations) {
    CHECK_H(h);
    return h->vt->mpc_run(h->impl, n_steps, d_applied_u, d_x_next, d_total_iterations, h->err);
}
int ilqr_set_profiling(ilqr_handle* h, int32_t on) { CHECK_H(h); return h->vt->set_profiling(h->impl, on, h->err); }
int ilqr_get_counters(ilqr_handle* h, int64_t* ticks, int64_t* launches, double kernel_ms[3], int64_t kernel_launches[3]) {
    CHECK_H(h);
    return h->vt->get_counters(h->impl, ticks, launches, kernel_ms, kernel_launches, h->err);
}

int ilqr_get_problem_ticks(ilqr_handle* h, int64_t* problem_ticks) { CHECK_H(h); return h->vt->get_problem_ticks(h->impl, problem_ticks, h->err); }
int ilqr_comm_unique_id(char id[ILQR_COMM_ID_BYTES]) {
    if (!id) return ILQR_EINVAL;
    const char* why = nullptr;
    const ilqr_nccl::Api* n = ilqr_nccl::api(&why);
    if (!n) return ILQR_ECUDA;
    ilqr_nccl::unique_id u;
    if (n->GetUniqueId(&u) != 0) return ILQR_ECUDA;
    memcpy(id, u.internal, ILQR_COMM_ID_BYTES);
    return 0;
}
int ilqr_comm_init(ilqr_handle* h, int32_t n_ranks, int32_t rank, const char id[ILQR_COMM_ID_BYTES]) {
    CHECK_H(h);
    return h->vt->comm_init(h->impl, n_ranks, rank, id, h->err);
}
int ilqr_gather(ilqr_handle* h, const void* d_local, void* d_all, size_t bytes_per_rank) {
    CHECK_H(h);
    return h->vt->gather(h->impl, d_local, d_all, bytes_per_rank, h->err);
}
int ilqr_solve_outer(ilqr_handle* h, int32_t restart, int32_t* n_paused) {
    CHECK_H(h);
    return h->vt->solve_outer(h->impl, restart, n_paused, h->err);
}
int ilqr_get_compactions(ilqr_handle* h, int64_t* compactions) { CHECK_H(h); return h->vt->get_compactions(h->impl, compactions, h->err); }

int ilqr_model_dims(const char* model_library, int32_t* n, int32_t* m, int32_t* p, int32_t* c_s, int32_t* c_T) {
    void* dl = nullptr;
    const ilqr_plugin_table* vt = nullptr;
    int rc = load_table(model_library, &dl, &vt, g_create_err);
    if (rc) return rc;
    if (n) *n = vt->n;
    if (m) *m = vt->m;
    if (p) *p = vt->p;
    if (c_s) *c_s = vt->c_s;
    if (c_T) *c_T = vt->c_T;
    return 0;
}

} /* extern "C" */

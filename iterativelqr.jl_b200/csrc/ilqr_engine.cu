/*
 * ilqr_engine.cu -- host side of one compiled MODEL plug-in: the batched device workspace
 * (the reference's Solver/ProblemData/PolicyData/SolverData, src/solver.jl:4-46, as
 * structure-of-arrays in HBM), the lock-step solve loop and the plug-in table that
 * libilqr_cuda.so (csrc/ilqr_front.cpp) forwards the C ABI to.
 *
 * Built per model:  nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -lineinfo
 *                        -include <generated model header> -shared ...   (build.py)
 */
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <map>
#include <vector>

#include "ilqr_cuda.h"
#include "ilqr_nccl_dyn.h"
#include "ilqr_plugin.h"
#ifndef ILQR_MODEL_GEN_H
#error "compile with -include <generated model header>"
#endif
#include "ilqr_kernels.cuh"

#ifndef ILQR_FWD_TMA_DEFAULT
#define ILQR_FWD_TMA_DEFAULT 3
#endif
#ifndef ILQR_TP_DEFAULT_MIN_WARPS_PER_SM
#define ILQR_TP_DEFAULT_MIN_WARPS_PER_SM 8
#endif

namespace ilqr {

static int fail(char* err, int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(err, ILQR_ERRLEN, fmt, ap);
    va_end(ap);
    return code;
}
#define CU(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            return fail(err, ILQR_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

struct Impl {
    Params P{};
    int device = 0;
    cudaStream_t stream = nullptr;     /* stream in use */
    cudaStream_t own_stream = nullptr; /* created with the handle */
    cudaStream_t side = nullptr;       /* side branch for k_refill (continuous batching) */
    cudaEvent_t ev_fork[2]{}, ev_join[2]{};
    bool refill_inflight[2] = {false, false};
    std::vector<void*> allocs;
    double* stage = nullptr;      /* device staging buffer for layout changes */
    double *roll_x = nullptr, *roll_u = nullptr; /* ilqr_rollout scratch */
    size_t stage_elems = 0;
    int32_t* h_active = nullptr;  /* pinned mirror of the active counters: 2 graphs x 8 ticks */
    struct GraphPair { cudaGraphExec_t g[2] = {nullptr, nullptr}; };
    std::map<long long, GraphPair> graphs; /* key = mode * 2^32 + blocks in the grid */
    ilqr_nccl::comm_t comm = nullptr;      /* ilqr_comm_init: the final gather's communicator */
    int comm_ranks = 0;
    int compact_fill_pct = 50;             /* compact when the running problems fill at most this percentage of the grid */
    long long compact_min_blocks = 0;      /* drain compaction never shrinks the grid below this many 32-problem blocks */
    int64_t compactions = 0;
    Job* d_job = nullptr;
    int32_t* d_next = nullptr;
    void* hs_buf = nullptr;       /* grow-only device arena for ilqr_solve_stream_host */
    size_t hs_bytes = 0;
    cudaEvent_t ev[8]{};
    std::vector<cudaEvent_t> pool; /* profiling events, 2 per timed launch, resolved after the solve */
    std::vector<int> pool_kind;
    size_t pool_used = 0;
    bool profiling = false;
    int64_t pt_acc = 0;
    int64_t ticks = 0, launches = 0, problem_ticks = 0;
    int num_sms = 148;
    long long tp_min_blocks = 0; /* grids of at least this many 32-problem warps take k_linback_tp */
    long long ft_min_blocks = 0; /* ... and k_forward_tp */
    bool fwd_wp = true;              /* wide models: k_forward_wp (warp per problem and trial) when the model allows; ILQR_FWD_WP=0 disables */
    int fwd_tma = 0;                 /* k_forward_tma (ONE ring per CTA) instead of k_forward: 1 = fed by TMA bulk copies, 2 = by
                                      * cp.async, 3 = by cp.async on dense grids only (default) */
    long long lb_dense_min_blocks = 0; /* grids of at least this many blocks take the two-CTAs-per-SM k_linback */
    double kernel_ms[3] = {0, 0, 0};
    int64_t kernel_launches[3] = {0, 0, 0};
    int rows() const { return (P.T - 1) * CS + CT; }
};

template <typename T>
static int dev_alloc(Impl* im, T** out, size_t count, char* err) {
    void* p = nullptr;
    const size_t bytes = (count > 0 ? count : 1) * sizeof(T);
    CU(cudaMalloc(&p, bytes));
    im->allocs.push_back(p);
    CU(cudaMemsetAsync(p, 0, bytes, im->stream));
    *out = (T*)p;
    return 0;
}

static int count_trials(const ilqr_options& o) { /* src/forward_pass.jl:28-29 */
    int n = 0;
    double a = 1.0;
    while (a >= o.min_step_size && n < 25) { ++n; a *= 0.5; }
    return n;
}

static int check_options(const ilqr_options* o, char* err) {
    if (!o) return fail(err, ILQR_EINVAL, "options is NULL");
    if (o->line_search != ILQR_LINE_SEARCH_ARMIJO && o->line_search != ILQR_LINE_SEARCH_NONE)
        return fail(err, ILQR_EINVAL, "line_search must be ILQR_LINE_SEARCH_ARMIJO or ILQR_LINE_SEARCH_NONE");
    if (o->max_iterations < 0 || o->max_dual_updates < 0) return fail(err, ILQR_EINVAL, "negative iteration limits");
    return 0;
}

static void drop_graphs(struct Impl* im);

template <typename T>
__global__ void k_fill(T* p, T v, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

static void plugin_destroy(void* impl) {
    Impl* im = (Impl*)impl;
    if (!im) return;
    cudaSetDevice(im->device);
    if (im->stream) cudaStreamSynchronize(im->stream);
    if (im->comm) { if (const ilqr_nccl::Api* n = ilqr_nccl::api(nullptr)) n->CommDestroy(im->comm); }
    drop_graphs(im);
    if (im->hs_buf) cudaFree(im->hs_buf);
    for (void* p : im->allocs) cudaFree(p);
    if (im->h_active) cudaFreeHost(im->h_active);
    for (auto& e : im->ev) if (e) cudaEventDestroy(e);
    for (auto& e : im->pool) if (e) cudaEventDestroy(e);
    for (int i = 0; i < 2; ++i) { if (im->ev_fork[i]) cudaEventDestroy(im->ev_fork[i]); if (im->ev_join[i]) cudaEventDestroy(im->ev_join[i]); }
    if (im->side) cudaStreamDestroy(im->side);
    if (im->own_stream) cudaStreamDestroy(im->own_stream);
    delete im;
}

static int plugin_create_inner(Impl* im, const ilqr_desc* desc, const ilqr_options* opt, char* err) {
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(err, ILQR_ECUDA, "no CUDA device available (%s); this engine has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    if (desc->device < 0 || desc->device >= ndev) return fail(err, ILQR_EINVAL, "device %d out of range (0..%d)", desc->device, ndev - 1);
    im->device = desc->device;
    CU(cudaSetDevice(im->device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, im->device));
    if (prop.major < 10) return fail(err, ILQR_ECUDA, "device %d is sm_%d%d; this engine is built for sm_100a only", im->device, prop.major, prop.minor);
    CU(cudaStreamCreateWithFlags(&im->own_stream, cudaStreamNonBlocking));
    im->stream = im->own_stream;
    CU(cudaStreamCreateWithFlags(&im->side, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
        CU(cudaEventCreateWithFlags(&im->ev_fork[i], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&im->ev_join[i], cudaEventDisableTiming));
    }
    for (auto& ev : im->ev) CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    CU(cudaMallocHost((void**)&im->h_active, 16 * sizeof(int32_t)));
#if !ILQR_LARGE
    if (BK_FUSED) {
        CU(cudaFuncSetAttribute(k_linback<LB_WARPS, 1, d1(LB_STAGES)>, cudaFuncAttributeMaxDynamicSharedMemorySize, LbGeom<d1(LB_STAGES)>::SMEM_BYTES));
        if (LB_DENSE_OK) {
            CU(cudaFuncSetAttribute(k_linback<LB_DENSE_WARPS, 2, d1(LB_DENSE_STAGES)>, cudaFuncAttributeMaxDynamicSharedMemorySize, LbGeom<d1(LB_DENSE_STAGES)>::SMEM_BYTES));
            CU(cudaFuncSetAttribute(k_linback<LB_DENSE_WARPS, 2, d1(LB_DENSE_STAGES)>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        }
    }
#endif
#if ILQR_LARGE
    CU(cudaFuncSetAttribute(k_backward, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RL_SMEM_BYTES));
#if ILQR_FWD_WP
    CU(cudaFuncSetAttribute(k_forward_wp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WP_DYN_SMEM));
#endif
#endif
    if (FWD_SMEM_BYTES > 0) {
        CU(cudaFuncSetAttribute(k_forward<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM_BYTES));
        CU(cudaFuncSetAttribute(k_forward<FWD_DENSE_CTAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM_BYTES));
    }
    CU(cudaDeviceGetAttribute(&im->num_sms, cudaDevAttrMultiProcessorCount, im->device));
    /* k_linback_tp pays once the machine holds several independent warps per SM sub-partition (measured cross-over
     * in profiles/README.md); ILQR_TP_MIN_BLOCKS overrides the threshold (0 = always, a huge value = never) */
    im->tp_min_blocks = (long long)ILQR_TP_DEFAULT_MIN_WARPS_PER_SM * im->num_sms;
    if (const char* e = getenv("ILQR_TP_MIN_BLOCKS")) im->tp_min_blocks = atoll(e);
    im->compact_min_blocks = im->num_sms; /* one 32-problem block per SM: below that a tick is pure latency anyway */
    if (const char* e = getenv("ILQR_COMPACT_MIN_BLOCKS")) im->compact_min_blocks = atoll(e); /* huge value = no compaction */
    if (ILQR_LARGE) im->compact_min_blocks = (1ll << 60); /* one CTA per problem there; the staged Jacobian blocks do not move */
    if (const char* e = getenv("ILQR_COMPACT_FILL")) { const int v = atoi(e); if (v >= 1 && v <= 95) im->compact_fill_pct = v; }
    /* k_forward_tp evaluates one step size per launch: 17 % more ticks per solve than k_forward's two, for no gain
     * per tick (both are DRAM-bound at ~10 ns per problem and tick) -- off unless asked for */
    im->ft_min_blocks = 1LL << 40;
    if (const char* e = getenv("ILQR_FT_MIN_BLOCKS")) im->ft_min_blocks = atoll(e);
#if !ILQR_LARGE
    /* default (3): the cp.async-fed shared ring on dense grids (measured -5..-9 % forward time from 14208 slots up, +6 % at
     * 4096), private rings below; ILQR_FWD_TMA = 0 private rings always, 1 TMA-fed ring always, 2 cp.async-fed ring always */
    im->fwd_tma = FWT_OK ? ILQR_FWD_TMA_DEFAULT : 0;
    if (const char* e = getenv("ILQR_FWD_TMA")) im->fwd_tma = FWT_OK ? atoi(e) : 0;
    if (FWT_OK) {
        CU(cudaFuncSetAttribute(k_forward_tma<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FWT_SMEM_BYTES));
        CU(cudaFuncSetAttribute(k_forward_tma<FWD_DENSE_CTAS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FWT_SMEM_BYTES));
        CU(cudaFuncSetAttribute(k_forward_tma<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FWT_SMEM_BYTES));
        CU(cudaFuncSetAttribute(k_forward_tma<FWD_DENSE_CTAS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FWT_SMEM_BYTES));
    }
#endif
    if (const char* e = getenv("ILQR_FWD_WP")) im->fwd_wp = atoi(e) != 0;
    im->lb_dense_min_blocks = (long long)im->num_sms * 3 / 2; /* beyond 1.5 CTAs per SM */
    if (const char* e = getenv("ILQR_LB_DENSE_MIN_BLOCKS")) im->lb_dense_min_blocks = atoll(e);
#if !ILQR_LARGE
    if (!FT_OK) im->ft_min_blocks = 1LL << 40; /* a step's rows do not fit the per-warp ring */
    else {
        CU(cudaFuncSetAttribute(k_forward_tp, cudaFuncAttributeMaxDynamicSharedMemorySize, FT_SMEM_BYTES));
        CU(cudaFuncSetAttribute(k_forward_tp, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    }
#endif

    Params& P = im->P;
    P.T = desc->T;
    P.B = desc->batch;
    P.Bp = (desc->batch + 31) / 32 * 32;
    P.cap = desc->history_cap > 0 ? desc->history_cap : 1000;
    P.o = *opt;
    P.n_alpha = count_trials(*opt);
    const size_t Bp = P.Bp, T = P.T;
    Dev& d = P.d;
    int rc = 0;
#define A(ptr, count) if ((rc = dev_alloc(im, &d.ptr, (size_t)(count) * Bp, err)) != 0) return rc
#define A2(ptr, count) if ((rc = dev_alloc(im, &d.ptr, (size_t)(count), err)) != 0) return rc
    A(xb, T * N); A(ub, (T - 1) * M); A(xc, T * N); A(uc, (T - 1) * M); A(w, T * NP);
#if ILQR_LARGE
    if (JAC_CONST) { if ((rc = dev_alloc(im, &d.fx, (size_t)JAC_BLOCK, err)) != 0) return rc; } /* one block for the batch */
    else A(fx, (T - 1) * (size_t)JAC_BLOCK); /* problem-major staged Jacobian blocks (ilqr_kernels.cuh) */
    A(fu, 1);
#else
    A(fx, (T - 1) * N * N); A(fu, (T - 1) * N * M);
#endif
    A(gx, T * N); A(gu, (T - 1) * M); A(gxx, T * N * N);
    A(guu, (HACC || HACC_L) ? 1 : (T - 1) * M * M); A(gux, (HACC || HACC_L) ? 1 : (T - 1) * M * N); A(hacc, (HACC || HACC_L) ? NH : 1);
    A(K, (T - 1) * M * N); A(k, (T - 1) * M); A(Lx, (T - 1) * N); A(Lu, (T - 1) * M);
    const size_t rows = (T - 1) * CS + CT;
    A(c, rows); A(lam, rows); A(rho, rows); A(act, rows);

    A(J, 1); A(obj_prev, 1); A(viol, 1); A(alpha, 1); A(gnorm, 1); A(dgp, 1); A(ls_base, 1);
    A(status, 1); A(iters, 1); A(iters0, 1); A(outer, 1); A(it, 1); A(phase, 1); A(kind, 1); A(inner_done, 1); A(flags, 1);
    A(h_cost, P.cap); A(h_gnorm, P.cap); A(h_viol, P.cap); A(h_alpha, P.cap); A(h_outer, P.cap); A(h_status, P.cap);
#undef A
    if ((rc = dev_alloc(im, &d.active, 8, err)) != 0) return rc;
    A2(pid, Bp); A2(done_list, 4 * Bp); A2(done_count, 4); A2(pending, Bp); A2(refilling, Bp); A2(mpc_step, Bp); A2(mpc_iters, Bp);
    A2(cmp_src, Bp); A2(cmp_dst, Bp); A2(cmp_n, 2);
    {   /* everything that belongs to a slot and is live between two ticks (k_compact_move copies these columns);
         * not in the list: gx/gu (written and read within one tick) */
        std::vector<MoveEntry> mv;
        auto add = [&](void* base, size_t rows_, int elsize) { if (rows_ > 0) mv.push_back(MoveEntry{(char*)base, (int32_t)rows_, (int32_t)elsize}); };
        add(d.xb, T * N, 8); add(d.ub, (T - 1) * M, 8); add(d.xc, T * N, 8); add(d.uc, (T - 1) * M, 8); add(d.w, T * NP, 8);
        if (!ILQR_LARGE) { add(d.fx, (T - 1) * N * N, 8); add(d.fu, (T - 1) * N * M, 8); } /* wide models: not column-organised, no compaction */
        add(d.gxx, T * N * N, 8);
        if (!(HACC || HACC_L)) { add(d.guu, (T - 1) * M * M, 8); add(d.gux, (T - 1) * M * N, 8); } else add(d.hacc, NH, 8);
        add(d.K, (T - 1) * M * N, 8); add(d.k, (T - 1) * M, 8); add(d.Lx, (T - 1) * N, 8); add(d.Lu, (T - 1) * M, 8);
        add(d.c, rows, 8); add(d.lam, rows, 8); add(d.rho, rows, 8); add(d.act, rows, 1);
        for (double* q : {d.J, d.obj_prev, d.viol, d.alpha, d.gnorm, d.dgp}) add(q, 1, 8);
        for (int32_t* q : {d.ls_base, d.status, d.iters, d.iters0, d.outer, d.it, d.phase, d.kind, d.inner_done, d.pid, d.pending,
                           d.refilling, d.mpc_step, d.mpc_iters}) add(q, 1, 4);
        add(d.flags, 1, 4);
        add(d.h_cost, P.cap, 8); add(d.h_gnorm, P.cap, 8); add(d.h_viol, P.cap, 8); add(d.h_alpha, P.cap, 8);
        add(d.h_outer, P.cap, 4); add(d.h_status, P.cap, 1);
        MoveEntry* dmv = nullptr;
        if ((rc = dev_alloc(im, &dmv, mv.size(), err)) != 0) return rc;
        CU(cudaMemcpyAsync(dmv, mv.data(), sizeof(MoveEntry) * mv.size(), cudaMemcpyHostToDevice, im->stream));
        CU(cudaStreamSynchronize(im->stream)); /* mv is a stack object */
        d.mv = dmv;
        d.n_mv = (int32_t)mv.size();
    }
    if ((rc = dev_alloc(im, &im->d_job, 1, err)) != 0) return rc;
    if ((rc = dev_alloc(im, &im->d_next, 1, err)) != 0) return rc;
    P.job = im->d_job;
    /* staging buffer: largest host-layout array that crosses the ABI */
    size_t mx = T * N > 8 ? T * N : 8; /* at least the six rows ilqr_get_stats stages at once */
    const size_t cands[] = {(T - 1) * (size_t)M, T * (size_t)NP, rows, (T - 1) * (size_t)M * N, (size_t)P.cap};
    for (size_t v : cands) mx = v > mx ? v : mx;
    im->stage_elems = mx * Bp;
    if ((rc = dev_alloc(im, &im->stage, im->stage_elems, err)) != 0) return rc;
    /* creation state: objective = Inf, step_size = 1, penalty = 1, active set = 1
     * (src/data/solver.jl:37-39, src/augmented_lagrangian.jl:17-22) */
    const unsigned tb = 256;
#if ILQR_LARGE
    if (JAC_CONST) k_jac_const_init<<<1, 1, 0, im->stream>>>(P);
#endif
    k_fill<double><<<(unsigned)((Bp + tb - 1) / tb), tb, 0, im->stream>>>(d.J, HUGE_VAL, Bp);
    k_fill<double><<<(unsigned)((Bp + tb - 1) / tb), tb, 0, im->stream>>>(d.alpha, 1.0, Bp);
    if (rows > 0) {
        k_fill<double><<<(unsigned)((rows * Bp + tb - 1) / tb), tb, 0, im->stream>>>(d.rho, 1.0, rows * Bp);
        k_fill<uint8_t><<<(unsigned)((rows * Bp + tb - 1) / tb), tb, 0, im->stream>>>(d.act, (uint8_t)1, rows * Bp);
    }
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(im->stream));
    return 0;
}

static int plugin_create(const ilqr_desc* desc, const ilqr_options* opt, void** impl, char* err) {
    if (!desc || !impl) return fail(err, ILQR_EINVAL, "desc/out is NULL");
    int rc = check_options(opt, err);
    if (rc) return rc;
    if (desc->n != N || desc->m != M || desc->p != NP || desc->c_s != CS || desc->c_T != CT)
        return fail(err, ILQR_EINVAL, "dimension mismatch: desc (n=%d m=%d p=%d c_s=%d c_T=%d) vs model '%s' (n=%d m=%d p=%d c_s=%d c_T=%d)",
                    desc->n, desc->m, desc->p, desc->c_s, desc->c_T, ILQR_MODEL_NAME, N, M, NP, CS, CT);
    if (desc->T < 2) return fail(err, ILQR_EINVAL, "T must be >= 2 (got %d)", desc->T);
    if (desc->batch < 1) return fail(err, ILQR_EINVAL, "batch must be >= 1 (got %d)", desc->batch);
    Impl* im = new Impl();
    rc = plugin_create_inner(im, desc, opt, err);
    if (rc) {
        plugin_destroy(im);
        return rc;
    }
    *impl = im;
    return 0;
}

static int plugin_set_options(void* impl, const ilqr_options* opt, char* err) {
    Impl* im = (Impl*)impl;
    int rc = check_options(opt, err);
    if (rc) return rc;
    im->P.o = *opt;
    im->P.n_alpha = count_trials(*opt);
    drop_graphs(im); /* the options are baked into the captured kernel parameters */
    return 0;
}

/* ---- layout changes across the ABI ------------------------------------------------- */
template <typename TD>
static int upload(Impl* im, const double* host, TD* dev, size_t rows, char* err, bool device_in = false) {
    if (!host) return fail(err, ILQR_EINVAL, "NULL input buffer");
    if (rows == 0) return 0;
    CU(cudaSetDevice(im->device));
    const Params& P = im->P;
    const double* src = host;
    if (!device_in) {
        CU(cudaMemcpyAsync(im->stage, host, sizeof(double) * rows * P.B, cudaMemcpyHostToDevice, im->stream));
        src = im->stage;
    }
    dim3 grid((unsigned)((rows + 31) / 32), (unsigned)((P.B + 31) / 32)), block(32, 8);
    k_to_soa<double, TD><<<grid, block, 0, im->stream>>>(src, dev, P.B, P.Bp, (int)rows);
    im->launches += 1;
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(im->stream));
    return 0;
}
/* stage_off: offset (in doubles) into the staging buffer, so that several small downloads can be in flight before ONE
 * synchronisation (sync = false on all but the last) */
template <typename TS, typename TH>
static int download(Impl* im, const TS* dev, TH* host, size_t rows, bool device_out, char* err, size_t stage_off = 0, bool sync = true) {
    if (!host || rows == 0) return 0;
    CU(cudaSetDevice(im->device));
    const Params& P = im->P;
    dim3 grid((unsigned)((rows + 31) / 32), (unsigned)((P.B + 31) / 32)), block(32, 8);
    im->launches += 1;
    if (device_out) {
        k_from_soa<TS, TH><<<grid, block, 0, im->stream>>>(dev, host, P.B, P.Bp, (int)rows);
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(im->stream));
        return 0;
    }
    TH* st = (TH*)(im->stage + stage_off);
    k_from_soa<TS, TH><<<grid, block, 0, im->stream>>>(dev, st, P.B, P.Bp, (int)rows);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(host, st, sizeof(TH) * rows * P.B, cudaMemcpyDeviceToHost, im->stream));
    if (sync) CU(cudaStreamSynchronize(im->stream));
    return 0;
}

static int plugin_initialize_controls(void* impl, const double* u, int device_in, char* err) {
    Impl* im = (Impl*)impl;
    return upload(im, u, im->P.d.ub, (size_t)(im->P.T - 1) * M, err, device_in != 0);
}
static int plugin_initialize_states(void* impl, const double* x, int device_in, char* err) {
    Impl* im = (Impl*)impl;
    return upload(im, x, im->P.d.xb, (size_t)im->P.T * N, err, device_in != 0);
}
static int plugin_set_parameters(void* impl, const double* w, char* err) {
    Impl* im = (Impl*)impl;
    if (NP == 0) return 0;
    return upload(im, w, im->P.d.w, (size_t)im->P.T * NP, err);
}

static int plugin_rollout(void* impl, const double* x1, const double* u, double* x_out, char* err) {
    Impl* im = (Impl*)impl;
    if (!x1 || !u || !x_out) return fail(err, ILQR_EINVAL, "NULL buffer");
    const Params& P = im->P;
    CU(cudaSetDevice(im->device));
    if (!im->roll_x) { /* scratch of ilqr_rollout, allocated on first use and kept with the handle */
        CU(cudaMalloc((void**)&im->roll_x, sizeof(double) * (size_t)P.T * N * P.Bp));
        im->allocs.push_back(im->roll_x);
        CU(cudaMalloc((void**)&im->roll_u, sizeof(double) * (size_t)(P.T - 1) * d1(M) * P.Bp));
        im->allocs.push_back(im->roll_u);
    }
    double *dx = im->roll_x, *du = im->roll_u;
    int rc = upload(im, x1, dx, (size_t)N, err);
    if (!rc) rc = upload(im, u, du, (size_t)(P.T - 1) * M, err);
    if (!rc) {
        k_rollout<<<(P.B + 127) / 128, 128, 0, im->stream>>>(P, dx, du);
        cudaError_t e3 = cudaGetLastError();
        if (e3 != cudaSuccess) rc = fail(err, ILQR_ECUDA, "k_rollout launch failed: %s", cudaGetErrorString(e3));
    }
    if (!rc) rc = download(im, dx, x_out, (size_t)P.T * N, false, err);
    return rc;
}

/* ---- the lock-step solve loop -------------------------------------------------------- */
static const int REFILL_CTAS = 64; /* k_refill grid: CTAs striding over the slots that finished in a tick */

static int launch_tick(Impl* im, unsigned nblk, char* err) {
    Params& P = im->P;
    const dim3 fb(32, FWD_TRIAL_WARPS + 2);
    const bool prof = im->profiling;
#define TIMED(kindex, launch)                                                   \
    do {                                                                        \
        cudaEvent_t e0_ = nullptr, e1_ = nullptr;                               \
        if (prof) {                                                             \
            if (im->pool_used + 2 > im->pool.size()) {                          \
                CU(cudaEventCreate(&e0_)); im->pool.push_back(e0_);             \
                CU(cudaEventCreate(&e1_)); im->pool.push_back(e1_);             \
            }                                                                   \
            e0_ = im->pool[im->pool_used]; e1_ = im->pool[im->pool_used + 1];   \
            im->pool_used += 2;                                                 \
            im->pool_kind.push_back(kindex);                                    \
            CU(cudaEventRecord(e0_, im->stream));                               \
        }                                                                       \
        launch;                                                                 \
        CU(cudaGetLastError());                                                 \
        if (prof) CU(cudaEventRecord(e1_, im->stream));                         \
        im->launches += 1;                                                      \
    } while (0)
    if (P.mode == MODE_STREAM && im->refill_inflight[P.tick & 1]) { /* join the k_refill of tick-2 */
        CU(cudaStreamWaitEvent(im->stream, im->ev_join[P.tick & 1], 0));
        im->refill_inflight[P.tick & 1] = false;
    }
#if !ILQR_LARGE
    if ((long long)nblk >= im->ft_min_blocks) {
        TIMED(0, (k_forward_tp<<<nblk, 32, FT_SMEM_BYTES, im->stream>>>(P)));
    } else
#elif ILQR_FWD_WP
    if (im->fwd_wp) { /* wide unconstrained model with table-mode dynamics: one CTA per problem */
        TIMED(0, (k_forward_wp<<<(unsigned)P.B, fb, WP_DYN_SMEM, im->stream>>>(P)));
    } else
#endif
#if !ILQR_LARGE
    if (im->fwd_tma == 1) {
        if (nblk > 2u * (unsigned)im->num_sms) {
            TIMED(0, (k_forward_tma<FWD_DENSE_CTAS, true><<<nblk, fb, FWT_SMEM_BYTES, im->stream>>>(P)));
        } else {
            TIMED(0, (k_forward_tma<1, true><<<nblk, fb, FWT_SMEM_BYTES, im->stream>>>(P)));
        }
    } else if (im->fwd_tma == 2 || (im->fwd_tma == 3 && nblk > 2u * (unsigned)im->num_sms)) {
        if (nblk > 2u * (unsigned)im->num_sms) {
            TIMED(0, (k_forward_tma<FWD_DENSE_CTAS, false><<<nblk, fb, FWT_SMEM_BYTES, im->stream>>>(P)));
        } else {
            TIMED(0, (k_forward_tma<1, false><<<nblk, fb, FWT_SMEM_BYTES, im->stream>>>(P)));
        }
    } else
#endif
    if (nblk > 2u * (unsigned)im->num_sms) { /* more than two CTAs per SM: the register-capped instantiation */
        TIMED(0, (k_forward<FWD_DENSE_CTAS><<<nblk, fb, FWD_SMEM_BYTES, im->stream>>>(P)));
    } else {
        TIMED(0, (k_forward<1><<<nblk, fb, FWD_SMEM_BYTES, im->stream>>>(P)));
    }
#if !ILQR_LARGE
    if ((long long)nblk >= im->tp_min_blocks) {
        TIMED(2, (k_linback_tp<<<nblk, 32, 0, im->stream>>>(P)));
    } else
#endif
#if !ILQR_LARGE
    if (BK_FUSED && LB_DENSE_OK && (long long)nblk >= im->lb_dense_min_blocks) { /* two CTAs per SM */
        TIMED(2, (k_linback<LB_DENSE_WARPS, 2, d1(LB_DENSE_STAGES)><<<nblk, dim3(32, LB_DENSE_WARPS), LbGeom<d1(LB_DENSE_STAGES)>::SMEM_BYTES, im->stream>>>(P)));
    } else if (BK_FUSED) {
        TIMED(2, (k_linback<LB_WARPS, 1, d1(LB_STAGES)><<<nblk, dim3(32, LB_WARPS), LbGeom<d1(LB_STAGES)>::SMEM_BYTES, im->stream>>>(P)));
    } else
#endif
    {
        const size_t threads = (size_t)P.T * P.Bp; /* (the unfused pair always covers the whole slot array) */
#if ILQR_LARGE
        if (LIN_COOP) { /* + the dynamics Jacobians by one CTA per (problem, step): listed with gradients! (kind 1) */
            TIMED(1, ([&] {
                k_linearize<<<(unsigned)((threads + 63) / 64), 64, 0, im->stream>>>(P);
                k_linearize_jac<<<(unsigned)(im->num_sms * LC_CTAS_PER_SM), LC_THREADS, 0, im->stream>>>(P);
                im->launches += 1;
            }()));
        } else {
            TIMED(1, (k_linearize<<<(unsigned)((threads + 63) / 64), 64, 0, im->stream>>>(P)));
        }
        TIMED(2, (k_backward<<<(unsigned)P.B, RL_CTA_THREADS, RL_SMEM_BYTES, im->stream>>>(P)));
#else
        TIMED(1, (k_linearize<<<(unsigned)((threads + 127) / 128), 128, 0, im->stream>>>(P)));
        TIMED(2, (k_backward<<<P.Bp / 32, 32, 0, im->stream>>>(P)));
#endif
    }
    if (P.mode == MODE_STREAM) {
        /* side branch: k_refill(tick) overlaps the kernels of tick+1 and must be complete before k_forward(tick+2) */
        const int par = P.tick & 1;
        CU(cudaEventRecord(im->ev_fork[par], im->stream));
        CU(cudaStreamWaitEvent(im->side, im->ev_fork[par], 0));
        k_refill<<<REFILL_CTAS, 128, 0, im->side>>>(P);
        CU(cudaGetLastError());
        CU(cudaEventRecord(im->ev_join[par], im->side));
        im->refill_inflight[par] = true;
        im->launches += 1;
    }
#undef TIMED
    return 0;
}

/* One CUDA graph = GRAPH_TICKS lock-step ticks + the copies of the per-tick "still running" counters into pinned
 * host memory.  Two instances alternate (they differ only in the host slots they report into), so the host can look
 * at the counters of graph g while graph g+1 is already running: the GPU never waits for the host.  Graphs are cached
 * per (mode, grid size): drain compaction shrinks the grid a handful of times per streamed job. */
static const int GRAPH_TICKS = 8;

static void drop_graphs(Impl* im) {
    for (auto& kv : im->graphs)
        for (int g = 0; g < 2; ++g)
            if (kv.second.g[g]) cudaGraphExecDestroy(kv.second.g[g]);
    im->graphs.clear();
}

static int get_graphs(Impl* im, unsigned nblk, cudaGraphExec_t** out, char* err) {
    Params& P = im->P;
    const long long key = ((long long)P.mode << 32) | ((long long)(P.pause_outer ? 1 : 0) << 31) | nblk;
    auto it = im->graphs.find(key);
    if (it != im->graphs.end()) { *out = it->second.g; return 0; }
    Impl::GraphPair pair;
    const long long launches_before = im->launches;
    for (int g = 0; g < 2; ++g) {
        cudaGraph_t graph = nullptr;
        im->refill_inflight[0] = im->refill_inflight[1] = false;
        CU(cudaStreamBeginCapture(im->stream, cudaStreamCaptureModeThreadLocal));
        int rc = 0;
        for (int j = 0; j < GRAPH_TICKS && !rc; ++j) {
            P.tick = j;
            rc = launch_tick(im, nblk, err);
            if (!rc) {
                cudaError_t e = cudaMemcpyAsync(&im->h_active[g * GRAPH_TICKS + j], &P.d.active[j], sizeof(int32_t),
                                                cudaMemcpyDeviceToHost, im->stream);
                if (e != cudaSuccess) rc = fail(err, ILQR_ECUDA, "capture memcpy failed: %s", cudaGetErrorString(e));
            }
        }
        for (int par = 0; par < 2; ++par) /* the graph ends when its last two k_refill branches have joined */
            if (im->refill_inflight[par]) {
                cudaStreamWaitEvent(im->stream, im->ev_join[par], 0);
                im->refill_inflight[par] = false;
            }
        cudaError_t e = cudaStreamEndCapture(im->stream, &graph);
        if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
        if (e != cudaSuccess) return fail(err, ILQR_ECUDA, "cudaStreamEndCapture failed: %s", cudaGetErrorString(e));
        e = cudaGraphInstantiate(&pair.g[g], graph, 0);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess) return fail(err, ILQR_ECUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(e));
    }
    im->launches = launches_before; /* capture does not launch */
    *out = im->graphs.emplace(key, pair).first->second.g;
    return 0;
}

static int resolve_profiling(Impl* im, char* err) {
    for (size_t i = 0; i < im->pool_kind.size(); ++i) {
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, im->pool[2 * i], im->pool[2 * i + 1]));
        im->kernel_ms[im->pool_kind[i]] += ms;
        im->kernel_launches[im->pool_kind[i]] += 1;
    }
    im->pool_used = 0;
    im->pool_kind.clear();
    return 0;
}

/* Drain compaction (see k_compact_plan): pack the running problems of a streamed job into the lowest slots; the grid
 * shrinks to `new_blocks`.  `last_tick` is the ring index of the tick whose "still running" counter is the bound.
 * Every k_refill branch must have been joined on im->stream. */
static int compact_slots(Impl* im, unsigned old_blocks, int last_tick, char* err) {
    Params& P = im->P;
    const int save = P.tick;
    P.tick = last_tick & 7;
    k_compact_reset<<<1, 1, 0, im->stream>>>(P);
    k_compact_plan<<<(old_blocks * 32 + 255) / 256, 256, 0, im->stream>>>(P, (int)(old_blocks * 32));
    if (getenv("ILQR_COMPACT_DEBUG")) { /* debugging aid: what is about to move, and in which state */
        CU(cudaStreamSynchronize(im->stream));
        int32_t n2[2], act[8];
        CU(cudaMemcpy(n2, P.d.cmp_n, sizeof(n2), cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(act, P.d.active, sizeof(act), cudaMemcpyDeviceToHost));
        const size_t Bp = P.Bp;
        std::vector<int32_t> src(Bp), dst(Bp), pid(Bp), phase(Bp), kind(Bp), lsb(Bp), pend(Bp), refl(Bp), it(Bp), indone(Bp);
        CU(cudaMemcpy(src.data(), P.d.cmp_src, 4 * Bp, cudaMemcpyDeviceToHost)); CU(cudaMemcpy(dst.data(), P.d.cmp_dst, 4 * Bp, cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(pid.data(), P.d.pid, 4 * Bp, cudaMemcpyDeviceToHost)); CU(cudaMemcpy(phase.data(), P.d.phase, 4 * Bp, cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(kind.data(), P.d.kind, 4 * Bp, cudaMemcpyDeviceToHost)); CU(cudaMemcpy(lsb.data(), P.d.ls_base, 4 * Bp, cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(pend.data(), P.d.pending, 4 * Bp, cudaMemcpyDeviceToHost)); CU(cudaMemcpy(refl.data(), P.d.refilling, 4 * Bp, cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(it.data(), P.d.it, 4 * Bp, cudaMemcpyDeviceToHost)); CU(cudaMemcpy(indone.data(), P.d.inner_done, 4 * Bp, cudaMemcpyDeviceToHost));
        int used = 0, used_beyond = 0;
        for (size_t b = 0; b < (size_t)old_blocks * 32; ++b) if (phase[b] != PH_DONE || pend[b] != 0) { ++used; if ((int)b >= act[last_tick & 7]) ++used_beyond; }
        fprintf(stderr, "[compact] grid %u blocks, A=%d, in use %d (beyond A: %d), sources %d, destinations %d\n", old_blocks, act[last_tick & 7], used, used_beyond, n2[0], n2[1]);
        for (int e = 0; e < n2[0]; ++e) {
            const int s_ = src[e], d_ = dst[e];
            fprintf(stderr, "[compact]   pid %d: slot %d -> %d  phase %d kind %d ls_base %d pending %d refilling %d it %d inner_done %d | dst: phase %d pending %d refilling %d pid %d\n",
                    pid[s_], s_, d_, phase[s_], kind[s_], lsb[s_], pend[s_], refl[s_], it[s_], indone[s_], phase[d_], pend[d_], refl[d_], pid[d_]);
        }
    }
    k_compact_move<<<4 * im->num_sms, 128, 0, im->stream>>>(P);
    P.tick = save;
    CU(cudaGetLastError());
    im->launches += 3;
    im->compactions += 1;
    return 0;
}

/* the grid a streamed job shrinks to once `active` problems are left (0 = keep the current one) */
static unsigned shrunk_blocks(const Impl* im, unsigned cur_blocks, int active) {
    if (im->P.mode != MODE_STREAM || (long long)cur_blocks <= im->compact_min_blocks) return 0;
    if ((long long)active * 100 > (long long)cur_blocks * 32 * im->compact_fill_pct) return 0; /* worth it once this share of the grid idles */
    long long nb = ((long long)active + 31) / 32;
    if (nb < im->compact_min_blocks) nb = im->compact_min_blocks;
    if (nb < 1) nb = 1;
    return nb < (long long)cur_blocks ? (unsigned)nb : 0;
}

/* Run lock-step ticks until no problem is running (or the bound is hit).  Graph mode: two graph
 * instances alternate so the GPU never waits for the host; profiling mode: plain launches with events. */
static int run_ticks(Impl* im, long long max_ticks, char* err) {
    Params& P = im->P;
    const bool use_graph = !im->profiling;
    im->refill_inflight[0] = im->refill_inflight[1] = false;
    unsigned nblk = P.Bp / 32;
    cudaGraphExec_t* gx = nullptr;
    if (use_graph) {
        int rc = get_graphs(im, nblk, &gx, err);
        if (rc) return rc;
    }
    long long tick = 0;
    bool finished = false;
    int last_active = P.B;
    im->pt_acc = P.mode == MODE_BATCH ? P.B : 0; /* tick 0 works on every problem; tick i+1 on those still running after tick i */
    auto per_tick = [&](unsigned blocks) {
        const bool two = BK_FUSED || (!ILQR_LARGE && (long long)blocks >= im->tp_min_blocks);
        return (two ? 2 : 3) + (P.mode == MODE_STREAM ? 1 : 0);
    };
    if (use_graph) {
        const long long max_graphs = (max_ticks + GRAPH_TICKS - 1) / GRAPH_TICKS;
        long long g = 0;
        unsigned shrink_to = 0;
        CU(cudaGraphLaunch(gx[0], im->stream));
        CU(cudaEventRecord(im->ev[0], im->stream));
        for (;; ++g) {
            const bool more = g + 1 < max_graphs;
            if (more) { /* keep the GPU fed before looking at graph g's counters */
                if (shrink_to) { /* decided on the counters of graph g-1; runs after graph g, whose refills are joined */
                    int rc = compact_slots(im, nblk, GRAPH_TICKS - 1, err);
                    if (rc) return rc;
                    nblk = shrink_to;
                    shrink_to = 0;
                    rc = get_graphs(im, nblk, &gx, err);
                    if (rc) return rc;
                }
                CU(cudaGraphLaunch(gx[(g + 1) & 1], im->stream));
                CU(cudaEventRecord(im->ev[(g + 1) & 1], im->stream));
            }
            CU(cudaEventSynchronize(im->ev[g & 1]));
            const int32_t* ha = im->h_active + (g & 1) * GRAPH_TICKS;
            for (int j = 0; j < GRAPH_TICKS; ++j) {
                if (last_active > 0) { tick += 1; im->launches += per_tick(nblk); }
                last_active = ha[j];
                im->pt_acc += ha[j];
            }
            if (last_active == 0) { finished = true; break; }
            if (!more) break;
            shrink_to = shrunk_blocks(im, nblk, last_active);
        }
        CU(cudaStreamSynchronize(im->stream));
    } else {
        const int LAG = 2; /* the host runs at most LAG ticks ahead of the last completion it has seen */
        for (; tick < max_ticks; ++tick) {
            if (tick >= LAG) {
                const int slot = (int)((tick - LAG) & 7);
                CU(cudaEventSynchronize(im->ev[slot]));
                if (im->h_active[slot] == 0) { finished = true; break; }
                im->pt_acc += im->h_active[slot];
                /* same drain-compaction policy as the graph path, at multiples of GRAPH_TICKS */
                const unsigned to = (tick % GRAPH_TICKS == 0) ? shrunk_blocks(im, nblk, im->h_active[slot]) : 0;
                if (to) {
                    for (int par = 0; par < 2; ++par)
                        if (im->refill_inflight[par]) {
                            CU(cudaStreamWaitEvent(im->stream, im->ev_join[par], 0));
                            im->refill_inflight[par] = false;
                        }
                    int rc = compact_slots(im, nblk, (int)((tick - 1) & 7), err);
                    if (rc) return rc;
                    nblk = to;
                }
            }
            P.tick = (int)(tick & 7);
            int rc = launch_tick(im, nblk, err);
            if (rc) return rc;
            const int slot = (int)(tick & 7);
            CU(cudaMemcpyAsync(&im->h_active[slot], &P.d.active[slot], sizeof(int32_t), cudaMemcpyDeviceToHost, im->stream));
            CU(cudaEventRecord(im->ev[slot], im->stream));
        }
        CU(cudaStreamSynchronize(im->stream));
        CU(cudaStreamSynchronize(im->side));
        im->refill_inflight[0] = im->refill_inflight[1] = false;
        if (!finished && tick > 0) last_active = im->h_active[(tick - 1) & 7];
        int rc = resolve_profiling(im, err);
        if (rc) return rc;
    }
    im->ticks += tick;
    im->problem_ticks += im->pt_acc;
    if (!finished && last_active != 0)
        return fail(err, ILQR_ESTATE, "solve loop hit its tick bound (%lld) with %d problems still running", max_ticks, last_active);
    return 0;
}

static long long ticks_per_solve_bound(const Params& P) {
    /* every inner solve costs 1 pre-loop tick + <= max_iterations iterations of <= `rounds` ticks each
     * (k_forward evaluates FWD_TRIAL_WARPS step sizes per launch) */
    const long long rounds = (P.n_alpha > 0 ? (P.n_alpha + FWD_TRIAL_WARPS - 1) / FWD_TRIAL_WARPS : 1) + 1; /* + the re-run of a deciding speculative trial */
    const long long inner = (long long)P.o.max_iterations * rounds + 1;
    return (CONSTRAINED ? (long long)P.o.max_dual_updates * inner : inner) + 2;
}

static int plugin_solve(void* impl, char* err) {
    Impl* im = (Impl*)impl;
    Params& P = im->P;
    CU(cudaSetDevice(im->device));
    P.mode = MODE_BATCH;
    CU(cudaMemsetAsync(P.d.active, 0, 8 * sizeof(int32_t), im->stream));
    k_solve_begin<<<(P.B + 127) / 128, 128, 0, im->stream>>>(P);
    CU(cudaGetLastError());
    im->launches += 1;
    return run_ticks(im, ticks_per_solve_bound(P), err);
}

/* ilqr_solve_outer: constrained_ilqr_solve! (src/solve.jl:88-129) one outer iteration at a time, so that the host can run
 * augmented_lagrangian_callback!(solver) (:125) between them.  restart != 0: the solve prologue, then every problem runs
 * until it has terminated or made its next dual update; restart == 0: the parked problems go on.  *n_paused = problems
 * waiting for the next call (0: the solve is complete). */
static int plugin_solve_outer(void* impl, int32_t restart, int32_t* n_paused, char* err) {
    Impl* im = (Impl*)impl;
    Params& P = im->P;
    CU(cudaSetDevice(im->device));
    P.mode = MODE_BATCH;
    P.pause_outer = 1;
    CU(cudaMemsetAsync(P.d.active, 0, 8 * sizeof(int32_t), im->stream));
    CU(cudaMemsetAsync(P.d.cmp_n, 0, 2 * sizeof(int32_t), im->stream));
    if (restart) k_solve_begin<<<(P.B + 127) / 128, 128, 0, im->stream>>>(P);
    else k_resume_paused<<<(P.B + 127) / 128, 128, 0, im->stream>>>(P, P.d.cmp_n);
    CU(cudaGetLastError());
    im->launches += 1;
    int rc = run_ticks(im, ticks_per_solve_bound(P), err);
    if (!rc) {
        k_count_paused<<<(P.B + 127) / 128, 128, 0, im->stream>>>(P, P.d.cmp_n + 1);
        int32_t n = 0;
        CU(cudaMemcpyAsync(&n, P.d.cmp_n + 1, sizeof(int32_t), cudaMemcpyDeviceToHost, im->stream));
        CU(cudaStreamSynchronize(im->stream));
        if (n_paused) *n_paused = n;
    }
    P.pause_outer = 0;
    return rc;
}

/* ilqr_solve_stream: n_total fresh problems through the handle's slots (continuous batching) */
static int plugin_solve_stream(void* impl, int32_t n_total, const double* d_x, const double* d_u, const double* d_w,
                               double* out_x, double* out_u, int32_t* out_iters, uint8_t* out_status, double* out_J,
                               double* out_viol, double* out_alpha, uint32_t* out_flags, char* err) {
    Impl* im = (Impl*)impl;
    Params& P = im->P;
    if (n_total < 1) return fail(err, ILQR_EINVAL, "n_problems must be >= 1");
    if (!d_x || !d_u) return fail(err, ILQR_EINVAL, "NULL input trajectory");
    if (NP > 0 && !d_w) return fail(err, ILQR_EINVAL, "this model has parameters: d_w must not be NULL");
    if (CONSTRAINED && P.o.max_dual_updates <= 0) return fail(err, ILQR_EINVAL, "streaming needs max_dual_updates >= 1");
    CU(cudaSetDevice(im->device));
    Job job{};
    job.n_total = n_total; job.next = im->d_next;
    job.in_x = d_x; job.in_u = d_u; job.in_w = d_w;
    job.out_x = out_x; job.out_u = out_u; job.out_iters = out_iters; job.out_status = out_status;
    job.out_J = out_J; job.out_viol = out_viol; job.out_alpha = out_alpha; job.out_flags = out_flags;
    CU(cudaMemcpyAsync(im->d_job, &job, sizeof(Job), cudaMemcpyHostToDevice, im->stream));
    CU(cudaStreamSynchronize(im->stream)); /* `job` is a stack object */
    P.mode = MODE_STREAM;
    /* the first min(B, n_total) problems enter through the transposing upload */
    const int first = n_total < P.B ? n_total : P.B;
    const int Bsave = P.B;
    int rc = 0;
    P.B = first; /* upload() transposes P.B problems */
    rc = upload(im, d_x, P.d.xb, (size_t)P.T * N, err, true);
    if (!rc) rc = upload(im, d_u, P.d.ub, (size_t)(P.T - 1) * M, err, true);
    if (!rc && NP > 0) rc = upload(im, d_w, P.d.w, (size_t)P.T * NP, err, true);
    P.B = Bsave;
    if (rc) { P.mode = MODE_BATCH; return rc; }
    CU(cudaMemsetAsync(P.d.xc, 0, sizeof(double) * (size_t)P.T * N * P.Bp, im->stream));
    CU(cudaMemsetAsync(P.d.uc, 0, sizeof(double) * (size_t)(P.T - 1) * M * P.Bp, im->stream));
    CU(cudaMemsetAsync(P.d.active, 0, 8 * sizeof(int32_t), im->stream));
    k_stream_begin<<<(P.Bp + 127) / 128, 128, 0, im->stream>>>(P);
    CU(cudaGetLastError());
    im->launches += 1;
    const long long rounds = ((long long)n_total + P.B - 1) / P.B + 1;
    rc = run_ticks(im, rounds * ticks_per_solve_bound(P), err);
    P.mode = MODE_BATCH;
    return rc;
}

/* ilqr_mpc_run: n_steps receding-horizon re-solves of every problem, each problem at its own pace */
static int plugin_mpc_run(void* impl, int32_t n_steps, double* d_applied_u, double* d_x_next, int32_t* d_total_iterations, char* err) {
    Impl* im = (Impl*)impl;
    Params& P = im->P;
    if (n_steps < 1) return fail(err, ILQR_EINVAL, "n_steps must be >= 1");
    if (CONSTRAINED && P.o.max_dual_updates <= 0) return fail(err, ILQR_EINVAL, "ilqr_mpc_run needs max_dual_updates >= 1");
    CU(cudaSetDevice(im->device));
    Job job{};
    job.mpc_steps = n_steps; job.mpc_u = d_applied_u; job.mpc_x = d_x_next; job.next = im->d_next;
    CU(cudaMemcpyAsync(im->d_job, &job, sizeof(Job), cudaMemcpyHostToDevice, im->stream));
    CU(cudaStreamSynchronize(im->stream));
    P.mode = MODE_MPC;
    CU(cudaMemsetAsync(P.d.active, 0, 8 * sizeof(int32_t), im->stream));
    k_mpc_begin<<<(P.B + 127) / 128, 128, 0, im->stream>>>(P);
    CU(cudaGetLastError());
    im->launches += 1;
    int rc = run_ticks(im, (long long)n_steps * (ticks_per_solve_bound(P) + 1) + 2, err);
    P.mode = MODE_BATCH;
    if (!rc && d_total_iterations)
        CU(cudaMemcpyAsync(d_total_iterations, P.d.mpc_iters, sizeof(int32_t) * P.B, cudaMemcpyDeviceToDevice, im->stream));
    CU(cudaStreamSynchronize(im->stream));
    return rc;
}

/* ilqr_solve_stream_host: the same job with HOST buffers (H2D of the inputs, D2H of the results inside the call) */
static int plugin_solve_stream_host(void* impl, int32_t n, const double* x, const double* u, const double* w, double* x_out,
                                    double* u_out, int32_t* iters, uint8_t* status, double* J, double* viol, double* alpha,
                                    uint32_t* flags, char* err) {
    Impl* im = (Impl*)impl;
    const Params& P = im->P;
    if (n < 1) return fail(err, ILQR_EINVAL, "n_problems must be >= 1");
    if (!x || !u) return fail(err, ILQR_EINVAL, "NULL input trajectory");
    CU(cudaSetDevice(im->device));
    const size_t nx = (size_t)n * P.T * N, nu = (size_t)n * (P.T - 1) * M, nw = (size_t)n * P.T * NP;
    auto al = [](size_t b) { return (b + 255) / 256 * 256; };
    const size_t sizes[] = {al(nx * 8), al(nu * 8), al(nw * 8), al(nx * 8), al(nu * 8), al((size_t)n * 4), al((size_t)n),
                            al((size_t)n * 8), al((size_t)n * 8), al((size_t)n * 8), al((size_t)n * 4)};
    size_t total = 0;
    for (size_t v : sizes) total += v;
    if (total > im->hs_bytes) {
        if (im->hs_buf) cudaFree(im->hs_buf);
        im->hs_buf = nullptr; im->hs_bytes = 0;
        CU(cudaMalloc(&im->hs_buf, total));
        im->hs_bytes = total;
    }
    char* base = (char*)im->hs_buf;
    void* p[11];
    size_t off = 0;
    for (int i = 0; i < 11; ++i) { p[i] = base + off; off += sizes[i]; }
    CU(cudaMemcpyAsync(p[0], x, nx * 8, cudaMemcpyHostToDevice, im->stream));
    CU(cudaMemcpyAsync(p[1], u, nu * 8, cudaMemcpyHostToDevice, im->stream));
    if (NP > 0) {
        if (!w) return fail(err, ILQR_EINVAL, "this model has parameters: w must not be NULL");
        CU(cudaMemcpyAsync(p[2], w, nw * 8, cudaMemcpyHostToDevice, im->stream));
    }
    int rc = plugin_solve_stream(impl, n, (double*)p[0], (double*)p[1], NP > 0 ? (double*)p[2] : nullptr, (double*)p[3], (double*)p[4],
                                 (int32_t*)p[5], (uint8_t*)p[6], (double*)p[7], (double*)p[8], (double*)p[9], (uint32_t*)p[10], err);
    if (rc) return rc;
    if (x_out) CU(cudaMemcpyAsync(x_out, p[3], nx * 8, cudaMemcpyDeviceToHost, im->stream));
    if (u_out) CU(cudaMemcpyAsync(u_out, p[4], nu * 8, cudaMemcpyDeviceToHost, im->stream));
    if (iters) CU(cudaMemcpyAsync(iters, p[5], (size_t)n * 4, cudaMemcpyDeviceToHost, im->stream));
    if (status) CU(cudaMemcpyAsync(status, p[6], (size_t)n, cudaMemcpyDeviceToHost, im->stream));
    if (J) CU(cudaMemcpyAsync(J, p[7], (size_t)n * 8, cudaMemcpyDeviceToHost, im->stream));
    if (viol) CU(cudaMemcpyAsync(viol, p[8], (size_t)n * 8, cudaMemcpyDeviceToHost, im->stream));
    if (alpha) CU(cudaMemcpyAsync(alpha, p[9], (size_t)n * 8, cudaMemcpyDeviceToHost, im->stream));
    if (flags) CU(cudaMemcpyAsync(flags, p[10], (size_t)n * 4, cudaMemcpyDeviceToHost, im->stream));
    CU(cudaStreamSynchronize(im->stream));
    return 0;
}

static int plugin_get_trajectory(void* impl, double* x, double* u, int current, int device_out, char* err) {
    Impl* im = (Impl*)impl;
    const Dev& d = im->P.d;
    int rc = download(im, current ? d.xc : d.xb, x, (size_t)im->P.T * N, device_out != 0, err);
    if (!rc) rc = download(im, current ? d.uc : d.ub, u, (size_t)(im->P.T - 1) * M, device_out != 0, err);
    return rc;
}

static int plugin_get_stats(void* impl, int32_t* iterations, uint8_t* status, double* objective, double* max_violation,
                            double* step_size, uint32_t* flags, char* err) {
    Impl* im = (Impl*)impl;
    const Dev& d = im->P.d;
    /* six one-row arrays through six disjoint pieces of the staging buffer, one synchronisation at the end */
    const size_t Bp = im->P.Bp;
    int rc = download(im, d.iters, iterations, 1, false, err, 0 * Bp, false);
    if (!rc) rc = download(im, d.status, status, 1, false, err, 1 * Bp, false);
    if (!rc) rc = download(im, d.J, objective, 1, false, err, 2 * Bp, false);
    if (!rc) rc = download(im, d.viol, max_violation, 1, false, err, 3 * Bp, false);
    if (!rc) rc = download(im, d.alpha, step_size, 1, false, err, 4 * Bp, false);
    if (!rc) rc = download(im, d.flags, flags, 1, false, err, 5 * Bp, false);
    if (!rc) { CU(cudaSetDevice(im->device)); CU(cudaStreamSynchronize(im->stream)); }
    return rc;
}

static int plugin_get_history(void* impl, int32_t cap, double* cost, double* gnorm, double* viol, double* alpha,
                              int32_t* outer, uint8_t* status, char* err) {
    Impl* im = (Impl*)impl;
    const Dev& d = im->P.d;
    if (cap < 0 || cap > im->P.cap) return fail(err, ILQR_EINVAL, "cap %d exceeds history_cap %d", cap, im->P.cap);
    int rc = download(im, d.h_cost, cost, (size_t)cap, false, err);
    if (!rc) rc = download(im, d.h_gnorm, gnorm, (size_t)cap, false, err);
    if (!rc) rc = download(im, d.h_viol, viol, (size_t)cap, false, err);
    if (!rc) rc = download(im, d.h_alpha, alpha, (size_t)cap, false, err);
    if (!rc) rc = download(im, d.h_outer, outer, (size_t)cap, false, err);
    if (!rc) rc = download(im, d.h_status, status, (size_t)cap, false, err);
    return rc;
}

static int plugin_get_duals(void* impl, double* dual, double* penalty, double* violations, int32_t* active, char* err) {
    Impl* im = (Impl*)impl;
    const Dev& d = im->P.d;
    const size_t rows = (size_t)im->rows();
    int rc = download(im, d.lam, dual, rows, false, err);
    if (!rc) rc = download(im, d.rho, penalty, rows, false, err);
    if (!rc) rc = download(im, d.c, violations, rows, false, err);
    if (!rc) rc = download(im, d.act, active, rows, false, err);
    return rc;
}

static int plugin_get_policy(void* impl, double* K, double* k, char* err) {
    Impl* im = (Impl*)impl;
    const Dev& d = im->P.d;
#if ILQR_LARGE
    {   /* wide models keep the policy problem-major: already the host layout [problem][step][...] */
        CU(cudaSetDevice(im->device));
        const size_t steps = (size_t)(im->P.T - 1) * im->P.B;
        if (K) CU(cudaMemcpyAsync(K, d.K, sizeof(double) * steps * M * N, cudaMemcpyDeviceToHost, im->stream));
        if (k) CU(cudaMemcpyAsync(k, d.k, sizeof(double) * steps * M, cudaMemcpyDeviceToHost, im->stream));
        CU(cudaStreamSynchronize(im->stream));
        return 0;
    }
#endif
    int rc = download(im, d.K, K, (size_t)(im->P.T - 1) * M * N, false, err);
    if (!rc) rc = download(im, d.k, k, (size_t)(im->P.T - 1) * M, false, err);
    return rc;
}

static int plugin_mpc_step(void* impl, double* applied_u, double* x_next, char* err) {
    Impl* im = (Impl*)impl;
    Params& P = im->P;
    CU(cudaSetDevice(im->device));
    /* borrow two rows-blocks of the staging buffer's tail is not safe (download uses it), so
     * use the Lagrangian-gradient blocks, which the next solve overwrites before reading */
    double* d_au = P.d.Lu; /* (T-1)*M rows >= M */
    double* d_xn = P.d.Lx; /* (T-1)*N rows >= N */
    k_mpc_shift<<<(P.B + 127) / 128, 128, 0, im->stream>>>(P, d_au, d_xn);
    CU(cudaGetLastError());
    im->launches += 1;
    int rc = download(im, d_au, applied_u, (size_t)M, false, err);
    if (!rc) rc = download(im, d_xn, x_next, (size_t)N, false, err);
    if (rc) return rc;
    return plugin_solve(impl, err);
}

static int plugin_set_profiling(void* impl, int32_t on, char*) {
    Impl* im = (Impl*)impl;
    im->profiling = on != 0;
    if (on) {
        im->ticks = im->launches = im->problem_ticks = 0;
        for (int i = 0; i < 3; ++i) { im->kernel_ms[i] = 0; im->kernel_launches[i] = 0; }
    }
    return 0;
}
static int plugin_get_counters(void* impl, int64_t* ticks, int64_t* launches, double* kernel_ms, int64_t* kernel_launches, char*) {
    Impl* im = (Impl*)impl;
    if (ticks) *ticks = im->ticks;
    if (launches) *launches = im->launches;
    for (int i = 0; i < 3; ++i) {
        if (kernel_ms) kernel_ms[i] = im->kernel_ms[i];
        if (kernel_launches) kernel_launches[i] = im->kernel_launches[i];
    }
    return 0;
}

static int plugin_set_stream(void* impl, void* cuda_stream, char* err) {
    Impl* im = (Impl*)impl;
    CU(cudaSetDevice(im->device));
    CU(cudaStreamSynchronize(im->stream));
    drop_graphs(im);
    im->stream = cuda_stream ? (cudaStream_t)cuda_stream : im->own_stream;
    return 0;
}
static int plugin_get_problem_ticks(void* impl, int64_t* pt, char*) {
    if (pt) *pt = ((Impl*)impl)->problem_ticks;
    return 0;
}
/* ---- the one collective of the path: the final gather ---------------------------------- */
static int plugin_comm_init(void* impl, int32_t n_ranks, int32_t rank, const char* id, char* err) {
    Impl* im = (Impl*)impl;
    if (!id || n_ranks < 1 || rank < 0 || rank >= n_ranks) return fail(err, ILQR_EINVAL, "comm_init: bad rank %d of %d or NULL id", rank, n_ranks);
    const char* why = nullptr;
    const ilqr_nccl::Api* n = ilqr_nccl::api(&why);
    if (!n) return fail(err, ILQR_ECUDA, "NCCL unavailable: %s", why);
    CU(cudaSetDevice(im->device));
    if (im->comm) { n->CommDestroy(im->comm); im->comm = nullptr; }
    ilqr_nccl::unique_id u;
    memcpy(u.internal, id, sizeof(u.internal));
    const int rc = n->CommInitRank(&im->comm, n_ranks, u, rank);
    if (rc != 0) { im->comm = nullptr; return fail(err, ILQR_ECUDA, "ncclCommInitRank failed: %s", n->GetErrorString(rc)); }
    im->comm_ranks = n_ranks;
    return 0;
}
static int plugin_gather(void* impl, const void* d_local, void* d_all, size_t bytes_per_rank, char* err) {
    Impl* im = (Impl*)impl;
    if (!im->comm) return fail(err, ILQR_ESTATE, "ilqr_gather before ilqr_comm_init");
    if (!d_local || !d_all) return fail(err, ILQR_EINVAL, "NULL buffer");
    const ilqr_nccl::Api* n = ilqr_nccl::api(nullptr);
    CU(cudaSetDevice(im->device));
    const int rc = n->AllGather(d_local, d_all, bytes_per_rank, ilqr_nccl::DT_CHAR, im->comm, (void*)im->stream);
    if (rc != 0) return fail(err, ILQR_ECUDA, "ncclAllGather failed: %s", n->GetErrorString(rc));
    return 0;
}
static int plugin_get_compactions(void* impl, int64_t* n, char*) {
    if (n) *n = ((Impl*)impl)->compactions;
    return 0;
}

} /* namespace ilqr */

extern "C" __attribute__((visibility("default"))) const ilqr_plugin_table ilqr_plugin_table_v7 = {
    ILQR_PLUGIN_VERSION,
    ILQR_N, ILQR_M, ILQR_P, ILQR_CS, ILQR_CT,
    ILQR_MODEL_NAME,
    ILQR_MODEL_HASH,
    ilqr::plugin_create,
    ilqr::plugin_destroy,
    ilqr::plugin_set_options,
    ilqr::plugin_initialize_controls,
    ilqr::plugin_initialize_states,
    ilqr::plugin_set_parameters,
    ilqr::plugin_rollout,
    ilqr::plugin_solve,
    ilqr::plugin_get_trajectory,
    ilqr::plugin_get_stats,
    ilqr::plugin_get_history,
    ilqr::plugin_get_duals,
    ilqr::plugin_get_policy,
    ilqr::plugin_mpc_step,
    ilqr::plugin_set_profiling,
    ilqr::plugin_get_counters,
    ilqr::plugin_get_problem_ticks,
    ilqr::plugin_set_stream,
    ilqr::plugin_solve_stream,
    ilqr::plugin_solve_stream_host,
    ilqr::plugin_mpc_run,
    ilqr::plugin_get_compactions,
    ilqr::plugin_comm_init,
    ilqr::plugin_gather,
    ilqr::plugin_solve_outer,
};

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
constexpr int LB_PRODUCERS = 0;
constexpr int LB_WARPS = 1;
constexpr int LB_SMEM_BYTES = 0;
__global__ void k_linback(const __grid_constant__ Params P) { (void)P; } /* fused path not used for large models */

/* -------------------------------------------------------------------------------------------- k_linearize */
__device__ __noinline__ void lin_dynamics(const Params& P, int b, int t, const double* x, const double* u, const double* wv) {
    const Dev& d = P.d;
    const size_t Bp = P.Bp;
    double fx[N * N], fu[d1(N * M)];
    ilqr_dyn_jac(fx, fu, x, u, wv);                                         /* src/dynamics.jl:41-50 */
    for (int r = 0; r < N * N; ++r) d.fx[((size_t)t * N * N + r) * Bp + b] = fx[r];
    for (int r = 0; r < N * M; ++r) d.fu[((size_t)t * N * M + r) * Bp + b] = fu[r];
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
    {
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

__global__ void __launch_bounds__(64) k_linearize(const __grid_constant__ Params P) {
    const Dev& d = P.d;
    const size_t Bp = P.Bp;
    const int T = P.T;
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int b = (int)(g % Bp);
    const int t = (int)(g / Bp);
    if (t >= T) return;
    const int kind = d.kind[b];
    if (kind == KIND_NONE) return;
    if (kind == KIND_ITER && P.o.line_search == ILQR_LINE_SEARCH_NONE) return; /* src/solve.jl:27 */
    const bool fresh = kind == KIND_PRELOOP;
    double x[N], u[d1(M)], wv[d1(NP)];
    for (int i = 0; i < N; ++i) x[i] = d.xb[((size_t)t * N + i) * Bp + b];
    for (int i = 0; i < NP; ++i) wv[i] = d.w[((size_t)t * NP + i) * Bp + b];
    if (t < T - 1) {
        for (int a = 0; a < M; ++a) u[a] = d.ub[((size_t)t * M + a) * Bp + b];
        lin_dynamics(P, b, t, x, u, wv);
        lin_cost_stage(P, b, t, fresh, x, u, wv);
    } else {
        for (int a = 0; a < M; ++a) u[a] = 0.0;
        lin_cost_terminal(P, b, fresh, x, u, wv);
    }
}

/* -------------------------------------------------------------------------------------------- k_backward */
constexpr int RL_THREADS = 256;
static_assert(N + M <= RL_THREADS && N <= 16 * 16 && M <= 16 * 16, "k_backward (large) thread mapping");
constexpr int TI = (N + 15) / 16;  /* outputs per thread along the first index for an N x N result on a 16 x 16 thread grid */
constexpr int TA = (M + 15) / 16;  /* ... for the M-row results */

struct RlSmem { /* carve-up of the dynamic shared memory, all doubles */
    double *P, *p, *fxT, *fuT, *xxhT, *uxhT, *Qxx, *Qux, *Quu, *uu, *K, *uxt, *Qx, *Qu, *kk, *rinv, *gxs, *gus;
};
constexpr size_t RL_SMEM_DOUBLES = (size_t)N * N /*P*/ + N /*p*/ + (size_t)N * N /*fxT*/ + (size_t)N * M /*fuT*/ + (size_t)N * N /*xxhT*/ +
                                   (size_t)M * N /*uxhT*/ + (size_t)N * N /*Qxx*/ + (size_t)M * N /*Qux*/ + (size_t)M * M /*Quu*/ +
                                   (size_t)M * M /*uu*/ + (size_t)M * N /*K*/ + (size_t)M * N /*uxt*/ + N + M + M + M + N + M;
constexpr size_t RL_SMEM_BYTES = RL_SMEM_DOUBLES * 8 + 64;

/* Layouts in shared memory (chosen so that the register-tiled loops read contiguous runs):
 *   P[k + l*N]        column-major like the reference
 *   fxT[k*N + i]  = fx[k, i]      ("k-major": the i's of one k are contiguous)
 *   fuT[k*M + a]  = fu[k, a]
 *   xxhT[l*N + i] = (fx' P)[i, l]
 *   uxhT[l*M + a] = (fu' P)[a, l]
 *   Qxx, Qux, Quu, K, uxt column-major (Qux[a + j*M], K[a + j*M], uxt[a + j*M]) */
__global__ void __launch_bounds__(RL_THREADS) k_backward(const __grid_constant__ Params P) {
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
    RlSmem s;
    {
        double* q = rl_smem;
        s.P = q; q += N * N; s.p = q; q += N; s.fxT = q; q += N * N; s.fuT = q; q += N * M; s.xxhT = q; q += N * N;
        s.uxhT = q; q += M * N; s.Qxx = q; q += N * N; s.Qux = q; q += M * N; s.Quu = q; q += M * M; s.uu = q; q += M * M;
        s.K = q; q += M * N; s.uxt = q; q += M * N; s.Qx = q; q += N; s.Qu = q; q += M; s.kk = q; q += M; s.rinv = q; q += M;
        s.gxs = q; q += N; s.gus = q; q += M;
    }
    double gn = 0.0;
    if (kind != KIND_NONE && !skip_ls_none) {
        if (tid == 0) s_cholfail = 0;
        /* terminal value function: src/backward_pass.jl:39-40 */
        for (int r = tid; r < N * N; r += RL_THREADS) s.P[r] = d.gxx[((size_t)(T - 1) * N * N + r) * Bp + b];
        for (int r = tid; r < N; r += RL_THREADS) s.p[r] = d.gx[((size_t)(T - 1) * N + r) * Bp + b];
        __syncthreads();
        const int ti = tid & 15, tl = tid >> 4; /* 16 x 16 thread grid */
        for (int t = T - 2; t >= 0; --t) {
            /* ---- A: this step's Jacobians into shared memory (k-major) */
            for (int r = tid; r < N * N; r += RL_THREADS) { /* r = k + i*N in global */
                const int k = r % N, i = r / N;
                s.fxT[k * N + i] = d.fx[((size_t)t * N * N + r) * Bp + b];
            }
            for (int r = tid; r < N * M; r += RL_THREADS) {
                const int k = r % N, a = r / N;
                s.fuT[k * M + a] = d.fu[((size_t)t * N * M + r) * Bp + b];
            }
            for (int r = tid; r < N; r += RL_THREADS) s.gxs[r] = d.gx[((size_t)t * N + r) * Bp + b];
            for (int r = tid; r < M; r += RL_THREADS) s.gus[r] = d.gu[((size_t)t * M + r) * Bp + b];
            __syncthreads();
            /* ---- B: xxh = fx' P (:52), uxh = fu' P (:57), Qx (:44-45), Qu (:48-49) */
            {
                double acc[TI][TI];
#pragma unroll
                for (int ii = 0; ii < TI; ++ii)
#pragma unroll
                    for (int ll = 0; ll < TI; ++ll) acc[ii][ll] = 0.0;
                for (int k = 0; k < N; ++k) {
                    double av[TI], bv[TI];
#pragma unroll
                    for (int ii = 0; ii < TI; ++ii) av[ii] = (ti + 16 * ii < N) ? s.fxT[k * N + ti + 16 * ii] : 0.0;
#pragma unroll
                    for (int ll = 0; ll < TI; ++ll) bv[ll] = (tl + 16 * ll < N) ? s.P[k + (tl + 16 * ll) * N] : 0.0;
#pragma unroll
                    for (int ii = 0; ii < TI; ++ii)
#pragma unroll
                        for (int ll = 0; ll < TI; ++ll)
                            acc[ii][ll] = (k == 0) ? av[ii] * bv[ll] : ilqr_fma(av[ii], bv[ll], acc[ii][ll]);
                }
#pragma unroll
                for (int ii = 0; ii < TI; ++ii)
#pragma unroll
                    for (int ll = 0; ll < TI; ++ll)
                        if (ti + 16 * ii < N && tl + 16 * ll < N) s.xxhT[(tl + 16 * ll) * N + ti + 16 * ii] = acc[ii][ll];
            }
            {
                double acc[TA][TI];
#pragma unroll
                for (int aa = 0; aa < TA; ++aa)
#pragma unroll
                    for (int ll = 0; ll < TI; ++ll) acc[aa][ll] = 0.0;
                for (int k = 0; k < N; ++k) {
                    double av[TA], bv[TI];
#pragma unroll
                    for (int aa = 0; aa < TA; ++aa) av[aa] = (ti + 16 * aa < M) ? s.fuT[k * M + ti + 16 * aa] : 0.0;
#pragma unroll
                    for (int ll = 0; ll < TI; ++ll) bv[ll] = (tl + 16 * ll < N) ? s.P[k + (tl + 16 * ll) * N] : 0.0;
#pragma unroll
                    for (int aa = 0; aa < TA; ++aa)
#pragma unroll
                        for (int ll = 0; ll < TI; ++ll)
                            acc[aa][ll] = (k == 0) ? av[aa] * bv[ll] : ilqr_fma(av[aa], bv[ll], acc[aa][ll]);
                }
#pragma unroll
                for (int aa = 0; aa < TA; ++aa)
#pragma unroll
                    for (int ll = 0; ll < TI; ++ll)
                        if (ti + 16 * aa < M && tl + 16 * ll < N) s.uxhT[(tl + 16 * ll) * M + ti + 16 * aa] = acc[aa][ll];
            }
            if (tid < N) {
                double acc = s.fxT[tid] * s.p[0];
                for (int k = 1; k < N; ++k) acc = ilqr_fma(s.fxT[k * N + tid], s.p[k], acc);
                s.Qx[tid] = acc + s.gxs[tid];
            } else if (tid >= RL_THREADS - M) {
                const int a = tid - (RL_THREADS - M);
                double acc = s.fuT[a] * s.p[0];
                for (int k = 1; k < N; ++k) acc = ilqr_fma(s.fuT[k * M + a], s.p[k], acc);
                s.Qu[a] = acc + s.gus[a];
            }
            __syncthreads();
            /* ---- C: Qxx = xxh fx + gxx (:53-54), Quu = uxh fu + guu (:58-59), Qux = uxh fx + gux (:63-64) */
            {
                double acc[TI][TI];
#pragma unroll
                for (int ii = 0; ii < TI; ++ii)
#pragma unroll
                    for (int jj = 0; jj < TI; ++jj) acc[ii][jj] = 0.0;
                for (int l = 0; l < N; ++l) {
                    double av[TI], bv[TI];
#pragma unroll
                    for (int ii = 0; ii < TI; ++ii) av[ii] = (ti + 16 * ii < N) ? s.xxhT[l * N + ti + 16 * ii] : 0.0;
#pragma unroll
                    for (int jj = 0; jj < TI; ++jj) bv[jj] = (tl + 16 * jj < N) ? s.fxT[l * N + tl + 16 * jj] : 0.0;
#pragma unroll
                    for (int ii = 0; ii < TI; ++ii)
#pragma unroll
                        for (int jj = 0; jj < TI; ++jj)
                            acc[ii][jj] = (l == 0) ? av[ii] * bv[jj] : ilqr_fma(av[ii], bv[jj], acc[ii][jj]);
                }
#pragma unroll
                for (int ii = 0; ii < TI; ++ii)
#pragma unroll
                    for (int jj = 0; jj < TI; ++jj) {
                        const int i = ti + 16 * ii, j = tl + 16 * jj;
                        if (i < N && j < N) s.Qxx[i + j * N] = acc[ii][jj] + d.gxx[((size_t)t * N * N + i + (size_t)j * N) * Bp + b];
                    }
            }
            {
                double acc[TA][TI];
#pragma unroll
                for (int aa = 0; aa < TA; ++aa)
#pragma unroll
                    for (int jj = 0; jj < TI; ++jj) acc[aa][jj] = 0.0;
                for (int l = 0; l < N; ++l) {
                    double av[TA], bv[TI];
#pragma unroll
                    for (int aa = 0; aa < TA; ++aa) av[aa] = (ti + 16 * aa < M) ? s.uxhT[l * M + ti + 16 * aa] : 0.0;
#pragma unroll
                    for (int jj = 0; jj < TI; ++jj) bv[jj] = (tl + 16 * jj < N) ? s.fxT[l * N + tl + 16 * jj] : 0.0;
#pragma unroll
                    for (int aa = 0; aa < TA; ++aa)
#pragma unroll
                        for (int jj = 0; jj < TI; ++jj)
                            acc[aa][jj] = (l == 0) ? av[aa] * bv[jj] : ilqr_fma(av[aa], bv[jj], acc[aa][jj]);
                }
#pragma unroll
                for (int aa = 0; aa < TA; ++aa)
#pragma unroll
                    for (int jj = 0; jj < TI; ++jj) {
                        const int a = ti + 16 * aa, j = tl + 16 * jj;
                        if (a < M && j < N) s.Qux[a + j * M] = acc[aa][jj] + d.gux[((size_t)t * M * N + a + (size_t)j * M) * Bp + b];
                    }
            }
            for (int o = tid; o < M * M; o += RL_THREADS) {
                const int a = o % M, e = o / M;
                double acc = s.uxhT[a] * s.fuT[e];
                for (int l = 1; l < N; ++l) acc = ilqr_fma(s.uxhT[l * M + a], s.fuT[l * M + e], acc);
                const double q = acc + d.guu[((size_t)t * M * M + o) * Bp + b];
                s.Quu[o] = q;
                s.uu[o] = q;                                                                  /* :68 */
            }
            __syncthreads();
            /* ---- D: Cholesky of Quu on warp 0, unblocked upper, stop at the first bad pivot (:69, Q3) */
            if (tid < 32) {
                bool ok = true;
                for (int j = 0; j < M; ++j) {
                    if (ok) {
                        double ajj = s.uu[j + j * M];
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
            __syncthreads();
            /* ---- E: K = -Quu \ Qux, k = -Quu \ Qu (:70-75): one thread per right-hand side */
            for (int col = tid; col < N + 1; col += RL_THREADS) {
                double bv[d1(M)];
                for (int a = 0; a < M; ++a) bv[a] = col < N ? s.Qux[a + col * M] : s.Qu[a];
                for (int i = 0; i < M; ++i) {
                    double sum = bv[i];
                    for (int k = 0; k < i; ++k) sum = ilqr_fma(-s.uu[k + i * M], bv[k], sum);
                    bv[i] = sum * s.rinv[i];
                }
                for (int i = M - 1; i >= 0; --i) {
                    double sum = bv[i];
                    for (int k = i + 1; k < M; ++k) sum = ilqr_fma(-s.uu[i + k * M], bv[k], sum);
                    bv[i] = sum * s.rinv[i];
                }
                if (col < N) {
                    for (int a = 0; a < M; ++a) {
                        s.K[a + col * M] = -bv[a];
                        d.K[((size_t)t * M * N + a + (size_t)col * M) * Bp + b] = -bv[a];
                    }
                } else {
                    for (int a = 0; a < M; ++a) {
                        s.kk[a] = -bv[a];
                        d.k[((size_t)t * M + a) * Bp + b] = -bv[a];
                    }
                }
            }
            __syncthreads();
            /* ---- F: uxt = Quu K (:79) */
            for (int o = tid; o < M * N; o += RL_THREADS) {
                const int a = o % M, j = o / M;
                double acc = s.Quu[a] * s.K[j * M];
                for (int e = 1; e < M; ++e) acc = ilqr_fma(s.Quu[a + e * M], s.K[e + j * M], acc);
                s.uxt[o] = acc;
            }
            __syncthreads();
            /* ---- G: P = K'uxt + K'Qux + Qux'K + Qxx (:81-84), p (:86-89), Lagrangian gradient (src/solve.jl:75-78) */
            {
                double newP[TI][TI];
#pragma unroll
                for (int ii = 0; ii < TI; ++ii)
#pragma unroll
                    for (int jj = 0; jj < TI; ++jj) {
                        const int i = ti + 16 * ii, j = tl + 16 * jj;
                        double v = 0.0;
                        if (i < N && j < N) {
                            double a1 = s.K[i * M] * s.uxt[j * M];
                            for (int a = 1; a < M; ++a) a1 = ilqr_fma(s.K[a + i * M], s.uxt[a + j * M], a1);
                            double a2 = s.K[i * M] * s.Qux[j * M];
                            for (int a = 1; a < M; ++a) a2 = ilqr_fma(s.K[a + i * M], s.Qux[a + j * M], a2);
                            double a3 = s.Qux[i * M] * s.K[j * M];
                            for (int a = 1; a < M; ++a) a3 = ilqr_fma(s.Qux[a + i * M], s.K[a + j * M], a3);
                            v = a1;
                            v = v + a2;
                            v = v + a3;
                            v = v + s.Qxx[i + j * N];
                        }
                        newP[ii][jj] = v;
                    }
                double newp = 0.0, lx = 0.0;
                if (tid < N) {
                    const int i = tid;
                    double a1 = s.uxt[i * M] * s.kk[0];
                    for (int a = 1; a < M; ++a) a1 = ilqr_fma(s.uxt[a + i * M], s.kk[a], a1);
                    double a2 = s.K[i * M] * s.Qu[0];
                    for (int a = 1; a < M; ++a) a2 = ilqr_fma(s.K[a + i * M], s.Qu[a], a2);
                    double a3 = s.Qux[i * M] * s.kk[0];
                    for (int a = 1; a < M; ++a) a3 = ilqr_fma(s.Qux[a + i * M], s.kk[a], a3);
                    double v = a1;
                    v = v + a2;
                    v = v + a3;
                    newp = v + s.Qx[i];
                    lx = s.Qx[i] - newp;
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
                __syncthreads(); /* everyone has read the old K, uxt, Qux, Qxx, Qx, p */
#pragma unroll
                for (int ii = 0; ii < TI; ++ii)
#pragma unroll
                    for (int jj = 0; jj < TI; ++jj) {
                        const int i = ti + 16 * ii, j = tl + 16 * jj;
                        if (i < N && j < N) s.P[i + j * N] = newP[ii][jj];
                    }
                if (tid < N) s.p[tid] = newp;
            }
            __syncthreads();
        }
        /* gradient norm: max over the CTA, NaN-propagating like norm(., Inf) */
        for (int off = 16; off > 0; off >>= 1) {
            const double o = __shfl_xor_sync(0xffffffffu, gn, off);
            if (o > gn || o != o) gn = o;
        }
        if ((tid & 31) == 0) s_gn[tid >> 5] = gn;
        __syncthreads();
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

// K7 — batched mode: many independent small LPs, ONE CTA PER LP, the whole revised-simplex
// loop of the reference (src/Simplex.cpp:255-353, Bland ratio test :66-84, both pricing rules
// :49-61 / :87-102) with all state in shared memory. HBM sees one read of (A, b, c, lists) and
// one write of (lists, x_B, objective, status) per LP; the per-pivot traffic is on chip.
//
// Shared-memory layout (doubles): A[N][lda] column-major, lda = m|1 (odd stride: conflict-free
// for "one thread per column"), Binv[m][ldb] row-major, ldb = m|1, then y, d, xB, prow, cB (m each).
#pragma once
#include <cfloat>
#include "sb200_state.cuh"

namespace sb200
{
    constexpr int BATCHED_THREADS = 256;

    struct BatchedParams
    {
        int m, N, batch;
        double eps;
        int rule;
        int iter_limit;
        const double * A;  // [batch][N][m]
        const double * b;  // [batch][m]
        const double * c;  // [batch][N]
        int * basic;       // [batch][m]
        int * nonbasic;    // [batch][N-m]
        double * xB;       // [batch][m]
        double * objective;
        int * term;
        int * iterations;
        int only_pending;  // 1: take only the LPs the fast kernel marked TERM_PENDING (term[lp] == -2)
        // finite upper bounds (nullptr: none, plain ratio test): ub[batch][N] (DBL_MAX = no bound, reference
        // src/LinearProgram.cpp:225), out: flip[batch][N] = ub_substitutions at the end, flips[batch] = iterations that
        // moved the entering variable to its bound instead of pivoting (src/Simplex.cpp:160-164)
        const double * ub;
        unsigned char * flip;
        int * flips;
    };

    __host__ __device__ inline size_t batched_smem_bytes(int m, int N, bool bounded = false)
    {
        const size_t lda = static_cast<size_t>(m) | 1, ldb = lda;
        const size_t doubles = static_cast<size_t>(N) * lda + static_cast<size_t>(m) * ldb + 6 * static_cast<size_t>(m) +
                               static_cast<size_t>(N) /* c */ + 8 + (bounded ? static_cast<size_t>(N) + 2 : 0) /* ub */;
        const size_t ints = static_cast<size_t>(N) + 8;  // basic + nonbasic
        return doubles * sizeof(double) + ints * sizeof(int) + (bounded ? static_cast<size_t>(N) + 8 : 0) /* flags */;
    }

    __global__ void __launch_bounds__(BATCHED_THREADS) k_batched(BatchedParams P)
    {
        extern __shared__ __align__(16) double smem[];
        __shared__ Cand scratch[33];
        __shared__ int sh_flag;

        const int m = P.m, N = P.N, nnb = N - m;
        const int lda = m | 1, ldb = m | 1;
        double * As = smem;
        double * Bi = As + static_cast<size_t>(N) * lda;
        double * y = Bi + static_cast<size_t>(m) * ldb;
        double * d = y + m;
        double * xB = d + m;
        double * prow = xB + m;
        double * bs = prow + m;
        double * tmp = bs + m;
        double * cs = tmp + m;
        const bool bounded = P.ub != nullptr;
        double * us = cs + N + (N & 1);                      // upper bounds (bounded only)
        int * basic = reinterpret_cast<int *>(us + (bounded ? N + (N & 1) : 0));
        int * nonbasic = basic + m;
        unsigned char * flipf = reinterpret_cast<unsigned char *>(nonbasic + nnb + ((nnb + m) & 1));  // ub_substitutions

        const int lp = blockIdx.x;
        if (P.only_pending && P.term[lp] != -2) return;
        const int tid = threadIdx.x, nt = blockDim.x;
        const double * gA = P.A + static_cast<size_t>(lp) * N * m;

        // ---- load (coalesced over the contiguous [N][m] block)
        for (int k = tid; k < N * m; k += nt)
        {
            const int col = k / m, row = k - col * m;
            As[static_cast<size_t>(col) * lda + row] = gA[k];
        }
        for (int k = tid; k < m; k += nt)
        {
            bs[k] = P.b[static_cast<size_t>(lp) * m + k];
            basic[k] = P.basic[static_cast<size_t>(lp) * m + k];
        }
        for (int k = tid; k < N; k += nt) cs[k] = P.c[static_cast<size_t>(lp) * N + k];
        if (bounded)
            for (int k = tid; k < N; k += nt)
            {
                us[k] = P.ub[static_cast<size_t>(lp) * N + k];
                flipf[k] = 0;  // Simplex.cpp:271
            }
        for (int k = tid; k < nnb; k += nt) nonbasic[k] = P.nonbasic[static_cast<size_t>(lp) * nnb + k];
        if (tid == 0) sh_flag = 0;
        __syncthreads();

        // ---- K0: B^-1 by Gauss-Jordan with partial pivoting on [A_B | I] kept as (W=tmp matrix in Bi)
        // W is built in place of Binv's storage is not possible (we need both), so W lives in the
        // A_B columns gathered into a scratch square carved from the end of A's padding? No: use a
        // diagonal fast path (crash basis) and a general path that inverts via unit-vector solves.
        {
            // diagonal check: column basic[i] must be a multiple of e_i
            for (int i = tid; i < m; i += nt)
            {
                const double * col = As + static_cast<size_t>(basic[i]) * lda;
                int bad = (col[i] == 0.0);
                for (int r = 0; r < m; ++r)
                    if (r != i && col[r] != 0.0) bad = 1;
                if (bad) atomicOr(&sh_flag, 1);
            }
            __syncthreads();
            const int general = sh_flag;
            for (int k = tid; k < m * ldb; k += nt) Bi[k] = 0.0;
            __syncthreads();
            if (!general)
            {
                for (int i = tid; i < m; i += nt) Bi[static_cast<size_t>(i) * ldb + i] = 1.0 / As[static_cast<size_t>(basic[i]) * lda + i];
            }
            else
            {
                // General basis: Gauss-Jordan on V (= Bi, starts as I) while W = A_B is applied
                // implicitly column by column from a copy kept in global-free scratch: we reuse the
                // d/prow/tmp vectors and re-derive W's column k on the fly:  W_k(step) = V(step) * a_{basic[k]}.
                for (int i = tid; i < m; i += nt) Bi[static_cast<size_t>(i) * ldb + i] = 1.0;
                __syncthreads();
                for (int k = 0; k < m; ++k)
                {
                    // w = V * a_k  (current transformed column k)
                    const double * ak = As + static_cast<size_t>(basic[k]) * lda;
                    for (int i = tid; i < m; i += nt)
                    {
                        double acc = 0.0;
                        const double * row = Bi + static_cast<size_t>(i) * ldb;
                        for (int j = 0; j < m; ++j) acc = fma(row[j], ak[j], acc);
                        tmp[i] = acc;
                    }
                    __syncthreads();
                    // partial pivot among rows >= k (first max)
                    Cand c = cand_none();
                    for (int i = k + tid; i < m; i += nt)
                    {
                        Cand o;
                        o.key = -fabs(tmp[i]);
                        o.aux = 0.0;
                        o.idx = i;
                        o.cls = 0;
                        if (cand_better(o, c)) c = o;
                    }
                    c = block_reduce_cand(c, scratch);
                    if (tid == 0) scratch[32] = c;
                    __syncthreads();
                    c = scratch[32];
                    const int r = c.idx;
                    if (r < 0 || c.key == 0.0)
                    {
                        if (tid == 0) sh_flag = 2;  // singular
                        break;
                    }
                    if (r != k)
                    {
                        for (int j = tid; j < m; j += nt)
                        {
                            const double t = Bi[static_cast<size_t>(k) * ldb + j];
                            Bi[static_cast<size_t>(k) * ldb + j] = Bi[static_cast<size_t>(r) * ldb + j];
                            Bi[static_cast<size_t>(r) * ldb + j] = t;
                        }
                        if (tid == 0)
                        {
                            const double t = tmp[k];
                            tmp[k] = tmp[r];
                            tmp[r] = t;
                        }
                    }
                    __syncthreads();
                    const double piv = tmp[k];
                    for (int j = tid; j < m; j += nt) prow[j] = Bi[static_cast<size_t>(k) * ldb + j] / piv;
                    __syncthreads();
                    for (int idx = tid; idx < m * m; idx += nt)
                    {
                        const int i = idx / m, j = idx - i * m;
                        if (i == k)
                            Bi[static_cast<size_t>(i) * ldb + j] = prow[j];
                        else
                            Bi[static_cast<size_t>(i) * ldb + j] = fma(-tmp[i], prow[j], Bi[static_cast<size_t>(i) * ldb + j]);
                    }
                    __syncthreads();
                }
                __syncthreads();
            }
        }
        __syncthreads();
        int status = TERM_RUNNING;
        if (sh_flag == 2) status = -1;  // singular start basis: reported as term = -1

        // x_B = B^-1 b
        for (int i = tid; i < m; i += nt)
        {
            double acc = 0.0;
            const double * row = Bi + static_cast<size_t>(i) * ldb;
            for (int j = 0; j < m; ++j) acc = fma(row[j], bs[j], acc);
            xB[i] = acc;
        }
        __syncthreads();

        int iterations = 0, nflips = 0;
        const double neg_eps = -P.eps;
        // LinearProgram::upper_bound_substitution (src/LinearProgram.cpp:154-169) on the shared-memory tableau:
        // c_j = -c_j, b -= a_j u (two roundings, as g++ compiles it), a_j = -a_j, flag toggled. All threads call it.
        auto substitute = [&](int col, double u) {
            double * a = As + static_cast<size_t>(col) * lda;
            for (int i = tid; i < m; i += nt)
            {
                const double v = a[i];
                bs[i] = __dsub_rn(bs[i], __dmul_rn(v, u));
                a[i] = -v;
            }
            if (tid == 0)
            {
                cs[col] = -cs[col];
                flipf[col] ^= 1;
            }
        };
        while (status == TERM_RUNNING)
        {
            iterations += 1;
            // ---- K1: y_j = sum_i cB_i Binv[i][j]
            for (int j = tid; j < m; j += nt)
            {
                double acc = 0.0;
                for (int i = 0; i < m; ++i) acc = fma(cs[basic[i]], Bi[static_cast<size_t>(i) * ldb + j], acc);
                y[j] = acc;
            }
            __syncthreads();
            // ---- K2: pricing, one thread per non-basic position
            Cand best = cand_none();
            for (int pos = tid; pos < nnb; pos += nt)
            {
                const int col = nonbasic[pos];
                const double * a = As + static_cast<size_t>(col) * lda;
                double acc = 0.0;
                for (int i = 0; i < m; ++i) acc = fma(a[i], y[i], acc);
                const double s = cs[col] - acc;
                if (s < neg_eps)
                {
                    Cand c;
                    c.key = (P.rule == RULE_MOST_NEG) ? s : 0.0;
                    c.aux = s;
                    c.idx = pos;
                    c.cls = 0;
                    if (cand_better(c, best)) best = c;
                }
            }
            best = block_reduce_cand(best, scratch);
            if (tid == 0) scratch[32] = best;
            __syncthreads();
            best = scratch[32];
            if (best.idx < 0)
            {
                status = (iterations >= P.iter_limit) ? TERM_ITER_LIMIT : TERM_OPTIMAL;
                break;
            }
            const int e = best.idx;
            const int q = nonbasic[e];
            // ---- K3: d = Binv a_q, ratio candidates (Bland, Simplex.cpp:66-84)
            const double * aq = As + static_cast<size_t>(q) * lda;
            Cand rb = cand_none();
            for (int i = tid; i < m; i += nt)
            {
                double acc = 0.0;
                const double * row = Bi + static_cast<size_t>(i) * ldb;
                for (int j = 0; j < m; ++j) acc = fma(row[j], aq[j], acc);
                d[i] = acc;
                if (acc > P.eps)
                {
                    const double t = xB[i] / acc;
                    if (t < DBL_MAX)
                    {
                        Cand c;
                        c.key = t;
                        c.aux = acc;
                        c.idx = i;
                        c.cls = 0;
                        if (cand_better(c, rb)) rb = c;
                    }
                }
                else if (bounded && acc < -P.eps)  // Simplex.cpp:136-153: the t2 pass
                {
                    const double t = (us[basic[i]] - xB[i]) / (-acc);
                    if (t < DBL_MAX)
                    {
                        Cand c;
                        c.key = t;
                        c.aux = acc;
                        c.idx = i;
                        c.cls = 1;
                        if (cand_better(c, rb)) rb = c;
                    }
                }
            }
            rb = block_reduce_cand(rb, scratch);
            if (tid == 0) scratch[32] = rb;
            __syncthreads();
            rb = scratch[32];
            if (bounded)
            {
                const double min_theta = (rb.idx >= 0) ? rb.key : DBL_MAX;
                const double ub_s = us[q];
                if (ub_s < min_theta)
                {
                    // Simplex.cpp:160-164: the entering variable reaches its own bound first: no basis change,
                    // x_B = B^-1 (b - a_q u) = x_B - u d
                    __syncthreads();  // everybody has read us[q], cs, a_q of this iteration
                    substitute(q, ub_s);
                    for (int i = tid; i < m; i += nt) xB[i] = fma(-ub_s, d[i], xB[i]);
                    nflips += 1;
                    if (iterations >= P.iter_limit) status = TERM_ITER_LIMIT;
                    __syncthreads();
                    continue;
                }
            }
            if (rb.idx < 0)
            {
                status = TERM_UNBOUNDED;
                break;
            }
            // ---- K4/K5: pivot
            const int p = rb.idx;
            const double dp = rb.aux;  // d_p as computed from the stored row p
            const double theta = rb.key;
            if (bounded && rb.cls == 1)
            {
                // Simplex.cpp:165-176: the leaving variable sits at its upper bound: substitute it first. Row p of B^-1
                // changes sign, which the scaled pivot row (-rho)/(-d) = rho/d never sees; x_p -> u - x_p = theta d_p'.
                __syncthreads();
                substitute(basic[p], us[basic[p]]);
            }
            for (int j = tid; j < m; j += nt) prow[j] = Bi[static_cast<size_t>(p) * ldb + j] / dp;
            __syncthreads();
            for (int idx = tid; idx < m * m; idx += nt)
            {
                const int i = idx / m, j = idx - i * m;
                if (i == p)
                    Bi[static_cast<size_t>(i) * ldb + j] = prow[j];
                else
                    Bi[static_cast<size_t>(i) * ldb + j] = fma(-d[i], prow[j], Bi[static_cast<size_t>(i) * ldb + j]);
            }
            for (int i = tid; i < m; i += nt) xB[i] = (i == p) ? theta : fma(-theta, d[i], xB[i]);
            if (tid == 0)
            {
                const int leaving = basic[p];
                basic[p] = q;
                nonbasic[e] = leaving;
            }
            if (iterations >= P.iter_limit) status = TERM_ITER_LIMIT;
            __syncthreads();
        }

        // ---- write back
        __syncthreads();
        for (int k = tid; k < m; k += nt)
        {
            P.basic[static_cast<size_t>(lp) * m + k] = basic[k];
            P.xB[static_cast<size_t>(lp) * m + k] = xB[k];
        }
        for (int k = tid; k < nnb; k += nt) P.nonbasic[static_cast<size_t>(lp) * nnb + k] = nonbasic[k];
        if (tid == 0)
        {
            double obj = 0.0;
            for (int i = 0; i < m; ++i) obj += cs[basic[i]] * xB[i];
            P.objective[lp] = obj;
            P.term[lp] = status;
            P.iterations[lp] = iterations;
            if (bounded && P.flips != nullptr) P.flips[lp] = nflips;
        }
        if (bounded && P.flip != nullptr)
            for (int k = tid; k < N; k += nt) P.flip[static_cast<size_t>(lp) * N + k] = flipf[k];
    }
}  // namespace sb200

// K7-fast — batched mode for small LPs (m <= 64): ONE CTA PER LP, the basis inverse lives in
// REGISTERS, the dense columns of A in shared memory, unit columns (slacks, artificials) are
// kept as (row, value) pairs. Same algorithm and tie-breaking as k_batched / the reference loop
// (src/Simplex.cpp:255-353 with the Bland ratio test :66-84 and both entering rules :49-61,
// :87-102); BASELINE configs[4] (65536 LPs of m=64, n=128) is the shape it is sized for.
//
// Why this layout: per pivot an LP of this size needs ~16k fp64 FMAs and, with everything in
// shared memory, ~150 KB of shared-memory traffic: the generic kernel is bound by shared-memory
// bandwidth and, with 135 KB per CTA, by having one CTA (8 warps) per SM. Here
//   * thread (r, k) = (tid >> 2, tid & 3) owns B^-1[r][k + 4*jj], jj = 0..15, in 32 registers:
//     FTRAN is 16 FMAs against the entering column (4 distinct shared addresses per warp load),
//     the rank-1 update is 16 FMAs against the scaled pivot row, no B^-1 traffic at all;
//   * pricing runs over non-basic POSITIONS, one half-warp per position (lane hl owns rows
//     hl, hl+16, hl+32, hl+48 and keeps those four duals in registers), 8 positions per half-warp
//     and trip, reduced by a transposed butterfly (8 shuffles for 8 dot products); a position
//     that holds a unit column costs one multiply;
//   * duals follow the recurrence y' = y + s_q * (new row p of B^-1) and are recomputed from
//     scratch every BF_DUAL_REFRESH pivots;
//   * shared memory per CTA is (N - m) * 65 * 8 B + ~12 KB (79 KB at config 4) and the kernel is
//     compiled for <= 128 registers, so two CTAs (16 warps) share an SM.
// An LP whose start basis is not made of unit columns, or that has more than N - m non-unit
// columns, is marked TERM_PENDING and left to the generic kernel (k_batched, sb200_batched.cuh).
#pragma once
#include <cfloat>
#include "sb200_batched.cuh"

namespace sb200
{
    constexpr int BF_THREADS = 256;
    constexpr int BF_ROWS = 64;          // rows the register layout covers (m <= 64, padded)
    constexpr int BF_LDA = 65;           // odd column stride: conflict-free half-warp reads
    constexpr int BF_DUAL_REFRESH = 32;  // pivots between full recomputations of the duals
    constexpr int TERM_PENDING = -2;     // fast kernel could not take this LP
    constexpr int BF_DENSE_MARK = 0x40000000;

    __host__ __device__ inline size_t batched_fast_smem_bytes(int N, int D)
    {
        const size_t doubles = static_cast<size_t>(D) * BF_LDA   // dense columns
                               + 2 * static_cast<size_t>(N)      // c, unit values
                               + 4 * BF_ROWS                     // y, prow, xB, b
                               + 8 * BF_ROWS;                    // dual refresh scratch
        const size_t ints = 2 * static_cast<size_t>(N) + BF_ROWS + 8;  // code, nonbasic, basic
        return doubles * sizeof(double) + ints * sizeof(int);
    }

    // (key, idx) arg-min with ties to the lower idx; idx < 0 = no candidate. aux rides along.
    struct ArgMin
    {
        double key, aux;
        int idx;
    };

    // Order-preserving map double -> uint64 (no NaNs reach it): unsigned order == numeric order.
    __device__ __forceinline__ unsigned long long ordered_bits(double v)
    {
        const unsigned long long u = static_cast<unsigned long long>(__double_as_longlong(v));
        return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
    }
    __device__ __forceinline__ double from_ordered_bits(unsigned long long u)
    {
        return __longlong_as_double(static_cast<long long>((u >> 63) ? (u & 0x7fffffffffffffffull) : ~u));
    }

    // Three redux.sync.min steps (high word, low word, index) instead of a 5-level shuffle tree.
    __device__ __forceinline__ ArgMin warp_argmin(double key, double aux, int idx)
    {
        const bool valid = idx >= 0;
        const unsigned long long u = valid ? ordered_bits(key) : ~0ull;  // valid keys are finite: below the sentinel
        const unsigned hi = static_cast<unsigned>(u >> 32), lo = static_cast<unsigned>(u);
        const unsigned hmin = __reduce_min_sync(0xffffffffu, hi);
        const unsigned lmin = __reduce_min_sync(0xffffffffu, (hi == hmin) ? lo : 0xffffffffu);
        const bool at_min = valid && hi == hmin && lo == lmin;
        const unsigned imin = __reduce_min_sync(0xffffffffu, at_min ? static_cast<unsigned>(idx) : 0xffffffffu);
        const unsigned win = __ballot_sync(0xffffffffu, at_min && static_cast<unsigned>(idx) == imin);
        ArgMin r;
        r.idx = (imin == 0xffffffffu) ? -1 : static_cast<int>(imin);
        r.key = from_ordered_bits((static_cast<unsigned long long>(hmin) << 32) | lmin);
        r.aux = __shfl_sync(0xffffffffu, aux, win ? (__ffs(win) - 1) : 0);
        return r;
    }

    // Block-wide arg-min with ONE barrier: every warp publishes its winner, then every warp reduces
    // the (<= 8) winners redundantly. `slots` must not be reused before the next barrier.
    __device__ __forceinline__ ArgMin block_argmin(double key, double aux, int idx, ArgMin * slots, int lane, int warp)
    {
        const ArgMin w = warp_argmin(key, aux, idx);
        if (lane == 0) slots[warp] = w;
        __syncthreads();
        ArgMin o;
        o.idx = -1;
        o.key = 0.0;
        o.aux = 0.0;
        if (lane < BF_THREADS / 32) o = slots[lane];
        return warp_argmin(o.key, o.aux, o.idx);
    }

    __global__ void __launch_bounds__(BF_THREADS, 2) k_batched_fast(BatchedParams P, int D, int * n_pending)
    {
        extern __shared__ __align__(16) double smem[];
        __shared__ ArgMin red_price[BF_THREADS / 32], red_ratio[BF_THREADS / 32];
        __shared__ int sh_info[4];  // 0: dense count, 1: start basis is not a unit basis

        const int m = P.m, N = P.N, nnb = N - m;
        double * As = smem;
        double * cs = As + static_cast<size_t>(D) * BF_LDA;
        double * uval = cs + N;
        double * y = uval + N;
        double * prow = y + BF_ROWS;
        double * xBs = prow + BF_ROWS;
        double * bs = xBs + BF_ROWS;
        double * yscr = bs + BF_ROWS;  // [8][64]
        int * code = reinterpret_cast<int *>(yscr + 8 * BF_ROWS);  // >= 0: dense slot; < 0: unit column, row = -code-1
        int * nonbasic = code + N;
        int * basic = nonbasic + N;

        const int lp = blockIdx.x;
        const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
        const int r = tid >> 2, k = tid & 3;    // B^-1 ownership
        const int hw = tid >> 4, hl = tid & 15;  // pricing: half-warp, lane within it
        const double * gA = P.A + static_cast<size_t>(lp) * N * m;

        // ---- lists, b, c
        for (int i = tid; i < BF_ROWS; i += BF_THREADS)
        {
            bs[i] = (i < m) ? P.b[static_cast<size_t>(lp) * m + i] : 0.0;
            basic[i] = (i < m) ? P.basic[static_cast<size_t>(lp) * m + i] : -1;
        }
        for (int j = tid; j < N; j += BF_THREADS) cs[j] = P.c[static_cast<size_t>(lp) * N + j];
        for (int j = tid; j < nnb; j += BF_THREADS) nonbasic[j] = P.nonbasic[static_cast<size_t>(lp) * nnb + j];
        if (tid < 4) sh_info[tid] = 0;

        // ---- pass 1: classify the columns (one warp per column; a column is 2 values per lane)
        for (int j = warp; j < N; j += BF_THREADS / 32)
        {
            const double v0 = (lane < m) ? gA[static_cast<size_t>(j) * m + lane] : 0.0;
            const double v1 = (lane + 32 < m) ? gA[static_cast<size_t>(j) * m + lane + 32] : 0.0;
            const unsigned nz0 = __ballot_sync(0xffffffffu, v0 != 0.0);
            const unsigned nz1 = __ballot_sync(0xffffffffu, v1 != 0.0);
            if (__popc(nz0) + __popc(nz1) == 1)
            {
                const int src = nz0 ? (__ffs(nz0) - 1) : (__ffs(nz1) - 1);
                const double val = __shfl_sync(0xffffffffu, nz0 ? v0 : v1, src);
                if (lane == 0)
                {
                    code[j] = -((nz0 ? src : src + 32) + 1);
                    uval[j] = val;
                }
            }
            else if (lane == 0)
            {
                code[j] = BF_DENSE_MARK;
                uval[j] = 0.0;
            }
        }
        __syncthreads();
        // ---- slots for the dense columns, in column order (warp 0, ballot prefix)
        if (warp == 0)
        {
            int base = 0;
            for (int j0 = 0; j0 < N; j0 += 32)
            {
                const int j = j0 + lane;
                const bool dense = j < N && code[j] == BF_DENSE_MARK;
                const unsigned mask = __ballot_sync(0xffffffffu, dense);
                if (dense) code[j] = base + __popc(mask & ((1u << lane) - 1u));
                base += __popc(mask);
            }
            if (lane == 0) sh_info[0] = base;
        }
        // the start basis must be a unit basis: column basic[i] = value * e_i
        if (tid < m)
        {
            const int col = basic[tid];
            // (code of a dense column is still BF_DENSE_MARK or a slot >= 0 here: both fail the test)
            if (col < 0 || col >= N || code[col] != -(tid + 1) || uval[col] == 0.0) atomicOr(&sh_info[1], 1);
        }
        __syncthreads();
        if (sh_info[0] > D || sh_info[1] != 0)
        {
            if (tid == 0)
            {
                P.term[lp] = TERM_PENDING;
                atomicAdd(n_pending, 1);
            }
            return;
        }
        // ---- pass 2: dense columns into their slots (rows m..63 zero)
        for (int j = warp; j < N; j += BF_THREADS / 32)
        {
            const int slot = code[j];
            if (slot < 0) continue;
            double * dst = As + static_cast<size_t>(slot) * BF_LDA;
            dst[lane] = (lane < m) ? gA[static_cast<size_t>(j) * m + lane] : 0.0;
            dst[lane + 32] = (lane + 32 < m) ? gA[static_cast<size_t>(j) * m + lane + 32] : 0.0;
        }

        // ---- K0: B^-1 = diag(1 / a_ii), x_B = B^-1 b
        double B[16];
        double xr = 0.0, cBr = 0.0;
        {
            const double inv = (r < m) ? 1.0 / uval[basic[r]] : 1.0;
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) B[jj] = (k + 4 * jj == r) ? inv : 0.0;
            if (r < m)
            {
                xr = inv * bs[r];
                cBr = cs[basic[r]];
            }
        }
        __syncthreads();

        int status = TERM_RUNNING;
        int iterations = 0, since_refresh = BF_DUAL_REFRESH;  // forces the first full dual solve
        const double eps = P.eps, neg_eps = -P.eps;
        double y0 = 0.0, y1 = 0.0, y2 = 0.0, y3 = 0.0;  // duals of rows hl, hl+16, hl+32, hl+48
        const int b3 = hl & 8, b2 = hl & 4, b1 = hl & 2;
        const int tsel = ((hl >> 3) & 1) * 4 + ((hl >> 2) & 1) * 2 + ((hl >> 1) & 1);
        // Three block barriers per pivot: after the pricing arg-min, after the ratio arg-min, after
        // the owners of row p have published the scaled pivot row.
        while (status == TERM_RUNNING)
        {
            iterations += 1;
            // ---- K1 (every BF_DUAL_REFRESH pivots): y_j = sum_r cB_r * Binv[r][j]
            if (since_refresh >= BF_DUAL_REFRESH)
            {
                since_refresh = 0;
#pragma unroll
                for (int jj = 0; jj < 16; ++jj)
                {
                    double v = cBr * B[jj];
                    v += __shfl_xor_sync(0xffffffffu, v, 4);
                    v += __shfl_xor_sync(0xffffffffu, v, 8);
                    v += __shfl_xor_sync(0xffffffffu, v, 16);
                    if (lane < 4) yscr[warp * BF_ROWS + k + 4 * jj] = v;
                }
                __syncthreads();
                if (tid < BF_ROWS)
                {
                    double v = 0.0;
#pragma unroll
                    for (int w = 0; w < 8; ++w) v += yscr[w * BF_ROWS + tid];
                    y[tid] = v;
                }
                __syncthreads();
                y0 = y[hl];
                y1 = y[hl + 16];
                y2 = y[hl + 32];
                y3 = y[hl + 48];
            }
            // ---- K2: pricing over positions, one half-warp per position
            double best_key = 0.0, best_aux = 0.0;
            int best_idx = -1;
            for (int base = 0; base < nnb; base += 128)
            {
                double v[8];
#pragma unroll
                for (int t = 0; t < 8; ++t)
                {
                    const int pos = base + hw + 16 * t;
                    double acc = 0.0;
                    if (pos < nnb)
                    {
                        const int col = nonbasic[pos];
                        const int cd = code[col];
                        if (cd >= 0)
                        {
                            const double * a = As + static_cast<size_t>(cd) * BF_LDA + hl;
                            acc = a[0] * y0;
                            acc = fma(a[16], y1, acc);
                            acc = fma(a[32], y2, acc);
                            acc = fma(a[48], y3, acc);
                        }
                        else
                        {
                            // unit column value * e_row: only the lane that owns `row` contributes
                            const int row = -cd - 1;
                            if ((row & 15) == hl)
                            {
                                const int g = row >> 4;
                                const double yr = (g == 0) ? y0 : (g == 1) ? y1 : (g == 2) ? y2 : y3;
                                acc = uval[col] * yr;
                            }
                        }
                    }
                    v[t] = acc;
                }
                // transposed butterfly: 8 sums over 16 lanes in 4 + 2 + 1 + 1 shuffles
                double w4[4], w2[2];
#pragma unroll
                for (int i = 0; i < 4; ++i)
                {
                    const double send = b3 ? v[i] : v[i + 4];
                    const double keep = b3 ? v[i + 4] : v[i];
                    w4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
                }
#pragma unroll
                for (int i = 0; i < 2; ++i)
                {
                    const double send = b2 ? w4[i] : w4[i + 2];
                    const double keep = b2 ? w4[i + 2] : w4[i];
                    w2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
                }
                double z;
                {
                    const double send = b1 ? w2[0] : w2[1];
                    const double keep = b1 ? w2[1] : w2[0];
                    z = keep + __shfl_xor_sync(0xffffffffu, send, 2);
                }
                z += __shfl_xor_sync(0xffffffffu, z, 1);
                const int pos = base + hw + 16 * tsel;
                if ((hl & 1) == 0 && pos < nnb)
                {
                    const double sj = cs[nonbasic[pos]] - z;
                    if (sj < neg_eps)
                    {
                        const double key = (P.rule == RULE_MOST_NEG) ? sj : 0.0;
                        if (best_idx < 0 || key < best_key)  // positions ascend with `base`: ties keep the earlier
                        {
                            best_key = key;
                            best_aux = sj;
                            best_idx = pos;
                        }
                    }
                }
            }
            const ArgMin ent = block_argmin(best_key, best_aux, best_idx, red_price, lane, warp);
            if (ent.idx < 0)
            {
                status = (iterations >= P.iter_limit) ? TERM_ITER_LIMIT : TERM_OPTIMAL;
                break;
            }
            const int e = ent.idx;
            const int q = nonbasic[e];
            const double s_q = ent.aux;
            // ---- K3: d_r = Binv[r][:] . a_q from registers, ratio candidates (Bland, :66-84)
            double dr;
            {
                const int cd = code[q];
                double acc = 0.0;
                if (cd >= 0)
                {
                    const double * a = As + static_cast<size_t>(cd) * BF_LDA + k;
#pragma unroll
                    for (int jj = 0; jj < 16; ++jj) acc = fma(B[jj], a[4 * jj], acc);
                }
                else
                {
                    const int row = -cd - 1;
                    const double uv = uval[q];
#pragma unroll
                    for (int jj = 0; jj < 16; ++jj) acc = fma(B[jj], (k + 4 * jj == row) ? uv : 0.0, acc);
                }
                acc += __shfl_xor_sync(0xffffffffu, acc, 1);
                acc += __shfl_xor_sync(0xffffffffu, acc, 2);
                dr = acc;
            }
            double rkey = 0.0;
            int ridx = -1;
            if (k == 0 && r < m && dr > eps)
            {
                const double t = xr / dr;
                if (t < DBL_MAX)
                {
                    rkey = t;
                    ridx = r;
                }
            }
            const ArgMin lv = block_argmin(rkey, dr, ridx, red_ratio, lane, warp);
            if (lv.idx < 0)
            {
                status = TERM_UNBOUNDED;
                break;
            }
            // ---- K4/K5: pivot on (p, q). The four owners of row p publish the scaled pivot row.
            const int p = lv.idx;
            const double dp = lv.aux, theta = lv.key;
            if (r == p)
            {
                // one division on the critical path of the block (the other 252 threads wait at the
                // barrier below for these four): row p times the reciprocal of the pivot element
                const double rdp = 1.0 / dp;
#pragma unroll
                for (int jj = 0; jj < 16; ++jj)
                {
                    B[jj] *= rdp;
                    prow[k + 4 * jj] = B[jj];
                }
                xr = theta;
                cBr = cs[q];
                if (k == 0)
                {
                    const int leaving = basic[p];
                    basic[p] = q;
                    nonbasic[e] = leaving;
                }
            }
            __syncthreads();
            if (r != p)
            {
#pragma unroll
                for (int jj = 0; jj < 16; ++jj) B[jj] = fma(-dr, prow[k + 4 * jj], B[jj]);
                xr = fma(-theta, dr, xr);
            }
            y0 = fma(s_q, prow[hl], y0);
            y1 = fma(s_q, prow[hl + 16], y1);
            y2 = fma(s_q, prow[hl + 32], y2);
            y3 = fma(s_q, prow[hl + 48], y3);
            since_refresh += 1;
            if (iterations >= P.iter_limit) status = TERM_ITER_LIMIT;
        }

        // ---- write back
        if (k == 0) xBs[r] = xr;
        __syncthreads();
        for (int i = tid; i < m; i += BF_THREADS)
        {
            P.basic[static_cast<size_t>(lp) * m + i] = basic[i];
            P.xB[static_cast<size_t>(lp) * m + i] = xBs[i];
        }
        for (int j = tid; j < nnb; j += BF_THREADS) P.nonbasic[static_cast<size_t>(lp) * nnb + j] = nonbasic[j];
        if (tid == 0)
        {
            double obj = 0.0;
            for (int i = 0; i < m; ++i) obj += cs[basic[i]] * xBs[i];
            P.objective[lp] = obj;
            P.term[lp] = status;
            P.iterations[lp] = iterations;
        }
    }
}  // namespace sb200

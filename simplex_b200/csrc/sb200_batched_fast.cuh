// K7-fast — batched mode for small LPs (m <= 64): ONE CTA PER LP, the basis inverse lives in
// REGISTERS, the dense columns of A in shared memory, unit columns (slacks, artificials) are
// kept as (row, value) pairs. Same algorithm and tie-breaking as k_batched / the reference loop
// (src/Simplex.cpp:255-353 with the Bland ratio test :66-84 and both entering rules :49-61,
// :87-102); BASELINE configs[4] (65536 LPs of m=64, n=128) is the shape it is sized for.
//
// Why this layout: per pivot an LP of this size needs ~16k fp64 FMAs and, with everything in
// shared memory, ~150 KB of shared-memory traffic: the generic kernel is bound by shared-memory
// bandwidth and, with 135 KB per CTA, by having one CTA (8 warps) per SM. Here
//   * thread (r, k) = (tid >> 2, tid & 3) owns B^-1[r][8*jj + 2*k + {0,1}], jj = 0..7, in 32
//     registers: FTRAN is 16 FMAs against the entering column (8 128-bit shared loads, 4 distinct
//     addresses per warp), the rank-1 update is 16 FMAs against the scaled pivot row, no B^-1
//     traffic at all;
//   * pricing runs over non-basic POSITIONS, one half-warp per position (lane hl owns rows
//     2hl, 2hl+1, 2hl+32, 2hl+33 and keeps those four duals in registers), 8 positions per
//     half-warp and trip, reduced by a transposed butterfly (8 shuffles for 8 dot products; lane l
//     visits its 8 positions in the order i ^ s(l), so the butterfly needs no selects); the shared
//     memory offsets of a thread's 8 columns live in registers when N - m <= 128 (one entry changes
//     per pivot); a position that holds a unit column costs one multiply;
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
    constexpr int BF_LDA = 66;           // even column stride: 128-bit shared loads stay 16-byte aligned
    constexpr int BF_DUAL_REFRESH = 32;  // pivots between full recomputations of the duals
    constexpr int TERM_PENDING = -2;     // fast kernel could not take this LP
    constexpr int BF_DENSE_MARK = 0x40000000;

    __host__ __device__ inline size_t batched_fast_smem_bytes(int N, int D)
    {
        const size_t doubles = static_cast<size_t>(D) * BF_LDA   // dense columns
                               + 4 * BF_ROWS                     // y, prow, xB, b
                               + 8 * BF_ROWS                     // dual refresh scratch
                               + 2 * static_cast<size_t>(N);     // c, unit values
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

    constexpr int BF_SKIP = static_cast<int>(0x80000000u);  // cached column offset: position beyond the list

    // kCached  : N - m <= 128, the offsets of a thread's 8 pricing columns live in registers
    // kAllDense: D >= N, every column (unit ones too) has a dense slot = its id: no classification
    //            pass and no unit/dense branches in the pivot loop (the shape of BASELINE configs[4])
    template <bool kCached, bool kAllDense>
    __global__ void __launch_bounds__(BF_THREADS, 2) k_batched_fast(BatchedParams P, int D, int * n_pending)
    {
        extern __shared__ __align__(16) double smem[];
        __shared__ ArgMin red_price[BF_THREADS / 32], red_ratio[BF_THREADS / 32];
        __shared__ int sh_info[4];  // 0: dense count, 1: start basis is not a unit basis

        const int m = P.m, N = P.N, nnb = N - m;
        double * As = smem;
        double * y = As + static_cast<size_t>(D) * BF_LDA;
        double * prow = y + BF_ROWS;
        double * xBs = prow + BF_ROWS;
        double * bs = xBs + BF_ROWS;
        double * yscr = bs + BF_ROWS;  // [8][64]
        double * cs = yscr + 8 * BF_ROWS;
        double * uval = cs + N;
        int * code = reinterpret_cast<int *>(uval + N);  // >= 0: dense slot; < 0: unit column, row = -code-1
        int * nonbasic = code + N;
        int * basic = nonbasic + N;

        const int lp = blockIdx.x;
        const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
        const int r = tid >> 2, k = tid & 3;    // B^-1 ownership: row r, columns 8*jj + 2*k + {0,1}
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

        if constexpr (kAllDense)
        {
            // ---- every column into the slot of its own id (rows m..63 zero); four columns in flight per warp
            for (int j0 = 4 * warp; j0 < N; j0 += 4 * (BF_THREADS / 32))
            {
                double v0[4], v1[4];
#pragma unroll
                for (int u = 0; u < 4; ++u)
                {
                    const int j = j0 + u;
                    v0[u] = (j < N && lane < m) ? gA[static_cast<size_t>(j) * m + lane] : 0.0;
                    v1[u] = (j < N && lane + 32 < m) ? gA[static_cast<size_t>(j) * m + lane + 32] : 0.0;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u)
                {
                    const int j = j0 + u;
                    if (j < N)
                    {
                        As[static_cast<size_t>(j) * BF_LDA + lane] = v0[u];
                        As[static_cast<size_t>(j) * BF_LDA + lane + 32] = v1[u];
                    }
                }
            }
            __syncthreads();
            // the start basis must be a unit basis: column basic[i] = value * e_i
            if (tid < m)
            {
                const int col = basic[tid];
                int bad = (col < 0 || col >= N);
                if (!bad)
                {
                    const double * a = As + static_cast<size_t>(col) * BF_LDA;
                    for (int i = 0; i < m; ++i) bad |= (i == tid) ? (a[i] == 0.0) : (a[i] != 0.0);
                }
                if (bad) atomicOr(&sh_info[1], 1);
            }
            __syncthreads();
            if (sh_info[1] != 0)
            {
                if (tid == 0)
                {
                    P.term[lp] = TERM_PENDING;
                    atomicAdd(n_pending, 1);
                }
                return;
            }
        }
        else
        {
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
        }

        // ---- K0: B^-1 = diag(1 / a_ii), x_B = B^-1 b
        double B[16];  // B[2*jj + h] = Binv[r][8*jj + 2*k + h]
        double xr = 0.0, cBr = 0.0;
        {
            double inv = 1.0;
            if (r < m) inv = 1.0 / (kAllDense ? As[static_cast<size_t>(basic[r]) * BF_LDA + r] : uval[basic[r]]);
#pragma unroll
            for (int i = 0; i < 16; ++i) B[i] = (8 * (i >> 1) + 2 * k + (i & 1) == r) ? inv : 0.0;
            if (r < m)
            {
                xr = inv * bs[r];
                cBr = cs[basic[r]];
            }
        }
        __syncthreads();

        // smem offset of the column at a non-basic position, as the pricing loop wants it:
        // >= 0 dense (offset of the slot), BF_SKIP beyond the list, otherwise -(col + 1) = unit column
        auto encode_col = [&](int col) -> int {
            if (kAllDense) return col * BF_LDA;
            const int cd = code[col];
            return (cd >= 0) ? cd * BF_LDA : -(col + 1);
        };
        const int tsel = ((hl >> 3) & 1) * 4 + ((hl >> 2) & 1) * 2 + ((hl >> 1) & 1);  // this lane's butterfly slot
        int coff[8];
        double mycost = 0.0;
        if (kCached)
        {
#pragma unroll
            for (int i = 0; i < 8; ++i)
            {
                const int pos = hw + 16 * (i ^ tsel);
                coff[i] = (pos < nnb) ? encode_col(nonbasic[pos]) : BF_SKIP;
            }
            if (hw + 16 * tsel < nnb) mycost = cs[nonbasic[hw + 16 * tsel]];
        }

        int status = TERM_RUNNING;
        int iterations = 0, since_refresh = BF_DUAL_REFRESH;  // forces the first full dual solve
        const double eps = P.eps, neg_eps = -P.eps;
        double2 ya = make_double2(0.0, 0.0), yb = ya;  // duals of rows (2hl, 2hl+1) and (2hl+32, 2hl+33)
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
                for (int i = 0; i < 16; ++i)
                {
                    double v = cBr * B[i];
                    v += __shfl_xor_sync(0xffffffffu, v, 4);
                    v += __shfl_xor_sync(0xffffffffu, v, 8);
                    v += __shfl_xor_sync(0xffffffffu, v, 16);
                    if (lane < 4) yscr[warp * BF_ROWS + 8 * (i >> 1) + 2 * k + (i & 1)] = v;
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
                ya = *reinterpret_cast<const double2 *>(y + 2 * hl);
                yb = *reinterpret_cast<const double2 *>(y + 2 * hl + 32);
            }
            // ---- K2: pricing over positions, one half-warp per position
            double best_key = 0.0, best_aux = 0.0;
            int best_idx = -1;
            for (int base = 0; base < nnb; base += 128)
            {
                if (!kCached)
                {
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                    {
                        const int pos = base + hw + 16 * (i ^ tsel);
                        coff[i] = (pos < nnb) ? encode_col(nonbasic[pos]) : BF_SKIP;
                    }
                    const int fpos = base + hw + 16 * tsel;
                    mycost = (fpos < nnb) ? cs[nonbasic[fpos]] : 0.0;
                }
                double v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i)
                {
                    const int o = coff[i];
                    double acc = 0.0;
                    if (kAllDense ? (o != BF_SKIP) : (o >= 0))
                    {
                        const double * a = As + o + 2 * hl;
                        const double2 a0 = *reinterpret_cast<const double2 *>(a);
                        const double2 a1 = *reinterpret_cast<const double2 *>(a + 32);
                        acc = a0.x * ya.x;
                        acc = fma(a0.y, ya.y, acc);
                        acc = fma(a1.x, yb.x, acc);
                        acc = fma(a1.y, yb.y, acc);
                    }
                    else if (!kAllDense && o != BF_SKIP)
                    {
                        // unit column value * e_row: only the lane that owns `row` contributes
                        const int col = -o - 1;
                        const int row = -code[col] - 1;
                        if (((row & 31) >> 1) == hl)
                        {
                            const double2 yy = (row & 32) ? yb : ya;
                            acc = uval[col] * ((row & 1) ? yy.y : yy.x);
                        }
                    }
                    v[i] = acc;
                }
                // transposed butterfly: 8 sums over 16 lanes in 4 + 2 + 1 + 1 shuffles. Lane l filled
                // v[i] with position slot i ^ tsel(l), so "keep the low half, send the high half" is
                // the same instruction in every lane.
                double w4[4], w2[2];
#pragma unroll
                for (int i = 0; i < 4; ++i) w4[i] = v[i] + __shfl_xor_sync(0xffffffffu, v[i + 4], 8);
#pragma unroll
                for (int i = 0; i < 2; ++i) w2[i] = w4[i] + __shfl_xor_sync(0xffffffffu, w4[i + 2], 4);
                double z = w2[0] + __shfl_xor_sync(0xffffffffu, w2[1], 2);
                z += __shfl_xor_sync(0xffffffffu, z, 1);
                const int pos = base + hw + 16 * tsel;
                if ((hl & 1) == 0 && pos < nnb)
                {
                    const double sj = mycost - z;
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
                const int cd = kAllDense ? q : code[q];
                double acc0 = 0.0, acc1 = 0.0;
                if (kAllDense || cd >= 0)
                {
                    const double * a = As + static_cast<size_t>(cd) * BF_LDA + 2 * k;
#pragma unroll
                    for (int jj = 0; jj < 8; ++jj)
                    {
                        const double2 av = *reinterpret_cast<const double2 *>(a + 8 * jj);
                        acc0 = fma(B[2 * jj], av.x, acc0);
                        acc1 = fma(B[2 * jj + 1], av.y, acc1);
                    }
                }
                else
                {
                    const int row = -cd - 1;
                    const double uv = uval[q];
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        acc0 = fma(B[i], (8 * (i >> 1) + 2 * k + (i & 1) == row) ? uv : 0.0, acc0);
                }
                double acc = acc0 + acc1;
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
            // Index lists (Basis::update, src/Basis.cpp:48-53). Cached mode: every thread needs the leaving
            // column (to patch its cached offsets), so all read basic[p] here and thread 0 swaps after
            // the barrier; nothing reads the lists again before the next barrier. Otherwise the next
            // pricing reads nonbasic[] right away, so the owner swaps BEFORE the barrier and is the only
            // reader of basic[p].
            int leaving = 0;
            if (kCached) leaving = basic[p];
            if (r == p)
            {
                // one division on the critical path of the block (the other 252 threads wait at the
                // barrier below for these four): row p times the reciprocal of the pivot element
                const double rdp = 1.0 / dp;
#pragma unroll
                for (int jj = 0; jj < 8; ++jj)
                {
                    B[2 * jj] *= rdp;
                    B[2 * jj + 1] *= rdp;
                    *reinterpret_cast<double2 *>(prow + 8 * jj + 2 * k) = make_double2(B[2 * jj], B[2 * jj + 1]);
                }
                xr = theta;
                cBr = cs[q];
                if (!kCached && k == 0)
                {
                    leaving = basic[p];
                    basic[p] = q;
                    nonbasic[e] = leaving;
                }
            }
            if (kCached)
            {
                // position e now holds the leaving column: patch the cached offset / cost that covers it
                if (hw == (e & 15))
                {
                    const int enc = encode_col(leaving);
                    const int slot = (e >> 4) ^ tsel;
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        if (i == slot) coff[i] = enc;
                    if (slot == 0) mycost = cs[leaving];
                }
            }
            __syncthreads();
            if (kCached && tid == 0)
            {
                basic[p] = q;
                nonbasic[e] = leaving;
            }
            if (r != p)
            {
#pragma unroll
                for (int jj = 0; jj < 8; ++jj)
                {
                    const double2 pr = *reinterpret_cast<const double2 *>(prow + 8 * jj + 2 * k);
                    B[2 * jj] = fma(-dr, pr.x, B[2 * jj]);
                    B[2 * jj + 1] = fma(-dr, pr.y, B[2 * jj + 1]);
                }
                xr = fma(-theta, dr, xr);
            }
            {
                const double2 pa = *reinterpret_cast<const double2 *>(prow + 2 * hl);
                const double2 pb = *reinterpret_cast<const double2 *>(prow + 2 * hl + 32);
                ya.x = fma(s_q, pa.x, ya.x);
                ya.y = fma(s_q, pa.y, ya.y);
                yb.x = fma(s_q, pb.x, yb.x);
                yb.y = fma(s_q, pb.y, yb.y);
            }
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

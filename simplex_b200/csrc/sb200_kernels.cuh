// __global__ wrappers of the staged engine (one kernel per phase) and the one-off set-up
// kernels (K0). See sb200_phases.cuh for the arithmetic and the reference lines it replaces.
#pragma once
#include "sb200_phases.cuh"

namespace sb200
{
    constexpr int PRICE_THREADS = 512;
    constexpr int FTRAN_THREADS = 512;
    constexpr int SELECT_THREADS = 1024;
    constexpr int UPDATE_THREADS = 256;
    constexpr int UPDATE_ROWS = 32;  // rows per rank-1 tile

    // Stage a length-n (multiple of 2) global vector into shared memory, all threads.
    __device__ __forceinline__ void stage_vector(double * dst, const double * __restrict__ src, int n)
    {
        for (int k = 2 * threadIdx.x; k < n; k += 2 * blockDim.x)
        {
            *reinterpret_cast<double2 *>(dst + k) = ld_keep2(src + k);
        }
    }

    // ---- K2 ------------------------------------------------------------------------------
    __global__ void __launch_bounds__(PRICE_THREADS) k_price(DevState S)
    {
        extern __shared__ __align__(16) double smem[];
        __shared__ Cand scratch[33];
        if (S.sc->status != TERM_RUNNING) return;
        stage_vector(smem, S.y, S.ld);
        __syncthreads();
        const int lane = threadIdx.x & 31;
        const int wpb = blockDim.x >> 5;
        const int gw = blockIdx.x * wpb + (threadIdx.x >> 5);
        Cand best = price_warp(S, smem, gw, gridDim.x * wpb, lane);
        if (lane != 0) best = cand_none();
        best = block_reduce_cand(best, scratch);
        if (threadIdx.x == 0) S.price_slots[blockIdx.x] = best;
    }

    // Step-API helper: commit the entering decision without running FTRAN.
    __global__ void k_pick_entering(DevState S)
    {
        __shared__ Cand scratch[33];
        if (S.sc->status != TERM_RUNNING) return;
        const Cand best = block_reduce_slots(S.price_slots, S.n_price_slots, scratch);
        if (threadIdx.x == 0) commit_entering(S, best);
    }

    // ---- K3 ------------------------------------------------------------------------------
    // kPick: the prologue reduces the pricing slots (every block redundantly, block 0 commits).
    template <bool kPick>
    __global__ void __launch_bounds__(FTRAN_THREADS) k_ftran(DevState S)
    {
        extern __shared__ __align__(16) double smem[];
        __shared__ Cand scratch[33];
        // the status is read ONCE per block: block 0 may write TERM_OPTIMAL (commit_entering) while other
        // blocks are still starting, and a block whose threads disagreed would run the barriers below short-handed
        __shared__ int status_at_entry;
        if (threadIdx.x == 0) status_at_entry = S.sc->status;
        __syncthreads();
        if (status_at_entry != TERM_RUNNING) return;
        int q;
        if (kPick)
        {
            const Cand best = block_reduce_slots(S.price_slots, S.n_price_slots, scratch);
            if (blockIdx.x == 0 && threadIdx.x == 0) commit_entering(S, best);
            if (best.idx < 0) return;
            q = S.nonbasic[best.idx];
        }
        else
        {
            if (S.sc->e_pos < 0) return;
            q = S.sc->q_col;
        }
        stage_vector(smem, S.A + static_cast<size_t>(q) * S.ld, S.ld);
        __syncthreads();
        const int lane = threadIdx.x & 31;
        const int wpb = blockDim.x >> 5;
        const int gw = blockIdx.x * wpb + (threadIdx.x >> 5);
        Cand best = ftran_warp(S, smem, gw, gridDim.x * wpb, lane);
        if (lane != 0) best = cand_none();
        best = block_reduce_cand(best, scratch);
        if (threadIdx.x == 0) S.ratio_slots[blockIdx.x] = best;
    }

    // ---- K4 ------------------------------------------------------------------------------
    __global__ void __launch_bounds__(SELECT_THREADS) k_select(DevState S)
    {
        __shared__ Cand scratch[33];
        if (S.sc->status != TERM_RUNNING)
        {
            // the loop has ended: make sure a left-over ACT_PIVOT is not applied twice by the
            // k_update launches that are still queued in the current chunk
            if (threadIdx.x == 0) S.sc->action = ACT_NONE;
            return;
        }
        select_block(S, scratch);
    }

    // ---- K5 ------------------------------------------------------------------------------
    __global__ void __launch_bounds__(UPDATE_THREADS) k_update(DevState S)
    {
        // status may already say ITER_LIMIT for the iteration whose update is still due
        if (S.sc->action != ACT_PIVOT) return;
        const int r0 = blockIdx.y * UPDATE_ROWS;
        const int r1 = min(r0 + UPDATE_ROWS, S.m);
        update_tile<8>(S, r0, r1, blockIdx.x * (2 * UPDATE_THREADS), S.sc->p_row);
    }

    // ---- K1 ------------------------------------------------------------------------------
    __global__ void __launch_bounds__(UPDATE_THREADS) k_dual_partial(DevState S, int rows_per_seg)
    {
        if (S.sc->status != TERM_RUNNING) return;
        const int seg = blockIdx.y;
        const int r0 = seg * rows_per_seg;
        const int r1 = min(r0 + rows_per_seg, S.m);
        dual_partial_tile(S, seg, r0, r1, blockIdx.x * (2 * UPDATE_THREADS));
    }

    __global__ void k_dual_reduce(DevState S)
    {
        if (S.sc->status != TERM_RUNNING) return;
        const int col = blockIdx.x * blockDim.x + threadIdx.x;
        if (col < S.ld) dual_reduce_col(S, col);
    }

    // ---- K0: set-up ------------------------------------------------------------------------
    // One warp per basis row i: is column basic[i] of A a multiple of e_i ? (slack / artificial
    // crash basis, src/InitialBasis.cpp:23-55). diag[i] = a_ii, *not_diag != 0 otherwise.
    __global__ void k_check_diag(DevState S, double * diag, int * not_diag)
    {
        const int lane = threadIdx.x & 31;
        const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
        if (i >= S.m) return;
        const double * col = S.A + static_cast<size_t>(S.basic[i]) * S.ld;
        int bad = 0;
        for (int r = lane; r < S.m; r += 32)
        {
            const double a = col[r];
            if (r == i)
            {
                if (a == 0.0) bad = 1;
            }
            else if (a != 0.0)
            {
                bad = 1;
            }
        }
        bad = __any_sync(0xffffffffu, bad);
        if (lane == 0)
        {
            diag[i] = col[i];
            if (bad) atomicOr(not_diag, 1);
        }
    }

    // One warp per column j of A: unit_row[j] = row of its only nonzero, unit_val[j] = that value;
    // unit_row[j] = -1 when the column has zero or more than one nonzero (incl. NaN entries).
    __global__ void k_classify_units(DevState S, int * unit_row, double * unit_val)
    {
        const int lane = threadIdx.x & 31;
        const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
        if (j >= S.N) return;
        const double * col = S.A + static_cast<size_t>(j) * S.ld;
        int count = 0, row = -1;
        double val = 0.0;
        for (int r = lane; r < S.m; r += 32)
        {
            const double a = col[r];
            if (a != 0.0)  // true for NaN as well
            {
                count += 1;
                row = r;
                val = a;
            }
        }
        const int total = __reduce_add_sync(0xffffffffu, count);
        const unsigned who = __ballot_sync(0xffffffffu, count > 0);
        if (total == 1)
        {
            const int src = __ffs(who) - 1;
            row = __shfl_sync(0xffffffffu, row, src);
            val = __shfl_sync(0xffffffffu, val, src);
        }
        if (lane == 0)
        {
            const bool unit = total == 1 && val == val;
            unit_row[j] = unit ? row : -1;
            unit_val[j] = unit ? val : 0.0;
        }
    }

    __global__ void k_set_diag_inverse(DevState S, const double * diag)
    {
        const int i = blockIdx.x * blockDim.x + threadIdx.x;
        if (i < S.m) S.Binv[static_cast<size_t>(i) * S.ld + i] = 1.0 / diag[i];
    }

    // W[i][j] = A[i, basic[j]] (row-major, ld), V = I.
    __global__ void k_gather_basis(DevState S, double * W, double * V)
    {
        const int j = blockIdx.x * blockDim.x + threadIdx.x;
        const int i = blockIdx.y;
        if (j >= S.ld) return;
        double w = 0.0;
        if (j < S.m) w = S.A[static_cast<size_t>(S.basic[j]) * S.ld + i];
        W[static_cast<size_t>(i) * S.ld + j] = w;
        V[static_cast<size_t>(i) * S.ld + j] = (i == j) ? 1.0 : 0.0;
    }

    // Gauss-Jordan step k, part 1 (one block): partial pivot search in column k (first row
    // attaining the maximum), row swap in W and V, scaling of the pivot row, multipliers out.
    __global__ void __launch_bounds__(1024) k_gj_pivot(int m, int ld, int k, double * W, double * V, double * fcol,
                                                       int * singular)
    {
        __shared__ Cand scratch[33];
        Cand c = cand_none();
        for (int i = k + threadIdx.x; i < m; i += blockDim.x)
        {
            Cand o;
            o.key = -fabs(W[static_cast<size_t>(i) * ld + k]);  // arg-max via arg-min of the negation
            o.aux = 0.0;
            o.idx = i;
            o.cls = 0;
            if (cand_better(o, c)) c = o;
        }
        c = block_reduce_cand(c, scratch);
        if (threadIdx.x == 0) scratch[32] = c;
        __syncthreads();
        c = scratch[32];
        const int r = c.idx;
        if (r < 0 || c.key == 0.0)
        {
            if (threadIdx.x == 0) *singular = 1;
            return;
        }
        double * wk = W + static_cast<size_t>(k) * ld;
        double * vk = V + static_cast<size_t>(k) * ld;
        if (r != k)
        {
            double * wr = W + static_cast<size_t>(r) * ld;
            double * vr = V + static_cast<size_t>(r) * ld;
            for (int j = threadIdx.x; j < ld; j += blockDim.x)
            {
                const double a = wk[j];
                wk[j] = wr[j];
                wr[j] = a;
                const double b = vk[j];
                vk[j] = vr[j];
                vr[j] = b;
            }
            __syncthreads();
        }
        const double piv = wk[k];
        __syncthreads();
        for (int j = threadIdx.x; j < ld; j += blockDim.x)
        {
            wk[j] = wk[j] / piv;
            vk[j] = vk[j] / piv;
        }
        for (int i = threadIdx.x; i < m; i += blockDim.x)
        {
            fcol[i] = (i == k) ? 0.0 : W[static_cast<size_t>(i) * ld + k];
        }
    }

    // Gauss-Jordan step k, part 2: row_i -= f_i * row_k for every i != k, in W and in V.
    __global__ void __launch_bounds__(UPDATE_THREADS) k_gj_elim(int m, int ld, int k, double * W, double * V,
                                                               const double * fcol, const int * singular)
    {
        if (*singular) return;
        const int col = blockIdx.x * (2 * UPDATE_THREADS) + 2 * threadIdx.x;
        if (col >= ld) return;
        const int r0 = blockIdx.y * UPDATE_ROWS;
        const int r1 = min(r0 + UPDATE_ROWS, m);
        const double2 wk = *reinterpret_cast<const double2 *>(W + static_cast<size_t>(k) * ld + col);
        const double2 vk = *reinterpret_cast<const double2 *>(V + static_cast<size_t>(k) * ld + col);
        for (int i = r0; i < r1; ++i)
        {
            const double f = fcol[i];
            if (f == 0.0) continue;  // also skips the pivot row itself
            double2 * wp = reinterpret_cast<double2 *>(W + static_cast<size_t>(i) * ld + col);
            double2 * vp = reinterpret_cast<double2 *>(V + static_cast<size_t>(i) * ld + col);
            double2 w = *wp, v = *vp;
            w.x = fma(-f, wk.x, w.x);
            w.y = fma(-f, wk.y, w.y);
            v.x = fma(-f, vk.x, v.x);
            v.y = fma(-f, vk.y, v.y);
            *wp = w;
            *vp = v;
        }
    }

    // out[i] = M[i,:] . v for a row-major M (x_B = B^-1 b). Warp per row, v staged in smem.
    __global__ void __launch_bounds__(FTRAN_THREADS) k_rows_dot(int m, int ld, const double * M, const double * v,
                                                                double * out)
    {
        extern __shared__ __align__(16) double smem[];
        for (int k = threadIdx.x; k < ld; k += blockDim.x) smem[k] = (k < m) ? v[k] : 0.0;
        __syncthreads();
        const int lane = threadIdx.x & 31;
        const int wpb = blockDim.x >> 5;
        for (int i = blockIdx.x * wpb + (threadIdx.x >> 5); i < m; i += gridDim.x * wpb)
        {
            const double r = warp_dot<false>(M + static_cast<size_t>(i) * ld, smem, ld, lane);
            if (lane == 0) out[i] = r;
        }
    }

    // ---- numerical hygiene, off the hot path. The reference has no counterpart: it
    // re-solves x_B = B^-1 b from a fresh LU in every iteration (src/Simplex.cpp:305-306).
    constexpr int HYG_SEGS = 32;

    // part[seg][i] = sum over the basis positions k of segment seg of A[i][basic[k]] * xB[k] (A column-major:
    // consecutive threads read consecutive rows of one column).
    __global__ void k_residual_partial(DevState S, double * part)
    {
        const int i = blockIdx.x * blockDim.x + threadIdx.x;
        const int per = (S.m + HYG_SEGS - 1) / HYG_SEGS;
        const int k0 = blockIdx.y * per;
        const int k1 = min(k0 + per, S.m);
        if (i >= S.m) return;
        double acc = 0.0;
        for (int k = k0; k < k1; ++k) acc = fma(S.A[static_cast<size_t>(S.basic[k]) * S.ld + i], S.xB[k], acc);
        part[static_cast<size_t>(blockIdx.y) * S.ld + i] = acc;
    }

    // res[i] = | b_i - sum of the partials in ascending order |; out[0] = max_i res[i], out[1] = max_i |b_i|
    // (bit patterns of non-negative doubles order like unsigned integers; a NaN residual saturates the maximum).
    __global__ void k_residual_reduce(DevState S, const double * part, double * res, unsigned long long * out)
    {
        const int i = blockIdx.x * blockDim.x + threadIdx.x;
        if (i >= S.m) return;
        double acc = 0.0;
        for (int s = 0; s < HYG_SEGS; ++s) acc += part[static_cast<size_t>(s) * S.ld + i];
        const double b = S.b[i];
        double r = fabs(b - acc);
        if (r != r) r = DBL_MAX;
        if (res != nullptr) res[i] = r;
        atomicMax(out, static_cast<unsigned long long>(__double_as_longlong(r)));
        atomicMax(out + 1, static_cast<unsigned long long>(__double_as_longlong(fabs(b))));
    }

    // Sampled rows of B^-1 A_B against the rows of I: block (x, s) stages row (first + s * stride) mod m of
    // B^-1 and its warps dot it with the basic columns; out[2] = max | e_r^T B^-1 A_B - e_r^T |.
    __global__ void __launch_bounds__(FTRAN_THREADS) k_inverse_residual(DevState S, int first, int stride,
                                                                        unsigned long long * out)
    {
        extern __shared__ __align__(16) double smem[];
        const int r = static_cast<int>((static_cast<long long>(first) + static_cast<long long>(blockIdx.y) * stride) % S.m);
        stage_vector(smem, S.Binv + static_cast<size_t>(r) * S.ld, S.ld);
        __syncthreads();
        const int lane = threadIdx.x & 31;
        const int wpb = blockDim.x >> 5;
        double worst = 0.0;
        for (int k = blockIdx.x * wpb + (threadIdx.x >> 5); k < S.m; k += gridDim.x * wpb)
        {
            const double v = warp_dot<false>(S.A + static_cast<size_t>(S.basic[k]) * S.ld, smem, S.ld, lane);
            double e = fabs(v - (k == r ? 1.0 : 0.0));
            if (e != e) e = DBL_MAX;
            worst = fmax(worst, e);
        }
        if (lane == 0) atomicMax(out + 2, static_cast<unsigned long long>(__double_as_longlong(worst)));
    }

    // An ending (optimal / unbounded) that the hygiene check did not accept is examined again with the fresh
    // inverse: the loop resumes at the iteration that reported it (not counted twice, src/Simplex.cpp:277).
    __global__ void k_resume(Scalars * sc)
    {
        sc->status = TERM_RUNNING;
        sc->iterations -= 1;
    }

    // Row i of B^-1 *= -1 for every basic variable whose flip flag is set, then clear all flags
    // (Phase I -> Phase II hand-off: the reference reverses the substitutions at optimality,
    // src/Simplex.cpp:194-198, and starts the next call with all flags false, :271).
    __global__ void k_unflip_rows(DevState S)
    {
        const int i = blockIdx.x;
        if (i >= S.m) return;
        if (!S.flip[S.basic[i]]) return;
        double * row = S.Binv + static_cast<size_t>(i) * S.ld;
        for (int j = threadIdx.x; j < S.ld; j += blockDim.x) row[j] = -row[j];
    }
}  // namespace sb200

// The phases of one revised-simplex iteration as __device__ functions. The staged engine wraps
// each in its own kernel; the persistent engine calls them back to back with grid barriers in
// between. Both therefore produce the same bits.
//
// Reference statements replaced (all in /root/reference/src):
//   dual_*      y = lu.transpose().solve(c_B)                      Simplex.cpp:307
//   price_*     s = c_N - A_N^T y  + select_entering_variable      Simplex.cpp:313, 318, 49-61, 87-102
//   ftran_*     d = lu.solve(A_entering) + ratio candidates        Simplex.cpp:326, 331, 66-84, 116-155
//   select_*    rest of select_leaving_variable(_SUB), Basis::update,
//               LinearProgram::upper_bound_substitution            Simplex.cpp:157-178, Basis.cpp:48-57,
//                                                                  LinearProgram.cpp:154-169
//   update_*    (no counterpart: the reference re-factorises A_B every iteration, Simplex.cpp:305)
#pragma once
#include <cfloat>
#include "sb200_state.cuh"

namespace sb200
{
    // ---------------------------------------------------------------- K2: pricing + arg-min
    // Every warp takes non-basic positions gw, gw+stride, ...; y must be in shared memory.
    // Returns this warp's best candidate in lane 0.
    __device__ __forceinline__ Cand price_warp(const DevState & S, const double * ysm, int gw, int stride, int lane)
    {
        Cand best = cand_none();
        const double neg_eps = -S.eps;
        for (int pos = gw; pos < S.nnb; pos += stride)
        {
            const int col = S.nonbasic[pos];
            const double dot = warp_dot<true>(S.A + static_cast<size_t>(col) * S.ld, ysm, S.ld, lane);
            const double s = S.c[col] - dot;  // Simplex.cpp:313
            if (s < neg_eps)                  // Simplex.cpp:55, 95
            {
                Cand c;
                c.key = (S.rule == RULE_MOST_NEG) ? s : 0.0;
                c.aux = s;
                c.idx = pos;
                c.cls = 0;
                if (cand_better(c, best)) best = c;
            }
        }
        return best;
    }

    // ---------------------------------------------------------------- K3: FTRAN + ratio candidates
    // aq (entering column) must be in shared memory. Rows gw, gw+stride, ... of B^-1.
    __device__ __forceinline__ Cand ftran_warp(const DevState & S, const double * aqsm, int gw, int stride, int lane)
    {
        Cand best = cand_none();
        for (int i = gw; i < S.m; i += stride)
        {
            const double di = warp_dot<false>(S.Binv + static_cast<size_t>(i) * S.ld, aqsm, S.ld, lane);
            if (lane == 0)
            {
                S.d[i] = di;
                const double xi = S.xB[i];
                Cand c = cand_none();
                if (di > S.eps)  // Simplex.cpp:73-79 / 125-132
                {
                    const double t = xi / di;
                    if (t < DBL_MAX)
                    {
                        c.key = t;
                        c.aux = di;
                        c.idx = i;
                        c.cls = 0;
                    }
                }
                else if (S.ub != nullptr && di < -S.eps)  // Simplex.cpp:136-153
                {
                    const double ubi = S.ub[S.basic[i]];
                    const double t = (ubi - xi) / (-di);
                    if (t < DBL_MAX)
                    {
                        c.key = t;
                        c.aux = di;
                        c.idx = i;
                        c.cls = 1;
                    }
                }
                if (cand_better(c, best)) best = c;
            }
        }
        return best;
    }

    // ---------------------------------------------------------------- entering decision
    // One thread. Accounts the iteration (Simplex.cpp:277) and records the entering column or
    // the optimal exit (Simplex.cpp:318-323).
    __device__ __forceinline__ void commit_entering(const DevState & S, const Cand & best)
    {
        Scalars * sc = S.sc;
        sc->iterations += 1;
        sc->action = ACT_NONE;
        if (best.idx < 0)
        {
            // The reference throws after the loop when the counter sits at the limit, even if
            // this very iteration found the optimum (Simplex.cpp:347-350).
            sc->status = (sc->iterations >= sc->iter_limit) ? TERM_ITER_LIMIT : TERM_OPTIMAL;
            sc->e_pos = -1;
            sc->last_ret = -1;
            return;
        }
        sc->e_pos = best.idx;
        sc->q_col = S.nonbasic[best.idx];
        sc->s_q = best.aux;
    }

    // ---------------------------------------------------------------- K4: leaving decision + bookkeeping
    // ONE block. Finishes select_leaving_variable(_SUB), applies bound flips, updates x_B and the
    // index lists, stages the scaled pivot row for the rank-1 kernel and (dual_update) the duals.
    __device__ __forceinline__ void upper_bound_substitution_block(const DevState & S, int col, double u)
    {
        // LinearProgram.cpp:154-169: c_j = -c_j ; b_i -= a_ij*u ; a_ij = -a_ij ; toggle flag.
        double * acol = S.A + static_cast<size_t>(col) * S.ld;
        for (int i = threadIdx.x; i < S.m; i += blockDim.x)
        {
            const double a = acol[i];
            S.b[i] = __dsub_rn(S.b[i], __dmul_rn(a, u));  // two roundings, as compiled g++ -O3 on x86-64
            acol[i] = -a;
        }
        if (threadIdx.x == 0)
        {
            S.c[col] = -S.c[col];
            S.flip[col] ^= 1;
        }
    }

    __device__ __forceinline__ void select_block(const DevState & S, Cand * scratch)
    {
        Scalars * sc = S.sc;
        const Cand best = block_reduce_slots(S.ratio_slots, S.n_ratio_slots, scratch);
        const int e = sc->e_pos;
        const int q = sc->q_col;
        const double min_theta = (best.idx >= 0) ? best.key : DBL_MAX;

        if (S.ub != nullptr)
        {
            const double ub_s = S.ub[q];
            if (ub_s < min_theta)  // Simplex.cpp:160-164: entering variable hits its own bound first
            {
                upper_bound_substitution_block(S, q, ub_s);
                // B^-1 is untouched; x_B = B^-1 (b - a_q u) = x_B - u d
                for (int i = threadIdx.x; i < S.m; i += blockDim.x) S.xB[i] = fma(-ub_s, S.d[i], S.xB[i]);
                if (threadIdx.x == 0)
                {
                    sc->flips += 1;
                    sc->last_ret = -2;
                    sc->action = ACT_NONE;
                    if (sc->iterations >= sc->iter_limit) sc->status = TERM_ITER_LIMIT;
                }
                return;
            }
        }
        if (best.idx < 0)
        {
            if (threadIdx.x == 0)  // Simplex.cpp:334-338
            {
                sc->status = TERM_UNBOUNDED;
                sc->last_ret = -1;
                sc->action = ACT_NONE;
            }
            return;
        }

        const int p = best.idx;
        double dp = best.aux;
        if (best.cls == 1)
        {
            // Simplex.cpp:165-176: the leaving basic variable sits at its upper bound: substitute
            // it first. Column k of A_B changes sign <=> row p of B^-1 changes sign; x_p -> u - x_p.
            const int k = S.basic[p];
            const double u = S.ub[k];
            upper_bound_substitution_block(S, k, u);
            __syncthreads();
            if (threadIdx.x == 0)
            {
                S.xB[p] = u - S.xB[p];
                S.d[p] = -dp;
            }
            dp = -dp;
            __syncthreads();
        }
        const double xp = S.xB[p];
        const double theta = xp / dp;  // IEEE division, same expression as the winning ratio
        __syncthreads();
        // scaled pivot row (new row p of B^-1). (-rho)/(-d) == rho/d exactly, so the sign change
        // of a substituted row needs no extra pass.
        const double * rowp = S.Binv + static_cast<size_t>(p) * S.ld;
        const double dp_for_row = (best.cls == 1) ? -dp : dp;
        const double sq = sc->s_q;
        for (int j = threadIdx.x; j < S.ld; j += blockDim.x)
        {
            const double r = rowp[j] / dp_for_row;
            S.prow[j] = r;
            if (S.dual_update) S.y[j] = fma(sq, r, S.y[j]);
        }
        for (int i = threadIdx.x; i < S.m; i += blockDim.x)
        {
            if (i != p) S.xB[i] = fma(-theta, S.d[i], S.xB[i]);
        }
        if (threadIdx.x == 0)
        {
            S.xB[p] = theta;
            const int leaving_col = S.basic[p];  // Basis.cpp:48-53
            S.basic[p] = q;
            S.nonbasic[e] = leaving_col;
            if (S.trace_cap > 0)
            {
                S.trace[sc->pivots % S.trace_cap] = make_int4(leaving_col, p, q, e);
            }
            sc->pivots += 1;
            sc->p_row = p;
            sc->theta = theta;
            sc->d_p = dp;
            sc->last_ret = p;
            sc->action = ACT_PIVOT;
            if (sc->iterations >= sc->iter_limit) sc->status = TERM_ITER_LIMIT;
        }
    }

    // ---------------------------------------------------------------- K5: product-form rank-1 update
    // Tile = rows [r0, r1) x columns [c0, c0 + 2*blockDim.x). row_i -= d_i * prow ; row_p = prow.
    // prow for this thread's two columns lives in registers; rows stream through 128-bit loads.
    template <int kUnroll>
    __device__ __forceinline__ void update_tile(const DevState & S, int r0, int r1, int c0, int p)
    {
        const int col = c0 + 2 * threadIdx.x;
        if (col >= S.ld) return;
        const double2 pr = *reinterpret_cast<const double2 *>(S.prow + col);
        int i = r0;
        for (; i + kUnroll <= r1; i += kUnroll)
        {
            double2 v[kUnroll];
            double di[kUnroll];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u)
            {
                v[u] = ld_keep2(S.Binv + static_cast<size_t>(i + u) * S.ld + col);
                di[u] = S.d[i + u];
            }
#pragma unroll
            for (int u = 0; u < kUnroll; ++u)
            {
                double2 w;
                if (i + u == p)
                {
                    w = pr;
                }
                else
                {
                    w.x = fma(-di[u], pr.x, v[u].x);
                    w.y = fma(-di[u], pr.y, v[u].y);
                }
                st2(S.Binv + static_cast<size_t>(i + u) * S.ld + col, w);
            }
        }
        for (; i < r1; ++i)
        {
            double2 v = ld_keep2(S.Binv + static_cast<size_t>(i) * S.ld + col);
            const double di = S.d[i];
            double2 w;
            if (i == p)
            {
                w = pr;
            }
            else
            {
                w.x = fma(-di, pr.x, v.x);
                w.y = fma(-di, pr.y, v.y);
            }
            st2(S.Binv + static_cast<size_t>(i) * S.ld + col, w);
        }
    }

    // ---------------------------------------------------------------- K1: duals
    // Partial column sums over rows [r0, r1) for this thread's two columns; fixed row order.
    __device__ __forceinline__ void dual_partial_tile(const DevState & S, int seg, int r0, int r1, int c0)
    {
        const int col = c0 + 2 * threadIdx.x;
        if (col >= S.ld) return;
        double2 acc = make_double2(0.0, 0.0);
        int i = r0;
        for (; i + 4 <= r1; i += 4)
        {
            double2 v[4];
            double cb[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
            {
                v[u] = ld_keep2(S.Binv + static_cast<size_t>(i + u) * S.ld + col);
                cb[u] = S.c[S.basic[i + u]];
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
            {
                acc.x = fma(cb[u], v[u].x, acc.x);
                acc.y = fma(cb[u], v[u].y, acc.y);
            }
        }
        for (; i < r1; ++i)
        {
            const double2 v = ld_keep2(S.Binv + static_cast<size_t>(i) * S.ld + col);
            const double cb = S.c[S.basic[i]];
            acc.x = fma(cb, v.x, acc.x);
            acc.y = fma(cb, v.y, acc.y);
        }
        *reinterpret_cast<double2 *>(S.ypart + static_cast<size_t>(seg) * S.ld + col) = acc;
    }

    // y[col] = sum over segments in ascending order (deterministic).
    __device__ __forceinline__ void dual_reduce_col(const DevState & S, int col)
    {
        double acc = 0.0;
        for (int s = 0; s < S.nseg; ++s) acc += S.ypart[static_cast<size_t>(s) * S.ld + col];
        S.y[col] = acc;
    }
}  // namespace sb200

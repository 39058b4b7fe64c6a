// Persistent engine: ONE cooperative kernel runs the whole revised-simplex loop
// (reference src/Simplex.cpp:273-345), one CTA per SM, grid barriers between the phases of an
// iteration. Only 24-byte (value,index) candidates cross CTAs; vectors never leave the device and
// the host only polls a status word between chunks of iterations.
//
// Ownership: CTA c owns basis rows [c*m/G, (c+1)*m/G) of B^-1 and of x_B for the whole run, so
// FTRAN, the ratio test, the rank-1 update and the fused dual partial of a row all happen on
// the SM that touched the row last; the non-basic positions are dealt round-robin to warps.
//
// Per iteration (RECOMPUTE duals: 4 barriers, UPDATE duals: 3):
//   price   s_j = c_j - a_j.y over this CTA's positions, block arg-min -> slot[c]      | barrier 1
//   ftran   every CTA reduces the slots (same answer everywhere), stages a_q in smem,
//           d_i = row_i . a_q for own rows, ratio candidates -> slot[c]                | barrier 2
//   update  every CTA reduces the slots, stages the OLD pivot row, scales it in smem,
//           row_i -= d_i * prow (128-bit loads/stores), x_i -= theta d_i, fused
//           partial duals sum_i cB_i row_i -> ypart[c]; CTA 0 does the bookkeeping      | barrier 3
//           owner of row p writes the new pivot row (deferred so nobody reads a torn row)
//   dual    CTA c sums its slice of columns over all partials -> y                     | barrier 4
#pragma once
#include "sb200_kernels.cuh"

namespace sb200
{
    constexpr int PERSIST_THREADS = 1024;

    __host__ __device__ inline int persist_max_own_rows(int m, int grid) { return (m + grid - 1) / grid + 1; }

    inline size_t persistent_smem_bytes(const DevState & S, int grid = 0)
    {
        // ysm[ld] + vsm[ld] + dsm[own rows] (+ slack); grid is only known at launch: assume >= 1 SM
        const int own = persist_max_own_rows(S.m, grid > 0 ? grid : 1);
        const int own_cap = own > 4096 ? 4096 : own;
        return (2 * static_cast<size_t>(S.ld) + static_cast<size_t>(own_cap) + 8) * sizeof(double);
    }

    __device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long * p)
    {
        unsigned long long v;
        asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
        return v;
    }

    // Monotonic-counter grid barrier; the host zeroes sc->barrier_count before each launch.
    __device__ __forceinline__ void grid_barrier(Scalars * sc, unsigned int & epoch)
    {
        __syncthreads();
        if (threadIdx.x == 0)
        {
            epoch += 1;
            const unsigned long long target = static_cast<unsigned long long>(epoch) * gridDim.x;
            __threadfence();
            atomicAdd(&sc->barrier_count, 1ULL);
            while (ld_acquire_u64(&sc->barrier_count) < target)
            {
            }
            __threadfence();
        }
        __syncthreads();
    }

    // y partial of this CTA's rows (used at kernel entry; inside the loop it is fused in the update)
    __device__ __forceinline__ void persist_dual_partial(const DevState & S, int r0, int r1)
    {
        double * out = S.ypart + static_cast<size_t>(blockIdx.x) * S.ld;
        for (int col = 2 * threadIdx.x; col < S.ld; col += 2 * blockDim.x)
        {
            double2 acc = make_double2(0.0, 0.0);
            for (int i = r0; i < r1; ++i)
            {
                const double2 v = ld_keep2(S.Binv + static_cast<size_t>(i) * S.ld + col);
                const double cb = S.c[S.basic[i]];
                acc.x = fma(cb, v.x, acc.x);
                acc.y = fma(cb, v.y, acc.y);
            }
            *reinterpret_cast<double2 *>(out + col) = acc;
        }
    }

    // y[col] for this CTA's slice of columns = sum of the per-CTA partials. One warp per column:
    // lane l adds partials l, l+32, ... in ascending order, then a butterfly; the order is a
    // function of the grid size only, so a run is reproducible bit for bit.
    __device__ __forceinline__ void persist_dual_reduce(const DevState & S)
    {
        const int G = gridDim.x;
        const int cpc = (S.ld + G - 1) / G;
        const int c0 = blockIdx.x * cpc;
        const int lane = threadIdx.x & 31;
        const int wpb = blockDim.x >> 5;
        for (int t = threadIdx.x >> 5; t < cpc; t += wpb)
        {
            const int col = c0 + t;
            if (col < S.ld)
            {
                double acc = 0.0;
                for (int s = lane; s < G; s += 32) acc += S.ypart[static_cast<size_t>(s) * S.ld + col];
                acc = warp_sum(acc);
                if (lane == 0) S.y[col] = acc;
            }
        }
    }

    // Per-phase cycle counters of CTA 0 (sb200_get_phase_cycles). They live in shared memory during a
    // launch and are added to Scalars::phase_cycles once, at the end: a read-modify-write of global memory
    // per tick is an L2 round trip (~0.3 us) on the critical path of the CTA every grid barrier waits for.
    __device__ __forceinline__ long long * phase_acc()
    {
        __shared__ long long acc[10];  // [0..7] phase cycles, [8] pivots so far, [9] next slot of the trace ring
        return acc;
    }
    // all threads, before the first barrier of the kernel (and before CTA 0 writes sc->pivots)
    __device__ __forceinline__ void phase_init(const Scalars * sc = nullptr, int trace_cap = 0)
    {
        if (threadIdx.x < 8) phase_acc()[threadIdx.x] = 0;
        if (threadIdx.x == 8 && sc != nullptr)
        {
            const long long pivots = sc->pivots;
            phase_acc()[8] = pivots;
            phase_acc()[9] = (trace_cap > 0) ? pivots % trace_cap : 0;
        }
    }
    // thread 0 of CTA 0: the pivot log entry and the pivot counter, stored to global memory, never read back
    __device__ __forceinline__ void record_pivot(const DevState & S, Scalars * sc, int4 entry)
    {
        long long * acc = phase_acc();
        if (S.trace_cap > 0)
        {
            const long long at = acc[9];
            S.trace[at] = entry;
            acc[9] = (at + 1 == S.trace_cap) ? 0 : at + 1;
        }
        const long long pivots = acc[8] + 1;
        acc[8] = pivots;
        sc->pivots = pivots;
    }
    __device__ __forceinline__ void phase_tick(Scalars * sc, int slot, long long & t_last, bool every_cta = false)
    {
        (void)sc;
        if ((blockIdx.x == 0 || every_cta) && threadIdx.x == 0)
        {
            const long long now = clock64();
            phase_acc()[slot] += now - t_last;
            t_last = now;
        }
    }
    __device__ __forceinline__ void phase_count_iteration()
    {
        if (blockIdx.x == 0 && threadIdx.x == 0) phase_acc()[7] += 1;
    }
    // thread 0 of CTA 0, after the loop
    __device__ __forceinline__ void phase_flush(Scalars * sc)
    {
        if (blockIdx.x == 0 && threadIdx.x == 0)
            for (int k = 0; k < 8; ++k) sc->phase_cycles[k] += phase_acc()[k];
    }

    __global__ void __launch_bounds__(PERSIST_THREADS, 1) k_persistent(DevState S, long long chunk, int recompute)
    {
        extern __shared__ __align__(16) double smem[];
        __shared__ Cand scratch[33];
        double * ysm = smem;
        double * vsm = smem + S.ld;
        double * dsm = smem + 2 * S.ld;

        Scalars * sc = S.sc;
        const int G = gridDim.x;
        const int cta = blockIdx.x;
        const int lane = threadIdx.x & 31;
        const int warp = threadIdx.x >> 5;
        const int wpb = blockDim.x >> 5;
        const int r0 = static_cast<int>((static_cast<long long>(cta) * S.m) / G);
        const int r1 = static_cast<int>((static_cast<long long>(cta + 1) * S.m) / G);
        unsigned int epoch = 0;

        if (sc->status != TERM_RUNNING) return;  // uniform: nobody has written yet
        phase_init();

        // ---- entry: y = c_B^T B^-1 from scratch (also the periodic refresh of UPDATE mode)
        persist_dual_partial(S, r0, r1);
        grid_barrier(sc, epoch);
        persist_dual_reduce(S);
        grid_barrier(sc, epoch);
        stage_vector(ysm, S.y, S.ld);
        __syncthreads();

        long long t_last = clock64();
        for (long long it = 0; it < chunk; ++it)
        {
            // ================= price
            {
                Cand best = price_warp(S, ysm, cta * wpb + warp, G * wpb, lane);
                if (lane != 0) best = cand_none();
                best = block_reduce_cand(best, scratch);
                if (threadIdx.x == 0) S.price_slots[cta] = best;
            }
            phase_tick(sc, 0, t_last);
            grid_barrier(sc, epoch);  // 1
            phase_tick(sc, 1, t_last);

            // ================= entering decision + ftran + ratio candidates
            const Cand ebest = block_reduce_slots(S.price_slots, G, scratch);
            if (cta == 0 && threadIdx.x == 0) commit_entering(S, ebest);
            if (ebest.idx < 0) break;  // optimal (uniform across the grid)
            const int e = ebest.idx;
            const int q = S.nonbasic[e];
            const double s_q = ebest.aux;
            stage_vector(vsm, S.A + static_cast<size_t>(q) * S.ld, S.ld);
            __syncthreads();
            {
                Cand best = cand_none();
                for (int i = r0 + warp; i < r1; i += wpb)
                {
                    const double di = warp_dot<false>(S.Binv + static_cast<size_t>(i) * S.ld, vsm, S.ld, lane);
                    if (lane == 0)
                    {
                        S.d[i] = di;
                        if (i - r0 < 4096) dsm[i - r0] = di;
                        const double xi = S.xB[i];
                        Cand c = cand_none();
                        if (di > S.eps)
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
                        else if (S.ub != nullptr && di < -S.eps)
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
                if (lane != 0) best = cand_none();
                best = block_reduce_cand(best, scratch);
                if (threadIdx.x == 0) S.ratio_slots[cta] = best;
            }
            phase_tick(sc, 2, t_last);
            grid_barrier(sc, epoch);  // 2
            phase_tick(sc, 3, t_last);

            // ================= leaving decision (every CTA computes the same answer)
            const Cand rbest = block_reduce_slots(S.ratio_slots, G, scratch);
            const double min_theta = (rbest.idx >= 0) ? rbest.key : DBL_MAX;
            bool flip_nb = false;
            double ub_s = 0.0;
            if (S.ub != nullptr)
            {
                ub_s = S.ub[q];
                flip_nb = ub_s < min_theta;  // Simplex.cpp:160
            }
            int deferred_p = -1;
            if (flip_nb)
            {
                // entering variable goes to its own bound: no basis change (Simplex.cpp:160-164)
                for (int i = r0 + threadIdx.x; i < r1; i += blockDim.x) S.xB[i] = fma(-ub_s, S.d[i], S.xB[i]);
                if (cta == 0)
                {
                    upper_bound_substitution_block(S, q, ub_s);
                    if (threadIdx.x == 0)
                    {
                        sc->flips += 1;
                        sc->last_ret = -2;
                        if (sc->iterations >= sc->iter_limit) sc->status = TERM_ITER_LIMIT;
                    }
                }
                if (recompute)
                {
                    // duals are unchanged by a non-basic flip; republish the same partial sums
                    // protocol so that the barrier count per iteration stays uniform
                }
            }
            else if (rbest.idx < 0)
            {
                if (cta == 0 && threadIdx.x == 0)  // Simplex.cpp:334-338
                {
                    sc->status = TERM_UNBOUNDED;
                    sc->last_ret = -1;
                }
                break;  // uniform
            }
            else
            {
                const int p = rbest.idx;
                const double dp_row = rbest.aux;                           // d_p as computed (sign of the stored row)
                const double dp = (rbest.cls == 1) ? -dp_row : dp_row;     // pivot element after a basic flip
                const double theta = rbest.key;                            // == x_p/d_p bit for bit
                int k_flip = -1;
                double u_flip = 0.0;
                if (rbest.cls == 1)
                {
                    k_flip = S.basic[p];
                    u_flip = S.ub[k_flip];
                }
                // stage the OLD pivot row and scale it: prow = rho_p / d_p
                const double * rowp = S.Binv + static_cast<size_t>(p) * S.ld;
                for (int j = 2 * threadIdx.x; j < S.ld; j += 2 * blockDim.x)
                {
                    double2 r = ld_keep2(rowp + j);
                    r.x = r.x / dp_row;
                    r.y = r.y / dp_row;
                    *reinterpret_cast<double2 *>(vsm + j) = r;
                    if (!recompute)
                    {
                        double2 yv = *reinterpret_cast<double2 *>(ysm + j);
                        yv.x = fma(s_q, r.x, yv.x);
                        yv.y = fma(s_q, r.y, yv.y);
                        *reinterpret_cast<double2 *>(ysm + j) = yv;
                    }
                }
                __syncthreads();

                // ---- rank-1 update of own rows, x_B, fused dual partial
                const int tcols = min(static_cast<int>(blockDim.x), S.ld / 2);  // threads across a row
                const int ngroups = blockDim.x / tcols;
                const int grp = threadIdx.x / tcols;
                const int tc = threadIdx.x - grp * tcols;
                const double c_q = S.c[q];
                // Uniform trip count for every thread of the CTA: the group-combine below
                // contains __syncthreads(), which must never sit in divergent code.
                for (int cbase = 0; cbase < S.ld; cbase += 2 * tcols)
                {
                    const int col = cbase + 2 * tc;
                    const bool active = (grp < ngroups) && (col < S.ld);
                    double2 acc = make_double2(0.0, 0.0);
                    if (active)
                    {
                        const double2 pr = *reinterpret_cast<const double2 *>(vsm + col);
                        int i = r0 + grp;
                        for (; i + 3 * ngroups < r1; i += 4 * ngroups)
                        {
                            double2 v[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u)
                                v[u] = ld_keep2(S.Binv + static_cast<size_t>(i + u * ngroups) * S.ld + col);
#pragma unroll
                            for (int u = 0; u < 4; ++u)
                            {
                                const int row = i + u * ngroups;
                                const double di = (row - r0 < 4096) ? dsm[row - r0] : S.d[row];
                                double2 w;
                                double cb;
                                if (row == p)
                                {
                                    w = pr;
                                    cb = c_q;
                                }
                                else
                                {
                                    w.x = fma(-di, pr.x, v[u].x);
                                    w.y = fma(-di, pr.y, v[u].y);
                                    cb = recompute ? S.c[S.basic[row]] : 0.0;
                                    st2(S.Binv + static_cast<size_t>(row) * S.ld + col, w);
                                }
                                acc.x = fma(cb, w.x, acc.x);
                                acc.y = fma(cb, w.y, acc.y);
                            }
                        }
                        for (; i < r1; i += ngroups)
                        {
                            const double2 v = ld_keep2(S.Binv + static_cast<size_t>(i) * S.ld + col);
                            const double di = (i - r0 < 4096) ? dsm[i - r0] : S.d[i];
                            double2 w;
                            double cb;
                            if (i == p)
                            {
                                w = pr;
                                cb = c_q;
                            }
                            else
                            {
                                w.x = fma(-di, pr.x, v.x);
                                w.y = fma(-di, pr.y, v.y);
                                cb = recompute ? S.c[S.basic[i]] : 0.0;
                                st2(S.Binv + static_cast<size_t>(i) * S.ld + col, w);
                            }
                            acc.x = fma(cb, w.x, acc.x);
                            acc.y = fma(cb, w.y, acc.y);
                        }
                    }
                    if (recompute)
                    {
                        // combine the row groups in group order (deterministic); ysm is free to
                        // be the accumulator because y is rebuilt from the partials afterwards
                        for (int g = 0; g < ngroups; ++g)
                        {
                            if (ngroups > 1) __syncthreads();
                            if (active && g == grp)
                            {
                                double2 cur =
                                    (g == 0) ? make_double2(0.0, 0.0) : *reinterpret_cast<double2 *>(ysm + col);
                                cur.x += acc.x;
                                cur.y += acc.y;
                                *reinterpret_cast<double2 *>(ysm + col) = cur;
                            }
                        }
                    }
                }
                if (p >= r0 && p < r1)
                {
                    deferred_p = p;
                    if (threadIdx.x == 0 && rbest.cls == 1) S.d[p] = dp;  // keep the read-out consistent
                }
                // x_B of own rows (Simplex.cpp:306 recomputes it; here it is carried)
                for (int i = r0 + threadIdx.x; i < r1; i += blockDim.x)
                {
                    if (i == p)
                        S.xB[i] = theta;
                    else
                        S.xB[i] = fma(-theta, S.d[i], S.xB[i]);
                }
                if (recompute)
                {
                    __syncthreads();
                    double * out = S.ypart + static_cast<size_t>(cta) * S.ld;
                    for (int j = 2 * threadIdx.x; j < S.ld; j += 2 * blockDim.x)
                        *reinterpret_cast<double2 *>(out + j) = *reinterpret_cast<double2 *>(ysm + j);
                }
                if (cta == 0)
                {
                    if (k_flip >= 0) upper_bound_substitution_block(S, k_flip, u_flip);  // Simplex.cpp:165-176
                    if (threadIdx.x == 0)
                    {
                        const int leaving_col = S.basic[p];  // Basis.cpp:48-53
                        S.basic[p] = q;
                        S.nonbasic[e] = leaving_col;
                        if (S.trace_cap > 0) S.trace[sc->pivots % S.trace_cap] = make_int4(leaving_col, p, q, e);
                        sc->pivots += 1;
                        sc->p_row = p;
                        sc->theta = theta;
                        sc->d_p = dp;
                        sc->last_ret = p;
                        if (sc->iterations >= sc->iter_limit) sc->status = TERM_ITER_LIMIT;
                    }
                }
            }
            phase_tick(sc, 4, t_last);
            grid_barrier(sc, epoch);  // 3
            phase_tick(sc, 5, t_last);

            if (deferred_p >= 0)
            {
                double * rowp = S.Binv + static_cast<size_t>(deferred_p) * S.ld;
                for (int j = 2 * threadIdx.x; j < S.ld; j += 2 * blockDim.x)
                    st2(rowp + j, *reinterpret_cast<const double2 *>(vsm + j));
            }
            if (recompute && !flip_nb)
            {
                persist_dual_reduce(S);
                grid_barrier(sc, epoch);  // 4
                stage_vector(ysm, S.y, S.ld);
            }
            __syncthreads();
            phase_tick(sc, 6, t_last);
            phase_count_iteration();
            if (sc->status != TERM_RUNNING) break;  // written before barrier 3: uniform
        }
        phase_flush(sc);

        // publish the duals this CTA has been carrying (UPDATE mode keeps them in smem only)
        if (cta == 0 && !recompute)
        {
            __syncthreads();
            for (int j = threadIdx.x; j < S.ld; j += blockDim.x) S.y[j] = ysm[j];
        }
    }
}  // namespace sb200

// Fused persistent engine (the headline path for LPs without finite upper bounds).
//
// Observation: FTRAN of iteration k+1 needs B^-1 AFTER the rank-1 update of iteration k, and the
// update needs nothing but (d_k, scaled pivot row). So the two are done in ONE pass over B^-1:
//     row_i  <-  row_i - d_k[i] * prow_k          (product-form update k, written back)
//     d_{k+1}[i] = row_i . a_q'                    (FTRAN k+1 on the freshly updated registers)
// The pricing that chooses q' sits between them and needs the duals of the updated basis; those
// follow from the recurrence y' = y + (s_q/d_p) * rho_p (rho_p = old pivot row), which touches
// only the pivot row. y is recomputed from scratch (K1: partial column sums + ordered reduction)
// at every kernel launch, i.e. every `chunk` iterations.
//
// DRAM traffic per pivot: A_N once (pricing) + B^-1 read once + written once = 8(m N_nb + 2 m^2),
// against the algorithmic 8(m N_nb + 4 m^2) of SURVEY.md §8d. Two grid barriers per iteration:
//   price -> [barrier 1] -> entering decision, fused update+FTRAN, ratio candidates
//         -> [barrier 2] -> leaving decision, x_B, stage + scale the pivot row, dual recurrence
// The swap of the two index lists (Basis::update, reference src/Basis.cpp:48-53) is applied to
// the global lists by CTA 0 one barrier later; until then every CTA patches the single stale
// entry locally, so no third barrier is needed.
//
// Reference statements replaced: src/Simplex.cpp:305-307 (LU + two solves), :313, :318 (pricing +
// entering), :326-333 (FTRAN + leaving), src/Basis.cpp:48-57.
#pragma once
#include <algorithm>
#include "sb200_persistent.cuh"

namespace sb200
{
    // 512 threads = 16 warps per SM with 128 registers each: no spills (1024 threads cap a thread at 64 registers:
    // 88 bytes of spills in the row and column loops). Measured on B200 (profiles/README.md, round 2), headline LP:
    // 1024 threads 130.5, 768: 139.1, 640: 130.4, 512: 126.2, 448: 128.2, 384: 135.8 us per pivot; 512 is also the
    // fastest at m = 1536 ... 3072.
#ifndef SB200_FUSED_THREADS
#define SB200_FUSED_THREADS 512
#endif
    // the sharded kernel likewise: 8 GPUs, m=8192, n=262144: 405.3 us per pivot with 1024 threads, 368.4 with 512
    // (the end of the pricing sweep is much less ragged: 20.4 -> 7.6 us of waiting for the rank's slowest CTA)
#ifndef SB200_SHARDED_THREADS
#define SB200_SHARDED_THREADS 512
#endif
    constexpr int FUSED_THREADS = SB200_FUSED_THREADS;
    constexpr int SHARDED_THREADS = SB200_SHARDED_THREADS;

    inline size_t fused_smem_bytes(const DevState & S, int grid)
    {
        const int own = std::max(persist_max_own_rows(S.m, grid > 0 ? grid : 1), S.tuned_own_rows);
        // ysm, aq, psm (3 vectors of ld) + d of own rows + mbarrier slot
        return (3 * static_cast<size_t>(S.ld) + static_cast<size_t>(own) + 16) * sizeof(double);
    }

    // ---- TMA-style bulk staging of a contiguous vector into shared memory (cp.async.bulk, one
    // elected thread issues, everybody waits on the mbarrier phase).
    __device__ __forceinline__ void mbar_init(unsigned long long * bar, int count)
    {
        const unsigned a = static_cast<unsigned>(__cvta_generic_to_shared(bar));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __device__ __forceinline__ void bulk_load_issue(double * dst_smem, const double * src, unsigned bytes,
                                                    unsigned long long * bar)
    {
        const unsigned d = static_cast<unsigned>(__cvta_generic_to_shared(dst_smem));
        const unsigned b = static_cast<unsigned>(__cvta_generic_to_shared(bar));
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d),
                     "l"(src), "r"(bytes), "r"(b)
                     : "memory");
    }
    // Orders this thread's generic-proxy accesses to shared memory before later async-proxy
    // (bulk copy) writes to the same buffer.
    __device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

    __device__ __forceinline__ void mbar_wait(unsigned long long * bar, unsigned parity)
    {
        const unsigned b = static_cast<unsigned>(__cvta_generic_to_shared(bar));
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "WAIT_LOOP:\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
            "@p bra DONE;\n"
            "bra WAIT_LOOP;\n"
            "DONE:\n"
            "}\n" ::"r"(b),
            "r"(parity)
            : "memory");
    }

    // One row of the fused pass, one warp. Applies the pending update (if any), writes the row
    // back, and returns row . aq with exactly the summation order of warp_dot().
    template <bool kPending>
    __device__ __forceinline__ double fused_row(double * __restrict__ row, const double * __restrict__ psm,
                                                const double * __restrict__ aq, double dprev, bool is_p, int n, int lane,
                                                unsigned long long pol)
    {
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        int k = lane * 2;
#pragma unroll 2
        for (; k + 192 < n; k += 256)
        {
            double2 v0 = ld_hint2(row + k, pol);
            double2 v1 = ld_hint2(row + k + 64, pol);
            double2 v2 = ld_hint2(row + k + 128, pol);
            double2 v3 = ld_hint2(row + k + 192, pol);
            if (kPending)
            {
                const double2 p0 = *reinterpret_cast<const double2 *>(psm + k);
                const double2 p1 = *reinterpret_cast<const double2 *>(psm + k + 64);
                const double2 p2 = *reinterpret_cast<const double2 *>(psm + k + 128);
                const double2 p3 = *reinterpret_cast<const double2 *>(psm + k + 192);
                if (is_p)
                {
                    v0 = p0;
                    v1 = p1;
                    v2 = p2;
                    v3 = p3;
                }
                else
                {
                    v0.x = fma(-dprev, p0.x, v0.x);
                    v0.y = fma(-dprev, p0.y, v0.y);
                    v1.x = fma(-dprev, p1.x, v1.x);
                    v1.y = fma(-dprev, p1.y, v1.y);
                    v2.x = fma(-dprev, p2.x, v2.x);
                    v2.y = fma(-dprev, p2.y, v2.y);
                    v3.x = fma(-dprev, p3.x, v3.x);
                    v3.y = fma(-dprev, p3.y, v3.y);
                }
                st_hint2(row + k, v0, pol);
                st_hint2(row + k + 64, v1, pol);
                st_hint2(row + k + 128, v2, pol);
                st_hint2(row + k + 192, v3, pol);
            }
            const double2 s0 = *reinterpret_cast<const double2 *>(aq + k);
            const double2 s1 = *reinterpret_cast<const double2 *>(aq + k + 64);
            const double2 s2 = *reinterpret_cast<const double2 *>(aq + k + 128);
            const double2 s3 = *reinterpret_cast<const double2 *>(aq + k + 192);
            a0 = fma(v0.x, s0.x, a0);
            a0 = fma(v0.y, s0.y, a0);
            a1 = fma(v1.x, s1.x, a1);
            a1 = fma(v1.y, s1.y, a1);
            a2 = fma(v2.x, s2.x, a2);
            a2 = fma(v2.y, s2.y, a2);
            a3 = fma(v3.x, s3.x, a3);
            a3 = fma(v3.y, s3.y, a3);
        }
        for (; k < n; k += 64)
        {
            double2 v0 = ld_hint2(row + k, pol);
            if (kPending)
            {
                const double2 p0 = *reinterpret_cast<const double2 *>(psm + k);
                if (is_p)
                {
                    v0 = p0;
                }
                else
                {
                    v0.x = fma(-dprev, p0.x, v0.x);
                    v0.y = fma(-dprev, p0.y, v0.y);
                }
                st_hint2(row + k, v0, pol);
            }
            const double2 s0 = *reinterpret_cast<const double2 *>(aq + k);
            a0 = fma(v0.x, s0.x, a0);
            a0 = fma(v0.y, s0.y, a0);
        }
        return warp_sum((a0 + a1) + (a2 + a3));
    }

    // A SHARE of fused_row(): the lane-local accumulators [first, first + kCount) of the four warp_dot() keeps
    // (accumulator a owns the elements k + 64 a of every 256-element trip; the trailing elements of a row
    // whose length is not a multiple of 256 all belong to accumulator 0). When a CTA owns fewer rows than it
    // has warps, 2 or 4 warps share a row this way: each streams (and updates) its own quarter(s) of the row,
    // keeps the summation order of its accumulators, and the lane-local partials are added in warp_dot()'s
    // order, (a0 + a1) + (a2 + a3), before the butterfly -- so d_i has the bits one warp alone would compute.
    // Returns the lane-local partial: kCount 4: (a0+a1)+(a2+a3); 2: a_first + a_first+1; 1: a_first.
    template <bool kPending, int kCount>
    __device__ __forceinline__ double fused_row_share(double * __restrict__ row, const double * __restrict__ psm,
                                                      const double * __restrict__ aq, double dprev, bool is_p, int n,
                                                      int lane, int first, unsigned long long pol)
    {
        double acc[kCount];
#pragma unroll
        for (int a = 0; a < kCount; ++a) acc[a] = 0.0;
        const int off = 64 * first;
        int k = lane * 2;
#pragma unroll 2
        for (; k + 192 < n; k += 256)
        {
            double2 v[kCount];
#pragma unroll
            for (int a = 0; a < kCount; ++a) v[a] = ld_hint2(row + k + off + 64 * a, pol);
            if (kPending)
            {
#pragma unroll
                for (int a = 0; a < kCount; ++a)
                {
                    const double2 pv = *reinterpret_cast<const double2 *>(psm + k + off + 64 * a);
                    if (is_p)
                        v[a] = pv;
                    else
                    {
                        v[a].x = fma(-dprev, pv.x, v[a].x);
                        v[a].y = fma(-dprev, pv.y, v[a].y);
                    }
                    st_hint2(row + k + off + 64 * a, v[a], pol);
                }
            }
#pragma unroll
            for (int a = 0; a < kCount; ++a)
            {
                const double2 sv = *reinterpret_cast<const double2 *>(aq + k + off + 64 * a);
                acc[a] = fma(v[a].x, sv.x, acc[a]);
                acc[a] = fma(v[a].y, sv.y, acc[a]);
            }
        }
        if (first == 0)
        {
            for (; k < n; k += 64)
            {
                double2 v0 = ld_hint2(row + k, pol);
                if (kPending)
                {
                    const double2 p0 = *reinterpret_cast<const double2 *>(psm + k);
                    if (is_p)
                        v0 = p0;
                    else
                    {
                        v0.x = fma(-dprev, p0.x, v0.x);
                        v0.y = fma(-dprev, p0.y, v0.y);
                    }
                    st_hint2(row + k, v0, pol);
                }
                const double2 s0 = *reinterpret_cast<const double2 *>(aq + k);
                acc[0] = fma(v0.x, s0.x, acc[0]);
                acc[0] = fma(v0.y, s0.y, acc[0]);
            }
        }
        if (kCount == 4) return (acc[0] + acc[1]) + (acc[2] + acc[3]);
        if (kCount == 2) return acc[0] + acc[1];
        return acc[0];
    }

    // Warps per row of the fused pass for a CTA that owns `rows` rows and has `wpb` warps.
    __device__ __forceinline__ int warps_per_row(int rows, int wpb) { return (4 * rows <= wpb) ? 4 : (2 * rows <= wpb) ? 2 : 1; }

    // Update-only pass for one row (flush of the pending update at kernel exit).
    __device__ __forceinline__ void flush_row(double * __restrict__ row, const double * __restrict__ psm, double dprev,
                                              bool is_p, int n, int lane)
    {
        for (int k = lane * 2; k < n; k += 64)
        {
            double2 v = ld_keep2(row + k);
            const double2 p = *reinterpret_cast<const double2 *>(psm + k);
            if (is_p)
            {
                v = p;
            }
            else
            {
                v.x = fma(-dprev, p.x, v.x);
                v.y = fma(-dprev, p.y, v.y);
            }
            st2(row + k, v);
        }
    }

    // Pricing with the single possibly-stale list entry patched locally.
    __device__ __forceinline__ Cand price_warp_patched(const DevState & S, const double * ysm, int gw, int stride, int lane,
                                                       int e_prev, int lc_prev)
    {
        Cand best = cand_none();
        const double neg_eps = -S.eps;
        for (int pos = gw; pos < S.nnb; pos += stride)
        {
            const int col = (pos == e_prev) ? lc_prev : S.nonbasic[pos];
            const double dot = warp_dot<true>(S.A + static_cast<size_t>(col) * S.ld, ysm, S.ld, lane);
            const double s = S.c[col] - dot;
            if (s < neg_eps)
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

    // Same arithmetic, different work distribution: CTA c owns the contiguous block of positions
    // [c*nnb/G, (c+1)*nnb/G) (equal bytes per SM) and its warps draw positions from a shared-memory
    // ticket counter, so no SM idles while another still has a column to read. The reduced cost of
    // a position does not depend on which warp evaluates it and the arg-min is a total order, so
    // the result is independent of the (timing dependent) draw order.
    // kBounded (LPs with finite upper bounds): a column that went through upper_bound_substitution
    // (reference src/LinearProgram.cpp:154-169) is NOT rewritten during a launch; S.neg[col] says that the
    // column and its cost count with the opposite sign, and (npcol, npneg) patches the one entry whose flag
    // CTA 0 has not published yet. (-c) - (-a).y == -(c - a.y) bit for bit, so the reduced cost is the one the
    // reference computes from its rewritten column.
    template <bool kBounded>
    __device__ __forceinline__ Cand price_cta_dynamic(const DevState & S, const double * ysm, int pos0, int pos1, int warp,
                                                      int wpb, int lane, int e_prev, int lc_prev, int * ticket,
                                                      int npcol = -1, int npneg = 0)
    {
        Cand best = cand_none();
        const double neg_eps = -S.eps;
        const bool units = S.unit_row != nullptr;
        int pos = pos0 + warp;
        int col = 0, urow = -1, cneg = 0;
        if (pos < pos1)
        {
            col = (pos == e_prev) ? lc_prev : S.nonbasic[pos];
            if (units) urow = S.unit_row[col];
            if (kBounded) cneg = (col == npcol) ? npneg : S.neg[col];
        }
        while (pos < pos1)
        {
            // draw the next position and look its column up now: both latencies hide behind this column
            int next = 0;
            if (lane == 0) next = atomicAdd(ticket, 1);
            next = __shfl_sync(0xffffffffu, next, 0);
            int ncol = 0, nurow = -1, nneg = 0;
            if (next < pos1)
            {
                ncol = (next == e_prev) ? lc_prev : S.nonbasic[next];
                if (units) nurow = S.unit_row[ncol];
                if (kBounded) nneg = (ncol == npcol) ? npneg : S.neg[ncol];
            }
            // A unit column value * e_row has the dot product value * y_row: the same bits warp_dot would
            // deliver (one product plus exact zeros) without streaming the column.
            const double dot = (urow >= 0) ? S.unit_val[col] * ysm[urow]
                                           : warp_dot<true>(S.A + static_cast<size_t>(col) * S.ld, ysm, S.ld, lane);
            double s = S.c[col] - dot;
            if (kBounded && cneg) s = -s;
            if (s < neg_eps)
            {
                Cand c;
                c.key = (S.rule == RULE_MOST_NEG) ? s : 0.0;
                c.aux = s;
                c.idx = pos;
                c.cls = 0;
                if (cand_better(c, best)) best = c;
            }
            pos = next;
            col = ncol;
            urow = nurow;
            cneg = nneg;
        }
        (void)wpb;
        return best;
    }

    // ---- K1 for the fused engines: y = c_B^T B^-1 with a summation order that depends on
    // NOTHING but m: rows are cut into blocks of DUAL_BLOCK_ROWS, a block is summed in row order,
    // the block partials are added in ascending block order. One GPU or eight, any grid size: the
    // same bits (this is what makes the sharded run walk the same pivot sequence as one GPU).
    constexpr int DUAL_BLOCK_ROWS = 32;

    // Partials of the local blocks [blk0, blk0+nblk) whose rows start at local row 0 of Binv_loc.
    // out[(blk - blk0)][ld]. global_row0 = first global row of Binv_loc (for basic[]).
    __device__ __forceinline__ void dual_block_partials(const DevState & S, const double * Binv_loc, int global_row0,
                                                        int nrows_loc, double * out)
    {
        const int nblk = (nrows_loc + DUAL_BLOCK_ROWS - 1) / DUAL_BLOCK_ROWS;
        for (int blk = blockIdx.x; blk < nblk; blk += gridDim.x)
        {
            const int i0 = blk * DUAL_BLOCK_ROWS;
            const int i1 = min(i0 + DUAL_BLOCK_ROWS, nrows_loc);
            for (int col = 2 * threadIdx.x; col < S.ld; col += 2 * blockDim.x)
            {
                double2 acc = make_double2(0.0, 0.0);
                for (int i = i0; i < i1; ++i)
                {
                    const double2 v = ld_keep2(Binv_loc + static_cast<size_t>(i) * S.ld + col);
                    const double cb = S.c[S.basic[global_row0 + i]];
                    acc.x = fma(cb, v.x, acc.x);
                    acc.y = fma(cb, v.y, acc.y);
                }
                *reinterpret_cast<double2 *>(out + static_cast<size_t>(blk) * S.ld + col) = acc;
            }
        }
    }

    // dst[col] = carry[col] (or 0) + sum of nblk partials in ascending order; columns split over the grid.
    __device__ __forceinline__ void dual_block_chain(const DevState & S, const double * parts, int nblk,
                                                     const double * carry_or_null, double * dst)
    {
        const int G = gridDim.x;
        const int cpc = (S.ld + G - 1) / G;
        const int c0 = blockIdx.x * cpc;
        for (int t = threadIdx.x; t < cpc; t += blockDim.x)
        {
            const int col = c0 + t;
            if (col < S.ld)
            {
                double acc = carry_or_null ? carry_or_null[col] : 0.0;
                for (int s = 0; s < nblk; ++s) acc += parts[static_cast<size_t>(s) * S.ld + col];
                dst[col] = acc;
            }
        }
    }

    // Grid-wide variant: the first wave is static (warp gw takes position gw), every later position
    // comes from ONE global ticket counter, so SMs that see more bandwidth simply take more columns.
    // Every processed column draws exactly one ticket, hence an iteration consumes exactly nnb
    // tickets and iteration `it` of this launch owns the ticket range [it*nnb, (it+1)*nnb).
    __device__ __forceinline__ Cand price_grid_dynamic(const DevState & S, const double * ysm, int gw, int nwarps,
                                                       int lane, int e_prev, int lc_prev, unsigned long long * ticket,
                                                       unsigned long long base)
    {
        Cand best = cand_none();
        const double neg_eps = -S.eps;
        int pos = gw;
        while (pos < S.nnb)
        {
            const int col = (pos == e_prev) ? lc_prev : S.nonbasic[pos];
            unsigned long long next = 0;
            if (lane == 0) next = atomicAdd(ticket, 1ULL);
            const double dot = warp_dot<true>(S.A + static_cast<size_t>(col) * S.ld, ysm, S.ld, lane);
            const double s = S.c[col] - dot;
            if (s < neg_eps)
            {
                Cand c;
                c.key = (S.rule == RULE_MOST_NEG) ? s : 0.0;
                c.aux = s;
                c.idx = pos;
                c.cls = 0;
                if (cand_better(c, best)) best = c;
            }
            next = __shfl_sync(0xffffffffu, next, 0);
            pos = nwarps + static_cast<int>(next - base);
        }
        return best;
    }

    // Shared tail of the pricing sweep (price_mode 3): positions [first, nnb) are handed out by ONE global
    // ticket counter. Every warp draws until its first ticket beyond the pool, so an iteration consumes
    // exactly pool + nwarps tickets and iteration `it` of a launch owns [it*(pool+nwarps), (it+1)*(pool+nwarps)).
    __device__ __forceinline__ Cand price_pool(const DevState & S, const double * ysm, int first, int nwarps, int lane,
                                               int e_prev, int lc_prev, unsigned long long * ticket, unsigned long long it)
    {
        Cand best = cand_none();
        const double neg_eps = -S.eps;
        const unsigned long long pool = static_cast<unsigned long long>(S.nnb - first);
        if (pool == 0) return best;
        const unsigned long long base = it * (pool + static_cast<unsigned long long>(nwarps));
        const bool units = S.unit_row != nullptr;
        unsigned long long t = 0;
        if (lane == 0) t = atomicAdd(ticket, 1ULL) - base;
        t = __shfl_sync(0xffffffffu, t, 0);
        while (t < pool)
        {
            unsigned long long tn = 0;
            if (lane == 0) tn = atomicAdd(ticket, 1ULL) - base;  // drawn early: hides behind the column
            const int pos = first + static_cast<int>(t);
            const int col = (pos == e_prev) ? lc_prev : S.nonbasic[pos];
            const int urow = units ? S.unit_row[col] : -1;
            const double dot = (urow >= 0) ? S.unit_val[col] * ysm[urow]
                                           : warp_dot<true>(S.A + static_cast<size_t>(col) * S.ld, ysm, S.ld, lane);
            const double s = S.c[col] - dot;
            if (s < neg_eps)
            {
                Cand c;
                c.key = (S.rule == RULE_MOST_NEG) ? s : 0.0;
                c.aux = s;
                c.idx = pos;
                c.cls = 0;
                if (cand_better(c, best)) best = c;
            }
            t = __shfl_sync(0xffffffffu, tn, 0);
        }
        return best;
    }

    // Exit epilogue of the bounded variant: the columns whose sign changed during this launch are rewritten
    // the way the reference rewrites them at substitution time (A[:,j] = -A[:,j], c_j = -c_j,
    // src/LinearProgram.cpp:156-166), so that between launches A, c and the unit-column table are what every
    // other engine, the re-inversion and the read-outs expect, and S.neg is all zero again.
    __device__ __forceinline__ void apply_pending_negations(const DevState & S, int npcol, int npneg, bool neg_dirty)
    {
        const int lane = threadIdx.x & 31;
        const int gw = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
        const int nw = gridDim.x * (blockDim.x >> 5);
        for (int base = gw * 32; base < S.N; base += nw * 32)
        {
            const int col = base + lane;
            int stored = 0, eff = 0;
            if (col < S.N)
            {
                stored = S.neg[col];
                eff = (neg_dirty && col == npcol) ? npneg : stored;
                if (stored) S.neg[col] = 0;
            }
            unsigned todo = __ballot_sync(0xffffffffu, eff != 0);
            while (todo)
            {
                const int l = __ffs(todo) - 1;
                todo &= todo - 1;
                double * acol = S.A + static_cast<size_t>(base + l) * S.ld;
                for (int k = 2 * lane; k < S.ld; k += 64)
                {
                    double2 v = ld_keep2(acol + k);
                    v.x = -v.x;
                    v.y = -v.y;
                    st2(acol + k, v);
                }
                if (lane == 0)
                {
                    S.c[base + l] = -S.c[base + l];
                    if (S.unit_val_rw != nullptr) S.unit_val_rw[base + l] = -S.unit_val_rw[base + l];
                }
            }
        }
    }

    // The loop. kBounded = the LP has finite upper bounds (reference: has_non_trivial_upper_bounds(),
    // src/Simplex.cpp:108): the ratio test gets the t2 candidates (d_i < -eps, :136-155), the leaving decision the
    // test of the entering variable's own bound (:157-164, a bound flip is an iteration without a pivot: B^-1 and y
    // stay, x_B -= u d, b -= a_q u) and the substitution of a leaving variable that sits at its bound (:165-176:
    // row p of B^-1 changes sign, which the scaled pivot row (-rho)/(-d) = rho/d never sees; x_p -> u - x_p).
    template <bool kBounded>
    __device__ __forceinline__ void fused_loop(const DevState & S, long long chunk)
    {
        extern __shared__ __align__(16) double smem[];
        __shared__ Cand scratch[33];
        __shared__ __align__(8) unsigned long long mbar;
        double * ysm = smem;
        double * aq = smem + S.ld;
        double * psm = smem + 2 * S.ld;
        double * dsm = smem + 3 * S.ld;

        Scalars * sc = S.sc;
        const int G = gridDim.x;
        const int cta = blockIdx.x;
        const int lane = threadIdx.x & 31;
        const int warp = threadIdx.x >> 5;
        const int wpb = blockDim.x >> 5;
        const int r0 = S.row_bounds ? S.row_bounds[cta] : static_cast<int>((static_cast<long long>(cta) * S.m) / G);
        const int r1 = S.row_bounds ? S.row_bounds[cta + 1] : static_cast<int>((static_cast<long long>(cta + 1) * S.m) / G);
        const unsigned vec_bytes = static_cast<unsigned>(S.ld * sizeof(double));
        unsigned int epoch = 0;
        unsigned int mphase = 0;

        if (sc->status != TERM_RUNNING) return;  // uniform: nobody has written yet
        __shared__ int ticket;
        // price_mode 3: the CTAs own contiguous slices of the first 7/8 of the positions (in-CTA tickets, as in
        // mode 1); the last 1/8 is a pool every warp of the grid drains through one global ticket counter
        // once its CTA's slice is done, which evens out the bandwidth differences between SMs
        const int price_mode = kBounded ? 1 : S.price_mode;
        const int n_sliced = (price_mode == 3) ? (S.nnb / 8) * 7 : S.nnb;
        const bool tuned_pos = S.pos_bounds != nullptr && price_mode == 1;
        const int pos0 = tuned_pos ? S.pos_bounds[cta] : static_cast<int>((static_cast<long long>(cta) * n_sliced) / G);
        const int pos1 = tuned_pos ? S.pos_bounds[cta + 1] : static_cast<int>((static_cast<long long>(cta + 1) * n_sliced) / G);
        if (threadIdx.x == 0)
        {
            mbar_init(&mbar, 1);
            ticket = pos0 + wpb;
        }
        const long long iters_at_entry = sc->iterations;  // read before anybody increments it
        const long long limit = sc->iter_limit;
        phase_init(sc, S.trace_cap);  // also: pivot counter and trace cursor of CTA 0, kept in shared memory
        const bool dbg = S.dbg_cycles != nullptr;  // every CTA keeps the phase counters (who waits for whom)

        // ---- entry: y = c_B^T B^-1 from scratch (K1; the periodic refresh of the recurrence)
        dual_block_partials(S, S.Binv, 0, S.m, S.ypart);
        grid_barrier(sc, epoch);
        dual_block_chain(S, S.ypart, (S.m + DUAL_BLOCK_ROWS - 1) / DUAL_BLOCK_ROWS, nullptr, S.y);
        grid_barrier(sc, epoch);
        if (threadIdx.x == 0) bulk_load_issue(ysm, S.y, vec_bytes, &mbar);
        mbar_wait(&mbar, mphase & 1);
        mphase++;
        __syncthreads();

        // pending pivot (update not yet applied to B^-1, swap not yet applied to the global lists)
        bool pending = false;
        int p_prev = -1, e_prev = -1, q_prev = -1, lc_prev = -1;
        bool lists_applied = true;
        int exit_status = TERM_RUNNING;
        long long local_iters = iters_at_entry;
        long long t_last = clock64();
        // L2 policies of the fused pass: the CTA's first keep_rows rows stay in L2 from pivot to pivot
        // (measured, profiles/README.md round 2: the gain comes with the policies on the STORES -- the kept rows stay
        // dirty in L2 instead of draining into the pricing sweep; on the loads alone they change nothing)
        const unsigned long long pol_keep = l2_policy_evict_last();
        const unsigned long long pol_stream = l2_policy_evict_first();
        const int keep_end = r0 + S.keep_rows;
        // kBounded: the column whose sign flag CTA 0 publishes one barrier late, and its new flag
        int npcol = -1, npneg = 0;
        bool neg_dirty = false;
        long long flips_local = 0;

        for (long long it = 0; it < chunk; ++it)
        {
            // ================= price
            {
                Cand best;
                if (price_mode == 2)
                    best = price_grid_dynamic(S, ysm, cta * wpb + warp, G * wpb, lane, lists_applied ? -1 : e_prev,
                                              lc_prev, &sc->ticket,
                                              static_cast<unsigned long long>(it) * static_cast<unsigned long long>(S.nnb));
                else if (price_mode == 1 || price_mode == 3)
                {
                    best = price_cta_dynamic<kBounded>(S, ysm, pos0, pos1, warp, wpb, lane, lists_applied ? -1 : e_prev,
                                                       lc_prev, &ticket, npcol, npneg);
                    if (price_mode == 3)
                    {
                        const Cand more = price_pool(S, ysm, n_sliced, G * wpb, lane, lists_applied ? -1 : e_prev, lc_prev,
                                                     &sc->ticket, static_cast<unsigned long long>(it));
                        if (cand_better(more, best)) best = more;
                    }
                }
                else
                    best = price_warp_patched(S, ysm, cta * wpb + warp, G * wpb, lane, lists_applied ? -1 : e_prev,
                                              lc_prev);
                if (lane != 0) best = cand_none();
                best = block_reduce_cand(best, scratch);
                if (threadIdx.x == 0) S.price_slots[cta] = best;
                // this CTA is done streaming A: pull its first rows of B^-1 towards L2 while it waits
                // (the first keep_rows rows are expected to be there already)
                if (lane == 0 && warp < S.prefetch_rows && keep_end + warp < r1)
                    prefetch_l2_bulk(S.Binv + static_cast<size_t>(keep_end + warp) * S.ld, vec_bytes);
            }
            phase_tick(sc, 0, t_last, dbg);
            grid_barrier(sc, epoch);  // 1
            phase_tick(sc, 1, t_last, dbg);
            if (threadIdx.x == 0) ticket = pos0 + wpb;  // re-arm for the next pricing (many syncs away)

            const Cand ebest = block_reduce_slots(S.price_slots, G, scratch);
            local_iters += 1;
            // entering column: read through the patch BEFORE CTA 0 brings the global list up to date
            int q = -1;
            if (ebest.idx >= 0) q = (!lists_applied && ebest.idx == e_prev) ? lc_prev : S.nonbasic[ebest.idx];
            // sign of the entering column (through the patch, for the same reason)
            int qneg = 0;
            if (kBounded && q >= 0) qneg = (q == npcol) ? npneg : S.neg[q];
            __syncthreads();
            if (cta == 0 && threadIdx.x == 0)
            {
                if (!lists_applied)
                {
                    S.basic[p_prev] = q_prev;  // Basis.cpp:48-53, one barrier late
                    S.nonbasic[e_prev] = lc_prev;
                }
                if (kBounded && neg_dirty) S.neg[npcol] = static_cast<unsigned char>(npneg);
                sc->iterations = local_iters;
                sc->action = ACT_NONE;
                if (ebest.idx >= 0)
                {
                    sc->e_pos = ebest.idx;
                    sc->q_col = q;
                    sc->s_q = ebest.aux;
                }
                else
                {
                    sc->e_pos = -1;
                    sc->last_ret = -1;
                }
            }
            lists_applied = true;
            neg_dirty = false;  // the patch (npcol, npneg) stays: it agrees with the published flag from now on
            if (ebest.idx < 0)
            {
                exit_status = (local_iters >= limit) ? TERM_ITER_LIMIT : TERM_OPTIMAL;  // Simplex.cpp:319-323, 347-350
                break;
            }
            const int e = ebest.idx;
            const double s_q = ebest.aux;

            // ================= fused: rank-1 update k-1 + FTRAN k + ratio candidates
            if (threadIdx.x == 0) bulk_load_issue(aq, S.A + static_cast<size_t>(q) * S.ld, vec_bytes, &mbar);
            mbar_wait(&mbar, mphase & 1);
            mphase++;
            {
                Cand best = cand_none();
                for (int i = r0 + warp; i < r1; i += wpb)
                {
                    double * row = S.Binv + static_cast<size_t>(i) * S.ld;
                    const double dprev = pending ? dsm[i - r0] : 0.0;
                    double di;
                    const unsigned long long pol = (i < keep_end) ? pol_keep : pol_stream;
                    if (pending)
                        di = fused_row<true>(row, psm, aq, dprev, i == p_prev, S.ld, lane, pol);
                    else
                        di = fused_row<false>(row, psm, aq, 0.0, false, S.ld, lane, pol);
                    if (kBounded && qneg) di = -di;  // B^-1 (-a_q) = -(B^-1 a_q) exactly
                    if (lane == 0)
                    {
                        S.d[i] = di;
                        const double xi = S.xB[i];
                        if (di > S.eps)  // Simplex.cpp:73-79 / 125-132
                        {
                            const double t = xi / di;
                            if (t < DBL_MAX)
                            {
                                Cand c;
                                c.key = t;
                                c.aux = di;
                                c.idx = i;
                                c.cls = 0;
                                if (cand_better(c, best)) best = c;
                            }
                        }
                        else if (kBounded && di < -S.eps)  // Simplex.cpp:136-153
                        {
                            // basic[] of the row pivoted last may be in flight (CTA 0 writes it behind barrier 1)
                            const int bi = (i == p_prev) ? q_prev : S.basic[i];
                            const double t = (S.ub[bi] - xi) / (-di);
                            if (t < DBL_MAX)
                            {
                                Cand c;
                                c.key = t;
                                c.aux = di;
                                c.idx = i;
                                c.cls = 1;
                                if (cand_better(c, best)) best = c;
                            }
                        }
                    }
                }
                if (lane != 0) best = cand_none();
                fence_async_smem();                       // aq / psm were read: the next bulk copies may overwrite them
                best = block_reduce_cand(best, scratch);  // (syncs: all d of own rows are written)
                if (threadIdx.x == 0) S.ratio_slots[cta] = best;
                // DRAM idles from here (barrier 2, leaving decision, pivot-row staging) to the next pricing
                // sweep: pull the first columns of this CTA's next wave (warp w starts with position
                // pos0 + w) towards L2. Position e is skipped: its new column is not known yet.
                if (lane == 0 && warp < S.prefetch_cols && pos0 + warp < pos1 && pos0 + warp != e)
                    prefetch_l2_bulk(S.A + static_cast<size_t>(S.nonbasic[pos0 + warp]) * S.ld, vec_bytes);
            }
            pending = false;
            phase_tick(sc, 2, t_last, dbg);
            grid_barrier(sc, epoch);  // 2
            phase_tick(sc, 3, t_last, dbg);

            // ================= leaving decision (same answer in every CTA)
            const Cand rbest = block_reduce_slots(S.ratio_slots, G, scratch);
            if (kBounded)
            {
                const double min_theta = (rbest.idx >= 0) ? rbest.key : DBL_MAX;
                const double ub_s = S.ub[q];
                if (ub_s < min_theta)
                {
                    // Simplex.cpp:160-164: the entering variable reaches its own bound first. No basis change:
                    // B^-1, y and the lists stay; x_B = B^-1 (b - a_q u) = x_B - u d; b -= a_q u; the column and
                    // its cost change sign (flag), the substitution flag toggles (LinearProgram.cpp:154-169).
                    for (int i = r0 + threadIdx.x; i < r1; i += blockDim.x)
                    {
                        const double a = qneg ? -aq[i] : aq[i];
                        S.b[i] = __dsub_rn(S.b[i], __dmul_rn(a, ub_s));  // two roundings, as g++ -O3 compiles it
                        S.xB[i] = fma(-ub_s, S.d[i], S.xB[i]);
                    }
                    npcol = q;
                    npneg = qneg ^ 1;
                    neg_dirty = true;
                    flips_local += 1;
                    if (cta == 0 && threadIdx.x == 0)
                    {
                        S.flip[q] ^= 1;
                        sc->last_ret = -2;
                    }
                    fence_async_smem();  // aq was read again: the next bulk copy may overwrite it
                    __syncthreads();
                    phase_tick(sc, 4, t_last, dbg);
                    phase_count_iteration();
                    if (local_iters >= limit)
                    {
                        exit_status = TERM_ITER_LIMIT;
                        break;
                    }
                    continue;
                }
            }
            if (rbest.idx < 0)
            {
                exit_status = TERM_UNBOUNDED;  // Simplex.cpp:334-338
                if (cta == 0 && threadIdx.x == 0) sc->last_ret = -1;
                break;
            }
            const int p = rbest.idx;
            const double dp = rbest.aux;     // d_p as computed from the stored row p
            const double theta = rbest.key;  // == x_p / d_p (t1) or (u - x_p) / (-d_p) (t2) bit for bit
            // stage the pivot row of the UPDATED matrix (TMA bulk copy), then scale it in place
            if (threadIdx.x == 0) bulk_load_issue(psm, S.Binv + static_cast<size_t>(p) * S.ld, vec_bytes, &mbar);
            const int lc = S.basic[p];  // not yet overwritten: CTA 0 applies swaps after the next barrier 1
            if (kBounded && rbest.cls == 1)
            {
                // Simplex.cpp:165-176: the leaving variable sits at its upper bound: substitute it first
                // (b -= a_k u, column and cost change sign); the normal pivot follows.
                const double u = S.ub[lc];
                const int lneg = (lc == npcol) ? npneg : S.neg[lc];
                const double * acol = S.A + static_cast<size_t>(lc) * S.ld;
                for (int i = r0 + threadIdx.x; i < r1; i += blockDim.x)
                {
                    const double a = lneg ? -acol[i] : acol[i];
                    S.b[i] = __dsub_rn(S.b[i], __dmul_rn(a, u));
                }
                npcol = lc;
                npneg = lneg ^ 1;
                neg_dirty = true;
                if (cta == 0 && threadIdx.x == 0) S.flip[lc] ^= 1;
            }
            // own rows: x_B and the d needed by the next fused pass
            for (int i = r0 + threadIdx.x; i < r1; i += blockDim.x)
            {
                const double di = S.d[i];
                dsm[i - r0] = di;
                S.xB[i] = (i == p) ? theta : fma(-theta, di, S.xB[i]);
                if (kBounded && i == p && rbest.cls == 1) S.d[i] = -di;  // read-out: the pivot element after the substitution
            }
            mbar_wait(&mbar, mphase & 1);
            mphase++;
            for (int j = 2 * threadIdx.x; j < S.ld; j += 2 * blockDim.x)
            {
                double2 r = *reinterpret_cast<double2 *>(psm + j);
                r.x = r.x / dp;  // (-rho_p)/(-d_p) == rho_p/d_p: a substituted row needs no extra pass
                r.y = r.y / dp;
                *reinterpret_cast<double2 *>(psm + j) = r;
                double2 yv = *reinterpret_cast<double2 *>(ysm + j);
                yv.x = fma(s_q, r.x, yv.x);  // y' = y + s_q * (rho_p / d_p)
                yv.y = fma(s_q, r.y, yv.y);
                *reinterpret_cast<double2 *>(ysm + j) = yv;
            }
            pending = true;
            lists_applied = false;
            p_prev = p;
            e_prev = e;
            q_prev = q;
            lc_prev = lc;
            if (cta == 0 && threadIdx.x == 0)
            {
                record_pivot(S, sc, make_int4(lc, p, q, e));
                sc->p_row = p;
                sc->theta = theta;
                sc->d_p = (kBounded && rbest.cls == 1) ? -dp : dp;
                sc->last_ret = p;
            }
            fence_async_smem();
            __syncthreads();
            phase_tick(sc, 4, t_last, dbg);
            phase_count_iteration();
            if (local_iters >= limit)
            {
                exit_status = TERM_ITER_LIMIT;
                break;
            }
        }

        // ---- exit: flush the pending update so that B^-1, the lists and y in HBM are consistent.
        // The barrier makes sure every CTA has finished staging the last pivot row from global
        // B^-1 before its owner overwrites it (inside the loop barrier 1 plays that role).
        grid_barrier(sc, epoch);
        if (pending)
        {
            for (int i = r0 + warp; i < r1; i += wpb)
                flush_row(S.Binv + static_cast<size_t>(i) * S.ld, psm, dsm[i - r0], i == p_prev, S.ld, lane);
        }
        if (kBounded) apply_pending_negations(S, npcol, npneg, neg_dirty);
        if (dbg && threadIdx.x < 8) S.dbg_cycles[cta * 8 + threadIdx.x] = phase_acc()[threadIdx.x];
        if (cta == 0)
        {
            __syncthreads();
            for (int j = threadIdx.x; j < S.ld; j += blockDim.x)
            {
                S.y[j] = ysm[j];
                if (pending) S.prow[j] = psm[j];
            }
            if (threadIdx.x == 0)
            {
                if (!lists_applied)
                {
                    S.basic[p_prev] = q_prev;
                    S.nonbasic[e_prev] = lc_prev;
                }
                if (kBounded) sc->flips += flips_local;
                if (exit_status != TERM_RUNNING) sc->status = exit_status;
                phase_flush(sc);
            }
        }
    }

    __global__ void __launch_bounds__(FUSED_THREADS, 1) k_persistent_fused(DevState S, long long chunk)
    {
        fused_loop<false>(S, chunk);
    }

    // the same loop for LPs with finite upper bounds (bounded ratio test, bound flips)
    __global__ void __launch_bounds__(FUSED_THREADS, 1) k_persistent_fused_bounded(DevState S, long long chunk)
    {
        fused_loop<true>(S, chunk);
    }
}  // namespace sb200

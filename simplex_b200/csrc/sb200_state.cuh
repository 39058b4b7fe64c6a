// Device-resident solver state and the small deterministic reductions shared by every kernel.
//
// Layout in HBM (all fp64 unless noted; see DESIGN.md "Data layout"):
//   A      [N][ld]   column-major tableau, ld = round_up(m,16); rows m..ld-1 are zero so that
//                    128-bit loops never need a tail (reference: Eigen::MatrixXd matrix_A,
//                    src/LinearProgram.h:89, column-major src/LinearProgram.cpp:101)
//   Binv   [m][ld]   dense basis inverse, ROW-major (pivot row contiguous; FTRAN = row dots)
//   b[m] c[N] ub[N]  right-hand side / costs / upper bounds (ub only if bounds were parsed)
//   xB[m] y[ld] d[ld] prow[ld]   basic solution, duals, FTRAN column, scaled pivot row
//   basic[m] nonbasic[N-m] (int32)  the two index lists of simplex::Basis (src/Basis.h:32-33)
//   flip[N] (uint8)  ub_substitutions (src/LinearProgram.h:98)
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace sb200
{
    constexpr int TERM_RUNNING = 0;
    constexpr int TERM_OPTIMAL = 1;
    constexpr int TERM_UNBOUNDED = 2;
    constexpr int TERM_ITER_LIMIT = 3;

    constexpr int ACT_NONE = 0;   // nothing for the rank-1 kernel to do (flip iteration / ended)
    constexpr int ACT_PIVOT = 1;  // apply the product-form update for (e_pos, p_row)

    constexpr int RULE_FIRST = 0;
    constexpr int RULE_MOST_NEG = 1;

    // An arg-min candidate. Total order (key, cls, idx); idx < 0 means "none".
    //   pricing: key = s_j (MOST_NEG) or 0 (FIRST), idx = non-basic POSITION, aux = s_j
    //   ratio  : key = x_i/d_i (cls 0) or (u-x_i)/(-d_i) (cls 1), idx = basis ROW, aux = d_i
    struct Cand
    {
        double key;
        double aux;
        int idx;
        int cls;
    };

    __host__ __device__ inline Cand cand_none()
    {
        Cand c;
        c.key = 0.0;
        c.aux = 0.0;
        c.idx = -1;
        c.cls = 0;
        return c;
    }

    // true iff a is strictly better than b. Mirrors the reference's strict '<' scans in index
    // order (src/Simplex.cpp:73-79, 95, 125-153): equal keys keep the earlier (lower) index, and
    // a t1 row (cls 0) beats a t2 row (cls 1) of equal ratio because the t1 pass runs first.
    __host__ __device__ inline bool cand_better(const Cand & a, const Cand & b)
    {
        if (a.idx < 0) return false;
        if (b.idx < 0) return true;
        if (a.key < b.key) return true;
        if (a.key > b.key) return false;
        if (a.cls != b.cls) return a.cls < b.cls;
        return a.idx < b.idx;
    }

    struct Scalars
    {
        int status;  // TERM_*
        int action;  // ACT_*
        int e_pos;   // entering position in the non-basic list
        int q_col;   // entering column id
        int p_row;   // leaving basis row
        int last_ret;  // what select_leaving_variable would have returned: row, -1 or -2
        double s_q;    // reduced cost of the entering column
        double theta;  // step length
        double d_p;    // pivot element
        long long iterations;
        long long pivots;
        long long flips;
        long long iter_limit;
        unsigned long long barrier_count;  // grid barrier (persistent engine), zeroed per launch
        unsigned long long ticket;         // pricing work queue (fused engine), zeroed per launch
        unsigned int barrier_gen;
        int pad;
        // SM-clock cycles CTA 0 spent per phase, summed over iterations (persistent engine):
        // 0 price, 1 barrier after price, 2 ftran, 3 barrier after ftran, 4 update,
        // 5 barrier after update, 6 dual reduce + barrier + restage, 7 iterations counted
        long long phase_cycles[8];
    };

    struct DevState
    {
        int m, N, nnb, ld;
        double * A;
        double * Binv;
        double * b;
        double * c;
        double * ub;  // nullptr: Bland ratio test
        double * xB;
        double * y;
        double * d;
        double * prow;
        int * basic;
        int * nonbasic;
        unsigned char * flip;
        double * ypart;  // [nseg][ld] partial duals
        int nseg;
        Cand * price_slots;
        int n_price_slots;
        Cand * ratio_slots;
        int n_ratio_slots;
        Scalars * sc;
        int4 * trace;
        int trace_cap;
        double eps;
        int rule;
        int dual_update;  // 1: y += s_q * prow in the select phase
        int price_mode;   // fused engine: 0 static strided, 1 per-CTA tickets, 2 grid-wide tickets
        // Unit columns (one nonzero: slacks, artificials), classified once per load when the LP has no
        // finite bounds: unit_row[j] = row of the nonzero (-1: dense column), unit_val[j] = its value.
        // The fused engine prices such a column as c_j - value * y_row instead of streaming m doubles
        // ("a legitimate traffic saving of up to 8 m^2 B").
        const int * unit_row;
        const double * unit_val;
        double * unit_val_rw;  // the same table, writable: the bounded fused engine negates the entry of a rewritten column
        // Fused engine with finite upper bounds: neg[j] != 0 <=> column j and its cost count with the opposite
        // sign until the exit of the running launch rewrites them (all zero between launches).
        unsigned char * neg;
        double * row_partials;  // sharded engine: lane partials of rows shared by several warps, [local row][3][32]
        int prefetch_cols;  // fused engine: columns of the next pricing wave pulled into L2 per CTA while DRAM idles
        int prefetch_rows;  // fused engine: own rows of B^-1 pulled into L2 per CTA before barrier 1
        int keep_rows;      // fused engine: the first keep_rows rows of every CTA are kept in L2 (evict-last policy)
        long long * dbg_cycles;  // fused engine: [G][8] per-CTA phase cycles of a launch (tuning, SB200_FUSED_DEBUG), else nullptr
        // fused engine: tuned partition of the work over the CTAs ([G+1] each, nullptr = equal shares): CTA c
        // prices positions [pos_bounds[c], pos_bounds[c+1]) and owns rows [row_bounds[c], row_bounds[c+1]) of
        // B^-1 for one launch. Results do not depend on the partition (every column / row is evaluated by one
        // warp in a fixed order, the arg-mins are total orders); only the time each SM needs does.
        const int * pos_bounds;
        const int * row_bounds;
        int tuned_own_rows;  // capacity of the per-CTA d buffer when row_bounds is set
    };

    // ------------------------------------------------------------------ warp / block reductions
    // Order-preserving map of a double onto an unsigned integer (no NaNs reach it).
    __device__ __forceinline__ unsigned long long sortable(double x)
    {
        const long long b = __double_as_longlong(x);
        return static_cast<unsigned long long>(b ^ ((b >> 63) | static_cast<long long>(0x8000000000000000ULL)));
    }

    // Warp arg-min on (key, idx), idx < 0 = no candidate, as three 32-bit REDUX steps (high word, low word,
    // index) instead of a five-step butterfly of 64-bit compares. Returns the winning lane, -1 if none.
    __device__ __forceinline__ int warp_argmin_lane(unsigned long long key, int idx)
    {
        const unsigned int full = 0xffffffffu;
        const bool valid = idx >= 0;
        const unsigned int hi = valid ? static_cast<unsigned int>(key >> 32) : 0xffffffffu;
        const unsigned int mh = __reduce_min_sync(full, hi);
        const bool a = valid && hi == mh;
        const unsigned int lo = a ? static_cast<unsigned int>(key) : 0xffffffffu;
        const unsigned int ml = __reduce_min_sync(full, lo);
        const bool b = a && lo == ml;
        const unsigned int ix = b ? static_cast<unsigned int>(idx) : 0xffffffffu;
        const unsigned int mi = __reduce_min_sync(full, ix);
        const unsigned int who = __ballot_sync(full, b && ix == mi);
        return who ? __ffs(who) - 1 : -1;
    }
    // (key, cls, idx) of a Cand as the two integers warp_argmin_lane() orders by: the key with -0.0 folded
    // onto +0.0 (the double comparison of cand_better() says they are equal), and cls above the index in
    // one 31-bit word (idx < 2^30: sb200_load's limit on N).
    __device__ __forceinline__ unsigned long long cand_key(const Cand & c) { return sortable(__dadd_rn(c.key, 0.0)); }
    __device__ __forceinline__ int cand_tie(const Cand & c) { return c.idx < 0 ? -1 : ((c.cls << 30) | c.idx); }

    // All threads of the block call this; result valid in thread 0 (in all of warp 0). scratch: >= 32 Cand
    // in smem. Two REDUX arg-mins (cand_better()'s order) instead of two five-step butterflies.
    __device__ __forceinline__ Cand block_reduce_cand(Cand c, Cand * scratch)
    {
        const int lane = threadIdx.x & 31;
        const int warp = threadIdx.x >> 5;
        const int nwarp = (blockDim.x + 31) >> 5;
        const int w = warp_argmin_lane(cand_key(c), cand_tie(c));
        __syncthreads();  // scratch may still be read from a previous use
        if (lane == (w < 0 ? 0 : w)) scratch[warp] = (w < 0) ? cand_none() : c;
        __syncthreads();
        Cand r = cand_none();
        if (warp == 0)
        {
            r = (lane < nwarp) ? scratch[lane] : cand_none();
            const int b = warp_argmin_lane(cand_key(r), cand_tie(r));
            r = (b < 0) ? cand_none() : scratch[b];
        }
        return r;
    }

    // Every thread of the block gets the best of slots[0..n) (deterministic: pure function of
    // the slot contents). scratch: >= 33 Cand in smem. Every warp repeats the reduction of the per-warp
    // winners, so nothing has to be broadcast.
    __device__ __forceinline__ Cand block_reduce_slots(const Cand * slots, int n, Cand * scratch)
    {
        const int lane = threadIdx.x & 31;
        const int warp = threadIdx.x >> 5;
        const int nwarp = (blockDim.x + 31) >> 5;
        Cand c = cand_none();
        for (int i = threadIdx.x; i < n; i += blockDim.x)
        {
            const Cand o = slots[i];
            if (cand_better(o, c)) c = o;
        }
        const int w = warp_argmin_lane(cand_key(c), cand_tie(c));
        __syncthreads();  // scratch may still be read from a previous use
        if (lane == (w < 0 ? 0 : w)) scratch[warp] = (w < 0) ? cand_none() : c;
        __syncthreads();
        const Cand r = (lane < nwarp) ? scratch[lane] : cand_none();
        const int b = warp_argmin_lane(cand_key(r), cand_tie(r));
        return (b < 0) ? cand_none() : scratch[b];
    }

    __device__ __forceinline__ double warp_sum(double v)
    {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        return v;
    }

    // ------------------------------------------------------------------ memory access helpers
    // Streaming 128-bit load (read once per pivot: the pricing sweep over A).
    __device__ __forceinline__ double2 ld_stream2(const double * p)
    {
        double2 r;
        asm volatile("ld.global.cs.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
        return r;
    }
    // Plain coherent 128-bit load (B^-1: re-read by the next phase, keep it in L2).
    __device__ __forceinline__ double2 ld_keep2(const double * p)
    {
        double2 r;
        asm volatile("ld.global.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
        return r;
    }
    // L2 prefetch of a contiguous run (hint only; bytes % 16 == 0): fills DRAM idle time behind the
    // grid barriers with bytes the next phase is going to read anyway.
    __device__ __forceinline__ void prefetch_l2_bulk(const void * p, unsigned bytes)
    {
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
    }
    __device__ __forceinline__ void st2(double * p, double2 v)
    {
        asm volatile("st.global.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
    }
    // 128-bit accesses with an L2 eviction policy (createpolicy): the rows of B^-1 a CTA wants to find in L2
    // again at the next pivot are read and written "evict last", everything that only streams through "evict
    // first", so that the 537 MB pricing sweep and the other rows of B^-1 replace each other instead of them.
    __device__ __forceinline__ unsigned long long l2_policy_evict_last()
    {
        unsigned long long p;
        asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
        return p;
    }
    __device__ __forceinline__ unsigned long long l2_policy_evict_first()
    {
        unsigned long long p;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
        return p;
    }
    __device__ __forceinline__ double2 ld_hint2(const double * p, unsigned long long policy)
    {
        double2 r;
        asm volatile("ld.global.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(r.x), "=d"(r.y) : "l"(p), "l"(policy));
        return r;
    }
    __device__ __forceinline__ void st_hint2(double * p, double2 v, unsigned long long policy)
    {
        asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(p), "d"(v.x), "d"(v.y), "l"(policy) : "memory");
    }

    // Warp-cooperative dot product of a global vector g[0..n) (16-byte aligned, n % 16 == 0,
    // zero padded) with a shared-memory vector s[0..n). The summation order is fixed by this
    // function alone (4 lane-local accumulators, then a butterfly), so a given column/row gives
    // the same bits whichever kernel, grid shape or GPU evaluates it.
#ifndef SB200_DOT_UNROLL
#define SB200_DOT_UNROLL 2
#endif
    constexpr int DOT_UNROLL = SB200_DOT_UNROLL;

    template <bool kStream>
    __device__ __forceinline__ double warp_dot(const double * __restrict__ g, const double * __restrict__ s, int n,
                                               int lane)
    {
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        int k = lane * 2;
        // 4 independent 128-bit loads in flight per lane and trip (x2 from unrolling)
#pragma unroll DOT_UNROLL
        for (; k + 192 < n; k += 256)
        {
            double2 v0, v1, v2, v3;
            if (kStream)
            {
                v0 = ld_stream2(g + k);
                v1 = ld_stream2(g + k + 64);
                v2 = ld_stream2(g + k + 128);
                v3 = ld_stream2(g + k + 192);
            }
            else
            {
                v0 = ld_keep2(g + k);
                v1 = ld_keep2(g + k + 64);
                v2 = ld_keep2(g + k + 128);
                v3 = ld_keep2(g + k + 192);
            }
            const double2 s0 = *reinterpret_cast<const double2 *>(s + k);
            const double2 s1 = *reinterpret_cast<const double2 *>(s + k + 64);
            const double2 s2 = *reinterpret_cast<const double2 *>(s + k + 128);
            const double2 s3 = *reinterpret_cast<const double2 *>(s + k + 192);
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
            const double2 v0 = kStream ? ld_stream2(g + k) : ld_keep2(g + k);
            const double2 s0 = *reinterpret_cast<const double2 *>(s + k);
            a0 = fma(v0.x, s0.x, a0);
            a0 = fma(v0.y, s0.y, a0);
        }
        return warp_sum((a0 + a1) + (a2 + a3));
    }
}  // namespace sb200

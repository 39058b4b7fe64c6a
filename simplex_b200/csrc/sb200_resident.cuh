// Shared-memory-resident engine: the whole tableau lives ON CHIP, spread over the SMs.
//
// For LPs whose dense columns and basis inverse fit the 148 x 227 KB of shared memory of a B200
// (BASELINE configs[1], m=1024, n=2048: 16.8 MB of structural columns + 8.4 MB of B^-1 = 170 KB per
// SM) the HBM-bound fused engine is the wrong tool: its working set sits in L2 and a pivot costs a
// chain of L2 round trips and two grid barriers. Here CTA c keeps for the whole
// launch
//   * its rows [c*m/G, (c+1)*m/G) of B^-1, x_B, the basic list and what it needs to know about the
//     basic column of each row (dense rank, unit row / value, cost),
//   * the dense columns of A whose dense rank k satisfies k % G == c (cyclic, by column id, whether
//     basic or not), with their cost and their current non-basic POSITION (-1 while basic),
//   * the slice [c*nnb/G, (c+1)*nnb/G) of the non-basic positions for the unit columns (slacks,
//     artificials), which are priced as c_j - v * y_r and never stored,
//   * y, the entering column and the scaled pivot row (replicated).
// Per pivot three things cross CTAs (two candidates and one row), through L2, NCCL-LL style: every 8-byte word carries 32 bits of
// payload and the 32-bit sequence number of the pivot, is written with one relaxed store and polled
// with relaxed loads until the number matches. 8-byte accesses are single-copy atomic, so a word is
// self-validating: no fence, no release/acquire pair and no barrier sits on the critical path.
// One pivot, in order (every CTA does the same, nobody waits for a barrier across CTAs):
//   1. price the CTA's non-basic dense columns (shared-memory dots with y) and its unit slice; block arg-min
//   2. exchange 1: push the candidate (4 words: reduced cost, position, column id) into slot [own CTA] of a
//      private mailbox of EVERY CTA. While that mail flies, apply the rank-1 update of the PREVIOUS pivot
//      to the CTA's rows (elementwise, all warps; it needs only that pivot's d and scaled row, not the
//      column this exchange is about to name). Then thread t < G validates slot t of the own mailbox
//      (polled early, in registers), five warps REDUX-reduce, one barrier, every warp reduces the five
//      winners: everybody knows the entering column q
//   3. read a_q from the (never modified) global A (8 m bytes from L2), FTRAN on the rows in shared
//      memory (one warp per row), ratio candidates, block arg-min
//   4. exchange 2: push (ratio, row) first, THEN stage the candidate row -- already divided by its pivot
//      element, 2 words per double, followed by what the others must learn about the leaving column --
//      so that the divisions and the 16 KB of stores overlap the flight of the mail; gather as in 2
//   5. poll the winner's staging line: that IS the scaled pivot row; y by the dual recurrence, x_B, caches
// Measured on B200 (profiles/README.md, round 1d): 7.8-8.1 us per pivot at m=1024, n=2048 against 13.6 us
// for the fused engine; a first version with one release/acquire slot per CTA and a fence after the row
// stores took 25.9 us. With 16 warps per SM and a strictly serial pivot every phase is bound by the length
// of its dependent-instruction path (fp64 latency), not by shared-memory or L2 bandwidth.
//
// The arithmetic is the fused engine's, operation by operation (warp_dot summation order, the
// fma forms of the update, the dual recurrence, the entry dual solve), so both engines produce the
// same bits and walk the same pivot sequence.
//
// LPs with finite upper bounds (k_resident_bounded; reference src/Simplex.cpp:116-179, LinearProgram.cpp:154-169)
// follow the bounded fused kernel: a column that goes through upper_bound_substitution is not rewritten, a sign
// flag says that it and its cost count negated (-(c - a.y) is exact). Every CTA sees every decision of every
// pivot, so each keeps its OWN copy of the N flags in shared memory and toggles it from the decisions: no flag
// ever crosses CTAs. The ratio candidates gain the t2 class (the tag bit of the row index orders a t1 row before
// a t2 row of equal ratio, as the reference's two passes do), the entering variable's own bound is checked by
// everybody on the gathered minimum (a flip is an iteration without a pivot), and a leaving variable at its upper
// bound is substituted by its row's owner. CTA 0 leaves the flags in S.neg; k_apply_negations (launched behind
// the kernel: nobody reads A any more) rewrites A, c and the unit table the way the reference does at once.
//
// kStream (k_resident_stream*): LPs whose rows of B^-1 still fit the shared memory but whose dense columns do not
// (1100 < m <= 1776 or so, A small enough to live in L2). The columns stay in global memory and the pricing dots
// read them from L2 (same lane partition, same summation order: same bits); everything else is unchanged.
//
// Reference statements replaced: src/Simplex.cpp:305-307, :313, :318, :326-333, src/Basis.cpp:48-57.
#pragma once
#include "sb200_fused.cuh"

namespace sb200
{
#ifndef RES_SHARED_MAIL
#define RES_SHARED_MAIL 0  // 1: one mailbox of G slots that every CTA polls; 0: a private mailbox per CTA, candidates pushed to all
#endif
#ifndef RES_EARLY_POLL
#define RES_EARLY_POLL 1  // 1: the first polls of an exchange are issued half way through the work that overlaps it
#endif
    constexpr int RES_THREADS = 512;
#ifndef RES_ROWS_PER_WARP
#define RES_ROWS_PER_WARP 1  // rows (FTRAN) a warp takes per trip; 2 shares the loads of a_q but doubles the warp's chain: slower (1.26 vs 1.07 us)
#endif
#ifndef RES_COLS_PER_WARP
#define RES_COLS_PER_WARP 2  // columns (pricing) a warp takes per trip, sharing the loads of y
#endif
    constexpr int RES_MAX_GRID = 192;    // most CTAs (= SMs) a mailbox can have: one polling thread per slot
    constexpr int RES_MAIL_STRIDE = 4;   // 8-byte words per CTA and exchange (4 used by price, 3 by ratio)
    constexpr int RES_ROW_EXTRAS = 6;    // doubles appended to a staged row (5 used)
    constexpr int TERM_RESIDENT_TIMEOUT = -4;
    constexpr long long RES_TIMEOUT_CYCLES = 4000000000LL;  // ~2 s at 1.9 GHz: a lost peer, not a slow one
    static_assert(RES_THREADS >= RES_MAX_GRID && RES_MAX_GRID % 32 == 0, "one thread per slot of a mailbox");
    static_assert(RES_THREADS % 4 == 0, "the push loops keep thread & 3 as the word index in every trip");

    struct ResidentInfo
    {
        const int * dense_ids;   // [n_dense] dense rank -> column id (ascending)
        const int * dense_rank;  // [N] column id -> dense rank, -1 for unit columns
        int n_dense;
        int max_rows;   // rows of B^-1 per CTA (capacity)
        int max_cols;   // dense columns per CTA (capacity)
        int max_slice;  // non-basic positions per CTA (capacity)
        unsigned long long * pmail;    // [G destinations][G sources][RES_MAIL_STRIDE]
        unsigned long long * rmail;    // [G destinations][G sources][RES_MAIL_STRIDE]
        unsigned long long * rowmail;  // [G][2 * (ld + RES_ROW_EXTRAS)]
        unsigned int seq0;             // sequence number of this launch's first pivot (never 0, never reused)
        int ticks;                     // 1: CTA 0 keeps the per-phase cycle counters (sb200_get_phase_cycles)
        long long * per_cta_cycles;    // [G][16] or nullptr: EVERY CTA keeps a finer set (SB200_RES_DEBUG: who waits
                                       // for whom, and where inside a phase); see RES_FINE_NAMES in sb200_abi.cu
    };

    __host__ __device__ inline int res_rows_cap(int m, int G) { return (m + G - 1) / G; }
    __host__ __device__ inline int res_cols_cap(int n_dense, int G) { return (n_dense + G - 1) / G; }
    __host__ __device__ inline int res_slice_cap(int nnb, int G) { return (nnb + G - 1) / G + 1; }
    __host__ __device__ inline size_t res_mail_words(int ld, int G)
    {
        return 2 * static_cast<size_t>(G) * G * RES_MAIL_STRIDE + static_cast<size_t>(G) * 2 * (ld + RES_ROW_EXTRAS);
    }

    // bounded_N > 0: the bounded variant, which also keeps the upper bound of every row's basic variable and one
    // sign flag per column (N bits)
    // stream_cols: the dense columns are not kept on chip (kStream)
    inline size_t resident_smem_bytes(int ld, int R, int C, int SL, int bounded_N = 0, bool stream_cols = false)
    {
        // doubles: y, a_q, pivot row | rows | columns | d, x_B, unit value and cost of the basic column per row |
        //          cost per column | unit value and cost per slice position
        size_t doubles = static_cast<size_t>(3 + R + (stream_cols ? 0 : C)) * ld + 4 * static_cast<size_t>(R) + C +
                         2 * static_cast<size_t>(SL);
        // ints: position, id, active list and its inverse per column | basic id, home, unit row per row |
        //       id and unit row per position
        size_t ints = 4 * static_cast<size_t>(C) + 3 * static_cast<size_t>(R) + 2 * static_cast<size_t>(SL);
        if (bounded_N > 0)
        {
            doubles += static_cast<size_t>(R);
            ints += (static_cast<size_t>(bounded_N) + 31) / 32;
        }
        return doubles * sizeof(double) + ((ints * sizeof(int) + 15) / 16) * 16;
    }

    // ---- candidates. Same total order as cand_better (sb200_state.cuh) with cls == 0 everywhere
    // (this engine never sees finite bounds): key, then index; idx < 0 = none.
    struct PCand  // pricing: v = reduced cost s_j, idx = non-basic position, col = column id
    {
        double v;
        int idx;
        int col;
    };
    struct RCand  // ratio test: t = x_i / d_i, idx = basis row, src = CTA that owns (and staged) the row
    {
        double t;
        int idx;
        int src;
    };
    // bounded variant: a t2 candidate ((u - x_i) / (-d_i), Simplex.cpp:136-153) carries this bit in idx, so that
    // at equal ratio every comparison below puts a t1 row first, then the lower row (cand_better's order)
    constexpr int RES_T2_TAG = 1 << 30;
    __device__ __forceinline__ int res_row_of(int tagged) { return tagged & (RES_T2_TAG - 1); }
    __device__ __forceinline__ bool res_flag(const unsigned int * negw, int col) { return (negw[col >> 5] >> (col & 31)) & 1u; }
    __device__ __forceinline__ PCand pnone() { return PCand{0.0, -1, -1}; }
    __device__ __forceinline__ RCand rnone() { return RCand{0.0, -1, -1}; }
    // most_neg: key = s_j; first eligible: key = 0 for everybody, so the position decides (Simplex.cpp:49-61)
    __device__ __forceinline__ bool pbetter(const PCand & a, const PCand & b, bool most_neg)
    {
        if (a.idx < 0) return false;
        if (b.idx < 0) return true;
        if (most_neg)
        {
            if (a.v < b.v) return true;
            if (a.v > b.v) return false;
        }
        return a.idx < b.idx;
    }
    __device__ __forceinline__ bool rbetter(const RCand & a, const RCand & b)
    {
        if (a.idx < 0) return false;
        if (b.idx < 0) return true;
        if (a.t < b.t) return true;
        if (a.t > b.t) return false;
        return a.idx < b.idx;
    }
    __device__ __forceinline__ unsigned long long pkey(const PCand & c, bool most_neg) { return most_neg ? sortable(c.v) : 0ULL; }
    // -0.0 and +0.0 are the same ratio (the double comparison of the other engines says so too)
    __device__ __forceinline__ unsigned long long rkey(const RCand & c) { return sortable(__dadd_rn(c.t, 0.0)); }

    // Block arg-min with ONE barrier: per-warp winners to shared memory, then every warp reduces them
    // again on its own. All threads call, all threads get the best. Two calls must be separated by
    // another __syncthreads() (they are: every exchange has one).
    // lanes_hold (warp-uniform): the lanes of this warp hold different candidates and are reduced first;
    // otherwise only lane 0's candidate counts (a warp that priced whole columns) and that step is skipped.
    __device__ __forceinline__ PCand block_allmin_p(const PCand & c, PCand * scratch, bool most_neg, bool lanes_hold)
    {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
        if (lanes_hold)
        {
            const int w = warp_argmin_lane(pkey(c, most_neg), c.idx);
            if (lane == (w < 0 ? 0 : w)) scratch[warp] = (w < 0) ? pnone() : c;
        }
        else if (lane == 0)
            scratch[warp] = c;
        __syncthreads();
        const PCand r = lane < nwarp ? scratch[lane] : pnone();
        const int b = warp_argmin_lane(pkey(r, most_neg), r.idx);
        return b < 0 ? pnone() : scratch[b];
    }
    // every warp's candidate is the one of its lane 0 (the lane that ran the ratio test of the warp's rows)
    __device__ __forceinline__ RCand block_allmin_r(const RCand & c, RCand * scratch)
    {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
        if (lane == 0) scratch[warp] = c;
        __syncthreads();
        const RCand r = lane < nwarp ? scratch[lane] : rnone();
        const int b = warp_argmin_lane(rkey(r), r.idx);
        return b < 0 ? rnone() : scratch[b];
    }

    // ---- self-validating words
    // (plain st.global instead of st.relaxed.gpu changes nothing measurable: profiles/README.md, round 1d)
    __device__ __forceinline__ void st_word(unsigned long long * p, unsigned long long v)
    {
        asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
    }
    __device__ __forceinline__ unsigned long long ld_word(const unsigned long long * p)
    {
        unsigned long long v;
        asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
        return v;
    }
    // two adjacent words in one 128-bit access; each half is validated on its own
    __device__ __forceinline__ void st_word2(unsigned long long * p, unsigned long long a, unsigned long long b)
    {
        asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
    }
    __device__ __forceinline__ void ld_word2(const unsigned long long * p, unsigned long long & a, unsigned long long & b)
    {
        asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
    }
    __device__ __forceinline__ unsigned long long word_of(unsigned int seq, unsigned int data)
    {
        return (static_cast<unsigned long long>(seq) << 32) | data;
    }
    __device__ __forceinline__ bool word_ok(unsigned long long w, unsigned int seq) { return static_cast<unsigned int>(w >> 32) == seq; }
    // a double as two words at p[0], p[1]
    __device__ __forceinline__ void st_double(unsigned long long * p, double v, unsigned int seq)
    {
        const unsigned long long b = static_cast<unsigned long long>(__double_as_longlong(v));
        st_word2(p, word_of(seq, static_cast<unsigned int>(b)), word_of(seq, static_cast<unsigned int>(b >> 32)));
    }
    __device__ __forceinline__ double double_of(unsigned long long lo, unsigned long long hi)
    {
        return __longlong_as_double(static_cast<long long>((hi << 32) | (lo & 0xffffffffULL)));
    }
    // Bounded wait for the four words of two adjacent doubles (both 128-bit loads in flight together).
    // false = timed out.
    __device__ __forceinline__ bool ld_double2(const unsigned long long * p, unsigned int seq, double2 & out)
    {
        unsigned long long a0, a1, b0, b1;
        ld_word2(p, a0, a1);
        ld_word2(p + 2, b0, b1);
        if (!(word_ok(a0, seq) && word_ok(a1, seq) && word_ok(b0, seq) && word_ok(b1, seq)))
        {
            const long long t0 = clock64();
            do
            {
                ld_word2(p, a0, a1);
                ld_word2(p + 2, b0, b1);
                if (clock64() - t0 > RES_TIMEOUT_CYCLES) return false;
            } while (!(word_ok(a0, seq) && word_ok(a1, seq) && word_ok(b0, seq) && word_ok(b1, seq)));
        }
        out.x = double_of(a0, a1);
        out.y = double_of(b0, b1);
        return true;
    }

    // One slot (the candidate of one CTA: 4 words, 32 bytes) per thread: thread t < G owns slot t of the
    // mailbox this CTA reads. poll_issue() fetches it EARLY -- the two 128-bit loads fly while the caller
    // does other work -- and poll_slot() validates what came back, re-polling words of an older pivot.
    struct Polled
    {
        unsigned long long a0, a1, b0, b1;
    };
    __device__ __forceinline__ Polled poll_issue(const unsigned long long * box, int G)
    {
        Polled p{0ULL, 0ULL, 0ULL, 0ULL};
        if (threadIdx.x < G)
        {
            ld_word2(box + threadIdx.x * RES_MAIL_STRIDE, p.a0, p.a1);
            ld_word2(box + threadIdx.x * RES_MAIL_STRIDE + 2, p.b0, p.b1);
        }
        return p;
    }
    // thread t < G: waits until words [0, W) of slot t carry this pivot's sequence number. false = timed out.
    template <int W>
    __device__ __forceinline__ bool poll_slot(const unsigned long long * box, unsigned int seq, Polled & p)
    {
        const unsigned long long * at = box + threadIdx.x * RES_MAIL_STRIDE;
        auto ok = [&]() {
            return word_ok(p.a0, seq) && word_ok(p.a1, seq) && word_ok(p.b0, seq) && (W < 4 || word_ok(p.b1, seq));
        };
        if (ok()) return true;
        const long long t0 = clock64();
        do
        {
            ld_word2(at, p.a0, p.a1);
            ld_word2(at + 2, p.b0, p.b1);
            if (clock64() - t0 > RES_TIMEOUT_CYCLES) return false;
        } while (!ok());
        return true;
    }

    // Exchange 1. Every CTA has a PRIVATE mailbox of G slots: a candidate is pushed into slot [own CTA] of
    // all G mailboxes (one 8-byte store per word and destination, spread over the block), and a CTA
    // polls only lines nobody else reads -- G CTAs spinning on the same few lines saturate the L2
    // slices that hold them and delay the very stores they wait for. Every warp reduces the G decoded
    // candidates on its own (a pure function of the mail: same answer in every CTA). Posting and gathering
    // are separate calls: the rank-1 update of the previous pivot runs between them (see the kernel).
    __device__ __forceinline__ void post_price(unsigned long long * mail, int G, unsigned int seq, const PCand & mine)
    {
        const unsigned long long b = static_cast<unsigned long long>(__double_as_longlong(mine.v));
        unsigned long long * mine_at = mail + static_cast<size_t>(blockIdx.x) * RES_MAIL_STRIDE;
        // word k of the candidate for this thread (k = thread & 3 in every trip): two selects, no branch
        const int k = threadIdx.x & 3;
        const unsigned int w01 = (k & 1) ? static_cast<unsigned int>(b >> 32) : static_cast<unsigned int>(b);
        const unsigned int w23 = (k & 1) ? static_cast<unsigned int>(mine.col) : static_cast<unsigned int>(mine.idx);
        const unsigned long long word = word_of(seq, (k & 2) ? w23 : w01);
        if (RES_SHARED_MAIL)
        {
            if (threadIdx.x < 4) st_word(mine_at + k, word);
        }
        else
        {
            const size_t stride = static_cast<size_t>(G) * RES_MAIL_STRIDE;
            for (int t = threadIdx.x; t < 4 * G; t += blockDim.x) st_word(mine_at + (t >> 2) * stride + k, word);
        }
    }
    // the mailbox this CTA reads
    __device__ __forceinline__ const unsigned long long * my_box(const unsigned long long * mail, int G)
    {
        return RES_SHARED_MAIL ? mail : mail + static_cast<size_t>(blockIdx.x) * G * RES_MAIL_STRIDE;
    }
    // All threads call. Thread t < G turns slot t into a candidate in registers; the ceil(G/32) warps that
    // hold candidates reduce them (REDUX arg-min), their winners go to shared memory, and after the one
    // barrier every warp reduces those few again: a pure function of the mail, the same answer in every CTA.
    __device__ __forceinline__ PCand gather_price(const unsigned long long * mail, int G, unsigned int seq, bool most_neg,
                                                  Polled first, PCand * scratch, int * timeout_flag)
    {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (G + 31) >> 5;
        PCand c = pnone();
        if (threadIdx.x < G)
        {
            if (!poll_slot<4>(my_box(mail, G), seq, first)) atomicExch(timeout_flag, 1);
            c = PCand{double_of(first.a0, first.a1), static_cast<int>(static_cast<unsigned int>(first.b0)),
                      static_cast<int>(static_cast<unsigned int>(first.b1))};
        }
        if (warp < nw)
        {
            const int w = warp_argmin_lane(pkey(c, most_neg), c.idx);
            if (lane == (w < 0 ? 0 : w)) scratch[warp] = (w < 0) ? pnone() : c;
        }
        __syncthreads();
        const PCand r = lane < nw ? scratch[lane] : pnone();
        const int b = warp_argmin_lane(pkey(r, most_neg), r.idx);
        return b < 0 ? pnone() : scratch[b];
    }

    // Exchange 2: (ratio, row); the slot index is the CTA that staged the row. Posting and gathering are
    // separate calls: the row is staged BETWEEN them, so its 16 KB of stores (and the divisions that
    // produce them) overlap the flight of the mail instead of queueing in front of it.
    __device__ __forceinline__ void post_ratio(unsigned long long * mail, int G, unsigned int seq, const RCand & mine)
    {
        const unsigned long long b = static_cast<unsigned long long>(__double_as_longlong(mine.t));
        unsigned long long * mine_at = mail + static_cast<size_t>(blockIdx.x) * RES_MAIL_STRIDE;
        const int k = threadIdx.x & 3;
        const unsigned int w01 = (k & 1) ? static_cast<unsigned int>(b >> 32) : static_cast<unsigned int>(b);
        const unsigned long long word = word_of(seq, (k & 2) ? static_cast<unsigned int>(mine.idx) : w01);
        if (RES_SHARED_MAIL)
        {
            if (threadIdx.x < 3) st_word(mine_at + k, word);
        }
        else
        {
            const size_t stride = static_cast<size_t>(G) * RES_MAIL_STRIDE;
            if (k < 3)
                for (int t = threadIdx.x; t < 4 * G; t += blockDim.x) st_word(mine_at + (t >> 2) * stride + k, word);
        }
    }
    __device__ __forceinline__ RCand gather_ratio(const unsigned long long * mail, int G, unsigned int seq, Polled first,
                                                  RCand * scratch, int * timeout_flag)
    {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (G + 31) >> 5;
        RCand c = rnone();
        if (threadIdx.x < G)
        {
            if (!poll_slot<3>(my_box(mail, G), seq, first)) atomicExch(timeout_flag, 1);
            c = RCand{double_of(first.a0, first.a1), static_cast<int>(static_cast<unsigned int>(first.b0)),
                      static_cast<int>(threadIdx.x)};
        }
        if (warp < nw)
        {
            const int w = warp_argmin_lane(rkey(c), c.idx);
            if (lane == (w < 0 ? 0 : w)) scratch[warp] = (w < 0) ? rnone() : c;
        }
        __syncthreads();
        const RCand r = lane < nw ? scratch[lane] : rnone();
        const int b = warp_argmin_lane(rkey(r), r.idx);
        return b < 0 ? rnone() : scratch[b];
    }

    // Where a dense column lives: (slot << 8) | owner CTA for dense rank k = slot * G + owner; -1 for a unit
    // column. Computed where the rank is fetched (off the critical path), so the tail does no division.
    __device__ __forceinline__ int home_of(int rank, int G) { return rank < 0 ? -1 : (((rank / G) << 8) | (rank % G)); }
    static_assert(RES_MAX_GRID <= 256, "home_of packs the owner CTA into 8 bits");

    // phase counters of CTA 0 in shared memory (no global read-modify-write on the critical path)
    __device__ __forceinline__ void res_tick(bool on, long long * ph, int slot, long long & t_last)
    {
        if (on && threadIdx.x == 0)
        {
            const long long now = clock64();
            ph[slot] += now - t_last;
            t_last = now;
        }
    }

    // warp_dot() for NC (1 or 2) columns in shared memory against the same shared vector, which is read
    // once for both. Per column: the accumulators, the order and therefore the bits of warp_dot().
    template <int NC>
    __device__ __forceinline__ void warp_dots_ss(const double * __restrict__ g0, const double * __restrict__ g1,
                                                 const double * __restrict__ s, int n, int lane, double & out0, double & out1)
    {
        double a[NC][4];
#pragma unroll
        for (int r = 0; r < NC; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) a[r][c] = 0.0;
        int k = lane * 2;
        for (; k + 192 < n; k += 256)
        {
            double2 sv[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) sv[c] = *reinterpret_cast<const double2 *>(s + k + 64 * c);
#pragma unroll
            for (int r = 0; r < NC; ++r)
            {
                const double * g = (r == 0) ? g0 : g1;
                double2 v[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) v[c] = *reinterpret_cast<const double2 *>(g + k + 64 * c);
#pragma unroll
                for (int c = 0; c < 4; ++c)
                {
                    a[r][c] = fma(v[c].x, sv[c].x, a[r][c]);
                    a[r][c] = fma(v[c].y, sv[c].y, a[r][c]);
                }
            }
        }
        for (; k < n; k += 64)
        {
            const double2 s0 = *reinterpret_cast<const double2 *>(s + k);
#pragma unroll
            for (int r = 0; r < NC; ++r)
            {
                const double2 v0 = *reinterpret_cast<const double2 *>(((r == 0) ? g0 : g1) + k);
                a[r][0] = fma(v0.x, s0.x, a[r][0]);
                a[r][0] = fma(v0.y, s0.y, a[r][0]);
            }
        }
        out0 = warp_sum((a[0][0] + a[0][1]) + (a[0][2] + a[0][3]));
        if (NC > 1) out1 = warp_sum((a[NC - 1][0] + a[NC - 1][1]) + (a[NC - 1][2] + a[NC - 1][3]));
    }

    // The same for columns in GLOBAL memory (kStream: A is read-only during a launch and sits in L2). All the loads
    // of a trip are issued before the first fma so that their round trips overlap.
    template <int NC>
    __device__ __forceinline__ void warp_dots_gs(const double * __restrict__ g0, const double * __restrict__ g1,
                                                 const double * __restrict__ s, int n, int lane, double & out0, double & out1)
    {
        double a[NC][4];
#pragma unroll
        for (int r = 0; r < NC; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) a[r][c] = 0.0;
        int k = lane * 2;
        for (; k + 192 < n; k += 256)
        {
            double2 v[NC][4];
#pragma unroll
            for (int r = 0; r < NC; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) v[r][c] = __ldg(reinterpret_cast<const double2 *>(((r == 0) ? g0 : g1) + k + 64 * c));
            double2 sv[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) sv[c] = *reinterpret_cast<const double2 *>(s + k + 64 * c);
#pragma unroll
            for (int r = 0; r < NC; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c)
                {
                    a[r][c] = fma(v[r][c].x, sv[c].x, a[r][c]);
                    a[r][c] = fma(v[r][c].y, sv[c].y, a[r][c]);
                }
        }
        for (; k < n; k += 64)
        {
            const double2 s0 = *reinterpret_cast<const double2 *>(s + k);
#pragma unroll
            for (int r = 0; r < NC; ++r)
            {
                const double2 v0 = __ldg(reinterpret_cast<const double2 *>(((r == 0) ? g0 : g1) + k));
                a[r][0] = fma(v0.x, s0.x, a[r][0]);
                a[r][0] = fma(v0.y, s0.y, a[r][0]);
            }
        }
        out0 = warp_sum((a[0][0] + a[0][1]) + (a[0][2] + a[0][3]));
        if (NC > 1) out1 = warp_sum((a[NC - 1][0] + a[NC - 1][1]) + (a[NC - 1][2] + a[NC - 1][3]));
    }

    // fused_row() on NR (1 or 2) rows that live in shared memory: applies the pending rank-1 update (if
    // any), stores the rows back and returns row . aq in the summation order of warp_dot(). The pivot
    // row and a_q are read once for both rows.
    template <bool kPending, int NR>
    __device__ __forceinline__ void resident_rows(double * __restrict__ rowA, double * __restrict__ rowB,
                                                  const double * __restrict__ psm, const double * __restrict__ aq, double dA,
                                                  double dB, bool pA, bool pB, int n, int lane, double & outA, double & outB)
    {
        double a[NR][4];
#pragma unroll
        for (int r = 0; r < NR; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) a[r][c] = 0.0;
        int k = lane * 2;
        for (; k + 192 < n; k += 256)
        {
            double2 sv[4], pv[4];
#pragma unroll
            for (int c = 0; c < 4; ++c)
            {
                sv[c] = *reinterpret_cast<const double2 *>(aq + k + 64 * c);
                if (kPending) pv[c] = *reinterpret_cast<const double2 *>(psm + k + 64 * c);
            }
#pragma unroll
            for (int r = 0; r < NR; ++r)
            {
                double * row = (r == 0) ? rowA : rowB;
                const double dprev = (r == 0) ? dA : dB;
                const bool is_p = (r == 0) ? pA : pB;
                double2 v[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) v[c] = *reinterpret_cast<const double2 *>(row + k + 64 * c);
                if (kPending)
                {
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                    {
                        if (is_p)
                            v[c] = pv[c];
                        else
                        {
                            v[c].x = fma(-dprev, pv[c].x, v[c].x);
                            v[c].y = fma(-dprev, pv[c].y, v[c].y);
                        }
                        *reinterpret_cast<double2 *>(row + k + 64 * c) = v[c];
                    }
                }
#pragma unroll
                for (int c = 0; c < 4; ++c)
                {
                    a[r][c] = fma(v[c].x, sv[c].x, a[r][c]);
                    a[r][c] = fma(v[c].y, sv[c].y, a[r][c]);
                }
            }
        }
        for (; k < n; k += 64)
        {
            const double2 s0 = *reinterpret_cast<const double2 *>(aq + k);
            double2 p0 = make_double2(0.0, 0.0);
            if (kPending) p0 = *reinterpret_cast<const double2 *>(psm + k);
#pragma unroll
            for (int r = 0; r < NR; ++r)
            {
                double * row = (r == 0) ? rowA : rowB;
                const double dprev = (r == 0) ? dA : dB;
                const bool is_p = (r == 0) ? pA : pB;
                double2 v0 = *reinterpret_cast<const double2 *>(row + k);
                if (kPending)
                {
                    if (is_p)
                        v0 = p0;
                    else
                    {
                        v0.x = fma(-dprev, p0.x, v0.x);
                        v0.y = fma(-dprev, p0.y, v0.y);
                    }
                    *reinterpret_cast<double2 *>(row + k) = v0;
                }
                a[r][0] = fma(v0.x, s0.x, a[r][0]);
                a[r][0] = fma(v0.y, s0.y, a[r][0]);
            }
        }
        outA = warp_sum((a[0][0] + a[0][1]) + (a[0][2] + a[0][3]));
        if (NR > 1) outB = warp_sum((a[NR - 1][0] + a[NR - 1][1]) + (a[NR - 1][2] + a[NR - 1][3]));
    }

    template <bool kBounded, bool kStream>
    __device__ __forceinline__ void resident_loop(const DevState & S, const ResidentInfo & R, long long chunk)
    {
        extern __shared__ __align__(16) double smem[];
        __shared__ PCand pscratch[32];
        __shared__ RCand rscratch[32];
        __shared__ PCand pgather[RES_MAX_GRID / 32];  // per-warp winners of the two gathers (their own arrays: a slow warp
        __shared__ RCand rgather[RES_MAX_GRID / 32];  // may still be reading pscratch / rscratch of the local arg-min)
        __shared__ long long ph[8];
        __shared__ long long phf[16];
        __shared__ int sh_timeout;
        __shared__ double qinfo_d[3];
        __shared__ int qinfo_i[2];
        __shared__ int sh_nact;
        __shared__ int sh_sub[2];  // bounded: the leaving column that is substituted in this pivot, and its flag before

        const int ld = S.ld;
        double * ysm = smem;
        double * aq = ysm + ld;
        double * psm = aq + ld;
        double * rows = psm + ld;
        double * cols = rows + static_cast<size_t>(R.max_rows) * ld;
        double * dsm = cols + (kStream ? 0 : static_cast<size_t>(R.max_cols) * ld);  // (kStream: no column on chip)
        double * xsm = dsm + R.max_rows;
        double * buval = xsm + R.max_rows;   // per row: unit value / cost of its basic column
        double * bcost = buval + R.max_rows;
        double * costsm = bcost + R.max_rows;
        double * suval = costsm + R.max_cols;
        double * scost = suval + R.max_slice;
        double * bub = scost + R.max_slice;  // kBounded: per row, the upper bound of its basic variable
        int * possm = reinterpret_cast<int *>(kBounded ? bub + R.max_rows : scost + R.max_slice);
        int * colid = possm + R.max_cols;
        int * act = colid + R.max_cols;      // slots of the non-basic dense columns (any order) and, per slot,
        int * actpos = act + R.max_cols;     // its index in that list
        int * bsm = actpos + R.max_cols;     // per row: basic column id, its home (home_of), its unit row
        int * bhome = bsm + R.max_rows;
        int * burow = bhome + R.max_rows;
        int * scol = burow + R.max_rows;
        int * surow = scol + R.max_slice;
        unsigned int * negw = reinterpret_cast<unsigned int *>(surow + R.max_slice);  // kBounded: one sign flag per column

        Scalars * sc = S.sc;
        const int G = gridDim.x;
        const int cta = blockIdx.x;
        const int tid = threadIdx.x;
        const int lane = tid & 31;
        const int warp = tid >> 5;
        const int wpb = blockDim.x >> 5;
        const int bk = blockDim.x - 32;  // the thread that keeps this CTA's position caches (tid 0 keeps the rest)
        const int r0 = static_cast<int>((static_cast<long long>(cta) * S.m) / G);
        const int r1 = static_cast<int>((static_cast<long long>(cta + 1) * S.m) / G);
        const int nrows = r1 - r0;
        const int ncols_own = (R.n_dense > cta) ? (R.n_dense - cta + G - 1) / G : 0;
        const int pos0 = static_cast<int>((static_cast<long long>(cta) * S.nnb) / G);
        const int pos1 = static_cast<int>((static_cast<long long>(cta + 1) * S.nnb) / G);
        const int nslice = pos1 - pos0;
        const int half = ld >> 1;
        const double neg_eps = -S.eps;
        const bool most_neg = S.rule == RULE_MOST_NEG;
        const bool ticks = R.ticks != 0 && blockIdx.x == 0;
        const size_t row_words = 2 * static_cast<size_t>(ld + RES_ROW_EXTRAS);
        unsigned int epoch = 0;

        if (sc->status != TERM_RUNNING) return;  // uniform: nobody has written yet
        if (tid == 0) sh_timeout = 0;
        if (tid < 8) ph[tid] = 0;
        if (tid < 16) phf[tid] = 0;
        const bool fine = R.per_cta_cycles != nullptr;
        long long tf_last = 0;
        const long long iters_at_entry = sc->iterations;  // read before anybody increments it
        const long long limit = sc->iter_limit;
        long long pivots_local = sc->pivots;  // only CTA 0 writes it, and only after the first exchange
        int trace_idx = (S.trace_cap > 0) ? static_cast<int>(pivots_local % S.trace_cap) : 0;

        // ---- entry: y = c_B^T B^-1 from scratch, exactly as the fused engine does it at every launch
        dual_block_partials(S, S.Binv, 0, S.m, S.ypart);
        grid_barrier(sc, epoch);
        dual_block_chain(S, S.ypart, (S.m + DUAL_BLOCK_ROWS - 1) / DUAL_BLOCK_ROWS, nullptr, S.y);
        grid_barrier(sc, epoch);

        // ---- bring this CTA's share of the tableau on chip
        for (int k = 2 * tid; k < ld; k += 2 * blockDim.x) *reinterpret_cast<double2 *>(ysm + k) = ld_keep2(S.y + k);
        for (int t = tid; t < nrows * half; t += blockDim.x)
        {
            const int i = t / half, k = 2 * (t - i * half);
            *reinterpret_cast<double2 *>(rows + static_cast<size_t>(i) * ld + k) =
                ld_keep2(S.Binv + static_cast<size_t>(r0 + i) * ld + k);
        }
        if constexpr (!kStream)
        {
            for (int t = tid; t < ncols_own * half; t += blockDim.x)
            {
                const int s = t / half, k = 2 * (t - s * half);
                const int col = R.dense_ids[s * G + cta];
                *reinterpret_cast<double2 *>(cols + static_cast<size_t>(s) * ld + k) =
                    ld_keep2(S.A + static_cast<size_t>(col) * ld + k);
            }
        }
        for (int i = tid; i < nrows; i += blockDim.x)
        {
            const int col = S.basic[r0 + i];
            xsm[i] = S.xB[r0 + i];
            dsm[i] = 0.0;
            bsm[i] = col;
            bhome[i] = home_of(R.dense_rank[col], G);
            burow[i] = S.unit_row[col];
            buval[i] = S.unit_val[col];
            bcost[i] = S.c[col];
            if constexpr (kBounded) bub[i] = S.ub[col];
        }
        if constexpr (kBounded)
        {
            // between launches every substitution is written into A, c and the unit table: all flags start clear
            for (int w = tid; w < (S.N + 31) / 32; w += blockDim.x) negw[w] = 0u;
        }
        for (int s = tid; s < ncols_own; s += blockDim.x)
        {
            const int col = R.dense_ids[s * G + cta];
            colid[s] = col;
            costsm[s] = S.c[col];
            possm[s] = -1;  // basic until the scan below says otherwise
        }
        __syncthreads();
        for (int pos = tid; pos < S.nnb; pos += blockDim.x)
        {
            const int k = R.dense_rank[S.nonbasic[pos]];
            if (k >= 0 && k % G == cta) possm[k / G] = pos;
        }
        for (int j = tid; j < nslice; j += blockDim.x)
        {
            const int col = S.nonbasic[pos0 + j];
            scol[j] = col;
            surow[j] = S.unit_row[col];
            suval[j] = S.unit_val[col];
            scost[j] = S.c[col];
        }
        __syncthreads();
        if (tid == 0)
        {
            int n = 0;
            for (int sl = 0; sl < ncols_own; ++sl)
                if (possm[sl] >= 0)
                {
                    act[n] = sl;
                    actpos[sl] = n;
                    n += 1;
                }
            sh_nact = n;
        }
        __syncthreads();

        bool pending = false;  // rank-1 update of the last pivot not yet applied to the rows
        bool have_d = false;
        int p_prev = -1;
        int exit_status = TERM_RUNNING;
        long long local_iters = iters_at_entry;
        long long flips_local = 0;
        long long t_last = clock64();
        tf_last = t_last;

        for (long long it = 0; it < chunk; ++it)
        {
            const unsigned int seq = R.seq0 + static_cast<unsigned int>(it);

            // ================= price (Simplex.cpp:313, :49-61 / :87-102)
            PCand best = pnone();
            {
                // the non-basic dense columns of this CTA, two per warp and trip: y is read once for both
                const int nact = sh_nact;
                for (int t = RES_COLS_PER_WARP * warp; t < nact; t += RES_COLS_PER_WARP * wpb)
                {
                    const int sa = act[t];
                    const bool two = RES_COLS_PER_WARP > 1 && t + 1 < nact;
                    const int sb = two ? act[t + 1] : sa;
                    double dota = 0.0, dotb = 0.0;
                    if constexpr (kStream)
                    {
                        const double * ga = S.A + static_cast<size_t>(colid[sa]) * ld;
                        const double * gb = S.A + static_cast<size_t>(colid[sb]) * ld;
                        if (two)
                            warp_dots_gs<2>(ga, gb, ysm, ld, lane, dota, dotb);
                        else
                            warp_dots_gs<1>(ga, ga, ysm, ld, lane, dota, dotb);
                    }
                    else if (two)
                        warp_dots_ss<2>(cols + static_cast<size_t>(sa) * ld, cols + static_cast<size_t>(sb) * ld, ysm, ld, lane, dota,
                                        dotb);
                    else
                        warp_dots_ss<1>(cols + static_cast<size_t>(sa) * ld, cols + static_cast<size_t>(sa) * ld, ysm, ld, lane, dota,
                                        dotb);
                    double sja = costsm[sa] - dota;
                    if constexpr (kBounded)
                        if (res_flag(negw, colid[sa])) sja = -sja;
                    if (sja < neg_eps)
                    {
                        const PCand c{sja, possm[sa], colid[sa]};
                        if (pbetter(c, best, most_neg)) best = c;
                    }
                    if (two)
                    {
                        double sjb = costsm[sb] - dotb;
                        if constexpr (kBounded)
                            if (res_flag(negw, colid[sb])) sjb = -sjb;
                        if (sjb < neg_eps)
                        {
                            const PCand c{sjb, possm[sb], colid[sb]};
                            if (pbetter(c, best, most_neg)) best = c;
                        }
                    }
                }
            }
            if (lane != 0) best = pnone();
            // unit columns of this CTA's slice of positions: on the LAST threads of the block, whose warps
            // have no column to price unless the CTA holds more than 2 x 15 non-basic dense columns
            for (int j = blockDim.x - 1 - tid; j < nslice; j += blockDim.x)
            {
                const int ur = surow[j];
                if (ur >= 0)
                {
                    // product rounded, then subtracted: the bits warp_dot would deliver for value * e_row
                    double sj = __dsub_rn(scost[j], __dmul_rn(suval[j], ysm[ur]));
                    if constexpr (kBounded)
                        if (res_flag(negw, scol[j])) sj = -sj;
                    if (sj < neg_eps)
                    {
                        const PCand c{sj, pos0 + j, scol[j]};
                        if (pbetter(c, best, most_neg)) best = c;
                    }
                }
            }
            res_tick(fine, phf, 0, tf_last);  // price loops
            best = block_allmin_p(best, pscratch, most_neg, static_cast<int>(blockDim.x) - 32 * (warp + 1) < nslice);
            res_tick(fine, phf, 1, tf_last);  // block arg-min
            res_tick(ticks, ph, 0, t_last);
            post_price(R.pmail, G, seq, best);
            res_tick(fine, phf, 11, tf_last);  // mail push
            // While the mail flies: the rank-1 update of the previous pivot on this CTA's rows. It needs
            // only (d, rho_p / d_p) of that pivot, not the entering column this exchange is about to
            // name, and it is elementwise, so all the warps share it evenly. The first polls of the
            // mailbox are issued half way through: their round trip hides behind the second half.
            Polled pfirst{0ULL, 0ULL, 0ULL, 0ULL};
            {
                // a thread keeps its column pair k and walks down the rows: rho_p[k] is read once per thread,
                // d_i once per row (shared-memory broadcasts)
                const int nr = pending ? nrows : 0;
                const int split = RES_EARLY_POLL ? nr / 2 : nr;
                for (int part = 0; part < 2; ++part)
                {
                    const int lo = part == 0 ? 0 : split, hi = part == 0 ? split : nr;
                    for (int k = 2 * tid; k < ld; k += 2 * blockDim.x)
                    {
                        const double2 pv = *reinterpret_cast<const double2 *>(psm + k);
                        for (int i = lo; i < hi; ++i)
                        {
                            double2 * at = reinterpret_cast<double2 *>(rows + static_cast<size_t>(i) * ld + k);
                            if (r0 + i == p_prev)
                                *at = pv;
                            else
                            {
                                const double dprev = dsm[i];
                                double2 v = *at;
                                v.x = fma(-dprev, pv.x, v.x);
                                v.y = fma(-dprev, pv.y, v.y);
                                *at = v;
                            }
                        }
                    }
                    if (part == 0) pfirst = poll_issue(my_box(R.pmail, G), G);
                }
                pending = false;
            }
            res_tick(fine, phf, 12, tf_last);  // rank-1 update of the previous pivot
            const PCand ebest = gather_price(R.pmail, G, seq, most_neg, pfirst, pgather, &sh_timeout);  // exchange 1
            if (sh_timeout != 0)
            {
                exit_status = TERM_RESIDENT_TIMEOUT;
                break;
            }
            res_tick(ticks, ph, 1, t_last);
            res_tick(fine, phf, 2, tf_last);  // exchange 1

            local_iters += 1;
            if (cta == 0 && tid == 0)
            {
                sc->iterations = local_iters;
                sc->action = ACT_NONE;
                if (ebest.idx >= 0)
                {
                    sc->e_pos = ebest.idx;
                    sc->q_col = ebest.col;
                    sc->s_q = ebest.v;
                }
                else
                {
                    sc->e_pos = -1;
                    sc->last_ret = -1;
                }
            }
            if (ebest.idx < 0)
            {
                exit_status = (local_iters >= limit) ? TERM_ITER_LIMIT : TERM_OPTIMAL;  // Simplex.cpp:319-323, 347-350
                break;
            }
            const int e = ebest.idx;
            const int q = ebest.col;
            const double s_q = ebest.v;
            // What the row that q is about to become basic in has to remember about it (used in the tail).
            // Fetched by the last warp, which owns no row unless a CTA holds more than 15 of them.
            if (tid == bk)
            {
                qinfo_i[0] = home_of(__ldg(R.dense_rank + q), G);
                qinfo_i[1] = __ldg(S.unit_row + q);
                qinfo_d[0] = __ldg(S.unit_val + q);
                qinfo_d[1] = __ldg(S.c + q);
                if constexpr (kBounded) qinfo_d[2] = __ldg(S.ub + q);
            }
            bool qneg = false;  // sign of the entering column: B^-1 (-a_q) = -(B^-1 a_q) exactly
            if constexpr (kBounded) qneg = res_flag(negw, q);

            // ================= fused: rank-1 update k-1 + FTRAN k + ratio candidates (Simplex.cpp:326-333, :66-84)
            {
                const double * src = S.A + static_cast<size_t>(q) * ld;  // A is never modified by this engine
                for (int k = 2 * tid; k < ld; k += 2 * blockDim.x)
                    *reinterpret_cast<double2 *>(aq + k) = __ldg(reinterpret_cast<const double2 *>(src + k));
            }
            __syncthreads();
            res_tick(fine, phf, 3, tf_last);  // bookkeeping, a_q fetch, barrier
            RCand rb = rnone();
            for (int i = RES_ROWS_PER_WARP * warp; i < nrows; i += RES_ROWS_PER_WARP * wpb)  // FTRAN: d_i = row_i . a_q
            {
                const double * rowa = rows + static_cast<size_t>(i) * ld;
                const bool two = RES_ROWS_PER_WARP > 1 && i + 1 < nrows;
                double da = 0.0, db = 0.0;
                if (two)
                    warp_dots_ss<2>(rowa, rowa + ld, aq, ld, lane, da, db);
                else
                    warp_dots_ss<1>(rowa, rowa, aq, ld, lane, da, db);
                if (lane == 0)
                {
                    for (int r = 0; r < (two ? 2 : 1); ++r)
                    {
                        double di = (r == 0) ? da : db;
                        if constexpr (kBounded)
                            if (qneg) di = -di;
                        dsm[i + r] = di;
                        const double xi = xsm[i + r];
                        if (di > S.eps)  // Simplex.cpp:73-79 / 125-132
                        {
                            const double t = xi / di;
                            if (t < DBL_MAX)
                            {
                                const RCand c{t, r0 + i + r, cta};
                                if (rbetter(c, rb)) rb = c;
                            }
                        }
                        else if (kBounded && di < -S.eps)  // Simplex.cpp:136-153
                        {
                            const double t = (bub[i + r] - xi) / (-di);
                            if (t < DBL_MAX)
                            {
                                const RCand c{t, (r0 + i + r) | RES_T2_TAG, cta};
                                if (rbetter(c, rb)) rb = c;
                            }
                        }
                    }
                }
            }
            res_tick(fine, phf, 4, tf_last);  // warp 0's rows of the pass
            have_d = true;
            rb = block_allmin_r(rb, rscratch);  // every thread: this CTA's best row (the barrier publishes dsm and the rows)
            res_tick(fine, phf, 5, tf_last);  // block arg-min (waits for the slowest warp of the pass)
            post_ratio(R.rmail, G, seq, rb);
            res_tick(fine, phf, 6, tf_last);  // mail push
            Polled rfirst{0ULL, 0ULL, 0ULL, 0ULL};
            if (!RES_EARLY_POLL || rb.idx < 0) rfirst = poll_issue(my_box(R.rmail, G), G);
            if (rb.idx >= 0)
            {
                // The only row of this CTA that can be the pivot row: stage it, ALREADY divided by its
                // pivot element (the division the readers would do: same operands, same bits), where
                // every CTA can poll it, followed by what the others must learn about the leaving column.
                const int i = (kBounded ? res_row_of(rb.idx) : rb.idx) - r0;
                const double di = dsm[i];
                // (a t2 row has d_i < 0: 0 / d_i is -0, so the shortcut below is for positive pivots only)
                const bool plain = kBounded && di < 0.0;
                const double * row = rows + static_cast<size_t>(i) * ld;
                unsigned long long * dst = R.rowmail + static_cast<size_t>(cta) * row_words;
                bool polled = !RES_EARLY_POLL;
                for (int k = 2 * tid; k < ld || !polled; k += 2 * blockDim.x)
                {
                    // 0 / d_i = 0 with the sign of the numerator (d_i > 0): the same bits without the slow path the
                    // IEEE division takes for a zero numerator -- and B^-1 is mostly zeros for many pivots
                    double2 v = make_double2(0.0, 0.0);
                    if (k < ld)
                    {
                        v = *reinterpret_cast<const double2 *>(row + k);
                        if (v.x != 0.0 || plain) v.x = v.x / di;
                        if (v.y != 0.0 || plain) v.y = v.y / di;
                    }
                    if (!polled)
                    {
                        // first polls of the mailbox: issued after the divisions, their round trip hides behind the stores
                        rfirst = poll_issue(my_box(R.rmail, G), G);
                        polled = true;
                    }
                    if (k < ld)
                    {
                        st_double(dst + 2 * k, v.x, seq);
                        st_double(dst + 2 * k + 2, v.y, seq);
                    }
                }
                if (tid < 5)
                {
                    unsigned long long * x = dst + 2 * (ld + tid);
                    if (tid == 0) st_double(x, di, seq);
                    if (tid == 1) st_double(x, buval[i], seq);
                    if (tid == 2) st_double(x, bcost[i], seq);
                    if (tid == 3) st_word2(x, word_of(seq, static_cast<unsigned int>(bsm[i])), word_of(seq, static_cast<unsigned int>(bhome[i])));
                    if (tid == 4) st_word2(x, word_of(seq, static_cast<unsigned int>(burow[i])), word_of(seq, 0u));
                }
            }
            res_tick(ticks, ph, 2, t_last);
            res_tick(fine, phf, 7, tf_last);  // row staging
            const RCand rbest = gather_ratio(R.rmail, G, seq, rfirst, rgather, &sh_timeout);  // exchange 2
            if (sh_timeout != 0)
            {
                exit_status = TERM_RESIDENT_TIMEOUT;
                break;
            }
            res_tick(ticks, ph, 3, t_last);
            res_tick(fine, phf, 8, tf_last);  // exchange 2

            // ================= leaving decision (same answer in every CTA), pivot row, dual recurrence, x_B
            if constexpr (kBounded)
            {
                const double min_theta = (rbest.idx >= 0) ? rbest.t : DBL_MAX;
                const double ub_s = qinfo_d[2];
                if (ub_s < min_theta)
                {
                    // Simplex.cpp:160-164: the entering variable reaches its own bound first. No basis change:
                    // B^-1, y, the lists and the staged rows stay; x_B -= u d; b -= a_q u; the flag of q toggles
                    for (int i = tid; i < nrows; i += blockDim.x)
                    {
                        const double a = qneg ? -aq[r0 + i] : aq[r0 + i];
                        S.b[r0 + i] = __dsub_rn(S.b[r0 + i], __dmul_rn(a, ub_s));  // two roundings, as g++ -O3 compiles it
                        xsm[i] = fma(-ub_s, dsm[i], xsm[i]);
                    }
                    if (tid == 0)
                    {
                        negw[q >> 5] ^= 1u << (q & 31);
                        if (cta == 0)
                        {
                            S.flip[q] ^= 1;
                            sc->last_ret = -2;
                        }
                    }
                    flips_local += 1;
                    __syncthreads();
                    res_tick(ticks, ph, 4, t_last);
                    res_tick(fine, phf, 10, tf_last);
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
                if (cta == 0 && tid == 0) sc->last_ret = -1;
                break;
            }
            const int p = kBounded ? res_row_of(rbest.idx) : rbest.idx;
            // Simplex.cpp:165-176: a t2 winner leaves at its upper bound and is substituted (b -= a u, its flag toggles)
            const bool sub = kBounded && (rbest.idx & RES_T2_TAG) != 0;
            const double theta = rbest.t;  // == x_p / d_p (t1) or (u - x_p) / (-d_p) (t2) bit for bit
            const unsigned long long * srow = R.rowmail + static_cast<size_t>(rbest.src) * row_words;
            bool got = true;
            // First attempt of everything this thread needs, all loads in flight together: its two doubles of
            // the row and, for the two threads that keep the caches, the extras. Validation (and the rare
            // retry) comes after.
            const bool keeper = tid == 0 || tid == bk;
            const int j0 = 2 * tid;
            const unsigned long long * x = srow + 2 * ld;
            unsigned long long ra0 = 0, ra1 = 0, rb0 = 0, rb1 = 0;
            unsigned long long a0 = 0, a1 = 0, b0 = 0, b1 = 0, c0 = 0, c1 = 0, w0 = 0, w1 = 0, w2 = 0, w3 = 0;
            if (j0 < ld)
            {
                ld_word2(srow + 2 * j0, ra0, ra1);
                ld_word2(srow + 2 * j0 + 2, rb0, rb1);
            }
            if (keeper)
            {
                ld_word2(x, a0, a1);
                ld_word2(x + 2, b0, b1);
                ld_word2(x + 4, c0, c1);
                ld_word2(x + 6, w0, w1);
                ld_word2(x + 8, w2, w3);
            }
            for (int j = j0; j < ld; j += 2 * blockDim.x)
            {
                double2 r;
                if (j == j0 && word_ok(ra0, seq) && word_ok(ra1, seq) && word_ok(rb0, seq) && word_ok(rb1, seq))
                {
                    r.x = double_of(ra0, ra1);
                    r.y = double_of(rb0, rb1);
                }
                else
                    got = ld_double2(srow + 2 * j, seq, r) && got;
                *reinterpret_cast<double2 *>(psm + j) = r;  // rho_p / d_p
                double2 yv = *reinterpret_cast<double2 *>(ysm + j);
                yv.x = fma(s_q, r.x, yv.x);  // y' = y + s_q * (rho_p / d_p)
                yv.y = fma(s_q, r.y, yv.y);
                *reinterpret_cast<double2 *>(ysm + j) = yv;
            }
            double dp = 0.0, lc_uval = 0.0, lc_cost = 0.0;
            int lc = -1, klc = -1, lc_urow = -1;
            if (keeper)
            {
                const long long t0 = clock64();
                while (!(word_ok(a0, seq) && word_ok(a1, seq) && word_ok(b0, seq) && word_ok(b1, seq) && word_ok(c0, seq) &&
                         word_ok(c1, seq) && word_ok(w0, seq) && word_ok(w1, seq) && word_ok(w2, seq)))
                {
                    if (clock64() - t0 > RES_TIMEOUT_CYCLES)
                    {
                        got = false;
                        break;
                    }
                    ld_word2(x, a0, a1);
                    ld_word2(x + 2, b0, b1);
                    ld_word2(x + 4, c0, c1);
                    ld_word2(x + 6, w0, w1);
                    ld_word2(x + 8, w2, w3);
                }
                dp = double_of(a0, a1);
                lc_uval = double_of(b0, b1);
                lc_cost = double_of(c0, c1);
                lc = static_cast<int>(static_cast<unsigned int>(w0));
                klc = static_cast<int>(static_cast<unsigned int>(w1));
                lc_urow = static_cast<int>(static_cast<unsigned int>(w2));
            }
            if (!got) atomicExch(&sh_timeout, 1);
            res_tick(fine, phf, 9, tf_last);  // row + extras polls, y update (thread 0's share)
            for (int i = tid; i < nrows; i += blockDim.x)
            {
                const double di = dsm[i];
                xsm[i] = (r0 + i == p) ? theta : fma(-theta, di, xsm[i]);
                if (kBounded && sub && r0 + i == p) dsm[i] = -di;  // read-out: the pivot element after the substitution
            }
            if (tid == 0)
            {
                // Basis::update (Basis.cpp:48-53) on the row caches this CTA owns
                if (p >= r0 && p < r1)
                {
                    const int i = p - r0;
                    bsm[i] = q;
                    bhome[i] = qinfo_i[0];
                    burow[i] = qinfo_i[1];
                    buval[i] = qinfo_d[0];
                    bcost[i] = qinfo_d[1];
                    if constexpr (kBounded) bub[i] = qinfo_d[2];
                }
                if constexpr (kBounded)
                {
                    if (sub)
                    {
                        sh_sub[0] = lc;
                        sh_sub[1] = res_flag(negw, lc) ? 1 : 0;
                        negw[lc >> 5] ^= 1u << (lc & 31);
                        if (cta == 0) S.flip[lc] ^= 1;
                    }
                }
                if (cta == 0)
                {
                    if (S.trace_cap > 0)
                    {
                        S.trace[trace_idx] = make_int4(lc, p, q, e);
                        trace_idx = (trace_idx + 1 == S.trace_cap) ? 0 : trace_idx + 1;
                    }
                    pivots_local += 1;
                    sc->pivots = pivots_local;
                    sc->p_row = p;
                    sc->theta = theta;
                    sc->d_p = sub ? -dp : dp;
                    sc->last_ret = p;
                    ph[7] += 1;
                }
            }
            if (tid == bk)
            {
                // ... and on the column / position caches: q leaves the non-basic set, lc joins it at position e
                const int hq = qinfo_i[0], hl = klc;
                if (hq >= 0 && (hq & 0xff) == cta)
                {
                    const int sl = hq >> 8, at = actpos[sl], last = act[sh_nact - 1];
                    possm[sl] = -1;
                    act[at] = last;
                    actpos[last] = at;
                    sh_nact -= 1;
                }
                if (hl >= 0 && (hl & 0xff) == cta)
                {
                    const int sl = hl >> 8;
                    possm[sl] = e;
                    act[sh_nact] = sl;
                    actpos[sl] = sh_nact;
                    sh_nact += 1;
                }
                if (e >= pos0 && e < pos1)
                {
                    const int j = e - pos0;
                    scol[j] = lc;
                    surow[j] = lc_urow;
                    suval[j] = lc_uval;
                    scost[j] = lc_cost;
                }
            }
            pending = true;
            p_prev = p;
            __syncthreads();
            if (sh_timeout != 0)
            {
                exit_status = TERM_RESIDENT_TIMEOUT;
                break;
            }
            if constexpr (kBounded)
            {
                if (sub)
                {
                    // b -= a_lc u on the rows this CTA owns (A is not modified during the launch; lc came with the row)
                    const int lcol = sh_sub[0];
                    const bool lneg = sh_sub[1] != 0;
                    const double u = __ldg(S.ub + lcol);
                    const double * acol = S.A + static_cast<size_t>(lcol) * ld;
                    for (int i = tid; i < nrows; i += blockDim.x)
                    {
                        const double a = lneg ? -__ldg(acol + r0 + i) : __ldg(acol + r0 + i);
                        S.b[r0 + i] = __dsub_rn(S.b[r0 + i], __dmul_rn(a, u));
                    }
                }
            }
            res_tick(ticks, ph, 4, t_last);
            res_tick(fine, phf, 10, tf_last);  // x_B, caches, barrier
            if (fine && tid == 0) phf[15] += 1;
            if (local_iters >= limit)
            {
                exit_status = TERM_ITER_LIMIT;
                break;
            }
        }

        // ---- exit: apply the pending update and write this CTA's share back to HBM
        __syncthreads();
        if (pending)
        {
            for (int i = warp; i < nrows; i += wpb)
            {
                double da, db;
                double * row = rows + static_cast<size_t>(i) * ld;
                resident_rows<true, 1>(row, row, psm, psm, dsm[i], 0.0, r0 + i == p_prev, false, ld, lane, da, db);
            }
            __syncthreads();
        }
        for (int t = tid; t < nrows * half; t += blockDim.x)
        {
            const int i = t / half, k = 2 * (t - i * half);
            st2(S.Binv + static_cast<size_t>(r0 + i) * ld + k,
                *reinterpret_cast<const double2 *>(rows + static_cast<size_t>(i) * ld + k));
        }
        for (int i = tid; i < nrows; i += blockDim.x)
        {
            S.xB[r0 + i] = xsm[i];
            S.basic[r0 + i] = bsm[i];  // the lists live on chip during the launch (Basis.cpp:48-53)
            if (have_d) S.d[r0 + i] = dsm[i];
        }
        for (int j = tid; j < nslice; j += blockDim.x) S.nonbasic[pos0 + j] = scol[j];
        if (fine && tid < 16) R.per_cta_cycles[cta * 16 + tid] = phf[tid];
        if (cta == 0)
        {
            for (int j = tid; j < ld; j += blockDim.x)
            {
                S.y[j] = ysm[j];
                if (pending) S.prow[j] = psm[j];
            }
            if (tid < 8 && R.ticks != 0) sc->phase_cycles[tid] += ph[tid];
            if constexpr (kBounded)
            {
                // the flags of this launch (the same in every CTA), for k_apply_negations behind the kernel
                for (int col = tid; col < S.N; col += blockDim.x) S.neg[col] = res_flag(negw, col) ? 1 : 0;
                if (tid == 0) sc->flips += flips_local;
            }
            if (tid == 0 && exit_status != TERM_RUNNING) sc->status = exit_status;
        }
    }

    __global__ void __launch_bounds__(RES_THREADS, 1) k_resident(DevState S, ResidentInfo R, long long chunk)
    {
        resident_loop<false, false>(S, R, chunk);
    }

    // the same loop for LPs with finite upper bounds (bounded ratio test, bound flips, substituted leaving variables)
    __global__ void __launch_bounds__(RES_THREADS, 1) k_resident_bounded(DevState S, ResidentInfo R, long long chunk)
    {
        resident_loop<true, false>(S, R, chunk);
    }

    // the dense columns stay in global memory (L2), only the rows of B^-1 live on chip (kStream)
    __global__ void __launch_bounds__(RES_THREADS, 1) k_resident_stream(DevState S, ResidentInfo R, long long chunk)
    {
        resident_loop<false, true>(S, R, chunk);
    }
    __global__ void __launch_bounds__(RES_THREADS, 1) k_resident_stream_bounded(DevState S, ResidentInfo R, long long chunk)
    {
        resident_loop<true, true>(S, R, chunk);
    }

    // Behind k_resident_bounded (stream order: no CTA reads A any more): the columns flagged in S.neg are rewritten
    // as the reference rewrites them at substitution time, and the flags are cleared (apply_pending_negations).
    __global__ void __launch_bounds__(256) k_apply_negations(DevState S)
    {
        apply_pending_negations(S, -1, 0, false);
    }
}  // namespace sb200

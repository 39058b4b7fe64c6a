// Sharded engine: ONE wide LP over G GPUs of one NVSwitch box (BASELINE configs[3]).
//
//   A       column-sharded: rank r holds the cyclic set of global columns {r + k*G} (each column
//           contiguous; cyclic keeps structural/slack columns and so the pricing work balanced)
//   B^-1    row-sharded:    rank r holds global rows    [r*rpr, (r+1)*rpr), row-major; x_B with it
//   y, lists, c, b          replicated (small), kept identical on every rank by construction
//
// Every rank runs the fused persistent kernel of sb200_fused.cuh on its shard. Per pivot the ranks
// exchange, THROUGH PEER-MAPPED HBM WINDOWS written by the kernels themselves (NVLink stores of self-validating
// 8-byte words: 32 bits of payload + the 32-bit iteration number, NCCL-LL style; no fence, no flag, no host round
// trip, no NCCL launch on the path):
//   1. the pricing candidate of each rank (key, aux, idx, cls) -> every rank takes the same lexicographic minimum
//      (value, position)                                                              [arg-min all-reduce]
//      and, BESIDE it, each rank's candidate COLUMN into its slot of every peer's window (before the winner is known)
//   2. the ratio-test candidate of each rank (ratio, class, row)                      [arg-min all-reduce]
//      and, beside it, each rank's candidate ROW of the updated B^-1
// so that neither the entering column nor the pivot row waits for an arg-min (round 1 broadcast them afterwards,
// with a system fence and a flag each: four serial NVLink round trips per pivot instead of two).
// Once per launch the dual solve is chained over the ranks (block partials are added in ascending row-block
// order across ranks, so y has the same bits as on one GPU).
//
// Because a column's reduced cost and a row's d_i are each computed with one fixed summation order (one warp,
// or 2 / 4 warps that each keep their accumulators of warp_dot, fused_row_share), and ties are broken on
// (value, position) / (ratio, class, row), the pivot sequence is independent of G.
//
// Pricing strides over the rank's own list of owned non-basic positions (ShardInfo::own_pos), kept up to date pivot
// by pivot. Every spin of the kernel honours an abort word, so an exchange failure ends the run on every CTA of
// every rank (TERM_EXCHANGE_TIMEOUT) instead of hanging it.
//
// Scope of this engine: with or without finite upper bounds (k_sharded_fused_bounded after sb200_shard_set_bounds;
// BASELINE configs[3] has none); most-negative or first-eligible pricing; start from a unit crash basis (sb200_shard_prepare), from the basis the previous run ended on with its
// B^-1 carried over (sb200_shard_continue: Phase I -> II), or from any basis (sb200_shard_prepare_general +
// sb200_shard_invert: every rank inverts the gathered basis matrix and keeps its rows).
#pragma once
#include "sb200_fused.cuh"

namespace sb200
{
    constexpr int MAX_WORLD = 8;
    constexpr int TERM_EXCHANGE_TIMEOUT = -3;
    constexpr unsigned long long XCHG_TIMEOUT_NS = 4000000000ULL;  // a peer that stays silent this long is gone

    // One candidate in flight, NCCL-LL style: six 8-byte words, each carrying 32 bits of payload
    // and the 32-bit sequence number, so that every word is self-validating (8-byte stores are
    // single-copy atomic over NVLink) and no fence / release-acquire pair sits on the critical path.
    // words: key.lo key.hi aux.lo aux.hi idx cls
    struct Mail
    {
        unsigned long long w[8];
    };

    // Lives at the start of every rank's window allocation; vector buffers follow at WIN_DATA_OFFSET.
    struct WindowHeader
    {
        Mail price[MAX_WORLD];  // [src rank]
        Mail ratio[MAX_WORLD];
        unsigned long long carry_flag;
        unsigned long long y_flag;
        unsigned long long abort_flag;  // raised (here and in every peer's window) by whoever gives up waiting
        unsigned long long pad;
    };
    constexpr size_t WIN_DATA_OFFSET = 2048;
    static_assert(sizeof(WindowHeader) <= WIN_DATA_OFFSET, "window header grew");

    // data layout, in units of ld 8-byte words: candidate columns [parity][src] and candidate rows [parity][src]
    // as self-validating words (2 ld words each: low and high half of every double beside the sequence number),
    // then the carry and the final y of the chained dual solve as plain doubles
    constexpr int WIN_VEC_AQ = 0;
    constexpr int WIN_VEC_ROW = 4 * MAX_WORLD;
    constexpr int WIN_VEC_LC = 8 * MAX_WORLD;   // bounded LPs: the column of a leaving variable that is substituted first
    constexpr int WIN_VEC_CARRY = 12 * MAX_WORLD;
    constexpr int WIN_VEC_Y = 12 * MAX_WORLD + 1;
    constexpr int WIN_VECS = 12 * MAX_WORLD + 2;
    __host__ __device__ inline size_t window_bytes(int ld) { return WIN_DATA_OFFSET + WIN_VECS * static_cast<size_t>(ld) * 8; }
    __host__ __device__ inline int win_slot(int base, int parity, int src) { return base + 2 * (parity * MAX_WORLD + src); }

    struct ShardInfo
    {
        int rank, world;
        int col0, ncols, cols_per_rank;  // local columns of A are global ids [col0, col0+ncols)
        int row0, nrows, rows_per_rank;  // local rows of B^-1 / x_B are global rows [row0, row0+nrows)
        unsigned long long launch_seq;   // 1-based launch counter (dual chain flags)
        unsigned long long timeout_ns;   // a peer that stays silent this long is gone (XCHG_TIMEOUT_NS unless overridden)
        // The non-basic POSITIONS whose column this rank owns, as a compact list kept up to date pivot by pivot
        // (a pivot changes the owner of one position at most): own_pos[0 .. *own_cnt), pos_slot[pos] = its slot or -1.
        // The pricing sweep strides over this list, so every warp of every rank gets the same number of columns
        // whatever the pivots have done to the order of the non-basic list.
        int * own_pos;
        int * pos_slot;
        int * own_cnt;
        unsigned char * win[MAX_WORLD];  // window of every rank (own entry = local pointer)
    };

    // dynamic shared memory of k_sharded_fused: y, a_q, pivot row, d of the CTA's rows. Nothing else: at m = 8192
    // the three vectors alone are 192 KB, and one more KB would push the SM to its largest shared-memory
    // carve-out, halving the L1 that holds the streaming loads in flight (measured: pricing 15 % slower).
    inline size_t sharded_smem_bytes(int ld, int nrows, int grid)
    {
        const int own = persist_max_own_rows(nrows, grid > 0 ? grid : 1);
        return (3 * static_cast<size_t>(ld) + static_cast<size_t>(own) + 16) * sizeof(double);
    }
    // lane partials of the rows that several warps share, in global memory (L2): [local row][3][32] doubles
    constexpr int ROW_PARTIAL_DOUBLES = 3 * 32;

    __device__ __forceinline__ WindowHeader * win_hdr(unsigned char * w) { return reinterpret_cast<WindowHeader *>(w); }
    __device__ __forceinline__ double * win_vec(unsigned char * w, int which, int ld)
    {
        return reinterpret_cast<double *>(w + WIN_DATA_OFFSET) + static_cast<size_t>(which) * ld;
    }

    __device__ __forceinline__ void st_release_sys(unsigned long long * p, unsigned long long v)
    {
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
    }
    __device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long * p)
    {
        unsigned long long v;
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
        return v;
    }
    __device__ __forceinline__ void st_relaxed_sys(unsigned long long * p, unsigned long long v)
    {
        asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
    }
    __device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long * p)
    {
        unsigned long long v;
        asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
        return v;
    }
    __device__ __forceinline__ unsigned long long now_ns()
    {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        return t;
    }

    // ---- giving up, for every CTA of every rank at once. Whoever waits too long raises abort_flag in its own
    // and in every peer's window; every spin of the kernel (grid barrier, mail poll, flag wait) looks at its own
    // window's flag every few hundred polls and leaves when it is up. After that no CTA enters a barrier again,
    // so nobody is left spinning for a CTA that has gone (the exit skips the barrier and the flush: the state
    // is not usable after an exchange failure, the host reports TERM_EXCHANGE_TIMEOUT).
    __device__ __forceinline__ void raise_abort(const ShardInfo & H)
    {
        for (int r = 0; r < H.world; ++r) st_relaxed_sys(&win_hdr(H.win[r])->abort_flag, 1ULL);
    }
    // called by a spinning thread every 256th poll; t0 = when it started to wait
    __device__ __forceinline__ bool must_leave(const ShardInfo & H, unsigned long long t0)
    {
        if (ld_relaxed_sys(&win_hdr(H.win[H.rank])->abort_flag) != 0) return true;
        if (now_ns() - t0 > H.timeout_ns)
        {
            raise_abort(H);
            return true;
        }
        return false;
    }

    // Grid barrier of the sharded kernel: the monotonic counter of grid_barrier(), with a way out.
    // Returns false (in every thread) when the run is being abandoned.
    __device__ __forceinline__ bool grid_barrier_or_leave(Scalars * sc, unsigned int & epoch, const ShardInfo & H, int * sh_ok)
    {
        __syncthreads();
        if (threadIdx.x == 0)
        {
            epoch += 1;
            const unsigned long long target = static_cast<unsigned long long>(epoch) * gridDim.x;
            __threadfence();
            atomicAdd(&sc->barrier_count, 1ULL);
            const unsigned long long t0 = now_ns();
            unsigned int polls = 0;
            while (ld_acquire_u64(&sc->barrier_count) < target)
            {
                if ((++polls & 255u) == 0u && must_leave(H, t0))
                {
                    *sh_ok = 0;
                    break;
                }
            }
            __threadfence();
        }
        __syncthreads();
        return *sh_ok != 0;
    }

    // Bounded wait on a flag in LOCAL memory written by a peer (release/acquire at system scope: the vector
    // the flag announces is complete when the flag is seen). false = the run is being abandoned.
    __device__ __forceinline__ bool wait_flag(const ShardInfo & H, const unsigned long long * p, unsigned long long want)
    {
        const unsigned long long t0 = now_ns();
        unsigned int polls = 0;
        while (ld_acquire_sys(p) < want)
        {
            if ((++polls & 255u) == 0u && must_leave(H, t0)) return false;
        }
        return true;
    }

    // Arg-min all-reduce of one candidate per rank. `mine` is this rank's candidate (same value in
    // every thread); kind 0 = pricing mailboxes, 1 = ratio mailboxes. Every CTA calls it; CTA 0
    // publishes (6 words to each rank, one thread per word, plain NVLink stores), every CTA polls
    // its LOCAL window. Returns the global best in every thread; *sh_ok = 0 when the run is abandoned.
    __device__ __forceinline__ Cand exchange_best(const ShardInfo & H, int kind, const Cand & mine,
                                                  unsigned long long seq, Cand * scratch, int * sh_ok)
    {
        __shared__ unsigned int words[MAX_WORLD][6];
        const unsigned int seq32 = static_cast<unsigned int>(seq);
        const int t = threadIdx.x;
        if (t < H.world * 6)
        {
            const int r = t / 6, k = t - 6 * r;
            if (blockIdx.x == 0)
            {
                const unsigned long long kb = static_cast<unsigned long long>(__double_as_longlong(mine.key));
                const unsigned long long ab = static_cast<unsigned long long>(__double_as_longlong(mine.aux));
                unsigned int data;
                switch (k)
                {
                    case 0: data = static_cast<unsigned int>(kb); break;
                    case 1: data = static_cast<unsigned int>(kb >> 32); break;
                    case 2: data = static_cast<unsigned int>(ab); break;
                    case 3: data = static_cast<unsigned int>(ab >> 32); break;
                    case 4: data = static_cast<unsigned int>(mine.idx); break;
                    default: data = static_cast<unsigned int>(mine.cls); break;
                }
                WindowHeader * dst = win_hdr(H.win[r]);
                Mail * mb = (kind == 0 ? dst->price : dst->ratio) + H.rank;
                st_relaxed_sys(&mb->w[k], (static_cast<unsigned long long>(seq32) << 32) | data);
            }
            WindowHeader * own = win_hdr(H.win[H.rank]);
            const Mail * mb = (kind == 0 ? own->price : own->ratio) + r;
            const unsigned long long t0 = now_ns();
            unsigned int polls = 0;
            for (;;)
            {
                const unsigned long long v = ld_relaxed_sys(&mb->w[k]);
                if (static_cast<unsigned int>(v >> 32) == seq32)
                {
                    words[r][k] = static_cast<unsigned int>(v);
                    break;
                }
                if ((++polls & 255u) == 0u && must_leave(H, t0))
                {
                    *sh_ok = 0;
                    words[r][k] = 0;
                    break;
                }
            }
        }
        __syncthreads();
        Cand c = cand_none();
        if (t < H.world && *sh_ok != 0)
        {
            const unsigned long long kb = (static_cast<unsigned long long>(words[t][1]) << 32) | words[t][0];
            const unsigned long long ab = (static_cast<unsigned long long>(words[t][3]) << 32) | words[t][2];
            c.key = __longlong_as_double(static_cast<long long>(kb));
            c.aux = __longlong_as_double(static_cast<long long>(ab));
            c.idx = static_cast<int>(words[t][4]);
            c.cls = static_cast<int>(words[t][5]);
        }
        Cand r = block_reduce_cand(c, scratch);
        if (threadIdx.x == 0) scratch[32] = r;
        __syncthreads();
        r = scratch[32];
        return r;
    }

    __device__ __forceinline__ void st_words2_sys(unsigned long long * p, unsigned long long a, unsigned long long b)
    {
        asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
    }
    __device__ __forceinline__ void ld_words2_sys(const unsigned long long * p, unsigned long long & a, unsigned long long & b)
    {
        asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
    }

    // SPECULATIVE vector broadcast, one peer per CTA: as soon as a rank knows ITS OWN candidate (before the
    // arg-min over the ranks is known) it copies the candidate's vector -- its column of A for the pricing
    // exchange, its row of B^-1 for the ratio exchange -- into slot [parity][this rank] of every peer's window.
    // The copy flies while the candidates of the other ranks are still in flight; whoever wins, its vector is
    // (about to be) in everybody's window when the winner is known, so neither broadcast waits for an arg-min.
    // The vector travels as self-validating 8-byte words like the candidates (low / high half of each double
    // beside the 32-bit sequence number of the iteration): no fence, no flag, the pusher goes on at once.
    // CTA 0 posts the candidate itself and never pushes; CTA 1 + c serves peer c. All threads of the CTA call it.
    __device__ __forceinline__ void push_vector(const ShardInfo & H, const double * src, int base, int parity,
                                                unsigned long long seq, int ld)
    {
        const unsigned long long tag = (seq & 0xffffffffULL) << 32;
        const int G = gridDim.x;
        for (int c = 0; c < H.world - 1; ++c)
        {
            if ((c + 1) % G != static_cast<int>(blockIdx.x)) continue;
            const int peer = (c < H.rank) ? c : c + 1;
            unsigned long long * dst = reinterpret_cast<unsigned long long *>(win_vec(H.win[peer], win_slot(base, parity, H.rank), ld));
            for (int k = 2 * threadIdx.x; k < ld; k += 2 * blockDim.x)
            {
                const double2 v = ld_keep2(src + k);
                const unsigned long long a = static_cast<unsigned long long>(__double_as_longlong(v.x));
                const unsigned long long b = static_cast<unsigned long long>(__double_as_longlong(v.y));
                st_words2_sys(dst + 2 * k, tag | (a & 0xffffffffULL), tag | (a >> 32));
                st_words2_sys(dst + 2 * k + 2, tag | (b & 0xffffffffULL), tag | (b >> 32));
            }
        }
    }

    // Receiver side: all threads of the CTA decode slot [parity][src] of the LOCAL window into shared memory,
    // polling every word until it carries this iteration's sequence number. *sh_ok = 0 when the run is abandoned.
    __device__ __forceinline__ void stage_pushed_vector(const ShardInfo & H, double * dst_smem, int base, int parity, int src,
                                                        unsigned long long seq, int ld, int * sh_ok)
    {
        const unsigned int seq32 = static_cast<unsigned int>(seq);
        const unsigned long long * w =
            reinterpret_cast<const unsigned long long *>(win_vec(H.win[H.rank], win_slot(base, parity, src), ld));
        for (int k = 2 * threadIdx.x; k < ld; k += 2 * blockDim.x)
        {
            unsigned long long a0, a1, b0, b1;
            const unsigned long long t0 = now_ns();
            unsigned int polls = 0;
            for (;;)
            {
                ld_words2_sys(w + 2 * k, a0, a1);
                ld_words2_sys(w + 2 * k + 2, b0, b1);
                if (static_cast<unsigned int>(a0 >> 32) == seq32 && static_cast<unsigned int>(a1 >> 32) == seq32 &&
                    static_cast<unsigned int>(b0 >> 32) == seq32 && static_cast<unsigned int>(b1 >> 32) == seq32)
                    break;
                if ((++polls & 255u) == 0u && must_leave(H, t0))
                {
                    *sh_ok = 0;
                    break;
                }
            }
            double2 v;
            v.x = __longlong_as_double(static_cast<long long>((a1 << 32) | (a0 & 0xffffffffULL)));
            v.y = __longlong_as_double(static_cast<long long>((b1 << 32) | (b0 & 0xffffffffULL)));
            *reinterpret_cast<double2 *>(dst_smem + k) = v;
        }
    }

    // One thread. Position `pos` of the non-basic list has changed hands (a pivot put another column there): the list
    // of the positions this rank owns follows. cnt = the count before the change.
    __device__ __forceinline__ void apply_owner_change(const ShardInfo & H, int pos, bool was_mine, bool is_mine, int cnt)
    {
        if (was_mine && !is_mine)
        {
            const int slot = H.pos_slot[pos];
            const int last = H.own_pos[cnt - 1];
            H.own_pos[slot] = last;
            H.pos_slot[last] = slot;
            H.pos_slot[pos] = -1;
        }
        else if (!was_mine && is_mine)
        {
            H.own_pos[cnt] = pos;
            H.pos_slot[pos] = cnt;
        }
    }

    // K0, sharded: for every basis row i whose column basic[i] lives on this rank, check that the
    // column is a multiple of e_i and record a_ii (0 for columns of other ranks: the host adds the
    // ranks' vectors up). One warp per basis row.
    __global__ void k_shard_diag(DevState S, ShardInfo H, double * diag, int * not_unit)
    {
        const int lane = threadIdx.x & 31;
        const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
        if (i >= S.m) return;
        const int bcol = S.basic[i];
        if (bcol % H.world != H.rank)
        {
            if (lane == 0) diag[i] = 0.0;
            return;
        }
        const int lcol = bcol / H.world;
        const double * col = S.A + static_cast<size_t>(lcol) * S.ld;
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
            if (bad) atomicOr(not_unit, 1);
        }
    }

    // Hygiene, sharded: out[i] = sum over the basis positions k whose column this rank owns of A[i][basic[k]] * x[k]
    // (x = the whole x_B, gathered by the host; the host adds the ranks' vectors up and compares with b).
    __global__ void k_shard_residual_partial(DevState S, ShardInfo H, const double * x_full, double * out)
    {
        const int i = blockIdx.x * blockDim.x + threadIdx.x;
        if (i >= S.m) return;
        double acc = 0.0;
        for (int k = 0; k < S.m; ++k)
        {
            const int col = S.basic[k];
            if (col % H.world == H.rank) acc = fma(S.A[static_cast<size_t>(col / H.world) * S.ld + i], x_full[k], acc);
        }
        out[i] = acc;
    }

    // W[i][j] = AB[j][i] (AB column-major m x m: column j = basic column j, gathered from its owner), V = I:
    // the start of the Gauss-Jordan inversion every rank runs on the whole basis (k_gj_pivot / k_gj_elim).
    __global__ void k_shard_load_basis(int m, int ld, const double * AB, double * W, double * V)
    {
        const int j = blockIdx.x * blockDim.x + threadIdx.x;
        const int i = blockIdx.y;
        if (j >= ld) return;
        W[static_cast<size_t>(i) * ld + j] = (j < m) ? AB[static_cast<size_t>(j) * m + i] : 0.0;
        V[static_cast<size_t>(i) * ld + j] = (i == j) ? 1.0 : 0.0;
    }

    // out[i] = M[i,:] . v for the nrows LOCAL rows of a row-major M with rows of length m (x_B = B^-1 b on a row
    // shard; v has ld entries, zero padded). Warp per row, v staged in shared memory, warp_dot's summation order.
    __global__ void __launch_bounds__(FTRAN_THREADS) k_shard_rows_dot(int nrows, int ld, const double * M, const double * v,
                                                                      double * out)
    {
        extern __shared__ __align__(16) double smem[];
        for (int k = threadIdx.x; k < ld; k += blockDim.x) smem[k] = v[k];
        __syncthreads();
        const int lane = threadIdx.x & 31;
        const int wpb = blockDim.x >> 5;
        for (int i = blockIdx.x * wpb + (threadIdx.x >> 5); i < nrows; i += gridDim.x * wpb)
        {
            const double r = warp_dot<false>(M + static_cast<size_t>(i) * ld, smem, ld, lane);
            if (lane == 0) out[i] = r;
        }
    }

    // Local rows of the diagonal start inverse and of x_B = B^-1 b (same expression as the
    // single-GPU path: (1/a_ii) * b_i).
    __global__ void k_shard_init_rows(DevState S, ShardInfo H, const double * diag)
    {
        const int li = blockIdx.x * blockDim.x + threadIdx.x;
        if (li >= H.nrows) return;
        const int gi = H.row0 + li;
        const double inv = 1.0 / diag[gi];
        S.Binv[static_cast<size_t>(li) * S.ld + gi] = inv;
        S.xB[li] = fma(inv, S.b[gi], 0.0);
    }

    // The fused pass over this CTA's rows with kWpr warps per row (fused_row_share): update k-1, FTRAN k, ratio
    // candidates. Returns this warp's best candidate in lane 0 of the warps that finish a row.
    template <int kWpr, bool kBounded>
    __device__ __forceinline__ Cand sharded_pass(const DevState & S, const ShardInfo & H, int lr0, int lr1, bool pending,
                                                 int p_prev, const double * psm, const double * aq, const double * dsm,
                                                 double * partial /* global: [local row][3][32] */, int warp, int wpb, int lane,
                                                 unsigned long long pol_keep, unsigned long long pol_stream, int keep_end,
                                                 int qneg = 0, int q_prev = -1)
    {
        constexpr int kCount = 4 / kWpr;
        Cand best = cand_none();
        const int ngroups = wpb / kWpr;
        const int grp = warp / kWpr;
        const int sub = warp - grp * kWpr;
        const int rounds = (lr1 - lr0 + ngroups - 1) / ngroups;
        for (int r = 0; r < rounds; ++r)
        {
            const int li = lr0 + grp + r * ngroups;
            const bool active = li < lr1;
            const int gi = H.row0 + li;
            double part = 0.0;
            if (active)
            {
                double * row = S.Binv + static_cast<size_t>(li) * S.ld;
                const unsigned long long pol = (li < keep_end) ? pol_keep : pol_stream;
                if (pending)
                    part = fused_row_share<true, kCount>(row, psm, aq, dsm[li - lr0], gi == p_prev, S.ld, lane, sub * kCount, pol);
                else
                    part = fused_row_share<false, kCount>(row, psm, aq, 0.0, false, S.ld, lane, sub * kCount, pol);
            }
            if (kWpr > 1)
            {
                // the partials cross the warps through L2 (no shared memory to spare, see sharded_smem_bytes): one
                // slot per row and pass, written before the barrier, read behind it past the (stale) L1
                double * buf = partial + static_cast<size_t>(active ? li : 0) * ROW_PARTIAL_DOUBLES;
                if (active && sub != 0) buf[(sub - 1) * 32 + lane] = part;
                __syncthreads();  // every warp runs the same number of rounds
                if (sub == 0 && active)
                {
                    if (kWpr == 2)
                        part = part + __ldcg(buf + lane);
                    else
                        part = (part + __ldcg(buf + lane)) + (__ldcg(buf + 32 + lane) + __ldcg(buf + 64 + lane));
                }
            }
            if (sub == 0 && active)
            {
                double di = warp_sum(part);
                if (kBounded && qneg) di = -di;  // B^-1 (-a_q) = -(B^-1 a_q) exactly
                if (lane == 0)
                {
                    S.d[li] = di;
                    const double xi = S.xB[li];
                    if (di > S.eps)
                    {
                        const double t = xi / di;
                        if (t < DBL_MAX)
                        {
                            Cand c;
                            c.key = t;
                            c.aux = di;
                            c.idx = gi;
                            c.cls = 0;
                            if (cand_better(c, best)) best = c;
                        }
                    }
                    else if (kBounded && di < -S.eps)  // Simplex.cpp:136-153
                    {
                        const int bi = (gi == p_prev) ? q_prev : S.basic[gi];  // (the last swap may be in flight)
                        const double t = (S.ub[bi] - xi) / (-di);
                        if (t < DBL_MAX)
                        {
                            Cand c;
                            c.key = t;
                            c.aux = di;
                            c.idx = gi;
                            c.cls = 1;
                            if (cand_better(c, best)) best = c;
                        }
                    }
                }
            }
        }
        return best;
    }

    // Exit epilogue of the bounded variant (see apply_pending_negations in sb200_fused.cuh): the flagged columns this
    // rank owns are rewritten, the (replicated) costs of ALL flagged columns change sign, the flags go back to zero.
    __device__ __forceinline__ void shard_apply_pending_negations(const DevState & S, const ShardInfo & H, int npcol, int npneg,
                                                                  bool neg_dirty)
    {
        const int lane = threadIdx.x & 31;
        const int gw = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
        const int nw = gridDim.x * (blockDim.x >> 5);
        for (int base = gw * 32; base < S.N; base += nw * 32)
        {
            const int col = base + lane;
            int eff = 0;
            if (col < S.N)
            {
                const int stored = S.neg[col];
                eff = (neg_dirty && col == npcol) ? npneg : stored;
                if (stored) S.neg[col] = 0;
            }
            unsigned todo = __ballot_sync(0xffffffffu, eff != 0);
            while (todo)
            {
                const int l = __ffs(todo) - 1;
                todo &= todo - 1;
                const int cc = base + l;
                if (cc % H.world == H.rank)
                {
                    double * acol = S.A + static_cast<size_t>(cc / H.world) * S.ld;
                    for (int k = 2 * lane; k < S.ld; k += 64)
                    {
                        double2 v = ld_keep2(acol + k);
                        v.x = -v.x;
                        v.y = -v.y;
                        st2(acol + k, v);
                    }
                }
                if (lane == 0) S.c[cc] = -S.c[cc];
            }
        }
    }

    template <bool kBounded>
    __device__ __forceinline__ void sharded_loop(const DevState & S, const ShardInfo & H, long long chunk)
    {
        extern __shared__ __align__(16) double smem[];
        __shared__ Cand scratch[33];
        __shared__ __align__(8) unsigned long long mbar;
        __shared__ int sh_ok;
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
        // local rows of this CTA, as LOCAL indices into S.Binv / S.xB
        const int lr0 = static_cast<int>((static_cast<long long>(cta) * H.nrows) / G);
        const int lr1 = static_cast<int>((static_cast<long long>(cta + 1) * H.nrows) / G);
        double * partial = S.row_partials;  // lane partials of rows shared by several warps (global memory)
        const int wpr = warps_per_row(lr1 - lr0, wpb);
        // the CTA's first keep_rows rows of B^-1 stay in L2 from pivot to pivot (see fused_loop)
        const unsigned long long pol_keep = l2_policy_evict_last();
        const unsigned long long pol_stream = l2_policy_evict_first();
        const int keep_end = lr0 + S.keep_rows;
        const unsigned vec_bytes = static_cast<unsigned>(S.ld * sizeof(double));
        unsigned char * mywin = H.win[H.rank];
        unsigned int epoch = 0;
        unsigned int mphase = 0;
        bool ok = true;

        if (sc->status != TERM_RUNNING) return;
        if (threadIdx.x == 0)
        {
            mbar_init(&mbar, 1);
            sh_ok = 1;
        }
        const long long iters_at_entry = sc->iterations;
        phase_init(sc, S.trace_cap);  // also: pivot counter and trace cursor of CTA 0, kept in shared memory
        const long long limit = sc->iter_limit;
        int exit_status = TERM_RUNNING;

        // ---- entry: chained dual solve. Rank r adds its block partials (ascending) onto the carry
        // of rank r-1; the last rank broadcasts y.
        {
            const int nblk_loc = (H.nrows + DUAL_BLOCK_ROWS - 1) / DUAL_BLOCK_ROWS;
            dual_block_partials(S, S.Binv, H.row0, H.nrows, S.ypart);
            ok = grid_barrier_or_leave(sc, epoch, H, &sh_ok);
            const double * carry = nullptr;
            if (ok && H.rank > 0)
            {
                if (threadIdx.x == 0 && !wait_flag(H, &win_hdr(mywin)->carry_flag, H.launch_seq)) sh_ok = 0;
                __syncthreads();
                carry = win_vec(mywin, WIN_VEC_CARRY, S.ld);
            }
            if (ok) dual_block_chain(S, S.ypart, nblk_loc, carry, S.y);  // S.y = running sum up to this rank
            if (ok) ok = grid_barrier_or_leave(sc, epoch, H, &sh_ok);
            if (ok && cta == 0)
            {
                if (H.rank + 1 < H.world)
                {
                    double * dst = win_vec(H.win[H.rank + 1], WIN_VEC_CARRY, S.ld);
                    for (int k = 2 * threadIdx.x; k < S.ld; k += 2 * blockDim.x) st2(dst + k, ld_keep2(S.y + k));
                    __threadfence_system();
                    __syncthreads();
                    if (threadIdx.x == 0) st_release_sys(&win_hdr(H.win[H.rank + 1])->carry_flag, H.launch_seq);
                }
                else
                {
                    for (int peer = 0; peer < H.world; ++peer)
                    {
                        double * dst = win_vec(H.win[peer], WIN_VEC_Y, S.ld);
                        for (int k = 2 * threadIdx.x; k < S.ld; k += 2 * blockDim.x) st2(dst + k, ld_keep2(S.y + k));
                    }
                    __threadfence_system();
                    __syncthreads();
                    if (threadIdx.x < H.world) st_release_sys(&win_hdr(H.win[threadIdx.x])->y_flag, H.launch_seq);
                }
            }
            if (ok)
            {
                if (threadIdx.x == 0)
                {
                    if (wait_flag(H, &win_hdr(mywin)->y_flag, H.launch_seq))
                        bulk_load_issue(ysm, win_vec(mywin, WIN_VEC_Y, S.ld), vec_bytes, &mbar);
                    else
                    {
                        sh_ok = 0;
                        bulk_load_issue(ysm, S.y, vec_bytes, &mbar);  // keep the mbarrier phases in step
                    }
                }
                mbar_wait(&mbar, mphase & 1);
                mphase++;
                __syncthreads();
                if (cta == 0)
                    for (int k = threadIdx.x; k < S.ld; k += blockDim.x) S.y[k] = ysm[k];
                ok = sh_ok != 0;
            }
        }

        bool pending = false;
        int p_prev = -1, e_prev = -1, q_prev = -1, lc_prev = -1;
        bool lists_applied = true;
        long long local_iters = iters_at_entry;
        long long t_last = clock64();
        int own_cnt = *H.own_cnt;                 // every CTA keeps the count in step (same arithmetic everywhere)
        // kBounded: the column whose sign flag CTA 0 publishes one barrier late, and its new flag (fused_loop<true>)
        int npcol = -1, npneg = 0;
        bool neg_dirty = false;
        long long flips_local = 0;
        unsigned long long ticket_base = 0;       // first ticket of the current sweep's pool (sc->ticket starts at 0)

        for (long long it = 0; it < chunk && ok; ++it)
        {
            const unsigned long long seq = static_cast<unsigned long long>(local_iters) + 1ULL;  // global iteration id
            const int parity = static_cast<int>(seq & 1ULL);
            // ================= price (local columns only)
            {
                Cand best = cand_none();
                const double neg_eps = -S.eps;
                // The pivot decided last has not reached the global lists yet (CTA 0 applies it behind barrier 1):
                // position e_prev now holds lc_prev. If that changed the position's owner, the list of owned positions
                // is patched locally for this one sweep: the entry is skipped, or one unit is added at the end.
                const bool stale = !lists_applied;
                const bool was_mine = stale && (q_prev % H.world) == H.rank;
                const bool is_mine = stale && (lc_prev % H.world) == H.rank;
                const int skip_pos = (was_mine && !is_mine) ? e_prev : -1;
                const int n_units = own_cnt + ((is_mine && !was_mine) ? 1 : 0);
                auto price_unit = [&](int u) {
                    const int pos = (u < own_cnt) ? H.own_pos[u] : e_prev;
                    if (pos == skip_pos) return;
                    const int col = (stale && pos == e_prev) ? lc_prev : S.nonbasic[pos];
                    const int lcol = col / H.world;
                    const double dot = warp_dot<true>(S.A + static_cast<size_t>(lcol) * S.ld, ysm, S.ld, lane);
                    double s = S.c[col] - dot;
                    // a substituted column counts with the opposite sign until the exit of the launch rewrites it
                    if (kBounded && ((col == npcol) ? npneg : S.neg[col])) s = -s;
                    if (s < neg_eps)
                    {
                        Cand c;
                        c.key = (S.rule == RULE_MOST_NEG) ? s : 0.0;
                        c.aux = s;
                        c.idx = pos;
                        c.cls = 0;
                        if (cand_better(c, best)) best = c;
                    }
                };
                // Most of the units (whole waves) are strided statically over the grid's warps; the rest is
                // a shared pool that warps drain through ONE global ticket counter once their static share
                // is done, so SMs that see more bandwidth take more columns and the sweep ends together.
                // Every warp draws until its first ticket beyond the pool: an iteration consumes exactly
                // pool + W tickets, so iteration `it` of this launch owns [ticket_base, ticket_base + pool + W).
                const int W = G * wpb;
                // (S.price_mode = 0: no pool; k > 0: the last 1/2^k of the units, default k = 3)
                const int n_static = (S.price_mode == 0) ? n_units
                                                         : static_cast<int>(((static_cast<long long>(n_units) - (n_units >> S.price_mode)) / W) * W);
                for (int u = cta * wpb + warp; u < n_static; u += W) price_unit(u);
                const unsigned long long pool = static_cast<unsigned long long>(n_units - n_static);
                if (pool > 0)
                {
                    unsigned long long t = 0;
                    if (lane == 0) t = atomicAdd(&sc->ticket, 1ULL) - ticket_base;
                    t = __shfl_sync(0xffffffffu, t, 0);
                    while (t < pool)
                    {
                        unsigned long long tn = 0;
                        if (lane == 0) tn = atomicAdd(&sc->ticket, 1ULL) - ticket_base;  // drawn early: hides behind the column
                        price_unit(n_static + static_cast<int>(t));
                        t = __shfl_sync(0xffffffffu, tn, 0);
                    }
                    ticket_base += pool + static_cast<unsigned long long>(W);  // (the pool size changes with own_cnt)
                }
                if (lane != 0) best = cand_none();
                best = block_reduce_cand(best, scratch);
                if (threadIdx.x == 0) S.price_slots[cta] = best;
            }
            phase_tick(sc, 0, t_last);
            ok = grid_barrier_or_leave(sc, epoch, H, &sh_ok);  // 1 (local)
            if (!ok) break;
            phase_tick(sc, 5, t_last);  // slot 5: wait for this rank's slowest CTA at the end of the pricing sweep
            const Cand emine = block_reduce_slots(S.price_slots, G, scratch);
            // this rank's candidate column goes to every peer NOW, beside the candidate itself
            int q_mine = -1;
            if (emine.idx >= 0)
            {
                q_mine = (!lists_applied && emine.idx == e_prev) ? lc_prev : S.nonbasic[emine.idx];
                push_vector(H, S.A + static_cast<size_t>(q_mine / H.world) * S.ld, WIN_VEC_AQ, parity, seq, S.ld);
            }
            const Cand ebest = exchange_best(H, 0, emine, seq, scratch, &sh_ok);  // NVLink arg-min
            phase_tick(sc, 1, t_last);
            ok = sh_ok != 0;
            if (!ok) break;

            local_iters += 1;
            int q = -1;
            if (ebest.idx >= 0) q = (!lists_applied && ebest.idx == e_prev) ? lc_prev : S.nonbasic[ebest.idx];
            int qneg = 0;
            if (kBounded && q >= 0) qneg = (q == npcol) ? npneg : S.neg[q];
            __syncthreads();
            if (kBounded && neg_dirty)
            {
                if (cta == 0 && threadIdx.x == 0) S.neg[npcol] = static_cast<unsigned char>(npneg);
                neg_dirty = false;  // the patch (npcol, npneg) stays: it agrees with the published flag from now on
            }
            if (!lists_applied)
            {
                // the swap of pivot k-1 reaches the global lists now (everybody is past the sweep that still needed
                // the old ones), and with it the list of owned positions
                const bool was_mine = (q_prev % H.world) == H.rank, is_mine = (lc_prev % H.world) == H.rank;
                if (cta == 0 && threadIdx.x == 0)
                {
                    S.basic[p_prev] = q_prev;
                    S.nonbasic[e_prev] = lc_prev;
                    apply_owner_change(H, e_prev, was_mine, is_mine, own_cnt);
                }
                own_cnt += (is_mine ? 1 : 0) - (was_mine ? 1 : 0);
            }
            if (cta == 0 && threadIdx.x == 0)
            {
                sc->iterations = local_iters;
                sc->action = ACT_NONE;
                sc->e_pos = ebest.idx;
                if (ebest.idx >= 0)
                {
                    sc->q_col = q;
                    sc->s_q = ebest.aux;
                }
                else
                {
                    sc->last_ret = -1;
                }
            }
            lists_applied = true;
            if (ebest.idx < 0)
            {
                exit_status = (local_iters >= limit) ? TERM_ITER_LIMIT : TERM_OPTIMAL;
                break;
            }
            const int e = ebest.idx;
            const double s_q = ebest.aux;

            // ================= entering column: from the own shard, or from the slot its owner pushed it to
            const int q_owner = q % H.world;
            if (q_owner == H.rank)
            {
                if (threadIdx.x == 0) bulk_load_issue(aq, S.A + static_cast<size_t>(q / H.world) * S.ld, vec_bytes, &mbar);
                mbar_wait(&mbar, mphase & 1);
                mphase++;
            }
            else
            {
                stage_pushed_vector(H, aq, WIN_VEC_AQ, parity, q_owner, seq, S.ld, &sh_ok);
                __syncthreads();
            }

            // ================= fused: rank-1 update k-1 + FTRAN k + ratio candidates (local rows)
            {
                Cand best;
                if (wpr == 4)
                    best = sharded_pass<4, kBounded>(S, H, lr0, lr1, pending, p_prev, psm, aq, dsm, partial, warp, wpb, lane, pol_keep, pol_stream, keep_end, qneg, q_prev);
                else if (wpr == 2)
                    best = sharded_pass<2, kBounded>(S, H, lr0, lr1, pending, p_prev, psm, aq, dsm, partial, warp, wpb, lane, pol_keep, pol_stream, keep_end, qneg, q_prev);
                else
                    best = sharded_pass<1, kBounded>(S, H, lr0, lr1, pending, p_prev, psm, aq, dsm, partial, warp, wpb, lane, pol_keep, pol_stream, keep_end, qneg, q_prev);
                if (lane != 0) best = cand_none();
                fence_async_smem();
                best = block_reduce_cand(best, scratch);
                if (threadIdx.x == 0) S.ratio_slots[cta] = best;
                // DRAM idles from here to the next pricing sweep (two exchanges, pivot-row broadcast): pull
                // the owned column of this warp's first pricing group towards L2. The group that holds
                // position e is skipped (its new column is not known yet).
                if (lane == 0 && warp < S.prefetch_cols)
                {
                    const int u = cta * wpb + warp;
                    if (u < own_cnt)
                    {
                        // (a hint only: the list may be mid-update by CTA 0, so the column is checked to be ours)
                        const int ppos = H.own_pos[u];
                        const int pcol = (ppos >= 0 && ppos < S.nnb && ppos != e && ppos != e_prev) ? S.nonbasic[ppos] : -1;
                        if (pcol >= 0 && (pcol % H.world) == H.rank)
                            prefetch_l2_bulk(S.A + static_cast<size_t>(pcol / H.world) * S.ld, vec_bytes);
                    }
                }
            }
            pending = false;
            phase_tick(sc, 2, t_last);
            ok = grid_barrier_or_leave(sc, epoch, H, &sh_ok);  // 2 (local)
            if (!ok) break;
            phase_tick(sc, 6, t_last);  // slot 6: wait for this rank's slowest CTA at the end of the fused pass
            const Cand rmine = block_reduce_slots(S.ratio_slots, G, scratch);
            // this rank's candidate row of the UPDATED B^-1 goes to every peer now (unscaled: the pivot element
            // travels with the candidate, every rank scales its copy itself)
            if (rmine.idx >= 0)
                push_vector(H, S.Binv + static_cast<size_t>(rmine.idx - H.row0) * S.ld, WIN_VEC_ROW, parity, seq, S.ld);
            const Cand rbest = exchange_best(H, 1, rmine, seq, scratch, &sh_ok);  // NVLink arg-min
            phase_tick(sc, 3, t_last);
            ok = sh_ok != 0;
            if (!ok) break;
            if (kBounded)
            {
                const double min_theta = (rbest.idx >= 0) ? rbest.key : DBL_MAX;
                const double ub_s = S.ub[q];
                if (ub_s < min_theta)
                {
                    // Simplex.cpp:160-164: the entering variable reaches its own bound first. No basis change: B^-1, y and
                    // the lists stay; x_B -= u d on the own rows; b (replicated) -= a_q u on EVERY rank, rows split over the
                    // CTAs (a_q is staged on every rank); the column and its cost change sign (flag), the flag toggles.
                    for (int li = lr0 + threadIdx.x; li < lr1; li += blockDim.x) S.xB[li] = fma(-ub_s, S.d[li], S.xB[li]);
                    const int b0 = static_cast<int>((static_cast<long long>(cta) * S.m) / G);
                    const int b1 = static_cast<int>((static_cast<long long>(cta + 1) * S.m) / G);
                    for (int i = b0 + threadIdx.x; i < b1; i += blockDim.x)
                    {
                        const double a = qneg ? -aq[i] : aq[i];
                        S.b[i] = __dsub_rn(S.b[i], __dmul_rn(a, ub_s));
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
                    fence_async_smem();
                    __syncthreads();
                    phase_tick(sc, 4, t_last);
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
                exit_status = TERM_UNBOUNDED;
                if (cta == 0 && threadIdx.x == 0) sc->last_ret = -1;
                break;
            }
            const int p = rbest.idx;  // GLOBAL row
            const double dp = rbest.aux;  // d_p as computed from the stored row p
            const double theta = rbest.key;
            if (kBounded && rbest.cls == 1)
            {
                // Simplex.cpp:165-176: the leaving variable sits at its upper bound and is substituted first. b -= a_k u
                // needs its column on every rank: the owner pushes it now (a third, rare exchange), everybody stages it
                // where the entering column was (no longer needed) and updates its replica of b.
                const int lck = S.basic[p];
                const double u = S.ub[lck];
                const int lneg = (lck == npcol) ? npneg : S.neg[lck];
                const int lc_owner = lck % H.world;
                __syncthreads();  // everybody is done with a_q
                if (lc_owner == H.rank)
                {
                    const double * src = S.A + static_cast<size_t>(lck / H.world) * S.ld;
                    push_vector(H, src, WIN_VEC_LC, parity, seq, S.ld);
                    for (int k = 2 * threadIdx.x; k < S.ld; k += 2 * blockDim.x)
                        *reinterpret_cast<double2 *>(aq + k) = ld_keep2(src + k);
                }
                else
                    stage_pushed_vector(H, aq, WIN_VEC_LC, parity, lc_owner, seq, S.ld, &sh_ok);
                __syncthreads();
                const int b0 = static_cast<int>((static_cast<long long>(cta) * S.m) / G);
                const int b1 = static_cast<int>((static_cast<long long>(cta + 1) * S.m) / G);
                for (int i = b0 + threadIdx.x; i < b1; i += blockDim.x)
                {
                    const double a = lneg ? -aq[i] : aq[i];
                    S.b[i] = __dsub_rn(S.b[i], __dmul_rn(a, u));
                }
                npcol = lck;
                npneg = lneg ^ 1;
                neg_dirty = true;
                if (cta == 0 && threadIdx.x == 0) S.flip[lck] ^= 1;
            }

            // ================= pivot row: from the own rows, or from the slot its owner pushed it to
            const int p_owner = p / H.rows_per_rank;
            if (p_owner == H.rank && threadIdx.x == 0)
                bulk_load_issue(psm, S.Binv + static_cast<size_t>(p - H.row0) * S.ld, vec_bytes, &mbar);
            for (int li = lr0 + threadIdx.x; li < lr1; li += blockDim.x)
            {
                const double di = S.d[li];
                dsm[li - lr0] = di;
                S.xB[li] = (H.row0 + li == p) ? theta : fma(-theta, di, S.xB[li]);
                if (kBounded && rbest.cls == 1 && H.row0 + li == p) S.d[li] = -di;  // read-out: the pivot element after the substitution
            }
            const int lc = S.basic[p];
            if (p_owner == H.rank)
            {
                mbar_wait(&mbar, mphase & 1);
                mphase++;
            }
            else
            {
                stage_pushed_vector(H, psm, WIN_VEC_ROW, parity, p_owner, seq, S.ld, &sh_ok);
                __syncthreads();
            }
            for (int j = 2 * threadIdx.x; j < S.ld; j += 2 * blockDim.x)
            {
                double2 r = *reinterpret_cast<double2 *>(psm + j);
                r.x = r.x / dp;
                r.y = r.y / dp;
                *reinterpret_cast<double2 *>(psm + j) = r;
                double2 yv = *reinterpret_cast<double2 *>(ysm + j);
                yv.x = fma(s_q, r.x, yv.x);
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
            ok = sh_ok != 0;
            phase_tick(sc, 4, t_last);
            phase_count_iteration();
            if (local_iters >= limit)
            {
                exit_status = TERM_ITER_LIMIT;
                break;
            }
        }

        if (!ok) exit_status = TERM_EXCHANGE_TIMEOUT;
        // Every CTA has staged the last pivot row (read from global B^-1) only once it is past this
        // barrier; without it the owner's flush below could overwrite the row under a late reader.
        // An abandoned run takes no barrier any more (see raise_abort) and leaves B^-1 as it is.
        if (ok) ok = grid_barrier_or_leave(sc, epoch, H, &sh_ok);
        if (ok && pending)
        {
            for (int li = lr0 + warp; li < lr1; li += wpb)
                flush_row(S.Binv + static_cast<size_t>(li) * S.ld, psm, dsm[li - lr0], H.row0 + li == p_prev, S.ld, lane);
        }
        if (!ok) exit_status = TERM_EXCHANGE_TIMEOUT;
        if (kBounded && ok) shard_apply_pending_negations(S, H, npcol, npneg, neg_dirty);
        if (cta == 0)
        {
            __syncthreads();
            for (int j = threadIdx.x; j < S.ld; j += blockDim.x) S.y[j] = ysm[j];
            if (threadIdx.x == 0)
            {
                if (ok && !lists_applied)
                {
                    const bool was_mine = (q_prev % H.world) == H.rank, is_mine = (lc_prev % H.world) == H.rank;
                    S.basic[p_prev] = q_prev;
                    S.nonbasic[e_prev] = lc_prev;
                    apply_owner_change(H, e_prev, was_mine, is_mine, own_cnt);
                    own_cnt += (is_mine ? 1 : 0) - (was_mine ? 1 : 0);
                }
                *H.own_cnt = own_cnt;
                if (kBounded) sc->flips += flips_local;
                if (exit_status != TERM_RUNNING) sc->status = exit_status;
                phase_flush(sc);
            }
        }
    }

    __global__ void __launch_bounds__(SHARDED_THREADS, 1) k_sharded_fused(DevState S, ShardInfo H, long long chunk)
    {
        sharded_loop<false>(S, H, chunk);
    }

    // the same loop for LPs with finite upper bounds (sb200_shard_set_bounds)
    __global__ void __launch_bounds__(SHARDED_THREADS, 1) k_sharded_fused_bounded(DevState S, ShardInfo H, long long chunk)
    {
        sharded_loop<true>(S, H, chunk);
    }
}  // namespace sb200

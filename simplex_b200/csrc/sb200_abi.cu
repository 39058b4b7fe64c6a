// C ABI of the B200 revised-simplex hot path (include/simplex_b200.h). Host side of the
// library: owns device memory, builds the per-iteration CUDA graph (staged engine) or launches
// the cooperative persistent kernel, polls a device status word between chunks.
//
// There is no CPU implementation in this file: without a CUDA device every compute entry
// point returns SB200_ERR_CUDA.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/simplex_b200.h"
#include "sb200_kernels.cuh"
#include "sb200_persistent.cuh"
#include "sb200_fused.cuh"
#include "sb200_resident.cuh"
#include "sb200_sharded.cuh"
#include "sb200_batched.cuh"
#include "sb200_batched_fast.cuh"

using namespace sb200;

namespace
{
    std::string g_create_error;

    inline int round_up(int64_t v, int q) { return static_cast<int>((v + q - 1) / q * q); }
}  // namespace

struct sb200_ctx
{
    sb200_opts opts{};
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    std::string err;
    int num_sms = 0;
    double sm_mhz = 0.0;  // cudaDevAttrClockRate: the device's nominal SM clock (debug printers only; a power-capped run is slower)
    int max_smem_optin = 0;
    int max_smem_per_sm = 0;
    bool coop_supported = false;

    DevState S{};
    bool loaded = false;
    bool prepared = false;
    int64_t N_loaded = 0;  // columns allocated (sb200_set_costs may shrink S.N)

    std::vector<void *> allocs;
    // Guard zones: every device buffer of the context sits between two 256-byte zones filled with a pattern;
    // sb200_debug_check_guards counts the zones a kernel has written into (compute-sanitizer is not available on
    // every pool: this is the library's own out-of-bounds-write check, run over the whole GPU test suite).
    struct Guarded
    {
        unsigned char * base;
        size_t payload;
    };
    std::vector<Guarded> guarded;
    double * W = nullptr;  // inversion scratch (lazy)
    double * V = nullptr;
    double * fcol = nullptr;
    double * shard_AB = nullptr;  // sharded mode: the whole basis matrix, column-major (sb200_shard_invert)
    double * shard_ub = nullptr;  // sharded mode: upper bounds (sb200_shard_set_bounds)
    double * diag = nullptr;
    int * flags = nullptr;  // [0] not_diag, [1] singular
    int * unit_row = nullptr;  // [N] row of the only nonzero of a unit column, -1 otherwise
    double * unit_val = nullptr;
    // shared-memory-resident engine (sb200_resident.cuh)
    int * dense_ids = nullptr;   // [N] dense rank -> column id
    int * dense_rank = nullptr;  // [N] column id -> dense rank, -1 for unit columns
    std::vector<int> h_dense_ids;
    unsigned long long * res_mail = nullptr;  // price mail, ratio mail, staged rows (res_mail_words)
    size_t res_mail_count = 0;
    int res_max_dynamic = 0;   // dynamic shared memory k_resident may ask for
    int res_max_dynamic_bounded = 0;  // ... and k_resident_bounded
    long long * fused_dbg = nullptr;  // [num_sms][8] per-CTA phase cycles of the last fused launch (tuning, SB200_FUSED_DEBUG)
    // fused engine: shares of the pricing positions / of the rows of B^-1 per CTA, tuned launch after launch
    // from the per-CTA phase times (sum 1 each; empty = equal shares), and their integer bounds on the device
    std::vector<double> tune_pos_w, tune_row_w;
    int * tune_bounds = nullptr;  // [2][num_sms + 1]
    long long * res_dbg = nullptr;  // [num_sms][8] per-CTA phase cycles of the last launch (SB200_RES_DEBUG)
    unsigned int res_seq = 1;  // sequence number of the next pivot: never 0, never reused while the mail is dirty
    int engine_used = SB200_ENGINE_USED_NONE;

    Scalars * h_sc = nullptr;  // pinned mirror
    cudaGraphExec_t graph_exec = nullptr;
    cudaGraph_t graph = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int64_t launches = 0;
    int64_t launches_per_chunk = 0;
    int64_t reinversions = 0;
    int64_t launches_in_run = 0;
    // numerical hygiene (sb200_opts.hygiene_*): scratch and the statistics of the last run
    unsigned long long * hyg_out = nullptr;  // [4] bit patterns: max residual, max |b|, max inverse defect
    int64_t hyg_checks = 0, hyg_refreshes = 0, hyg_counter = 0;
    double hyg_worst = 0.0;
    // The basis list B^-1 and x_B on the device are valid for (as the last run left them); empty = none. A call that
    // starts from exactly that basis (Phase II after Phase I with A resident, src/InitialBasis.cpp:85-121) carries
    // them over instead of inverting the basis again.
    std::vector<int64_t> carry_basic;
    bool carried = false;  // the last prepare carried B^-1 over
    bool units_stale = false;  // a kernel that rewrites columns without maintaining the unit table has run
    std::string l2_note;

    // sharded mode
    bool sharded = false;
    bool connected = false;
    ShardInfo H{};
    int shard_grid = 0;  // CTAs of the sharded kernel (all SMs; SMs / ranks when several ranks share one device)
    unsigned char * window = nullptr;
    std::vector<void *> peer_maps;
    unsigned long long launch_seq = 0;

    void * batch_buffers = nullptr;  // BatchBuffers of sb200_run_batched (persistent, grow-only)

    // launch shapes
    int price_grid = 1, ftran_grid = 1, rows_per_seg = 1;
    int persistent_grid = 0;
    size_t vec_smem = 0;
};

namespace
{
    int fail(sb200_ctx * ctx, int rc, const std::string & msg)
    {
        if (ctx)
            ctx->err = msg;
        else
            g_create_error = msg;
        return rc;
    }

#define CU_TRY(ctx, expr)                                                                             \
    do                                                                                                \
    {                                                                                                 \
        cudaError_t e__ = (expr);                                                                     \
        if (e__ != cudaSuccess)                                                                       \
        {                                                                                             \
            return fail(ctx, (e__ == cudaErrorMemoryAllocation) ? SB200_ERR_NOMEM : SB200_ERR_CUDA,   \
                        std::string(#expr) + ": " + cudaGetErrorString(e__));                         \
        }                                                                                             \
    } while (0)

    void free_all(sb200_ctx * ctx)
    {
        for (void * p : ctx->peer_maps) cudaIpcCloseMemHandle(p);
        ctx->peer_maps.clear();
        ctx->window = nullptr;
        ctx->sharded = ctx->connected = false;
        for (void * p : ctx->allocs) cudaFree(p);
        ctx->allocs.clear();
        ctx->guarded.clear();
        ctx->W = ctx->V = ctx->fcol = ctx->diag = ctx->shard_AB = ctx->shard_ub = nullptr;
        ctx->flags = nullptr;
        ctx->hyg_out = nullptr;
        ctx->dense_ids = ctx->dense_rank = nullptr;
        ctx->h_dense_ids.clear();
        ctx->res_mail = nullptr;
        ctx->res_dbg = nullptr;
        ctx->fused_dbg = nullptr;
        ctx->tune_bounds = nullptr;
        ctx->tune_pos_w.clear();
        ctx->tune_row_w.clear();
        ctx->res_mail_count = 0;
        ctx->res_seq = 1;
        if (ctx->graph_exec) cudaGraphExecDestroy(ctx->graph_exec);
        if (ctx->graph) cudaGraphDestroy(ctx->graph);
        ctx->graph_exec = nullptr;
        ctx->graph = nullptr;
        ctx->loaded = ctx->prepared = false;
    }

    constexpr size_t GUARD_BYTES = 256;
    constexpr int GUARD_PATTERN = 0xA5;

    template <class T>
    cudaError_t dalloc(sb200_ctx * ctx, T ** p, size_t count)
    {
        // [256 B guard][payload, rounded up to 256 B][256 B guard]
        void * q = nullptr;
        const size_t payload = (count * sizeof(T) + 255) & ~size_t(255);
        cudaError_t e = cudaMalloc(&q, payload + 2 * GUARD_BYTES);
        if (e != cudaSuccess) return e;
        ctx->allocs.push_back(q);
        unsigned char * base = static_cast<unsigned char *>(q);
        e = cudaMemsetAsync(base, GUARD_PATTERN, GUARD_BYTES, ctx->stream);
        if (e == cudaSuccess) e = cudaMemsetAsync(base + GUARD_BYTES + payload, GUARD_PATTERN, GUARD_BYTES, ctx->stream);
        if (e != cudaSuccess) return e;
        ctx->guarded.push_back({ base, payload });
        *p = reinterpret_cast<T *>(base + GUARD_BYTES);
        return cudaSuccess;
    }

    int allocate_state(sb200_ctx * ctx, int64_t m, int64_t N, bool has_ub)
    {
        if (ctx->loaded && !ctx->sharded && ctx->S.m == m && ctx->N_loaded == N && (ctx->S.ub != nullptr) == has_ub)
        {
            // same shape as the resident problem: keep the allocations (a re-load per step must
            // not pay cudaMalloc/cudaFree), only reset what a previous run may have left behind
            DevState & S = ctx->S;
            S.N = static_cast<int>(N);
            S.nnb = static_cast<int>(N - m);
            const int wpb0 = PRICE_THREADS / 32;
            ctx->price_grid = std::max(1, std::min(2 * ctx->num_sms, (S.nnb + wpb0 - 1) / wpb0));
            S.n_price_slots = ctx->price_grid;
            if (S.ld != S.m)
                CU_TRY(ctx, cudaMemsetAsync(S.A, 0, static_cast<size_t>(S.ld) * N * sizeof(double), ctx->stream));
            CU_TRY(ctx, cudaMemsetAsync(S.b, 0, static_cast<size_t>(S.ld) * sizeof(double), ctx->stream));
            CU_TRY(ctx, cudaMemsetAsync(S.neg, 0, static_cast<size_t>(N), ctx->stream));
            ctx->prepared = false;
            return SB200_OK;
        }
        free_all(ctx);
        if (m <= 0 || N < m || N > (int64_t(1) << 30) || m > (int64_t(1) << 20))
            return fail(ctx, SB200_ERR_INVALID, "sb200_load: need 0 < m <= N (m <= 2^20, N <= 2^30)");
        DevState & S = ctx->S;
        S = DevState{};
        S.m = static_cast<int>(m);
        S.N = static_cast<int>(N);
        S.nnb = static_cast<int>(N - m);
        S.ld = round_up(m, 16);
        S.eps = ctx->opts.eps;
        S.rule = ctx->opts.pricing;
        S.dual_update = (ctx->opts.dual_mode == SB200_DUAL_UPDATE) ? 1 : 0;
        ctx->N_loaded = N;
        const size_t ld = S.ld;
        if (ld * sizeof(double) > static_cast<size_t>(ctx->max_smem_optin) - 4096)
            return fail(ctx, SB200_ERR_UNSUPPORTED, "m too large for the shared-memory staged vectors");
        CU_TRY(ctx, dalloc(ctx, &S.A, ld * N));
        CU_TRY(ctx, dalloc(ctx, &S.Binv, ld * m));
        CU_TRY(ctx, dalloc(ctx, &S.b, ld));
        CU_TRY(ctx, dalloc(ctx, &S.c, static_cast<size_t>(N)));
        if (has_ub) CU_TRY(ctx, dalloc(ctx, &S.ub, static_cast<size_t>(N)));
        CU_TRY(ctx, dalloc(ctx, &S.xB, ld));
        CU_TRY(ctx, dalloc(ctx, &S.y, ld));
        CU_TRY(ctx, dalloc(ctx, &S.d, ld));
        CU_TRY(ctx, dalloc(ctx, &S.prow, ld));
        CU_TRY(ctx, dalloc(ctx, &S.basic, static_cast<size_t>(m)));
        CU_TRY(ctx, dalloc(ctx, &S.nonbasic, static_cast<size_t>(N - m) + 1));
        CU_TRY(ctx, dalloc(ctx, &S.flip, static_cast<size_t>(N)));
        CU_TRY(ctx, dalloc(ctx, &S.neg, static_cast<size_t>(N)));
        CU_TRY(ctx, dalloc(ctx, &S.sc, 1));
        CU_TRY(ctx, dalloc(ctx, &ctx->diag, ld));
        CU_TRY(ctx, dalloc(ctx, &ctx->fcol, ld));
        CU_TRY(ctx, dalloc(ctx, &ctx->flags, 4));
        CU_TRY(ctx, dalloc(ctx, &ctx->hyg_out, 4));
        CU_TRY(ctx, dalloc(ctx, &ctx->unit_row, static_cast<size_t>(N)));
        CU_TRY(ctx, dalloc(ctx, &ctx->unit_val, static_cast<size_t>(N)));
        CU_TRY(ctx, dalloc(ctx, &ctx->dense_ids, static_cast<size_t>(N)));
        CU_TRY(ctx, dalloc(ctx, &ctx->dense_rank, static_cast<size_t>(N)));
        ctx->res_mail_count = res_mail_words(static_cast<int>(ld), ctx->num_sms);
        CU_TRY(ctx, dalloc(ctx, &ctx->res_mail, ctx->res_mail_count));
        CU_TRY(ctx, cudaMemsetAsync(ctx->res_mail, 0, ctx->res_mail_count * sizeof(unsigned long long), ctx->stream));
        ctx->res_seq = 1;
        if (const char * s0 = std::getenv("SB200_RES_SEQ0"))  // tests: start close to the wrap of the 32-bit sequence numbers
            ctx->res_seq = std::max(1u, static_cast<unsigned int>(std::strtoul(s0, nullptr, 0)));
        CU_TRY(ctx, dalloc(ctx, &ctx->fused_dbg, 8 * static_cast<size_t>(ctx->num_sms)));
        CU_TRY(ctx, dalloc(ctx, &ctx->tune_bounds, 2 * (static_cast<size_t>(ctx->num_sms) + 1)));
        if (std::getenv("SB200_RES_DEBUG")) CU_TRY(ctx, dalloc(ctx, &ctx->res_dbg, 16 * static_cast<size_t>(ctx->num_sms)));

        // launch shapes
        const int sms = ctx->num_sms;
        ctx->vec_smem = ld * sizeof(double);
        const int wpb = PRICE_THREADS / 32;
        ctx->price_grid = std::max(1, std::min(2 * sms, (S.nnb + wpb - 1) / wpb));
        ctx->ftran_grid = std::max(1, std::min(2 * sms, (S.m + wpb - 1) / wpb));
        S.nseg = std::max(1, std::min(64, (S.m + 7) / 8));
        ctx->rows_per_seg = (S.m + S.nseg - 1) / S.nseg;
        S.nseg = (S.m + ctx->rows_per_seg - 1) / ctx->rows_per_seg;
        ctx->persistent_grid = sms;
        const int max_slots = std::max(std::max(ctx->price_grid, ctx->ftran_grid), ctx->persistent_grid);
        S.n_price_slots = ctx->price_grid;
        S.n_ratio_slots = ctx->ftran_grid;
        CU_TRY(ctx, dalloc(ctx, &S.price_slots, static_cast<size_t>(max_slots)));
        CU_TRY(ctx, dalloc(ctx, &S.ratio_slots, static_cast<size_t>(max_slots)));
        const int ypart_rows =
            std::max(std::max(S.nseg, ctx->persistent_grid), (S.m + DUAL_BLOCK_ROWS - 1) / DUAL_BLOCK_ROWS);
        CU_TRY(ctx, dalloc(ctx, &S.ypart, static_cast<size_t>(ypart_rows) * ld));
        S.trace_cap = ctx->opts.trace_capacity;
        if (S.trace_cap > 0) CU_TRY(ctx, dalloc(ctx, &S.trace, static_cast<size_t>(S.trace_cap)));

        cudaStream_t st = ctx->stream;
        CU_TRY(ctx, cudaMemsetAsync(S.A, 0, ld * N * sizeof(double), st));
        CU_TRY(ctx, cudaMemsetAsync(S.b, 0, ld * sizeof(double), st));
        CU_TRY(ctx, cudaMemsetAsync(S.xB, 0, ld * sizeof(double), st));
        CU_TRY(ctx, cudaMemsetAsync(S.y, 0, ld * sizeof(double), st));
        CU_TRY(ctx, cudaMemsetAsync(S.d, 0, ld * sizeof(double), st));
        CU_TRY(ctx, cudaMemsetAsync(S.prow, 0, ld * sizeof(double), st));
        CU_TRY(ctx, cudaMemsetAsync(S.flip, 0, static_cast<size_t>(N), st));
        CU_TRY(ctx, cudaMemsetAsync(S.neg, 0, static_cast<size_t>(N), st));
        CU_TRY(ctx, cudaMemsetAsync(S.sc, 0, sizeof(Scalars), st));
        return SB200_OK;
    }

    int classify_units(sb200_ctx * ctx)
    {
        DevState & S = ctx->S;
        const long long blocks = (static_cast<long long>(ctx->N_loaded) + 7) / 8;  // one warp per column, 8 warps per block
        DevState all = S;
        all.N = static_cast<int>(ctx->N_loaded);  // every loaded column, also those sb200_set_costs has dropped
        k_classify_units<<<static_cast<unsigned>(blocks), 256, 0, ctx->stream>>>(all, ctx->unit_row, ctx->unit_val);
        ctx->launches += 1;
        S.unit_row = ctx->unit_row;
        S.unit_val = ctx->unit_val;
        S.unit_val_rw = ctx->unit_val;
        ctx->units_stale = false;
        CU_TRY(ctx, cudaGetLastError());
        return SB200_OK;
    }

    int load_common(sb200_ctx * ctx, int64_t m, int64_t N, const double * A, int64_t lda, const double * b,
                    const double * c, const double * ub, cudaMemcpyKind kind)
    {
        if (!ctx) return SB200_ERR_INVALID;
        if (!A || !b || !c || lda < m) return fail(ctx, SB200_ERR_INVALID, "sb200_load: null pointer or lda < m");
        CU_TRY(ctx, cudaSetDevice(ctx->opts.device));
        ctx->carry_basic.clear();
        int rc = allocate_state(ctx, m, N, ub != nullptr);
        if (rc != SB200_OK) return rc;
        DevState & S = ctx->S;
        cudaStream_t st = ctx->stream;
        CU_TRY(ctx, cudaMemcpy2DAsync(S.A, static_cast<size_t>(S.ld) * sizeof(double), A, lda * sizeof(double),
                                      m * sizeof(double), N, kind, st));
        CU_TRY(ctx, cudaMemcpyAsync(S.b, b, m * sizeof(double), kind, st));
        CU_TRY(ctx, cudaMemcpyAsync(S.c, c, N * sizeof(double), kind, st));
        if (ub) CU_TRY(ctx, cudaMemcpyAsync(S.ub, ub, N * sizeof(double), kind, st));
        // unit columns (slacks, artificials): classified once per load. Only the engines that never rewrite
        // a column inside a launch use the table (fused, resident); the bounded fused kernel keeps it in step
        // with the columns it rewrites at its exit, other engines mark it stale (classify_units again).
        if (std::getenv("SB200_NO_UNIT_COLUMNS") == nullptr)
        {
            rc = classify_units(ctx);
            if (rc != SB200_OK) return rc;
        }
        else
        {
            S.unit_row = nullptr;
            S.unit_val = nullptr;
            S.unit_val_rw = nullptr;
        }
        CU_TRY(ctx, cudaStreamSynchronize(st));
        ctx->h_dense_ids.clear();
        if (S.unit_row)
        {
            // dense rank <-> column id tables of the shared-memory-resident engine (a few KB; ascending
            // ids, so the dense columns among the first N' columns are a prefix: sb200_set_costs)
            std::vector<int> urow(static_cast<size_t>(N)), rank(static_cast<size_t>(N));
            CU_TRY(ctx, cudaMemcpy(urow.data(), ctx->unit_row, N * sizeof(int), cudaMemcpyDeviceToHost));
            for (int64_t j = 0; j < N; ++j)
            {
                if (urow[j] < 0)
                {
                    rank[j] = static_cast<int>(ctx->h_dense_ids.size());
                    ctx->h_dense_ids.push_back(static_cast<int>(j));
                }
                else
                    rank[j] = -1;
            }
            CU_TRY(ctx, cudaMemcpy(ctx->dense_rank, rank.data(), N * sizeof(int), cudaMemcpyHostToDevice));
            if (!ctx->h_dense_ids.empty())
                CU_TRY(ctx, cudaMemcpy(ctx->dense_ids, ctx->h_dense_ids.data(), ctx->h_dense_ids.size() * sizeof(int),
                                       cudaMemcpyHostToDevice));
        }
        ctx->loaded = true;
        return SB200_OK;
    }

    // ------------------------------------------------------------------ staged engine
    void launch_dual(sb200_ctx * ctx)
    {
        const DevState & S = ctx->S;
        dim3 g1((S.ld + 2 * UPDATE_THREADS - 1) / (2 * UPDATE_THREADS), S.nseg);
        k_dual_partial<<<g1, UPDATE_THREADS, 0, ctx->stream>>>(S, ctx->rows_per_seg);
        k_dual_reduce<<<(S.ld + 255) / 256, 256, 0, ctx->stream>>>(S);
    }
    void launch_price(sb200_ctx * ctx)
    {
        k_price<<<ctx->price_grid, PRICE_THREADS, ctx->vec_smem, ctx->stream>>>(ctx->S);
    }
    void launch_ftran(sb200_ctx * ctx, bool pick)
    {
        if (pick)
            k_ftran<true><<<ctx->ftran_grid, FTRAN_THREADS, ctx->vec_smem, ctx->stream>>>(ctx->S);
        else
            k_ftran<false><<<ctx->ftran_grid, FTRAN_THREADS, ctx->vec_smem, ctx->stream>>>(ctx->S);
    }
    void launch_select(sb200_ctx * ctx) { k_select<<<1, SELECT_THREADS, 0, ctx->stream>>>(ctx->S); }
    void launch_update(sb200_ctx * ctx)
    {
        const DevState & S = ctx->S;
        dim3 g((S.ld + 2 * UPDATE_THREADS - 1) / (2 * UPDATE_THREADS), (S.m + UPDATE_ROWS - 1) / UPDATE_ROWS);
        k_update<<<g, UPDATE_THREADS, 0, ctx->stream>>>(S);
    }

    int build_graph(sb200_ctx * ctx)
    {
        if (ctx->graph_exec) cudaGraphExecDestroy(ctx->graph_exec);
        if (ctx->graph) cudaGraphDestroy(ctx->graph);
        ctx->graph_exec = nullptr;
        ctx->graph = nullptr;
        const int chunk = static_cast<int>(std::max<int64_t>(1, ctx->opts.chunk_iters));
        const bool recompute = ctx->opts.dual_mode == SB200_DUAL_RECOMPUTE;
        const int refresh = ctx->opts.dual_refresh;
        CU_TRY(ctx, cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
        int64_t n = 0;
        for (int it = 0; it < chunk; ++it)
        {
            if (recompute || it == 0 || (refresh > 0 && it % refresh == 0))
            {
                launch_dual(ctx);
                n += 2;
            }
            launch_price(ctx);
            launch_ftran(ctx, true);
            launch_select(ctx);
            launch_update(ctx);
            n += 4;
        }
        cudaError_t e = cudaStreamEndCapture(ctx->stream, &ctx->graph);
        if (e != cudaSuccess) return fail(ctx, SB200_ERR_CUDA, std::string("graph capture: ") + cudaGetErrorString(e));
        CU_TRY(ctx, cudaGraphInstantiate(&ctx->graph_exec, ctx->graph, 0));
        ctx->launches_per_chunk = n;
        return SB200_OK;
    }

    int set_func_attrs(sb200_ctx * ctx)
    {
        const int optin = ctx->max_smem_optin;
        CU_TRY(ctx, cudaFuncSetAttribute(k_price, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - 2048));
        CU_TRY(ctx, cudaFuncSetAttribute(k_ftran<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - 2048));
        CU_TRY(ctx, cudaFuncSetAttribute(k_ftran<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - 2048));
        CU_TRY(ctx, cudaFuncSetAttribute(k_rows_dot, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - 2048));
        CU_TRY(ctx, cudaFuncSetAttribute(k_inverse_residual, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - 2048));
        CU_TRY(ctx, cudaFuncSetAttribute(k_persistent, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - 2048));
        CU_TRY(ctx, cudaFuncSetAttribute(k_persistent_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - 2048));
        CU_TRY(ctx, cudaFuncSetAttribute(k_persistent_fused_bounded, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - 2048));
        {
            // k_resident has a few KB of static shared memory (mail decode buffers): what is left is dynamic
            cudaFuncAttributes fa{};
            CU_TRY(ctx, cudaFuncGetAttributes(&fa, k_resident));
            ctx->res_max_dynamic = optin - static_cast<int>(fa.sharedSizeBytes);
            CU_TRY(ctx, cudaFuncSetAttribute(k_resident, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx->res_max_dynamic));
            cudaFuncAttributes fb{};
            CU_TRY(ctx, cudaFuncGetAttributes(&fb, k_resident_bounded));
            ctx->res_max_dynamic_bounded = optin - static_cast<int>(fb.sharedSizeBytes);
            CU_TRY(ctx, cudaFuncSetAttribute(k_resident_bounded, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             ctx->res_max_dynamic_bounded));
            // (the stream-columns variants have the same static shared memory as their on-chip twins)
            CU_TRY(ctx, cudaFuncSetAttribute(k_resident_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx->res_max_dynamic));
            CU_TRY(ctx, cudaFuncSetAttribute(k_resident_stream_bounded, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             ctx->res_max_dynamic_bounded));
        }
        CU_TRY(ctx, cudaFuncSetAttribute(k_sharded_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - 2048));
        CU_TRY(ctx, cudaFuncSetAttribute(k_sharded_fused_bounded, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - 2048));
        CU_TRY(ctx, cudaFuncSetAttribute(k_shard_rows_dot, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - 2048));
        CU_TRY(ctx, cudaFuncSetAttribute(k_batched, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - 2048));
        return SB200_OK;
    }

    int invert_basis_general(sb200_ctx * ctx)
    {
        DevState & S = ctx->S;
        const size_t ld = S.ld, m = S.m;
        if (!ctx->W)
        {
            CU_TRY(ctx, dalloc(ctx, &ctx->W, ld * m));
            CU_TRY(ctx, dalloc(ctx, &ctx->V, ld * m));
        }
        cudaStream_t st = ctx->stream;
        dim3 gg((S.ld + 255) / 256, S.m);
        k_gather_basis<<<gg, 256, 0, st>>>(S, ctx->W, ctx->V);
        dim3 ge((S.ld + 2 * UPDATE_THREADS - 1) / (2 * UPDATE_THREADS), (S.m + UPDATE_ROWS - 1) / UPDATE_ROWS);
        for (int k = 0; k < S.m; ++k)
        {
            k_gj_pivot<<<1, 1024, 0, st>>>(S.m, S.ld, k, ctx->W, ctx->V, ctx->fcol, ctx->flags + 1);
            k_gj_elim<<<ge, UPDATE_THREADS, 0, st>>>(S.m, S.ld, k, ctx->W, ctx->V, ctx->fcol, ctx->flags + 1);
        }
        ctx->launches += 1 + 2 * static_cast<int64_t>(S.m);
        int singular = 0;
        CU_TRY(ctx, cudaMemcpyAsync(&singular, ctx->flags + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
        CU_TRY(ctx, cudaStreamSynchronize(st));
        if (singular) return fail(ctx, SB200_ERR_SINGULAR, "basis matrix is singular");
        CU_TRY(ctx, cudaMemcpyAsync(S.Binv, ctx->V, ld * m * sizeof(double), cudaMemcpyDeviceToDevice, st));
        return SB200_OK;
    }

    // Periodic re-inversion (sb200_opts.reinvert_period): B^-1 from the resident A_B by Gauss-Jordan
    // with partial pivoting, x_B = B^-1 b. The duals follow at the next launch (both persistent
    // kernels start with a full dual solve; the staged graph starts with K1).
    int refresh_inverse(sb200_ctx * ctx)
    {
        int rc = invert_basis_general(ctx);
        if (rc != SB200_OK) return rc;
        DevState & S = ctx->S;
        k_rows_dot<<<ctx->ftran_grid, FTRAN_THREADS, ctx->vec_smem, ctx->stream>>>(S.m, S.ld, S.Binv, S.b, S.xB);
        ctx->launches += 1;
        ctx->reinversions += 1;
        return SB200_OK;
    }

    // r_primal = max_i |b_i - (A_B x_B)_i|, b_max = max_i |b_i|, r_inverse = max over `rows` sampled rows of
    // |e_i^T B^-1 A_B - e_i^T|_inf (rows == 0: not measured). S.ypart is free between two launches.
    int hygiene_measure(sb200_ctx * ctx, int rows, int first, double * r_primal, double * b_max, double * r_inverse)
    {
        const DevState & S = ctx->S;
        cudaStream_t st = ctx->stream;
        CU_TRY(ctx, cudaMemsetAsync(ctx->hyg_out, 0, 4 * sizeof(unsigned long long), st));
        const dim3 gp((S.m + 127) / 128, HYG_SEGS);
        k_residual_partial<<<gp, 128, 0, st>>>(S, S.ypart);
        k_residual_reduce<<<(S.m + 127) / 128, 128, 0, st>>>(S, S.ypart, nullptr, ctx->hyg_out);
        ctx->launches += 2;
        rows = std::max(0, std::min(std::min(rows, 64), S.m));
        if (rows > 0)
        {
            const int wpb = FTRAN_THREADS / 32;
            const int stride = std::max(1, S.m / rows);
            const dim3 gi(std::max(1, std::min(2 * ctx->num_sms, (S.m + wpb - 1) / wpb)), rows);
            k_inverse_residual<<<gi, FTRAN_THREADS, ctx->vec_smem, st>>>(S, ((first % S.m) + S.m) % S.m, stride, ctx->hyg_out);
            ctx->launches += 1;
        }
        unsigned long long h[4] = { 0, 0, 0, 0 };
        CU_TRY(ctx, cudaMemcpyAsync(h, ctx->hyg_out, sizeof h, cudaMemcpyDeviceToHost, st));
        CU_TRY(ctx, cudaStreamSynchronize(st));
        CU_TRY(ctx, cudaGetLastError());
        double v[3];
        std::memcpy(v, h, sizeof v);
        *r_primal = v[0];
        *b_max = v[1];
        *r_inverse = v[2];
        return SB200_OK;
    }

    // The hygiene gate behind every launch of a loop engine (sb200_opts.hygiene_*). *again: launch once more.
    int hygiene_gate(sb200_ctx * ctx, bool * again, int * ending_refreshes)
    {
        const int status = ctx->h_sc->status;
        *again = status == TERM_RUNNING;
        const double tol = ctx->opts.hygiene_tol;
        if (!(tol > 0.0)) return SB200_OK;
        const bool ending = status == TERM_OPTIMAL || status == TERM_UNBOUNDED;
        const bool periodic = status == TERM_RUNNING && ctx->opts.hygiene_period > 0 &&
                              ctx->launches_in_run % ctx->opts.hygiene_period == 0;
        if (!ending && !periodic) return SB200_OK;
        if (ending && *ending_refreshes >= 2) return SB200_OK;  // two fresh inverses have said the same thing
        double rp = 0.0, bm = 0.0, ri = 0.0;
        ctx->hyg_counter += 1;
        int rc = hygiene_measure(ctx, ctx->opts.hygiene_rows, static_cast<int>((ctx->hyg_counter * 7919) % (1 << 30)), &rp, &bm, &ri);
        if (rc != SB200_OK) return rc;
        const double r = std::max(rp / (1.0 + bm), ri);
        ctx->hyg_checks += 1;
        ctx->hyg_worst = std::max(ctx->hyg_worst, r);
        if (r > tol)
        {
            rc = refresh_inverse(ctx);
            if (rc != SB200_OK) return rc;
            ctx->hyg_refreshes += 1;
            if (ending)
            {
                k_resume<<<1, 1, 0, ctx->stream>>>(ctx->S.sc);
                ctx->launches += 1;
                *ending_refreshes += 1;
                *again = true;
            }
        }
        return SB200_OK;
    }

    int upload_lists(sb200_ctx * ctx, const int64_t * basic, int64_t nb, const int64_t * nonbasic, int64_t nn)
    {
        DevState & S = ctx->S;
        if (nb != S.m) return fail(ctx, SB200_ERR_BASIS_SIZE, "Basis size must be equal to M=" + std::to_string(S.m));
        if (nn != S.nnb)
            return fail(ctx, SB200_ERR_INVALID, "non-basic list size must be N-M=" + std::to_string(S.nnb));
        std::vector<int> hb(nb), hn(nn);
        std::vector<char> seen(S.N, 0);
        for (int64_t i = 0; i < nb; ++i)
        {
            if (basic[i] < 0 || basic[i] >= S.N || seen[basic[i]])
                return fail(ctx, SB200_ERR_INVALID, "basic list: column out of range or repeated");
            seen[basic[i]] = 1;
            hb[i] = static_cast<int>(basic[i]);
        }
        for (int64_t i = 0; i < nn; ++i)
        {
            if (nonbasic[i] < 0 || nonbasic[i] >= S.N || seen[nonbasic[i]])
                return fail(ctx, SB200_ERR_INVALID, "non-basic list: column out of range or repeated");
            seen[nonbasic[i]] = 1;
            hn[i] = static_cast<int>(nonbasic[i]);
        }
        CU_TRY(ctx, cudaMemcpyAsync(S.basic, hb.data(), nb * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        if (nn > 0)
            CU_TRY(ctx, cudaMemcpyAsync(S.nonbasic, hn.data(), nn * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        return SB200_OK;
    }

    // Keep (part of) B^-1 resident in the 126 MB L2 across the phases of a pivot: the pricing
    // sweep streams 4x more bytes of A through L2 than B^-1 occupies and would otherwise evict it
    // before FTRAN and the rank-1 update re-read it. The window marks B^-1 accesses "persisting"
    // for a fraction of its lines that fits the L2 set-aside; A keeps the streaming policy.
    // SB200_L2_PERSIST=<fraction of the device maximum>. MEASURED ON B200 (profiles/README.md): it
    // makes the pivot SLOWER (173 -> 189/194/240 us at 0.5/0.75/1.0: the carve-out starves the
    // streaming sweep and the update phase doubles), so the default is 0 = off.
    int configure_l2(sb200_ctx * ctx)
    {
        const char * env = std::getenv("SB200_L2_PERSIST");
        const double frac = env ? std::atof(env) : 0.0;
        int max_persist = 0, max_window = 0;
        cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, ctx->opts.device);
        cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, ctx->opts.device);
        cudaStreamAttrValue attr{};
        const DevState & S = ctx->S;
        const size_t bytes = static_cast<size_t>(S.ld) * S.m * sizeof(double);
        if (frac <= 0.0 || max_persist <= 0 || max_window <= 0)
        {
            attr.accessPolicyWindow.base_ptr = nullptr;
            attr.accessPolicyWindow.num_bytes = 0;
            attr.accessPolicyWindow.hitRatio = 0.f;
            attr.accessPolicyWindow.hitProp = cudaAccessPropertyNormal;
            attr.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
            cudaStreamSetAttribute(ctx->stream, cudaStreamAttributeAccessPolicyWindow, &attr);
            ctx->l2_note = "off";
            return SB200_OK;
        }
        const size_t setaside = static_cast<size_t>(static_cast<double>(max_persist) * std::min(frac, 1.0));
        CU_TRY(ctx, cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, setaside));
        const size_t window = std::min(bytes, static_cast<size_t>(max_window));
        attr.accessPolicyWindow.base_ptr = S.Binv;
        attr.accessPolicyWindow.num_bytes = window;
        attr.accessPolicyWindow.hitRatio = static_cast<float>(std::min(1.0, static_cast<double>(setaside) / window));
        attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        CU_TRY(ctx, cudaStreamSetAttribute(ctx->stream, cudaStreamAttributeAccessPolicyWindow, &attr));
        char buf[160];
        std::snprintf(buf, sizeof buf, "set-aside %zu B of max %d, window %zu B of max %d, hitRatio %.3f", setaside,
                      max_persist, window, max_window, attr.accessPolicyWindow.hitRatio);
        ctx->l2_note = buf;
        if (std::getenv("SB200_VERBOSE")) std::fprintf(stderr, "[sb200] L2 persistence: %s\n", buf);
        return SB200_OK;
    }

    int prepare_impl(sb200_ctx * ctx, const int64_t * basic, int64_t nb, const int64_t * nonbasic, int64_t nn)
    {
        if (!ctx->loaded) return fail(ctx, SB200_ERR_STATE, "sb200_prepare: nothing loaded");
        if (ctx->sharded)
            return fail(ctx, SB200_ERR_STATE, "this context holds a column shard (sb200_load_shard): use the sb200_shard_* entry points");
        CU_TRY(ctx, cudaSetDevice(ctx->opts.device));
        int rc = upload_lists(ctx, basic, nb, nonbasic, nn);
        if (rc != SB200_OK) return rc;
        DevState & S = ctx->S;
        cudaStream_t st = ctx->stream;
        rc = configure_l2(ctx);
        if (rc != SB200_OK) return rc;
        // Simplex.cpp:271: every call starts with all substitutions cleared
        CU_TRY(ctx, cudaMemsetAsync(S.flip, 0, static_cast<size_t>(ctx->N_loaded), st));
        CU_TRY(ctx, cudaMemsetAsync(ctx->flags, 0, 4 * sizeof(int), st));
        // K0: B^-1 and x_B = B^-1 b (Simplex.cpp:305-306) -- unless the device already holds them for exactly this
        // basis (the run that ended on it left them there; the hygiene gate watches their drift)
        ctx->carried = static_cast<int64_t>(ctx->carry_basic.size()) == nb && std::getenv("SB200_NO_CARRY") == nullptr &&
                       std::equal(ctx->carry_basic.begin(), ctx->carry_basic.end(), basic);
        ctx->carry_basic.clear();
        if (!ctx->carried)
        {
            k_check_diag<<<(S.m * 32 + 255) / 256, 256, 0, st>>>(S, ctx->diag, ctx->flags);
            ctx->launches += 1;
            int not_diag = 0;
            CU_TRY(ctx, cudaMemcpyAsync(&not_diag, ctx->flags, sizeof(int), cudaMemcpyDeviceToHost, st));
            CU_TRY(ctx, cudaStreamSynchronize(st));
            if (!not_diag)
            {
                CU_TRY(ctx, cudaMemsetAsync(S.Binv, 0, static_cast<size_t>(S.ld) * S.m * sizeof(double), st));
                k_set_diag_inverse<<<(S.m + 255) / 256, 256, 0, st>>>(S, ctx->diag);
                ctx->launches += 1;
            }
            else
            {
                rc = invert_basis_general(ctx);
                if (rc != SB200_OK) return rc;
            }
            k_rows_dot<<<ctx->ftran_grid, FTRAN_THREADS, ctx->vec_smem, st>>>(S.m, S.ld, S.Binv, S.b, S.xB);
            ctx->launches += 1;
        }
        Scalars init{};
        init.status = TERM_RUNNING;
        init.e_pos = -1;
        init.iter_limit = ctx->opts.iter_limit;
        CU_TRY(ctx, cudaMemcpyAsync(S.sc, &init, sizeof(Scalars), cudaMemcpyHostToDevice, st));
        // y = c_B^T B^-1 (Simplex.cpp:307) so that a step_price right after prepare is meaningful
        launch_dual(ctx);
        ctx->launches += 2;
        CU_TRY(ctx, cudaStreamSynchronize(st));
        CU_TRY(ctx, cudaGetLastError());
        ctx->prepared = true;
        return SB200_OK;
    }

    int fetch_scalars(sb200_ctx * ctx)
    {
        CU_TRY(ctx, cudaMemcpyAsync(ctx->h_sc, ctx->S.sc, sizeof(Scalars), cudaMemcpyDeviceToHost, ctx->stream));
        CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        return SB200_OK;
    }

    int run_staged(sb200_ctx * ctx)
    {
        int rc = build_graph(ctx);
        if (rc != SB200_OK) return rc;
        ctx->engine_used = SB200_ENGINE_USED_STAGED;
        if (ctx->S.ub != nullptr) ctx->units_stale = true;  // k_select rewrites flipped columns in place
        const long long reinvert = ctx->opts.reinvert_period;
        long long last_reinvert = 0;
        int ending_refreshes = 0;
        ctx->launches_in_run = 0;
        ctx->hyg_checks = ctx->hyg_refreshes = 0;
        ctx->hyg_worst = 0.0;
        CU_TRY(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
        for (;;)
        {
            CU_TRY(ctx, cudaGraphLaunch(ctx->graph_exec, ctx->stream));
            ctx->launches += ctx->launches_per_chunk;
            ctx->launches_in_run += 1;
            rc = fetch_scalars(ctx);
            if (rc != SB200_OK) return rc;
            bool again = false;
            rc = hygiene_gate(ctx, &again, &ending_refreshes);
            if (rc != SB200_OK) return rc;
            if (!again) break;
            if (reinvert > 0 && ctx->h_sc->pivots - last_reinvert >= reinvert)
            {
                last_reinvert = ctx->h_sc->pivots;
                rc = refresh_inverse(ctx);
                if (rc != SB200_OK) return rc;
            }
        }
        CU_TRY(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
        CU_TRY(ctx, cudaEventSynchronize(ctx->ev1));
        return SB200_OK;
    }

    int run_persistent(sb200_ctx * ctx)
    {
        if (!ctx->coop_supported) return fail(ctx, SB200_ERR_UNSUPPORTED, "device lacks cooperative launch");
        DevState S = ctx->S;
        const size_t smem = persistent_smem_bytes(S, ctx->persistent_grid);
        int per_sm = 0;
        CU_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_persistent, PERSIST_THREADS, smem));
        if (per_sm < 1) return fail(ctx, SB200_ERR_UNSUPPORTED, "persistent kernel does not fit on an SM");
        const int grid = ctx->persistent_grid;
        S.n_price_slots = grid;
        S.n_ratio_slots = grid;
        S.nseg = grid;
        long long chunk = std::max<int64_t>(1, ctx->opts.chunk_iters);
        int recompute = ctx->opts.dual_mode == SB200_DUAL_RECOMPUTE ? 1 : 0;
        // Fused update+FTRAN engine: needs the dual recurrence (the duals must be known before
        // B^-1 is touched). LPs with finite upper bounds run its bounded variant (t2 candidates, bound flips).
        // Tuned partition (fused engine, HBM-bound shapes): the SMs do not all see the same DRAM bandwidth and the
        // same ones are slow launch after launch (profiles/README.md, round 1d), so the shares of the pricing
        // sweep and of the rows of B^-1 follow the per-CTA phase times of the previous launches.
        const bool tune_shape = static_cast<size_t>(S.ld) * S.N * sizeof(double) > (size_t(256) << 20);
        const char * tune_env = std::getenv("SB200_TUNE");
        const bool tune = tune_shape && (tune_env ? std::atoi(tune_env) != 0 : false);
        const int own_uniform = persist_max_own_rows(S.m, grid);
        const int own_cap = own_uniform + std::max(2, own_uniform / 8);
        S.tuned_own_rows = tune ? own_cap : 0;
        const size_t smem_fused = fused_smem_bytes(S, grid);
        const bool bounded = S.ub != nullptr;
        const bool fused = !recompute && smem_fused <= static_cast<size_t>(ctx->max_smem_optin) - 2048 &&
                           std::getenv("SB200_NO_FUSED") == nullptr &&
                           !(bounded && std::getenv("SB200_NO_FUSED_BOUNDED") != nullptr);
        if (fused && bounded && ctx->units_stale && S.unit_row != nullptr)
        {
            int crc = classify_units(ctx);
            if (crc != SB200_OK) return crc;
            S = ctx->S;
            S.n_price_slots = grid;
            S.n_ratio_slots = grid;
            S.nseg = grid;
            S.tuned_own_rows = tune ? own_cap : 0;
        }
        if (!fused && bounded) ctx->units_stale = true;  // the 4-phase kernel rewrites flipped columns in place
        {
            const char * pm = std::getenv("SB200_PRICE_MODE");
            // measured on B200 (profiles/README.md): all three modes tie at m=4096 (HBM-bound), the
            // per-CTA tickets win at m=1024 (15.9 vs 17.1 / 18.4 us per pivot)
            S.price_mode = pm ? std::atoi(pm) : 1;
            const char * pc = std::getenv("SB200_PREFETCH_COLS");
            const char * pr = std::getenv("SB200_PREFETCH_ROWS");
            // measured on B200 at m=4096, n=16384 (profiles/README.md): 8 columns + 8 rows per CTA fill part
            // of the DRAM idle time behind the two grid barriers (143.5 -> 137.9 us per pivot); more
            // only moves the wait from the pricing phase to barrier 2. Useless when A lives in L2.
            const bool hbm_bound = static_cast<size_t>(S.ld) * S.N * sizeof(double) > (size_t(256) << 20);
            S.prefetch_cols = pc ? std::atoi(pc) : (hbm_bound ? 8 : 0);
            S.prefetch_rows = pr ? std::atoi(pr) : (hbm_bound ? 8 : 0);
            if (S.price_mode != 1 && S.price_mode != 3) S.prefetch_cols = 0;  // the prefetch assumes per-CTA slices
            // Rows of B^-1 kept in L2 from pivot to pivot (evict-last loads and stores; every other row and the whole
            // pricing sweep evict-first). Measured on B200 at m=4096, n=16384 (profiles/README.md, round 2): about
            // 75 MB of kept rows (16 per CTA) is the optimum, 137.0 -> 130.1 us per pivot together with dropping the
            // row prefetch; more rows thrash. Useless when the tableau lives in L2 anyway.
            const char * kr = std::getenv("SB200_L2_KEEP_ROWS");
            const int own_rows = (S.m + grid - 1) / grid;
            const int keep_auto = static_cast<int>((size_t(75) << 20) / (static_cast<size_t>(S.ld) * sizeof(double) * grid));
            S.keep_rows = kr ? std::max(0, std::atoi(kr)) : (hbm_bound ? std::min(own_rows, keep_auto) : 0);
            if (!pr && S.keep_rows > 0) S.prefetch_rows = 0;
        }
        // Shared-memory-resident engine: same preconditions as the fused one, plus the CTA's share of
        // B^-1 rows and dense columns must fit its shared memory (sb200_resident.cuh).
        ResidentInfo RI{};
        size_t smem_res = 0;
        bool resident = false;
        const void * res_kernel = bounded ? (const void *)k_resident_bounded : (const void *)k_resident;
        // (finite upper bounds: k_resident_bounded; SB200_RESIDENT_BOUNDED=0 keeps such LPs on the fused kernel)
        const char * rb_env = std::getenv("SB200_RESIDENT_BOUNDED");
        const bool resident_bounded_ok = rb_env ? std::atoi(rb_env) != 0 : true;
        if (fused && (!bounded || resident_bounded_ok) && S.unit_row != nullptr && grid <= RES_MAX_GRID &&
            ctx->opts.engine != SB200_ENGINE_STREAMING && std::getenv("SB200_NO_RESIDENT") == nullptr)
        {
            const auto & ids = ctx->h_dense_ids;
            RI.n_dense = static_cast<int>(std::lower_bound(ids.begin(), ids.end(), S.N) - ids.begin());
            RI.max_rows = res_rows_cap(S.m, grid);
            RI.max_cols = res_cols_cap(RI.n_dense, grid);
            RI.max_slice = res_slice_cap(S.nnb, grid);
            RI.dense_ids = ctx->dense_ids;
            RI.dense_rank = ctx->dense_rank;
            {
                const char * tk = std::getenv("SB200_RES_TICKS");
                RI.ticks = tk ? std::atoi(tk) : 1;
            }
            RI.per_cta_cycles = ctx->res_dbg;
            RI.pmail = ctx->res_mail;
            RI.rmail = ctx->res_mail + static_cast<size_t>(grid) * grid * RES_MAIL_STRIDE;
            RI.rowmail = ctx->res_mail + 2 * static_cast<size_t>(grid) * grid * RES_MAIL_STRIDE;
            const size_t res_budget = static_cast<size_t>(bounded ? ctx->res_max_dynamic_bounded : ctx->res_max_dynamic);
            smem_res = resident_smem_bytes(S.ld, RI.max_rows, RI.max_cols, RI.max_slice, bounded ? S.N : 0);
            resident = smem_res <= res_budget;
            // The rows of B^-1 fit but the dense columns do not (1100 < m <= 1776 or so): the columns stay in global
            // memory and are priced out of L2. Measured on B200 against the fused kernel (profiles/README.md, round 2,
            // scripts/stream_cols_sweep.py): 1.4-2.2 x faster while A lives in L2 (29 ... 114 MB, m = 64 ... 1776),
            // 7-13 % beyond (136 / 258 MB), where the fused kernel's HBM streaming is kept instead.
            // SB200_RES_STREAM_COLS=0 switches the variant off, =2 forces it on whenever it fits (tests).
            const char * sc_env = std::getenv("SB200_RES_STREAM_COLS");
            const int sc_mode = sc_env ? std::atoi(sc_env) : 1;
            const bool a_in_l2 = static_cast<size_t>(S.ld) * S.N * sizeof(double) <= (size_t(120) << 20);
            if ((sc_mode == 2) || (sc_mode == 1 && !resident && a_in_l2))
            {
                const size_t smem_stream = resident_smem_bytes(S.ld, RI.max_rows, RI.max_cols, RI.max_slice, bounded ? S.N : 0, true);
                if (smem_stream <= res_budget)
                {
                    smem_res = smem_stream;
                    resident = true;
                    res_kernel = bounded ? (const void *)k_resident_stream_bounded : (const void *)k_resident_stream;
                }
            }
            if (resident)
            {
                int res_per_sm = 0;
                CU_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&res_per_sm, res_kernel, RES_THREADS, smem_res));
                resident = res_per_sm >= 1;
            }
        }
        ctx->engine_used = resident ? SB200_ENGINE_USED_RESIDENT
                                    : (fused ? SB200_ENGINE_USED_FUSED : SB200_ENGINE_USED_PERSISTENT);
        ctx->launches_in_run = 0;
        ctx->hyg_checks = ctx->hyg_refreshes = 0;
        ctx->hyg_worst = 0.0;
        int ending_refreshes = 0;
        bool again = false;
        static_assert(offsetof(Scalars, ticket) == offsetof(Scalars, barrier_count) + sizeof(unsigned long long),
                      "barrier_count and ticket are zeroed together");
        const long long reinvert = ctx->opts.reinvert_period;
        long long last_reinvert = 0;
        CU_TRY(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
        for (;;)
        {
            if (reinvert > 0 && ctx->h_sc->pivots - last_reinvert >= reinvert && ctx->launches_in_run > 0)
            {
                last_reinvert = ctx->h_sc->pivots;
                int irc = refresh_inverse(ctx);
                if (irc != SB200_OK) return irc;
            }
            ctx->launches_in_run += 1;
            CU_TRY(ctx, cudaMemsetAsync(reinterpret_cast<char *>(S.sc) + offsetof(Scalars, barrier_count), 0,
                                        2 * sizeof(unsigned long long), ctx->stream));
            if (resident)
            {
                // every pivot of the context's life gets its own sequence number; before the 32 bits run
                // out the mail is wiped and the count restarts
                long long rchunk = std::min<long long>(chunk, 1LL << 28);
                if (static_cast<unsigned long long>(ctx->res_seq) + static_cast<unsigned long long>(rchunk) > 0xF0000000ULL)
                {
                    CU_TRY(ctx, cudaMemsetAsync(ctx->res_mail, 0, ctx->res_mail_count * sizeof(unsigned long long), ctx->stream));
                    ctx->res_seq = 1;
                }
                RI.seq0 = ctx->res_seq;
                ctx->res_seq += static_cast<unsigned int>(rchunk);
                void * rargs[] = { &S, &RI, &rchunk };
                CU_TRY(ctx, cudaLaunchCooperativeKernel(res_kernel, dim3(grid), dim3(RES_THREADS), rargs, smem_res, ctx->stream));
                ctx->launches += 1;
                if (bounded)
                {
                    // the substitutions of this launch, written into A, c and the unit table (nobody reads A any more)
                    k_apply_negations<<<grid, 256, 0, ctx->stream>>>(S);
                    CU_TRY(ctx, cudaGetLastError());
                    ctx->launches += 1;
                }
                int rrc = fetch_scalars(ctx);
                if (rrc != SB200_OK) return rrc;
                if (ctx->res_dbg)
                {
                    // per-CTA phase times of this launch: min / mean / max over the CTAs, in us per pivot
                    std::vector<long long> pc(16 * static_cast<size_t>(grid));
                    CU_TRY(ctx, cudaMemcpy(pc.data(), ctx->res_dbg, pc.size() * sizeof(long long), cudaMemcpyDeviceToHost));
                    static const char * RES_FINE_NAMES[13] = { "price loops", "price argmin", "exchange 1 wait", "a_q fetch", "ftran (warp 0)",
                                                                "ratio argmin", "ratio push", "row staging", "exchange 2 wait", "row polls + y",
                                                                "x_B + caches", "price push", "rank-1 update" };
                    const double piv = std::max<double>(1.0, static_cast<double>(pc[15]));
                    double total = 0.0;
                    for (int ph = 0; ph < 13; ++ph)
                    {
                        double lo = 1e30, hi = 0.0, sum = 0.0;
                        int argmin = 0, argmax = 0;
                        for (int c = 0; c < grid; ++c)
                        {
                            const double v = static_cast<double>(pc[16 * c + ph]) / piv / ctx->sm_mhz;
                            if (v < lo) { lo = v; argmin = c; }
                            if (v > hi) { hi = v; argmax = c; }
                            sum += v;
                        }
                        total += sum / grid;
                        std::fprintf(stderr, "[resident] %-14s min %.2f (cta %3d)  mean %.2f  max %.2f (cta %3d) us/pivot\n",
                                     RES_FINE_NAMES[ph], lo, argmin, sum / grid, hi, argmax);
                    }
                    std::fprintf(stderr, "[resident] sum of means %.2f us/pivot over %.0f pivots (thread 0 of every CTA, at the nominal SM clock of %.0f MHz)\n",
                                 total, piv, ctx->sm_mhz);
                }
                if (ctx->h_sc->status == TERM_RESIDENT_TIMEOUT)
                    return fail(ctx, SB200_ERR_CUDA, "resident engine: a CTA did not answer within the exchange timeout");
                rrc = hygiene_gate(ctx, &again, &ending_refreshes);
                if (rrc != SB200_OK) return rrc;
                if (!again) break;
                continue;
            }
            if (fused)
            {
                const bool fdbg = std::getenv("SB200_FUSED_DEBUG") != nullptr;
                S.dbg_cycles = (tune || fdbg) ? ctx->fused_dbg : nullptr;
                std::vector<int> hb;
                if (tune)
                {
                    // integer bounds from the current shares (cumulative rounding: monotone, exact totals)
                    if (static_cast<int>(ctx->tune_pos_w.size()) != grid) ctx->tune_pos_w.assign(static_cast<size_t>(grid), 1.0 / grid);
                    if (static_cast<int>(ctx->tune_row_w.size()) != grid) ctx->tune_row_w.assign(static_cast<size_t>(grid), 1.0 / grid);
                    hb.assign(2 * (static_cast<size_t>(grid) + 1), 0);
                    auto fill = [&](const std::vector<double> & w, int total, int * out) {
                        double cum = 0.0;
                        out[0] = 0;
                        for (int c = 0; c < grid; ++c)
                        {
                            cum += w[static_cast<size_t>(c)];
                            int b = static_cast<int>(std::llround(cum * total));
                            b = std::max(out[c], std::min(total, b));
                            out[c + 1] = b;
                        }
                        out[grid] = total;
                    };
                    fill(ctx->tune_pos_w, S.nnb, hb.data());
                    fill(ctx->tune_row_w, S.m, hb.data() + grid + 1);
                    bool rows_ok = true;
                    for (int c = 0; c < grid; ++c) rows_ok = rows_ok && hb[static_cast<size_t>(grid) + 1 + c + 1] - hb[static_cast<size_t>(grid) + 1 + c] <= own_cap;
                    if (!rows_ok)
                    {
                        ctx->tune_row_w.assign(static_cast<size_t>(grid), 1.0 / grid);
                        for (int c = 0; c <= grid; ++c) hb[static_cast<size_t>(grid) + 1 + c] = static_cast<int>((static_cast<long long>(c) * S.m) / grid);
                    }
                    CU_TRY(ctx, cudaMemcpyAsync(ctx->tune_bounds, hb.data(), hb.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
                    S.pos_bounds = ctx->tune_bounds;
                    S.row_bounds = ctx->tune_bounds + grid + 1;
                }
                void * fargs[] = { &S, &chunk };
                CU_TRY(ctx, cudaLaunchCooperativeKernel(bounded ? (void *)k_persistent_fused_bounded : (void *)k_persistent_fused,
                                                        dim3(grid), dim3(FUSED_THREADS), fargs, smem_fused, ctx->stream));
                ctx->launches += 1;
                int frc = fetch_scalars(ctx);
                if (frc != SB200_OK) return frc;
                if (tune)
                {
                    // new shares: what each CTA got through per cycle of its own phase, damped
                    std::vector<long long> pc(8 * static_cast<size_t>(grid));
                    CU_TRY(ctx, cudaMemcpy(pc.data(), ctx->fused_dbg, pc.size() * sizeof(long long), cudaMemcpyDeviceToHost));
                    auto retune = [&](std::vector<double> & w, const int * b, int phase) {
                        std::vector<double> thr(static_cast<size_t>(grid), 0.0);
                        double sum = 0.0;
                        for (int c = 0; c < grid; ++c)
                        {
                            const double units = b[c + 1] - b[c], t = static_cast<double>(pc[8 * static_cast<size_t>(c) + phase]);
                            if (units <= 0.0 || t <= 0.0) return;  // nothing to learn from this launch
                            thr[static_cast<size_t>(c)] = units / t;
                            sum += thr[static_cast<size_t>(c)];
                        }
                        for (int c = 0; c < grid; ++c) w[static_cast<size_t>(c)] = 0.5 * w[static_cast<size_t>(c)] + 0.5 * thr[static_cast<size_t>(c)] / sum;
                    };
                    if (pc[7] >= 16)  // a launch of a few pivots says little
                    {
                        retune(ctx->tune_pos_w, hb.data(), 0);
                        retune(ctx->tune_row_w, hb.data() + grid + 1, 2);
                    }
                }
                if (fdbg)
                {
                    // per-CTA phase times of this launch (us per pivot): spread over the CTAs, and which CTAs
                    // are the slowest / fastest in the pricing sweep (the same ones launch after launch?)
                    std::vector<long long> pc(8 * static_cast<size_t>(grid));
                    CU_TRY(ctx, cudaMemcpy(pc.data(), ctx->fused_dbg, pc.size() * sizeof(long long), cudaMemcpyDeviceToHost));
                    const double piv = std::max<double>(1.0, static_cast<double>(pc[7]));
                    const char * names[5] = { "price", "barrier 1", "fused pass", "barrier 2", "tail" };
                    for (int ph = 0; ph < 5; ++ph)
                    {
                        double lo = 1e30, hi = 0.0, sum = 0.0;
                        for (int c = 0; c < grid; ++c)
                        {
                            const double v = static_cast<double>(pc[8 * c + ph]) / piv / ctx->sm_mhz;
                            lo = std::min(lo, v);
                            hi = std::max(hi, v);
                            sum += v;
                        }
                        std::fprintf(stderr, "[fused] %-10s min %.2f  mean %.2f  max %.2f us/pivot\n", names[ph], lo, sum / grid, hi);
                    }
                    std::vector<int> order(static_cast<size_t>(grid));
                    for (int c = 0; c < grid; ++c) order[static_cast<size_t>(c)] = c;
                    std::sort(order.begin(), order.end(), [&](int a, int b) { return pc[8 * a] < pc[8 * b]; });
                    std::fprintf(stderr, "[fused] fastest pricing CTAs:");
                    for (int k = 0; k < 8 && k < grid; ++k) std::fprintf(stderr, " %d", order[static_cast<size_t>(k)]);
                    std::fprintf(stderr, "   slowest:");
                    for (int k = 0; k < 8 && k < grid; ++k) std::fprintf(stderr, " %d", order[static_cast<size_t>(grid - 1 - k)]);
                    std::fprintf(stderr, "   (%.0f pivots)\n", piv);
                }
                frc = hygiene_gate(ctx, &again, &ending_refreshes);
                if (frc != SB200_OK) return frc;
                if (!again) break;
                continue;
            }
            void * args[] = { &S, &chunk, &recompute };
            CU_TRY(ctx, cudaLaunchCooperativeKernel((void *)k_persistent, dim3(grid), dim3(PERSIST_THREADS), args, smem,
                                                    ctx->stream));
            ctx->launches += 1;
            int rc = fetch_scalars(ctx);
            if (rc != SB200_OK) return rc;
            rc = hygiene_gate(ctx, &again, &ending_refreshes);
            if (rc != SB200_OK) return rc;
            if (!again) break;
        }
        CU_TRY(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
        CU_TRY(ctx, cudaEventSynchronize(ctx->ev1));
        return SB200_OK;
    }
}  // namespace

namespace
{
    // Persistent buffers of the batched mode (grow-only, owned by the context, freed with it): two device
    // buffers of `chunk` LPs each, so that the host->device copy of chunk k+1 runs beside the kernel of chunk k,
    // and two pinned staging buffers for callers whose inputs live in pageable memory.
    struct BatchBuffers
    {
        size_t chunk_lps = 0;
        int m = 0, N = 0;
        double *dA[2] = {}, *db[2] = {}, *dc[2] = {}, *dx[2] = {}, *dobj[2] = {};
        int *dbas[2] = {}, *dnon[2] = {}, *dterm[2] = {}, *dit[2] = {};
        double * dub[2] = {};           // bounded runs only (allocated on first use)
        unsigned char * dflip[2] = {};
        int * dflips[2] = {};
        int * d_pending = nullptr;  // [2]: LPs the fast kernel declined, per slot
        int * h_pending = nullptr;  // pinned mirror
        unsigned char * pinned[2] = {};
        size_t pinned_bytes = 0;
        cudaStream_t copy_stream = nullptr;
        cudaEvent_t h2d_done[2] = {}, slot_free[2] = {};
        std::vector<void *> allocs;
    };

    void free_batch_buffers(BatchBuffers * B)
    {
        if (!B) return;
        for (void * p : B->allocs) cudaFree(p);
        for (int s = 0; s < 2; ++s)
        {
            if (B->pinned[s]) cudaFreeHost(B->pinned[s]);
            if (s == 0 && B->h_pending) cudaFreeHost(B->h_pending);
            if (B->h2d_done[s]) cudaEventDestroy(B->h2d_done[s]);
            if (B->slot_free[s]) cudaEventDestroy(B->slot_free[s]);
        }
        if (B->copy_stream) cudaStreamDestroy(B->copy_stream);
        delete B;
    }

    bool host_pointer_is_pinned(const void * p)
    {
        cudaPointerAttributes at{};
        if (cudaPointerGetAttributes(&at, p) != cudaSuccess)
        {
            cudaGetLastError();
            return false;
        }
        return at.type == cudaMemoryTypeHost;
    }

    // memcpy with a few threads (one thread tops out near 10 GB/s, PCIe 5 x16 moves 50)
    void parallel_memcpy(void * dst, const void * src, size_t bytes)
    {
        const int nthreads = bytes > (size_t(32) << 20) ? 4 : 1;
        if (nthreads == 1)
        {
            std::memcpy(dst, src, bytes);
            return;
        }
        std::vector<std::thread> th;
        const size_t per = (bytes / nthreads + 4095) & ~size_t(4095);
        for (int t = 0; t < nthreads; ++t)
        {
            const size_t o = std::min(bytes, per * t), e = std::min(bytes, per * (t + 1));
            if (e > o) th.emplace_back([=]() { std::memcpy(static_cast<char *>(dst) + o, static_cast<const char *>(src) + o, e - o); });
        }
        for (auto & t : th) t.join();
    }
}  // namespace


// ======================================================================================= ABI
extern "C" {

int sb200_abi_version(void) { return SB200_ABI_VERSION; }

void sb200_default_opts(sb200_opts * o)
{
    if (!o) return;
    std::memset(o, 0, sizeof(*o));
    o->abi_version = SB200_ABI_VERSION;
    o->device = 0;
    o->pricing = SB200_PRICE_FIRST_ELIGIBLE;  // Simplex.cpp:249
    o->engine = SB200_ENGINE_PERSISTENT;
    o->dual_mode = SB200_DUAL_UPDATE;  // refreshed by a full dual solve at every launch; same pivot sequences
    o->dual_refresh = 0;
    o->reinvert_period = 0;
    o->trace_capacity = 0;
    o->eps = 1e-7;                 // Utils.h:9
    o->iter_limit = 1000000000LL;  // Simplex.cpp:274
    o->chunk_iters = 64;
    o->stream = nullptr;
    o->rank = 0;
    o->world = 1;
    o->hygiene_tol = 1e-9;
    o->hygiene_period = 16;
    o->hygiene_rows = 8;
}

const char * sb200_last_error(const sb200_ctx * ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int sb200_create(sb200_ctx ** out, const sb200_opts * opts)
{
    if (!out) return SB200_ERR_INVALID;
    *out = nullptr;
    sb200_opts o;
    if (opts)
        o = *opts;
    else
        sb200_default_opts(&o);
    if (o.abi_version != SB200_ABI_VERSION) return fail(nullptr, SB200_ERR_INVALID, "sb200_create: ABI version mismatch");
    if (!(o.eps >= 0.0) || o.iter_limit <= 0) return fail(nullptr, SB200_ERR_INVALID, "sb200_create: bad eps/iter_limit");
    if (o.pricing != SB200_PRICE_FIRST_ELIGIBLE && o.pricing != SB200_PRICE_MOST_NEGATIVE)
        return fail(nullptr, SB200_ERR_INVALID, "sb200_create: unknown pricing rule");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0)
        return fail(nullptr, SB200_ERR_CUDA,
                    std::string("no CUDA device (this library has no CPU path): ") + cudaGetErrorString(e));
    if (o.device < 0 || o.device >= ndev) return fail(nullptr, SB200_ERR_INVALID, "sb200_create: bad device ordinal");
    sb200_ctx * ctx = new sb200_ctx();
    ctx->opts = o;
    auto bail = [&](const char * what, cudaError_t ce) {
        g_create_error = std::string(what) + ": " + cudaGetErrorString(ce);
        delete ctx;
        return SB200_ERR_CUDA;
    };
    if ((e = cudaSetDevice(o.device)) != cudaSuccess) return bail("cudaSetDevice", e);
    cudaDeviceProp prop{};
    if ((e = cudaGetDeviceProperties(&prop, o.device)) != cudaSuccess) return bail("cudaGetDeviceProperties", e);
    ctx->num_sms = prop.multiProcessorCount;
    {
        int khz = 0;
        cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, o.device);
        ctx->sm_mhz = khz > 0 ? khz / 1000.0 : 1.0;
    }
    ctx->max_smem_optin = static_cast<int>(prop.sharedMemPerBlockOptin);
    ctx->max_smem_per_sm = static_cast<int>(prop.sharedMemPerMultiprocessor);
    ctx->coop_supported = prop.cooperativeLaunch != 0;
    if (o.stream)
    {
        ctx->stream = static_cast<cudaStream_t>(o.stream);
    }
    else
    {
        if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess)
            return bail("cudaStreamCreate", e);
        ctx->own_stream = true;
    }
    if ((e = cudaEventCreate(&ctx->ev0)) != cudaSuccess) return bail("cudaEventCreate", e);
    if ((e = cudaEventCreate(&ctx->ev1)) != cudaSuccess) return bail("cudaEventCreate", e);
    if ((e = cudaMallocHost(&ctx->h_sc, sizeof(Scalars))) != cudaSuccess) return bail("cudaMallocHost", e);
    int rc = set_func_attrs(ctx);
    if (rc != SB200_OK)
    {
        g_create_error = ctx->err;
        sb200_destroy(ctx);
        return rc;
    }
    *out = ctx;
    return SB200_OK;
}

void sb200_destroy(sb200_ctx * ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->opts.device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    free_batch_buffers(static_cast<BatchBuffers *>(ctx->batch_buffers));
    ctx->batch_buffers = nullptr;
    free_all(ctx);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->h_sc) cudaFreeHost(ctx->h_sc);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int sb200_load(sb200_ctx * ctx, int64_t m, int64_t N, const double * A, int64_t lda, const double * b, const double * c,
               const double * ub)
{
    return load_common(ctx, m, N, A, lda, b, c, ub, cudaMemcpyHostToDevice);
}

int sb200_load_device(sb200_ctx * ctx, int64_t m, int64_t N, const double * dA, int64_t lda, const double * db,
                      const double * dc, const double * dub)
{
    return load_common(ctx, m, N, dA, lda, db, dc, dub, cudaMemcpyDeviceToDevice);
}

int sb200_set_costs(sb200_ctx * ctx, int64_t N_new, const double * c)
{
    if (!ctx) return SB200_ERR_INVALID;
    if (!ctx->loaded) return fail(ctx, SB200_ERR_STATE, "sb200_set_costs: nothing loaded");
    if (ctx->sharded) return fail(ctx, SB200_ERR_STATE, "sb200_set_costs: sharded context (use sb200_shard_set_costs)");
    if (!c || N_new < ctx->S.m || N_new > ctx->N_loaded)
        return fail(ctx, SB200_ERR_INVALID, "sb200_set_costs: need m <= N_new <= loaded N");
    CU_TRY(ctx, cudaSetDevice(ctx->opts.device));
    ctx->S.N = static_cast<int>(N_new);
    ctx->S.nnb = static_cast<int>(N_new - ctx->S.m);
    const int wpb = PRICE_THREADS / 32;
    ctx->price_grid = std::max(1, std::min(2 * ctx->num_sms, (ctx->S.nnb + wpb - 1) / wpb));
    ctx->S.n_price_slots = ctx->price_grid;
    CU_TRY(ctx, cudaMemcpyAsync(ctx->S.c, c, N_new * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->prepared = false;
    return SB200_OK;
}

int sb200_prepare(sb200_ctx * ctx, const int64_t * basic, int64_t nb, const int64_t * nonbasic, int64_t nn)
{
    if (!ctx) return SB200_ERR_INVALID;
    if (!basic || (!nonbasic && nn > 0)) return fail(ctx, SB200_ERR_INVALID, "sb200_prepare: null list");
    return prepare_impl(ctx, basic, nb, nonbasic, nn);
}

int sb200_run(sb200_ctx * ctx, int64_t * basic, int64_t nb, int64_t * nonbasic, int64_t nn, sb200_result * out)
{
    if (!ctx) return SB200_ERR_INVALID;
    if (!basic || (!nonbasic && nn > 0) || !out) return fail(ctx, SB200_ERR_INVALID, "sb200_run: null argument");
    int rc = prepare_impl(ctx, basic, nb, nonbasic, nn);
    if (rc != SB200_OK) return rc;
    if (ctx->opts.engine == SB200_ENGINE_PERSISTENT || ctx->opts.engine == SB200_ENGINE_STREAMING)
        rc = run_persistent(ctx);
    else if (ctx->opts.engine == SB200_ENGINE_STAGED)
        rc = run_staged(ctx);
    else
        rc = fail(ctx, SB200_ERR_INVALID, "unknown engine");
    if (rc != SB200_OK) return rc;
    CU_TRY(ctx, cudaGetLastError());

    const DevState & S = ctx->S;
    float ms = 0.f;
    CU_TRY(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    std::vector<int> hb(S.m), hn(S.nnb);
    std::vector<double> hx(S.m), hc(S.N);
    CU_TRY(ctx, cudaMemcpyAsync(hb.data(), S.basic, S.m * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    if (S.nnb > 0)
        CU_TRY(ctx, cudaMemcpyAsync(hn.data(), S.nonbasic, S.nnb * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaMemcpyAsync(hx.data(), S.xB, S.m * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaMemcpyAsync(hc.data(), S.c, S.N * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    double obj = 0.0;
    for (int i = 0; i < S.m; ++i)
    {
        basic[i] = hb[i];
        obj += hc[hb[i]] * hx[i];  // Simplex.cpp:210-211 (before shifts)
    }
    for (int i = 0; i < S.nnb; ++i) nonbasic[i] = hn[i];
    if (ctx->h_sc->status == TERM_OPTIMAL || ctx->h_sc->status == TERM_UNBOUNDED || ctx->h_sc->status == TERM_ITER_LIMIT)
        ctx->carry_basic.assign(basic, basic + S.m);  // B^-1 and x_B on the device belong to this basis
    out->term = ctx->h_sc->status;
    out->reserved = 0;
    out->iterations = ctx->h_sc->iterations;
    out->pivots = ctx->h_sc->pivots;
    out->bound_flips = ctx->h_sc->flips;
    out->objective = obj;
    out->loop_seconds = ms * 1e-3;
    return SB200_OK;
}

#define GETTER_PRE(name)                                                              \
    if (!ctx) return SB200_ERR_INVALID;                                               \
    if (!ctx->prepared) return fail(ctx, SB200_ERR_STATE, name ": no prepared state"); \
    if (!dst) return fail(ctx, SB200_ERR_INVALID, name ": null buffer");              \
    CU_TRY(ctx, cudaSetDevice(ctx->opts.device))

int sb200_get_xB(sb200_ctx * ctx, double * dst)
{
    GETTER_PRE("sb200_get_xB");
    if (ctx->sharded) return fail(ctx, SB200_ERR_STATE, "sb200_get_xB: x_B is row-sharded on this context (sb200_shard_get_xB_local)");
    CU_TRY(ctx, cudaMemcpyAsync(dst, ctx->S.xB, ctx->S.m * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return SB200_OK;
}
int sb200_get_y(sb200_ctx * ctx, double * dst)
{
    GETTER_PRE("sb200_get_y");
    if (ctx->sharded) return fail(ctx, SB200_ERR_STATE, "sb200_get_y: not available on a sharded context");
    CU_TRY(ctx, cudaMemcpyAsync(dst, ctx->S.y, ctx->S.m * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return SB200_OK;
}
int sb200_get_d(sb200_ctx * ctx, double * dst)
{
    GETTER_PRE("sb200_get_d");
    CU_TRY(ctx, cudaMemcpyAsync(dst, ctx->S.d, ctx->S.m * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return SB200_OK;
}
int sb200_get_b(sb200_ctx * ctx, double * dst)
{
    GETTER_PRE("sb200_get_b");
    CU_TRY(ctx, cudaMemcpyAsync(dst, ctx->S.b, ctx->S.m * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return SB200_OK;
}
int sb200_get_flip_flags(sb200_ctx * ctx, uint8_t * dst)
{
    GETTER_PRE("sb200_get_flip_flags");
    CU_TRY(ctx, cudaMemcpyAsync(dst, ctx->S.flip, ctx->S.N, cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return SB200_OK;
}
int sb200_get_Binv(sb200_ctx * ctx, double * dst)
{
    GETTER_PRE("sb200_get_Binv");
    if (ctx->sharded) return fail(ctx, SB200_ERR_STATE, "sb200_get_Binv: B^-1 is row-sharded on this context");
    const DevState & S = ctx->S;
    CU_TRY(ctx, cudaMemcpy2DAsync(dst, S.m * sizeof(double), S.Binv, S.ld * sizeof(double), S.m * sizeof(double), S.m,
                                  cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return SB200_OK;
}

int sb200_get_trace(sb200_ctx * ctx, sb200_pivot * dst, int64_t cap, int64_t * n_total)
{
    if (!ctx) return SB200_ERR_INVALID;
    if (!ctx->prepared) return fail(ctx, SB200_ERR_STATE, "sb200_get_trace: no prepared state");
    CU_TRY(ctx, cudaSetDevice(ctx->opts.device));
    int rc = fetch_scalars(ctx);
    if (rc != SB200_OK) return rc;
    const int64_t total = ctx->h_sc->pivots;
    if (n_total) *n_total = total;
    const DevState & S = ctx->S;
    if (!dst || cap <= 0 || S.trace_cap <= 0) return SB200_OK;
    const int64_t kept = std::min<int64_t>(total, S.trace_cap);
    const int64_t n = std::min<int64_t>(kept, cap);
    std::vector<int4> ring(S.trace_cap);
    CU_TRY(ctx, cudaMemcpyAsync(ring.data(), S.trace, sizeof(int4) * S.trace_cap, cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    const int64_t first = total - kept;  // oldest pivot still in the ring
    for (int64_t k = 0; k < n; ++k)
    {
        const int4 v = ring[(first + k) % S.trace_cap];
        dst[k].leaving_col = v.x;
        dst[k].row = v.y;
        dst[k].entering_col = v.z;
        dst[k].pos = v.w;
    }
    return SB200_OK;
}

// ---- single-phase entry points -------------------------------------------------------------
#define STEP_PRE(name)                                                                 \
    if (!ctx) return SB200_ERR_INVALID;                                                \
    if (!ctx->prepared) return fail(ctx, SB200_ERR_STATE, name ": call sb200_prepare first"); \
    if (ctx->sharded) return fail(ctx, SB200_ERR_STATE, name ": not available on a sharded context"); \
    CU_TRY(ctx, cudaSetDevice(ctx->opts.device))

int sb200_step_dual(sb200_ctx * ctx)
{
    STEP_PRE("sb200_step_dual");
    launch_dual(ctx);
    ctx->launches += 2;
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    CU_TRY(ctx, cudaGetLastError());
    return SB200_OK;
}

int sb200_step_price(sb200_ctx * ctx, int64_t * pos, double * value)
{
    STEP_PRE("sb200_step_price");
    launch_price(ctx);
    k_pick_entering<<<1, 256, 0, ctx->stream>>>(ctx->S);
    ctx->launches += 2;
    int rc = fetch_scalars(ctx);
    if (rc != SB200_OK) return rc;
    CU_TRY(ctx, cudaGetLastError());
    if (pos) *pos = ctx->h_sc->e_pos;
    if (value) *value = ctx->h_sc->e_pos >= 0 ? ctx->h_sc->s_q : 0.0;
    return SB200_OK;
}

int sb200_step_ftran(sb200_ctx * ctx, int64_t pos)
{
    STEP_PRE("sb200_step_ftran");
    int rc = fetch_scalars(ctx);
    if (rc != SB200_OK) return rc;
    if (pos != ctx->h_sc->e_pos) return fail(ctx, SB200_ERR_INVALID, "sb200_step_ftran: pos is not the priced position");
    launch_ftran(ctx, false);
    ctx->launches += 1;
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    CU_TRY(ctx, cudaGetLastError());
    return SB200_OK;
}

int sb200_step_ratio(sb200_ctx * ctx, int64_t pos, int64_t * row, double * theta)
{
    STEP_PRE("sb200_step_ratio");
    ctx->carry_basic.clear();
    if (ctx->S.ub != nullptr) ctx->units_stale = true;
    int rc = fetch_scalars(ctx);
    if (rc != SB200_OK) return rc;
    if (pos != ctx->h_sc->e_pos) return fail(ctx, SB200_ERR_INVALID, "sb200_step_ratio: pos is not the priced position");
    launch_select(ctx);
    ctx->launches += 1;
    rc = fetch_scalars(ctx);
    if (rc != SB200_OK) return rc;
    CU_TRY(ctx, cudaGetLastError());
    if (row) *row = ctx->h_sc->last_ret;
    if (theta) *theta = ctx->h_sc->last_ret >= 0 ? ctx->h_sc->theta : 0.0;
    return SB200_OK;
}

int sb200_step_update(sb200_ctx * ctx)
{
    STEP_PRE("sb200_step_update");
    ctx->carry_basic.clear();
    launch_update(ctx);
    ctx->launches += 1;
    // consume the action so that a second call is a no-op
    const int none = ACT_NONE;
    CU_TRY(ctx, cudaMemcpyAsync(reinterpret_cast<char *>(ctx->S.sc) + offsetof(Scalars, action), &none, sizeof(int),
                                cudaMemcpyHostToDevice, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    CU_TRY(ctx, cudaGetLastError());
    return SB200_OK;
}

// ---- batched mode ---------------------------------------------------------------------------
static int run_batched_impl(sb200_ctx * ctx, int64_t batch, int64_t m, int64_t N, const double * A, const double * b,
                            const double * c, const double * ub, int32_t * basic, int32_t * nonbasic, double * xB,
                            double * objective, int32_t * term, int32_t * iterations, uint8_t * flip_flags, int32_t * bound_flips,
                            int device_resident, double * seconds);

int sb200_run_batched(sb200_ctx * ctx, int64_t batch, int64_t m, int64_t N, const double * A, const double * b,
                      const double * c, int32_t * basic, int32_t * nonbasic, double * xB, double * objective,
                      int32_t * term, int32_t * iterations, int device_resident, double * seconds)
{
    return run_batched_impl(ctx, batch, m, N, A, b, c, nullptr, basic, nonbasic, xB, objective, term, iterations, nullptr, nullptr,
                            device_resident, seconds);
}

int sb200_run_batched_bounded(sb200_ctx * ctx, int64_t batch, int64_t m, int64_t N, const double * A, const double * b,
                              const double * c, const double * ub, int32_t * basic, int32_t * nonbasic, double * xB,
                              double * objective, int32_t * term, int32_t * iterations, uint8_t * flip_flags,
                              int32_t * bound_flips, int device_resident, double * seconds)
{
    if (ctx && (!ub || !flip_flags || !bound_flips))
        return fail(ctx, SB200_ERR_INVALID, "sb200_run_batched_bounded: ub, flip_flags and bound_flips are required");
    return run_batched_impl(ctx, batch, m, N, A, b, c, ub, basic, nonbasic, xB, objective, term, iterations, flip_flags, bound_flips,
                            device_resident, seconds);
}

static int run_batched_impl(sb200_ctx * ctx, int64_t batch, int64_t m, int64_t N, const double * A, const double * b,
                            const double * c, const double * ub, int32_t * basic, int32_t * nonbasic, double * xB,
                            double * objective, int32_t * term, int32_t * iterations, uint8_t * flip_flags, int32_t * bound_flips,
                            int device_resident, double * seconds)
{
    if (!ctx) return SB200_ERR_INVALID;
    const bool bounded = ub != nullptr;
    if (batch <= 0 || m <= 0 || N < m || !A || !b || !c || !basic || !nonbasic || !xB || !objective || !term ||
        !iterations)
        return fail(ctx, SB200_ERR_INVALID, "sb200_run_batched: bad argument");
    CU_TRY(ctx, cudaSetDevice(ctx->opts.device));
    BatchedParams P{};
    P.m = static_cast<int>(m);
    P.N = static_cast<int>(N);
    P.eps = ctx->opts.eps;
    P.rule = ctx->opts.pricing;
    P.iter_limit = static_cast<int>(std::min<int64_t>(ctx->opts.iter_limit, 0x7fffffff));
    const size_t smem = batched_smem_bytes(P.m, P.N, bounded);
    if (smem > static_cast<size_t>(ctx->max_smem_optin) - 2048)
        return fail(ctx, SB200_ERR_UNSUPPORTED, "sb200_run_batched: LP does not fit in shared memory (one CTA per LP)");
    cudaStream_t st = ctx->stream;
    const size_t nn1 = static_cast<size_t>(N - m);

    // fast kernel (B^-1 in registers, two CTAs per SM) for m <= 64; LPs it cannot take (start basis not
    // made of unit columns, too many dense columns) are marked TERM_PENDING and go to the generic one
    // Dense capacity: all N columns when that still lets two CTAs share an SM (configs[4]: 108 KB each),
    // otherwise N - m (enough for any LP that starts from a unit basis; unit columns stay implicit).
    const size_t two_per_sm = (static_cast<size_t>(ctx->max_smem_per_sm) - 2 * 1024) / 2 - 1024;
    const bool all_dense = batched_fast_smem_bytes(P.N, P.N) <= two_per_sm && std::getenv("SB200_BATCHED_HYBRID") == nullptr;
    const int dense_cap = all_dense ? P.N : static_cast<int>(N - m);
    const size_t smem_fast = batched_fast_smem_bytes(P.N, dense_cap);
    // (finite upper bounds: the generic kernel, whose whole tableau sits in shared memory and is rewritten in place)
    const bool use_fast = !bounded && m <= BF_ROWS && smem_fast <= static_cast<size_t>(ctx->max_smem_optin) - 2048 &&
                          std::getenv("SB200_BATCHED_GENERIC") == nullptr;
    // N - m <= 128: one pricing trip per pivot, the column offsets of a thread's 8 positions stay in registers
    const bool cached = N - m <= 128 && std::getenv("SB200_BATCHED_NOCACHE") == nullptr;
    if (use_fast)
    {
        for (auto fn : { (const void *)k_batched_fast<true, true>, (const void *)k_batched_fast<false, true>,
                         (const void *)k_batched_fast<true, false>, (const void *)k_batched_fast<false, false> })
        {
            CU_TRY(ctx, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_fast)));
            CU_TRY(ctx, cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        }
    }
    // persistent buffers (first call allocates, later calls of the same or a smaller shape reuse)
    const size_t want_chunk = device_resident ? 1 : static_cast<size_t>(std::min<int64_t>(batch, 4096));
    BatchBuffers * B = static_cast<BatchBuffers *>(ctx->batch_buffers);
    if (!B || B->m != P.m || B->N != P.N || B->chunk_lps < want_chunk || (bounded && !device_resident && B->dub[0] == nullptr))
    {
        CU_TRY(ctx, cudaStreamSynchronize(st));
        free_batch_buffers(B);
        ctx->batch_buffers = nullptr;
        B = new BatchBuffers();
        ctx->batch_buffers = B;
        B->m = P.m;
        B->N = P.N;
        B->chunk_lps = want_chunk;
        auto balloc = [&](void ** p, size_t bytes) {
            cudaError_t e = cudaMalloc(p, bytes + 256);
            if (e == cudaSuccess) B->allocs.push_back(*p);
            return e;
        };
        CU_TRY(ctx, balloc(reinterpret_cast<void **>(&B->d_pending), 2 * sizeof(int)));
        CU_TRY(ctx, cudaMallocHost(reinterpret_cast<void **>(&B->h_pending), 2 * sizeof(int)));
        B->h_pending[0] = B->h_pending[1] = 0;
        if (!device_resident)
        {
            const size_t L = want_chunk;
            for (int s2 = 0; s2 < 2; ++s2)
            {
                CU_TRY(ctx, balloc(reinterpret_cast<void **>(&B->dA[s2]), L * N * m * sizeof(double)));
                CU_TRY(ctx, balloc(reinterpret_cast<void **>(&B->db[s2]), L * m * sizeof(double)));
                CU_TRY(ctx, balloc(reinterpret_cast<void **>(&B->dc[s2]), L * N * sizeof(double)));
                CU_TRY(ctx, balloc(reinterpret_cast<void **>(&B->dx[s2]), L * m * sizeof(double)));
                CU_TRY(ctx, balloc(reinterpret_cast<void **>(&B->dobj[s2]), L * sizeof(double)));
                CU_TRY(ctx, balloc(reinterpret_cast<void **>(&B->dbas[s2]), L * m * sizeof(int)));
                CU_TRY(ctx, balloc(reinterpret_cast<void **>(&B->dnon[s2]), (L * nn1 + 1) * sizeof(int)));
                CU_TRY(ctx, balloc(reinterpret_cast<void **>(&B->dterm[s2]), L * sizeof(int)));
                CU_TRY(ctx, balloc(reinterpret_cast<void **>(&B->dit[s2]), L * sizeof(int)));
                if (bounded)
                {
                    CU_TRY(ctx, balloc(reinterpret_cast<void **>(&B->dub[s2]), L * N * sizeof(double)));
                    CU_TRY(ctx, balloc(reinterpret_cast<void **>(&B->dflip[s2]), L * N));
                    CU_TRY(ctx, balloc(reinterpret_cast<void **>(&B->dflips[s2]), L * sizeof(int)));
                }
                CU_TRY(ctx, cudaEventCreateWithFlags(&B->h2d_done[s2], cudaEventDisableTiming));
                CU_TRY(ctx, cudaEventCreateWithFlags(&B->slot_free[s2], cudaEventDisableTiming));
            }
            CU_TRY(ctx, cudaStreamCreateWithFlags(&B->copy_stream, cudaStreamNonBlocking));
        }
    }
    auto launch_fast = [&](const BatchedParams & Q, int * pending) {
        const dim3 grid(static_cast<unsigned>(Q.batch)), block(BF_THREADS);
        if (cached && all_dense)
            k_batched_fast<true, true><<<grid, block, smem_fast, st>>>(Q, dense_cap, pending);
        else if (cached)
            k_batched_fast<true, false><<<grid, block, smem_fast, st>>>(Q, dense_cap, pending);
        else if (all_dense)
            k_batched_fast<false, true><<<grid, block, smem_fast, st>>>(Q, dense_cap, pending);
        else
            k_batched_fast<false, false><<<grid, block, smem_fast, st>>>(Q, dense_cap, pending);
        ctx->launches += 1;
    };

    if (device_resident)
    {
        P.batch = static_cast<int>(batch);
        P.A = A;
        P.b = b;
        P.c = c;
        P.basic = basic;
        P.nonbasic = nonbasic;
        P.xB = xB;
        P.objective = objective;
        P.term = term;
        P.iterations = iterations;
        P.ub = ub;
        P.flip = flip_flags;
        P.flips = bound_flips;
        int h_pending = 0;
        CU_TRY(ctx, cudaMemsetAsync(B->d_pending, 0, sizeof(int), st));
        CU_TRY(ctx, cudaEventRecord(ctx->ev0, st));
        if (use_fast)
        {
            launch_fast(P, B->d_pending);
            CU_TRY(ctx, cudaMemcpyAsync(&h_pending, B->d_pending, sizeof(int), cudaMemcpyDeviceToHost, st));
            CU_TRY(ctx, cudaStreamSynchronize(st));
        }
        if (!use_fast || h_pending > 0)
        {
            P.only_pending = use_fast ? 1 : 0;
            k_batched<<<static_cast<unsigned>(batch), BATCHED_THREADS, smem, st>>>(P);
            ctx->launches += 1;
        }
        CU_TRY(ctx, cudaEventRecord(ctx->ev1, st));
        CU_TRY(ctx, cudaStreamSynchronize(st));
        CU_TRY(ctx, cudaGetLastError());
        float ms = 0.f;
        CU_TRY(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        if (seconds) *seconds = ms * 1e-3;
        return SB200_OK;
    }

    // ---- host buffers: chunks of LPs through two device buffers. Chunk k+1 is copied in (copy stream) while
    // the kernel of chunk k runs (compute stream); results go back behind each kernel. Inputs in pinned memory
    // are DMA-ed straight from the caller's buffer; pageable inputs pass through the pinned staging buffers.
    const size_t L = B->chunk_lps;
    const size_t bytesA = static_cast<size_t>(N) * m * sizeof(double), bytesb = m * sizeof(double), bytesc = N * sizeof(double);
    const size_t bytesbas = m * sizeof(int), bytesnon = nn1 * sizeof(int);
    const size_t bytesub = bounded ? bytesc : 0;
    const size_t per_lp = bytesA + bytesb + bytesc + bytesbas + bytesnon + bytesub;
    const bool direct = host_pointer_is_pinned(A);   // A is 98 % of the bytes
    if (!direct && B->pinned_bytes < L * per_lp + 64)
    {
        for (int s2 = 0; s2 < 2; ++s2)
        {
            if (B->pinned[s2]) cudaFreeHost(B->pinned[s2]);
            B->pinned[s2] = nullptr;
            CU_TRY(ctx, cudaMallocHost(reinterpret_cast<void **>(&B->pinned[s2]), L * per_lp + 64));
        }
        B->pinned_bytes = L * per_lp + 64;
    }
    cudaStream_t cs = B->copy_stream;
    CU_TRY(ctx, cudaEventRecord(ctx->ev0, st));
    const int64_t nchunks = (batch + static_cast<int64_t>(L) - 1) / static_cast<int64_t>(L);
    BatchedParams slotQ[2] = {};
    size_t slot_lp0[2] = { 0, 0 };
    auto read_back = [&](const BatchedParams & Q, size_t lp0) -> int {
        const size_t cnt = static_cast<size_t>(Q.batch);
        CU_TRY(ctx, cudaMemcpyAsync(basic + lp0 * m, Q.basic, cnt * bytesbas, cudaMemcpyDeviceToHost, st));
        if (nn1 > 0) CU_TRY(ctx, cudaMemcpyAsync(nonbasic + lp0 * nn1, Q.nonbasic, cnt * bytesnon, cudaMemcpyDeviceToHost, st));
        CU_TRY(ctx, cudaMemcpyAsync(xB + lp0 * m, Q.xB, cnt * m * sizeof(double), cudaMemcpyDeviceToHost, st));
        CU_TRY(ctx, cudaMemcpyAsync(objective + lp0, Q.objective, cnt * sizeof(double), cudaMemcpyDeviceToHost, st));
        CU_TRY(ctx, cudaMemcpyAsync(term + lp0, Q.term, cnt * sizeof(int), cudaMemcpyDeviceToHost, st));
        CU_TRY(ctx, cudaMemcpyAsync(iterations + lp0, Q.iterations, cnt * sizeof(int), cudaMemcpyDeviceToHost, st));
        if (bounded)
        {
            CU_TRY(ctx, cudaMemcpyAsync(flip_flags + lp0 * N, Q.flip, cnt * static_cast<size_t>(N), cudaMemcpyDeviceToHost, st));
            CU_TRY(ctx, cudaMemcpyAsync(bound_flips + lp0, Q.flips, cnt * sizeof(int), cudaMemcpyDeviceToHost, st));
        }
        return SB200_OK;
    };
    // LPs the fast kernel declined (TERM_PENDING: start basis not made of unit columns, ...) are left to the generic
    // kernel. The count comes back with the results of the chunk; it is looked at when the slot is about to be
    // reused (or after the last chunk), which needs that chunk finished anyway: no extra round trip in the pipeline.
    auto finish_slot = [&](int s2) -> int {
        CU_TRY(ctx, cudaEventSynchronize(B->slot_free[s2]));
        if (use_fast && B->h_pending[s2] > 0)
        {
            BatchedParams Q = slotQ[s2];
            Q.only_pending = 1;
            k_batched<<<static_cast<unsigned>(Q.batch), BATCHED_THREADS, smem, st>>>(Q);
            ctx->launches += 1;
            int rrc = read_back(Q, slot_lp0[s2]);
            if (rrc != SB200_OK) return rrc;
            CU_TRY(ctx, cudaStreamSynchronize(st));
            B->h_pending[s2] = 0;
        }
        return SB200_OK;
    };
    // copy of chunk k into slot k & 1 (copy stream); returns with the H2D queued
    auto issue_h2d = [&](int64_t k) -> int {
        const int s2 = static_cast<int>(k & 1);
        const size_t lp0 = static_cast<size_t>(k) * L;
        const size_t cnt = std::min(L, static_cast<size_t>(batch) - lp0);
        // the device buffers of this slot are free once the results of chunk k-2 have gone back
        if (k >= 2)
        {
            int frc = finish_slot(s2);
            if (frc != SB200_OK) return frc;
        }
        const double * srcA = A + lp0 * N * m;
        const double * srcb = b + lp0 * m;
        const double * srcc = c + lp0 * N;
        const int * srcbas = basic + lp0 * m;
        const int * srcnon = nonbasic + lp0 * nn1;
        const double * srcub = bounded ? ub + lp0 * N : nullptr;
        if (!direct)
        {
            // (the staging buffer of this slot is free as well: its copy to the device ended before that chunk's kernel)
            unsigned char * pz = B->pinned[s2];
            parallel_memcpy(pz, srcA, cnt * bytesA);
            srcA = reinterpret_cast<const double *>(pz);
            pz += cnt * bytesA;
            std::memcpy(pz, srcb, cnt * bytesb);
            srcb = reinterpret_cast<const double *>(pz);
            pz += cnt * bytesb;
            std::memcpy(pz, srcc, cnt * bytesc);
            srcc = reinterpret_cast<const double *>(pz);
            pz += cnt * bytesc;
            std::memcpy(pz, srcbas, cnt * bytesbas);
            srcbas = reinterpret_cast<const int *>(pz);
            pz += cnt * bytesbas;
            std::memcpy(pz, srcnon, cnt * bytesnon);
            srcnon = reinterpret_cast<const int *>(pz);
            pz += cnt * bytesnon;
            pz = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(pz) + 7) & ~uintptr_t(7));
            if (bounded)
            {
                std::memcpy(pz, srcub, cnt * bytesub);
                srcub = reinterpret_cast<const double *>(pz);
            }
        }
        CU_TRY(ctx, cudaMemcpyAsync(B->dA[s2], srcA, cnt * bytesA, cudaMemcpyHostToDevice, cs));
        CU_TRY(ctx, cudaMemcpyAsync(B->db[s2], srcb, cnt * bytesb, cudaMemcpyHostToDevice, cs));
        CU_TRY(ctx, cudaMemcpyAsync(B->dc[s2], srcc, cnt * bytesc, cudaMemcpyHostToDevice, cs));
        CU_TRY(ctx, cudaMemcpyAsync(B->dbas[s2], srcbas, cnt * bytesbas, cudaMemcpyHostToDevice, cs));
        if (nn1 > 0) CU_TRY(ctx, cudaMemcpyAsync(B->dnon[s2], srcnon, cnt * bytesnon, cudaMemcpyHostToDevice, cs));
        if (bounded) CU_TRY(ctx, cudaMemcpyAsync(B->dub[s2], srcub, cnt * bytesub, cudaMemcpyHostToDevice, cs));
        CU_TRY(ctx, cudaEventRecord(B->h2d_done[s2], cs));
        return SB200_OK;
    };
    {
        int hrc = issue_h2d(0);
        if (hrc != SB200_OK) return hrc;
    }
    for (int64_t k = 0; k < nchunks; ++k)
    {
        const int s2 = static_cast<int>(k & 1);
        const size_t lp0 = static_cast<size_t>(k) * L;
        const size_t cnt = std::min(L, static_cast<size_t>(batch) - lp0);
        CU_TRY(ctx, cudaStreamWaitEvent(st, B->h2d_done[s2], 0));
        BatchedParams Q = P;
        Q.batch = static_cast<int>(cnt);
        Q.A = B->dA[s2];
        Q.b = B->db[s2];
        Q.c = B->dc[s2];
        Q.basic = B->dbas[s2];
        Q.nonbasic = B->dnon[s2];
        Q.xB = B->dx[s2];
        Q.objective = B->dobj[s2];
        Q.term = B->dterm[s2];
        Q.iterations = B->dit[s2];
        Q.ub = bounded ? B->dub[s2] : nullptr;
        Q.flip = bounded ? B->dflip[s2] : nullptr;
        Q.flips = bounded ? B->dflips[s2] : nullptr;
        Q.only_pending = 0;
        slotQ[s2] = Q;
        slot_lp0[s2] = lp0;
        if (use_fast)
        {
            CU_TRY(ctx, cudaMemsetAsync(B->d_pending + s2, 0, sizeof(int), st));
            launch_fast(Q, B->d_pending + s2);
            CU_TRY(ctx, cudaMemcpyAsync(B->h_pending + s2, B->d_pending + s2, sizeof(int), cudaMemcpyDeviceToHost, st));
        }
        else
        {
            k_batched<<<static_cast<unsigned>(cnt), BATCHED_THREADS, smem, st>>>(Q);
            ctx->launches += 1;
        }
        // the next chunk's copy is queued BEFORE this chunk's results are read back: a read-back into pageable
        // memory holds the host until the kernel has finished, and the copy engine would idle meanwhile
        if (k + 1 < nchunks)
        {
            int hrc = issue_h2d(k + 1);
            if (hrc != SB200_OK) return hrc;
        }
        int rrc = read_back(Q, lp0);
        if (rrc != SB200_OK) return rrc;
        CU_TRY(ctx, cudaEventRecord(B->slot_free[s2], st));
    }
    for (int64_t k = std::max<int64_t>(0, nchunks - 2); k < nchunks; ++k)
    {
        int frc = finish_slot(static_cast<int>(k & 1));
        if (frc != SB200_OK) return frc;
    }
    CU_TRY(ctx, cudaEventRecord(ctx->ev1, st));
    CU_TRY(ctx, cudaStreamSynchronize(cs));
    CU_TRY(ctx, cudaStreamSynchronize(st));
    CU_TRY(ctx, cudaGetLastError());
    float ms = 0.f;
    CU_TRY(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    if (seconds) *seconds = ms * 1e-3;
    return SB200_OK;
}

// ---- sharded mode (one wide LP over the GPUs of one box; kernels in sb200_sharded.cuh) --------
int sb200_load_shard(sb200_ctx * ctx, int64_t m, int64_t N, int64_t col0, int64_t ncols, const double * A_cols,
                     int64_t lda, const double * b, const double * c, int device_resident)
{
    if (!ctx) return SB200_ERR_INVALID;
    const int world = ctx->opts.world, rank = ctx->opts.rank;
    if (world < 1 || world > MAX_WORLD || rank < 0 || rank >= world)
        return fail(ctx, SB200_ERR_INVALID, "sb200_load_shard: need 0 <= rank < world <= 8 in sb200_opts");
    if (m <= 0 || N < m || !A_cols || !b || !c || lda < m || ncols < 0)
        return fail(ctx, SB200_ERR_INVALID, "sb200_load_shard: bad argument");
    const int64_t cpr = (N + world - 1) / world;
    if (col0 != rank || ncols != (N > rank ? (N - rank + world - 1) / world : 0))
        return fail(ctx, SB200_ERR_INVALID,
                    "sb200_load_shard: the shard must be the cyclic column set {rank + k*world} of this rank "
                    "(col0 = rank, ncols = ceil((N - rank)/world))");
    CU_TRY(ctx, cudaSetDevice(ctx->opts.device));
    free_all(ctx);
    DevState & S = ctx->S;
    S = DevState{};
    S.m = static_cast<int>(m);
    S.N = static_cast<int>(N);
    S.nnb = static_cast<int>(N - m);
    S.ld = round_up(m, 16);
    S.eps = ctx->opts.eps;
    S.rule = ctx->opts.pricing;
    S.dual_update = 1;
    ctx->N_loaded = N;
    ShardInfo & H = ctx->H;
    H = ShardInfo{};
    H.rank = rank;
    H.world = world;
    H.col0 = static_cast<int>(col0);
    H.ncols = static_cast<int>(ncols);
    H.cols_per_rank = static_cast<int>(cpr);
    // row blocks are multiples of DUAL_BLOCK_ROWS so that the dual-solve blocks never straddle ranks
    H.rows_per_rank = round_up((m + world - 1) / world, DUAL_BLOCK_ROWS);
    H.row0 = std::min<int64_t>(m, static_cast<int64_t>(rank) * H.rows_per_rank);
    H.nrows = static_cast<int>(std::max<int64_t>(0, std::min<int64_t>(H.rows_per_rank, m - H.row0)));
    const size_t ld = S.ld;
    const int grid = ctx->num_sms;
    ctx->shard_grid = grid;
    if (sharded_smem_bytes(S.ld, H.nrows, grid) > static_cast<size_t>(ctx->max_smem_optin) - 2048)
        return fail(ctx, SB200_ERR_UNSUPPORTED, "m too large for the shared-memory staged vectors");
    CU_TRY(ctx, dalloc(ctx, &S.A, ld * std::max<int64_t>(ncols, 1)));
    CU_TRY(ctx, dalloc(ctx, &S.Binv, ld * std::max(H.nrows, 1)));
    CU_TRY(ctx, dalloc(ctx, &S.b, ld));
    CU_TRY(ctx, dalloc(ctx, &S.c, static_cast<size_t>(N)));
    CU_TRY(ctx, dalloc(ctx, &S.xB, ld));
    CU_TRY(ctx, dalloc(ctx, &S.y, ld));
    CU_TRY(ctx, dalloc(ctx, &S.d, ld));
    CU_TRY(ctx, dalloc(ctx, &S.prow, ld));
    CU_TRY(ctx, dalloc(ctx, &S.basic, static_cast<size_t>(m)));
    CU_TRY(ctx, dalloc(ctx, &S.nonbasic, static_cast<size_t>(N - m) + 1));
    CU_TRY(ctx, dalloc(ctx, &S.flip, static_cast<size_t>(N)));
    CU_TRY(ctx, dalloc(ctx, &S.sc, 1));
    CU_TRY(ctx, dalloc(ctx, &ctx->diag, ld));
    CU_TRY(ctx, dalloc(ctx, &ctx->flags, 4));
    CU_TRY(ctx, dalloc(ctx, &S.price_slots, static_cast<size_t>(grid)));
    CU_TRY(ctx, dalloc(ctx, &S.ratio_slots, static_cast<size_t>(grid)));
    S.n_price_slots = S.n_ratio_slots = grid;
    const int nblk = (H.nrows + DUAL_BLOCK_ROWS - 1) / DUAL_BLOCK_ROWS;
    CU_TRY(ctx, dalloc(ctx, &S.ypart, static_cast<size_t>(std::max(nblk, 1)) * ld));
    CU_TRY(ctx, dalloc(ctx, &S.row_partials, static_cast<size_t>(std::max(H.nrows, 1)) * ROW_PARTIAL_DOUBLES));
    CU_TRY(ctx, dalloc(ctx, &H.own_pos, static_cast<size_t>(N - m) + 2));
    CU_TRY(ctx, dalloc(ctx, &H.pos_slot, static_cast<size_t>(N - m) + 2));
    CU_TRY(ctx, dalloc(ctx, &H.own_cnt, 1));
    S.trace_cap = ctx->opts.trace_capacity;
    if (S.trace_cap > 0) CU_TRY(ctx, dalloc(ctx, &S.trace, static_cast<size_t>(S.trace_cap)));
    // the exchange window: its own allocation, so that one IPC handle maps exactly it
    {
        // (no guard zones here: the IPC handle must map exactly this allocation from its first byte)
        void * w = nullptr;
        CU_TRY(ctx, cudaMalloc(&w, window_bytes(S.ld)));
        ctx->allocs.push_back(w);
        ctx->window = static_cast<unsigned char *>(w);
        CU_TRY(ctx, cudaMemset(w, 0, window_bytes(S.ld)));
        for (int r = 0; r < MAX_WORLD; ++r) H.win[r] = nullptr;
        H.win[rank] = ctx->window;
        ctx->connected = (world == 1);
    }
    cudaStream_t st = ctx->stream;
    const cudaMemcpyKind kind = device_resident ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    CU_TRY(ctx, cudaMemsetAsync(S.A, 0, ld * std::max<int64_t>(ncols, 1) * sizeof(double), st));
    CU_TRY(ctx, cudaMemsetAsync(S.b, 0, ld * sizeof(double), st));
    CU_TRY(ctx, cudaMemsetAsync(S.xB, 0, ld * sizeof(double), st));
    CU_TRY(ctx, cudaMemsetAsync(S.y, 0, ld * sizeof(double), st));
    CU_TRY(ctx, cudaMemsetAsync(S.d, 0, ld * sizeof(double), st));
    CU_TRY(ctx, cudaMemsetAsync(S.flip, 0, static_cast<size_t>(N), st));
    CU_TRY(ctx, cudaMemsetAsync(S.sc, 0, sizeof(Scalars), st));
    if (ncols > 0)
        CU_TRY(ctx, cudaMemcpy2DAsync(S.A, ld * sizeof(double), A_cols, lda * sizeof(double), m * sizeof(double), ncols,
                                      kind, st));
    CU_TRY(ctx, cudaMemcpyAsync(S.b, b, m * sizeof(double), kind, st));
    CU_TRY(ctx, cudaMemcpyAsync(S.c, c, N * sizeof(double), kind, st));
    CU_TRY(ctx, cudaStreamSynchronize(st));
    ctx->loaded = true;
    ctx->sharded = true;
    return SB200_OK;
}

int sb200_shard_window_handle(sb200_ctx * ctx, void * handle64)
{
    if (!ctx || !handle64) return SB200_ERR_INVALID;
    if (!ctx->sharded) return fail(ctx, SB200_ERR_STATE, "sb200_shard_window_handle: call sb200_load_shard first");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is 64 bytes");
    CU_TRY(ctx, cudaSetDevice(ctx->opts.device));
    cudaIpcMemHandle_t h;
    CU_TRY(ctx, cudaIpcGetMemHandle(&h, ctx->window));
    std::memcpy(handle64, &h, 64);
    return SB200_OK;
}

int sb200_shard_connect(sb200_ctx * ctx, const void * handles)
{
    if (!ctx || !handles) return SB200_ERR_INVALID;
    if (!ctx->sharded) return fail(ctx, SB200_ERR_STATE, "sb200_shard_connect: call sb200_load_shard first");
    CU_TRY(ctx, cudaSetDevice(ctx->opts.device));
    ShardInfo & H = ctx->H;
    for (int r = 0; r < H.world; ++r)
    {
        if (r == H.rank) continue;
        cudaIpcMemHandle_t h;
        std::memcpy(&h, static_cast<const char *>(handles) + 64 * r, 64);
        void * p = nullptr;
        CU_TRY(ctx, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        H.win[r] = static_cast<unsigned char *>(p);
        ctx->peer_maps.push_back(p);
    }
    ctx->connected = true;
    return SB200_OK;
}

int sb200_shard_connect_local(sb200_ctx * ctx, sb200_ctx * const * peers, int64_t n_peers)
{
    if (!ctx || !peers) return SB200_ERR_INVALID;
    if (!ctx->sharded) return fail(ctx, SB200_ERR_STATE, "sb200_shard_connect_local: call sb200_load_shard first");
    ShardInfo & H = ctx->H;
    if (n_peers != H.world) return fail(ctx, SB200_ERR_INVALID, "sb200_shard_connect_local: need one context per rank");
    CU_TRY(ctx, cudaSetDevice(ctx->opts.device));
    int same_device = 0;
    for (int r = 0; r < H.world; ++r)
    {
        sb200_ctx * p = peers[r];
        if (!p || !p->sharded || p->H.rank != r || p->H.world != H.world || p->S.m != ctx->S.m || p->S.N != ctx->S.N)
            return fail(ctx, SB200_ERR_INVALID, "sb200_shard_connect_local: peers[r] must be rank r's loaded context of the same LP");
        if (p->opts.device == ctx->opts.device)
            same_device += 1;
        else
        {
            int can = 0;
            CU_TRY(ctx, cudaDeviceCanAccessPeer(&can, ctx->opts.device, p->opts.device));
            if (!can) return fail(ctx, SB200_ERR_UNSUPPORTED, "sb200_shard_connect_local: no peer access between the devices");
            cudaError_t e = cudaDeviceEnablePeerAccess(p->opts.device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CU_TRY(ctx, e);
            cudaGetLastError();
        }
        H.win[r] = p->window;
    }
    // ranks that share a device share its SMs: every rank's cooperative grid must be resident at the same time
    ctx->shard_grid = std::max(1, ctx->num_sms / same_device);
    ctx->connected = true;
    return SB200_OK;
}

namespace
{
    // The positions of the non-basic list whose column this rank owns (cyclic shards: column % world == rank), as the
    // compact list the pricing sweep strides over; the kernel keeps it up to date pivot by pivot.
    int upload_owned_positions(sb200_ctx * ctx, const int64_t * nonbasic, int64_t nn)
    {
        const ShardInfo & H = ctx->H;
        std::vector<int> own, slot(static_cast<size_t>(nn) + 1, -1);
        own.reserve(static_cast<size_t>(nn / H.world) + 2);
        for (int64_t pos = 0; pos < nn; ++pos)
        {
            if (nonbasic[pos] % H.world == H.rank)
            {
                slot[static_cast<size_t>(pos)] = static_cast<int>(own.size());
                own.push_back(static_cast<int>(pos));
            }
        }
        const int cnt = static_cast<int>(own.size());
        own.push_back(0);  // room for the one entry a pivot may add
        cudaStream_t st = ctx->stream;
        CU_TRY(ctx, cudaMemcpyAsync(H.own_pos, own.data(), own.size() * sizeof(int), cudaMemcpyHostToDevice, st));
        CU_TRY(ctx, cudaMemcpyAsync(H.pos_slot, slot.data(), slot.size() * sizeof(int), cudaMemcpyHostToDevice, st));
        CU_TRY(ctx, cudaMemcpyAsync(H.own_cnt, &cnt, sizeof(int), cudaMemcpyHostToDevice, st));
        // Simplex.cpp:271: every call starts with all substitutions cleared
        CU_TRY(ctx, cudaMemsetAsync(ctx->S.flip, 0, static_cast<size_t>(ctx->N_loaded), st));
        CU_TRY(ctx, cudaStreamSynchronize(st));
        return SB200_OK;
    }
}  // namespace

int sb200_shard_prepare(sb200_ctx * ctx, const int64_t * basic, int64_t nb, const int64_t * nonbasic, int64_t nn,
                        double * diag_partial)
{
    if (!ctx || !basic || !diag_partial) return SB200_ERR_INVALID;
    if (!ctx->sharded) return fail(ctx, SB200_ERR_STATE, "sb200_shard_prepare: call sb200_load_shard first");
    CU_TRY(ctx, cudaSetDevice(ctx->opts.device));
    int rc = upload_lists(ctx, basic, nb, nonbasic, nn);
    if (rc != SB200_OK) return rc;
    rc = upload_owned_positions(ctx, nonbasic, nn);
    if (rc != SB200_OK) return rc;
    DevState & S = ctx->S;
    cudaStream_t st = ctx->stream;
    CU_TRY(ctx, cudaMemsetAsync(ctx->flags, 0, 4 * sizeof(int), st));
    // flags, mail and vector slots back to 0: the sequence numbers restart with this call's iteration count
    CU_TRY(ctx, cudaMemsetAsync(ctx->window, 0, window_bytes(S.ld), st));
    k_shard_diag<<<(S.m * 32 + 255) / 256, 256, 0, st>>>(S, ctx->H, ctx->diag, ctx->flags);
    ctx->launches += 1;
    int not_unit = 0;
    CU_TRY(ctx, cudaMemcpyAsync(&not_unit, ctx->flags, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU_TRY(ctx, cudaMemcpyAsync(diag_partial, ctx->diag, S.m * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU_TRY(ctx, cudaStreamSynchronize(st));
    CU_TRY(ctx, cudaGetLastError());
    if (not_unit)
        return fail(ctx, SB200_ERR_UNSUPPORTED,
                    "sharded mode starts from a unit (slack/artificial) basis only: a basic column owned by this rank "
                    "is not a multiple of a unit vector");
    return SB200_OK;
}

int sb200_shard_finish_prepare(sb200_ctx * ctx, const double * diag)
{
    if (!ctx || !diag) return SB200_ERR_INVALID;
    if (!ctx->sharded) return fail(ctx, SB200_ERR_STATE, "sb200_shard_finish_prepare: call sb200_load_shard first");
    CU_TRY(ctx, cudaSetDevice(ctx->opts.device));
    DevState & S = ctx->S;
    cudaStream_t st = ctx->stream;
    CU_TRY(ctx, cudaMemcpyAsync(ctx->diag, diag, S.m * sizeof(double), cudaMemcpyHostToDevice, st));
    CU_TRY(ctx, cudaMemsetAsync(S.Binv, 0, static_cast<size_t>(S.ld) * std::max(ctx->H.nrows, 1) * sizeof(double), st));
    k_shard_init_rows<<<(ctx->H.nrows + 255) / 256 + 1, 256, 0, st>>>(S, ctx->H, ctx->diag);
    ctx->launches += 1;
    Scalars init{};
    init.status = TERM_RUNNING;
    init.e_pos = -1;
    init.iter_limit = ctx->opts.iter_limit;
    CU_TRY(ctx, cudaMemcpyAsync(S.sc, &init, sizeof(Scalars), cudaMemcpyHostToDevice, st));
    CU_TRY(ctx, cudaStreamSynchronize(st));
    CU_TRY(ctx, cudaGetLastError());
    ctx->launch_seq = 0;
    ctx->prepared = true;
    return SB200_OK;
}

// ---- sharded mode, any start basis / re-inversion. B^-1 is row-sharded, but inverting a basis is a one-off of
// O(m^3) on m x m doubles (512 MB at m = 8192): every rank gets the whole A_B (each basic column from its owner,
// gathered by the host), runs the single-GPU Gauss-Jordan on it -- the same kernels on the same bits, hence the
// same inverse on every rank and on one GPU -- and keeps its own rows.
int sb200_shard_prepare_general(sb200_ctx * ctx, const int64_t * basic, int64_t nb, const int64_t * nonbasic, int64_t nn)
{
    if (!ctx || !basic || (!nonbasic && nn > 0)) return SB200_ERR_INVALID;
    if (!ctx->sharded) return fail(ctx, SB200_ERR_STATE, "sb200_shard_prepare_general: call sb200_load_shard first");
    CU_TRY(ctx, cudaSetDevice(ctx->opts.device));
    int rc = upload_lists(ctx, basic, nb, nonbasic, nn);
    if (rc != SB200_OK) return rc;
    rc = upload_owned_positions(ctx, nonbasic, nn);
    if (rc != SB200_OK) return rc;
    DevState & S = ctx->S;
    cudaStream_t st = ctx->stream;
    CU_TRY(ctx, cudaMemsetAsync(ctx->window, 0, window_bytes(S.ld), st));
    Scalars init{};
    init.status = TERM_RUNNING;
    init.e_pos = -1;
    init.iter_limit = ctx->opts.iter_limit;
    CU_TRY(ctx, cudaMemcpyAsync(S.sc, &init, sizeof(Scalars), cudaMemcpyHostToDevice, st));
    CU_TRY(ctx, cudaStreamSynchronize(st));
    ctx->launch_seq = 0;
    ctx->prepared = false;  // until sb200_shard_invert has delivered B^-1 and x_B
    return SB200_OK;
}

int sb200_shard_get_basic_columns(sb200_ctx * ctx, double * cols, int64_t * positions, int64_t cap, int64_t * n_owned)
{
    if (!ctx || !n_owned) return SB200_ERR_INVALID;
    if (!ctx->sharded || !ctx->loaded) return fail(ctx, SB200_ERR_STATE, "sb200_shard_get_basic_columns: call sb200_load_shard first");
    CU_TRY(ctx, cudaSetDevice(ctx->opts.device));
    const DevState & S = ctx->S;
    const ShardInfo & H = ctx->H;
    std::vector<int> hb(static_cast<size_t>(S.m));
    CU_TRY(ctx, cudaMemcpy(hb.data(), S.basic, S.m * sizeof(int), cudaMemcpyDeviceToHost));
    int64_t n = 0;
    for (int k = 0; k < S.m; ++k)
    {
        if (hb[static_cast<size_t>(k)] % H.world != H.rank) continue;
        if (cols && positions)
        {
            if (n >= cap) return fail(ctx, SB200_ERR_INVALID, "sb200_shard_get_basic_columns: buffer too small");
            CU_TRY(ctx, cudaMemcpyAsync(cols + n * S.m, S.A + static_cast<size_t>(hb[static_cast<size_t>(k)] / H.world) * S.ld,
                                        S.m * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
            positions[n] = k;
        }
        n += 1;
    }
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    *n_owned = n;
    return SB200_OK;
}

int sb200_shard_invert(sb200_ctx * ctx, const double * AB_colmajor)
{
    if (!ctx || !AB_colmajor) return SB200_ERR_INVALID;
    if (!ctx->sharded || !ctx->loaded) return fail(ctx, SB200_ERR_STATE, "sb200_shard_invert: call sb200_load_shard first");
    CU_TRY(ctx, cudaSetDevice(ctx->opts.device));
    DevState & S = ctx->S;
    const ShardInfo & H = ctx->H;
    const size_t ld = S.ld, m = S.m;
    cudaStream_t st = ctx->stream;
    if (!ctx->W)
    {
        CU_TRY(ctx, dalloc(ctx, &ctx->W, ld * m));
        CU_TRY(ctx, dalloc(ctx, &ctx->V, ld * m));
        CU_TRY(ctx, dalloc(ctx, &ctx->fcol, ld));
        CU_TRY(ctx, dalloc(ctx, &ctx->shard_AB, m * m));
    }
    CU_TRY(ctx, cudaMemsetAsync(ctx->flags, 0, 4 * sizeof(int), st));
    CU_TRY(ctx, cudaMemcpyAsync(ctx->shard_AB, AB_colmajor, m * m * sizeof(double), cudaMemcpyHostToDevice, st));
    const dim3 gl((S.ld + 255) / 256, S.m);
    k_shard_load_basis<<<gl, 256, 0, st>>>(S.m, S.ld, ctx->shard_AB, ctx->W, ctx->V);
    const dim3 ge((S.ld + 2 * UPDATE_THREADS - 1) / (2 * UPDATE_THREADS), (S.m + UPDATE_ROWS - 1) / UPDATE_ROWS);
    for (int k = 0; k < S.m; ++k)
    {
        k_gj_pivot<<<1, 1024, 0, st>>>(S.m, S.ld, k, ctx->W, ctx->V, ctx->fcol, ctx->flags + 1);
        k_gj_elim<<<ge, UPDATE_THREADS, 0, st>>>(S.m, S.ld, k, ctx->W, ctx->V, ctx->fcol, ctx->flags + 1);
    }
    ctx->launches += 1 + 2 * static_cast<int64_t>(S.m);
    int singular = 0;
    CU_TRY(ctx, cudaMemcpyAsync(&singular, ctx->flags + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU_TRY(ctx, cudaStreamSynchronize(st));
    if (singular) return fail(ctx, SB200_ERR_SINGULAR, "basis matrix is singular");
    if (H.nrows > 0)
    {
        // this rank's rows of the inverse, and x_B = B^-1 b on them (the expression of the single-GPU path)
        CU_TRY(ctx, cudaMemcpyAsync(S.Binv, ctx->V + static_cast<size_t>(H.row0) * ld, ld * H.nrows * sizeof(double),
                                    cudaMemcpyDeviceToDevice, st));
        const int wpb = FTRAN_THREADS / 32;
        const int grid = std::max(1, std::min(2 * ctx->num_sms, (H.nrows + wpb - 1) / wpb));
        k_shard_rows_dot<<<grid, FTRAN_THREADS, ld * sizeof(double), st>>>(H.nrows, S.ld, S.Binv, S.b, S.xB);
        ctx->launches += 1;
    }
    CU_TRY(ctx, cudaStreamSynchronize(st));
    CU_TRY(ctx, cudaGetLastError());
    ctx->reinversions += 1;
    ctx->prepared = true;
    return SB200_OK;
}

int sb200_shard_resume(sb200_ctx * ctx, int64_t new_iter_limit)
{
    if (!ctx || new_iter_limit <= 0) return SB200_ERR_INVALID;
    if (!ctx->sharded || !ctx->prepared) return fail(ctx, SB200_ERR_STATE, "sb200_shard_resume: no run to resume");
    CU_TRY(ctx, cudaSetDevice(ctx->opts.device));
    int rc = fetch_scalars(ctx);
    if (rc != SB200_OK) return rc;
    if (ctx->h_sc->status != TERM_ITER_LIMIT && ctx->h_sc->status != TERM_RUNNING)
        return fail(ctx, SB200_ERR_STATE, "sb200_shard_resume: the last run did not stop at its iteration limit");
    if (new_iter_limit <= ctx->h_sc->iterations)
        return fail(ctx, SB200_ERR_INVALID, "sb200_shard_resume: the new limit must lie beyond the iterations already made");
    ctx->opts.iter_limit = new_iter_limit;
    const long long lim = new_iter_limit;
    const int running = TERM_RUNNING;
    CU_TRY(ctx, cudaMemcpyAsync(reinterpret_cast<char *>(ctx->S.sc) + offsetof(Scalars, iter_limit), &lim, sizeof(long long),
                                cudaMemcpyHostToDevice, ctx->stream));
    CU_TRY(ctx, cudaMemcpyAsync(reinterpret_cast<char *>(ctx->S.sc) + offsetof(Scalars, status), &running, sizeof(int),
                                cudaMemcpyHostToDevice, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return SB200_OK;
}

int sb200_shard_set_bounds(sb200_ctx * ctx, const double * ub)
{
    if (!ctx) return SB200_ERR_INVALID;
    if (!ctx->sharded || !ctx->loaded) return fail(ctx, SB200_ERR_STATE, "sb200_shard_set_bounds: call sb200_load_shard first");
    CU_TRY(ctx, cudaSetDevice(ctx->opts.device));
    DevState & S = ctx->S;
    if (!ub)
    {
        S.ub = nullptr;  // plain ratio test again (the allocation stays with the context)
        return SB200_OK;
    }
    if (!ctx->shard_ub)
    {
        CU_TRY(ctx, dalloc(ctx, &ctx->shard_ub, static_cast<size_t>(ctx->N_loaded)));
        CU_TRY(ctx, dalloc(ctx, &S.neg, static_cast<size_t>(ctx->N_loaded)));
    }
    S.ub = ctx->shard_ub;
    CU_TRY(ctx, cudaMemcpyAsync(S.ub, ub, static_cast<size_t>(ctx->N_loaded) * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CU_TRY(ctx, cudaMemsetAsync(S.neg, 0, static_cast<size_t>(ctx->N_loaded), ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return SB200_OK;
}

int sb200_shard_set_costs(sb200_ctx * ctx, int64_t N_new, const double * c)
{
    if (!ctx || !c) return SB200_ERR_INVALID;
    if (!ctx->sharded || !ctx->loaded) return fail(ctx, SB200_ERR_STATE, "sb200_shard_set_costs: call sb200_load_shard first");
    if (N_new < ctx->S.m || N_new > ctx->N_loaded)
        return fail(ctx, SB200_ERR_INVALID, "sb200_shard_set_costs: need m <= N_new <= loaded N");
    CU_TRY(ctx, cudaSetDevice(ctx->opts.device));
    // the cyclic column shards make dropping the trailing columns a matter of the bound alone: no position of the
    // new non-basic list names a dropped column, so no rank ever prices one
    ctx->S.N = static_cast<int>(N_new);
    ctx->S.nnb = static_cast<int>(N_new - ctx->S.m);
    CU_TRY(ctx, cudaMemcpyAsync(ctx->S.c, c, N_new * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return SB200_OK;
}

int sb200_shard_continue(sb200_ctx * ctx, const int64_t * basic, int64_t nb, const int64_t * nonbasic, int64_t nn)
{
    if (!ctx || !basic || (!nonbasic && nn > 0)) return SB200_ERR_INVALID;
    if (!ctx->sharded || !ctx->prepared)
        return fail(ctx, SB200_ERR_STATE, "sb200_shard_continue: needs the state a previous sb200_shard_run left behind");
    CU_TRY(ctx, cudaSetDevice(ctx->opts.device));
    DevState & S = ctx->S;
    if (nb != S.m) return fail(ctx, SB200_ERR_BASIS_SIZE, "Basis size must be equal to M=" + std::to_string(S.m));
    std::vector<int> prev(static_cast<size_t>(S.m));
    CU_TRY(ctx, cudaMemcpy(prev.data(), S.basic, S.m * sizeof(int), cudaMemcpyDeviceToHost));
    for (int i = 0; i < S.m; ++i)
        if (prev[static_cast<size_t>(i)] != basic[i])
            return fail(ctx, SB200_ERR_INVALID, "sb200_shard_continue: B^-1 and x_B are carried over, so the basic list must be "
                                                "the one the previous run ended on");
    int rc = upload_lists(ctx, basic, nb, nonbasic, nn);
    if (rc != SB200_OK) return rc;
    rc = upload_owned_positions(ctx, nonbasic, nn);
    if (rc != SB200_OK) return rc;
    cudaStream_t st = ctx->stream;
    CU_TRY(ctx, cudaMemsetAsync(ctx->window, 0, window_bytes(S.ld), st));
    Scalars init{};
    init.status = TERM_RUNNING;
    init.e_pos = -1;
    init.iter_limit = ctx->opts.iter_limit;
    CU_TRY(ctx, cudaMemcpyAsync(S.sc, &init, sizeof(Scalars), cudaMemcpyHostToDevice, st));
    CU_TRY(ctx, cudaStreamSynchronize(st));
    ctx->launch_seq = 0;
    return SB200_OK;
}

int sb200_shard_residual_partial(sb200_ctx * ctx, const double * xB_full, double * partial)
{
    if (!ctx || !xB_full || !partial) return SB200_ERR_INVALID;
    if (!ctx->sharded || !ctx->prepared) return fail(ctx, SB200_ERR_STATE, "sb200_shard_residual_partial: no prepared state");
    CU_TRY(ctx, cudaSetDevice(ctx->opts.device));
    DevState & S = ctx->S;
    cudaStream_t st = ctx->stream;
    // S.d (m doubles, scratch between runs) receives x, S.y's neighbour prow the result
    CU_TRY(ctx, cudaMemcpyAsync(S.d, xB_full, S.m * sizeof(double), cudaMemcpyHostToDevice, st));
    k_shard_residual_partial<<<(S.m + 127) / 128, 128, 0, st>>>(S, ctx->H, S.d, S.prow);
    ctx->launches += 1;
    CU_TRY(ctx, cudaMemcpyAsync(partial, S.prow, S.m * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU_TRY(ctx, cudaStreamSynchronize(st));
    CU_TRY(ctx, cudaGetLastError());
    return SB200_OK;
}

int sb200_shard_run(sb200_ctx * ctx, int64_t * basic, int64_t * nonbasic, sb200_result * out)
{
    if (!ctx || !basic || !out) return SB200_ERR_INVALID;
    if (!ctx->sharded || !ctx->prepared) return fail(ctx, SB200_ERR_STATE, "sb200_shard_run: prepare first");
    if (!ctx->connected) return fail(ctx, SB200_ERR_STATE, "sb200_shard_run: call sb200_shard_connect first");
    if (!ctx->coop_supported) return fail(ctx, SB200_ERR_UNSUPPORTED, "device lacks cooperative launch");
    CU_TRY(ctx, cudaSetDevice(ctx->opts.device));
    DevState S = ctx->S;
    const int grid = ctx->shard_grid > 0 ? ctx->shard_grid : ctx->num_sms;
    const size_t smem = sharded_smem_bytes(S.ld, ctx->H.nrows, grid);
    if (smem > static_cast<size_t>(ctx->max_smem_optin) - 2048)
        return fail(ctx, SB200_ERR_UNSUPPORTED, "sb200_shard_run: the staged vectors and the CTA's rows do not fit shared memory "
                                                "(too many rows per CTA for this many SMs per rank)");
    long long chunk = std::max<int64_t>(1, ctx->opts.chunk_iters);
    {
        const char * pm = std::getenv("SB200_PRICE_MODE");
        // 0: static sweep only; k > 0: the last 1/2^k of the pricing groups form a shared ticket pool (default 1/8)
        S.price_mode = pm ? std::max(0, std::min(6, std::atoi(pm))) : 3;
        // L2 prefetch of the first pricing wave behind the exchanges: about 8 warps x 148 CTAs x 8 m bytes
        // (78 MB at m = 8192) fit the 126 MB L2 next to the B^-1 shard's working set
        const char * pc = std::getenv("SB200_PREFETCH_COLS");
        const bool hbm_bound = static_cast<size_t>(S.ld) * ctx->H.ncols * sizeof(double) > (size_t(256) << 20);
        S.prefetch_cols = pc ? std::atoi(pc) : (hbm_bound ? 8 : 0);
        // rows of this rank's B^-1 shard kept in L2 (about 75 MB of them, as on one GPU; at 8 GPUs the whole shard)
        const char * kr = std::getenv("SB200_L2_KEEP_ROWS");
        const int own_rows = (ctx->H.nrows + grid - 1) / grid;
        const int keep_auto = static_cast<int>((size_t(75) << 20) / (static_cast<size_t>(S.ld) * sizeof(double) * grid));
        S.keep_rows = kr ? std::max(0, std::atoi(kr)) : (hbm_bound ? std::min(own_rows, keep_auto) : 0);
    }
    CU_TRY(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    for (;;)
    {
        CU_TRY(ctx, cudaMemsetAsync(reinterpret_cast<char *>(S.sc) + offsetof(Scalars, barrier_count), 0,
                                    2 * sizeof(unsigned long long), ctx->stream));
        ctx->launch_seq += 1;
        ShardInfo H = ctx->H;
        H.launch_seq = ctx->launch_seq;
        H.timeout_ns = XCHG_TIMEOUT_NS;
        if (const char * tmo = std::getenv("SB200_XCHG_TIMEOUT_MS")) H.timeout_ns = 1000000ULL * std::strtoull(tmo, nullptr, 10);
        void * args[] = { &S, &H, &chunk };
        CU_TRY(ctx, cudaLaunchCooperativeKernel(S.ub != nullptr ? (void *)k_sharded_fused_bounded : (void *)k_sharded_fused,
                                                dim3(grid), dim3(SHARDED_THREADS), args, smem, ctx->stream));
        ctx->launches += 1;
        int rc = fetch_scalars(ctx);
        if (rc != SB200_OK) return rc;
        if (ctx->h_sc->status != TERM_RUNNING) break;
    }
    CU_TRY(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    CU_TRY(ctx, cudaEventSynchronize(ctx->ev1));
    CU_TRY(ctx, cudaGetLastError());
    if (ctx->h_sc->status == TERM_EXCHANGE_TIMEOUT)
        return fail(ctx, SB200_ERR_CUDA, "sharded run: a peer did not answer within the exchange timeout");
    float ms = 0.f;
    CU_TRY(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    const ShardInfo & H = ctx->H;
    std::vector<int> hb(S.m), hn(S.nnb);
    std::vector<double> hx(std::max(H.nrows, 1)), hc(S.N);
    CU_TRY(ctx, cudaMemcpyAsync(hb.data(), S.basic, S.m * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    if (S.nnb > 0)
        CU_TRY(ctx, cudaMemcpyAsync(hn.data(), S.nonbasic, S.nnb * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    if (H.nrows > 0)
        CU_TRY(ctx, cudaMemcpyAsync(hx.data(), S.xB, H.nrows * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaMemcpyAsync(hc.data(), S.c, S.N * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    double obj = 0.0;  // this rank's rows only: the caller adds the ranks up
    for (int li = 0; li < H.nrows; ++li) obj += hc[hb[H.row0 + li]] * hx[li];
    for (int i = 0; i < S.m; ++i) basic[i] = hb[i];
    if (nonbasic)
        for (int i = 0; i < S.nnb; ++i) nonbasic[i] = hn[i];
    if (ctx->h_sc->status == TERM_OPTIMAL || ctx->h_sc->status == TERM_UNBOUNDED || ctx->h_sc->status == TERM_ITER_LIMIT)
        ctx->carry_basic.assign(basic, basic + S.m);  // B^-1 and x_B on the device belong to this basis
    out->term = ctx->h_sc->status;
    out->reserved = 0;
    out->iterations = ctx->h_sc->iterations;
    out->pivots = ctx->h_sc->pivots;
    out->bound_flips = ctx->h_sc->flips;
    out->objective = obj;
    out->loop_seconds = ms * 1e-3;
    return SB200_OK;
}

int sb200_shard_get_xB_local(sb200_ctx * ctx, double * xB_local, int64_t * row0, int64_t * nrows)
{
    if (!ctx || !row0 || !nrows) return SB200_ERR_INVALID;
    if (!ctx->sharded || !ctx->prepared) return fail(ctx, SB200_ERR_STATE, "sb200_shard_get_xB_local: no prepared state");
    CU_TRY(ctx, cudaSetDevice(ctx->opts.device));
    *row0 = ctx->H.row0;
    *nrows = ctx->H.nrows;
    if (xB_local && ctx->H.nrows > 0)
    {
        CU_TRY(ctx, cudaMemcpyAsync(xB_local, ctx->S.xB, ctx->H.nrows * sizeof(double), cudaMemcpyDeviceToHost,
                                    ctx->stream));
        CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return SB200_OK;
}

double sb200_algorithmic_bytes_per_pivot(int64_t m, int64_t n_nonbasic)
{
    return 8.0 * (static_cast<double>(m) * static_cast<double>(n_nonbasic) + 4.0 * static_cast<double>(m) * m);
}

int64_t sb200_launch_count(const sb200_ctx * ctx) { return ctx ? ctx->launches : 0; }

int sb200_inverse_carried(const sb200_ctx * ctx) { return (ctx && ctx->carried) ? 1 : 0; }

int sb200_engine_used(const sb200_ctx * ctx) { return ctx ? ctx->engine_used : SB200_ENGINE_USED_NONE; }

int sb200_primal_residual(sb200_ctx * ctx, double * max_abs)
{
    if (!ctx || !max_abs) return SB200_ERR_INVALID;
    if (!ctx->prepared) return fail(ctx, SB200_ERR_STATE, "sb200_primal_residual: no basis prepared");
    if (ctx->sharded) return fail(ctx, SB200_ERR_STATE, "sb200_primal_residual: sharded context (use sb200_shard_residual)");
    CU_TRY(ctx, cudaSetDevice(ctx->opts.device));
    double rp = 0.0, bm = 0.0, ri = 0.0;
    int rc = hygiene_measure(ctx, 0, 0, &rp, &bm, &ri);
    if (rc != SB200_OK) return rc;
    *max_abs = (rp >= DBL_MAX) ? std::nan("") : rp;
    return SB200_OK;
}

int sb200_inverse_residual(sb200_ctx * ctx, int64_t n_rows, int64_t first, double * max_abs)
{
    if (!ctx || !max_abs || n_rows <= 0) return SB200_ERR_INVALID;
    if (!ctx->prepared) return fail(ctx, SB200_ERR_STATE, "sb200_inverse_residual: no basis prepared");
    if (ctx->sharded) return fail(ctx, SB200_ERR_STATE, "sb200_inverse_residual: sharded context");
    CU_TRY(ctx, cudaSetDevice(ctx->opts.device));
    double rp = 0.0, bm = 0.0, ri = 0.0;
    int rc = hygiene_measure(ctx, static_cast<int>(std::min<int64_t>(n_rows, 64)), static_cast<int>(first % (1 << 30)), &rp, &bm, &ri);
    if (rc != SB200_OK) return rc;
    *max_abs = (ri >= DBL_MAX) ? std::nan("") : ri;
    return SB200_OK;
}

int sb200_hygiene_stats(const sb200_ctx * ctx, int64_t * checks, int64_t * refreshes, double * worst_r)
{
    if (!ctx) return SB200_ERR_INVALID;
    if (checks) *checks = ctx->hyg_checks;
    if (refreshes) *refreshes = ctx->hyg_refreshes;
    if (worst_r) *worst_r = ctx->hyg_worst;
    return SB200_OK;
}

int64_t sb200_reinversion_count(const sb200_ctx * ctx) { return ctx ? ctx->reinversions : 0; }

int sb200_debug_check_guards(sb200_ctx * ctx, int64_t * n_buffers, int64_t * n_damaged)
{
    if (!ctx || !n_damaged) return SB200_ERR_INVALID;
    CU_TRY(ctx, cudaSetDevice(ctx->opts.device));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    int64_t bad = 0;
    std::vector<unsigned char> h(2 * GUARD_BYTES);
    for (const auto & g : ctx->guarded)
    {
        CU_TRY(ctx, cudaMemcpy(h.data(), g.base, GUARD_BYTES, cudaMemcpyDeviceToHost));
        CU_TRY(ctx, cudaMemcpy(h.data() + GUARD_BYTES, g.base + GUARD_BYTES + g.payload, GUARD_BYTES, cudaMemcpyDeviceToHost));
        bool damaged = false;
        for (unsigned char v : h) damaged = damaged || v != static_cast<unsigned char>(GUARD_PATTERN);
        bad += damaged ? 1 : 0;
    }
    if (n_buffers) *n_buffers = static_cast<int64_t>(ctx->guarded.size());
    *n_damaged = bad;
    return SB200_OK;
}

int sb200_get_phase_cycles(sb200_ctx * ctx, int64_t * out8)
{
    if (!ctx || !out8) return SB200_ERR_INVALID;
    if (!ctx->prepared) return fail(ctx, SB200_ERR_STATE, "sb200_get_phase_cycles: no prepared state");
    CU_TRY(ctx, cudaSetDevice(ctx->opts.device));
    int rc = fetch_scalars(ctx);
    if (rc != SB200_OK) return rc;
    for (int i = 0; i < 8; ++i) out8[i] = ctx->h_sc->phase_cycles[i];
    return SB200_OK;
}

}  // extern "C"

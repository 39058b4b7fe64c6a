/*
 * simplex_b200 — C ABI of the B200-native revised-simplex hot path.
 *
 * This header is the drop-in boundary for ONE path of maciejdrwal/simplex: the body of
 *     int simplex::Simplex::simplex(Basis & arg_basis)         (reference src/Simplex.h:25,
 *                                                               src/Simplex.cpp:255-353)
 * which the reference reaches twice per solve (src/InitialBasis.cpp:82 for Phase I,
 * src/Simplex.cpp:41 for Phase II). The reference has no FFI of its own; a maintainer
 * replaces the loop body by a ~30-line marshal + the calls below (INTEGRATION.md shows it).
 *
 * Conventions
 *   - plain C types only; the caller owns every host buffer, the library owns all device
 *     memory; nothing throws across this boundary; every entry point returns an sb200_rc
 *     (0 = ok) and sb200_last_error() gives the text of the last failure on that context.
 *   - one context is used from one host thread at a time (the reference is single-threaded).
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with
 *     SB200_ERR_CUDA.
 *   - matrices are dense fp64. A is column-major with leading dimension lda (exactly the
 *     storage of the reference's Eigen::MatrixXd matrix_A, src/LinearProgram.h:89) and is
 *     the POST-SCALING tableau the reference loop sees (src/LinearProgram.cpp:171-195).
 *   - index lists are int64 (Eigen::Index, src/Basis.h:32-33); ORDER IS SIGNIFICANT: entering
 *     candidates are keyed by POSITION in the non-basic list, leaving candidates by basis row
 *     (src/Simplex.cpp:49-61, 66-84, 87-102; src/Basis.cpp:48-53).
 */
#ifndef SIMPLEX_B200_H
#define SIMPLEX_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SB200_ABI_VERSION 2

typedef struct sb200_ctx sb200_ctx;

/* return codes */
typedef enum sb200_rc {
    SB200_OK = 0,
    SB200_ERR_INVALID = 1,    /* bad argument */
    SB200_ERR_CUDA = 2,       /* CUDA runtime failure or no device */
    SB200_ERR_NOMEM = 3,      /* allocation failed */
    SB200_ERR_STATE = 4,      /* call order wrong (e.g. run before load) */
    SB200_ERR_BASIS_SIZE = 5, /* basic list size != M: the reference throws std::string here
                                 (src/Simplex.cpp:260-263) */
    SB200_ERR_SINGULAR = 6,   /* the supplied basis matrix could not be inverted */
    SB200_ERR_UNSUPPORTED = 7
} sb200_rc;

/* entering-variable rule (src/Simplex.cpp:247-251) */
typedef enum sb200_pricing {
    SB200_PRICE_FIRST_ELIGIBLE = 0, /* stock reference: first POSITION with s < -eps (:49-61) */
    SB200_PRICE_MOST_NEGATIVE = 1   /* "Dantzig": most negative s, strict <, lowest position wins
                                       ties (:87-102; disabled by a comment at :250) */
} sb200_pricing;

/* how the loop ended (the reference only logs these, src/Simplex.cpp:319-338, 347-350) */
typedef enum sb200_term {
    SB200_TERM_RUNNING = 0,
    SB200_TERM_OPTIMAL = 1,    /* no eligible entering position (:319-323) */
    SB200_TERM_UNBOUNDED = 2,  /* ratio test found no row (:334-338) */
    SB200_TERM_ITER_LIMIT = 3  /* iteration cap hit; the reference throws const char* (:347-350) */
} sb200_term;

/* loop engines; both run the same device functions */
typedef enum sb200_engine {
    SB200_ENGINE_PERSISTENT = 0, /* one cooperative persistent kernel, grid barriers between phases */
    SB200_ENGINE_STAGED = 1,     /* one kernel per phase (K1..K5), stream ordered, CUDA-graph replayed */
    SB200_ENGINE_STREAMING = 2   /* SB200_ENGINE_PERSISTENT without its shared-memory-resident kernel: the
                                    tableau is streamed from HBM / L2 every pivot whatever its size */
} sb200_engine;

/* how the duals are kept */
typedef enum sb200_dual_mode {
    SB200_DUAL_RECOMPUTE = 0, /* y = c_B^T B^-1 every iteration (src/Simplex.cpp:307) */
    SB200_DUAL_UPDATE = 1     /* y += (s_q/d_p) * (pivot row); recomputed from scratch at every device launch
                                 (every chunk_iters iterations) by the persistent engines, every dual_refresh
                                 pivots by the staged engine. Default: it is what lets the fused and the
                                 shared-memory-resident kernels run, and it walks the same pivot sequence */
} sb200_dual_mode;

typedef struct sb200_opts {
    int32_t abi_version;     /* SB200_ABI_VERSION */
    int32_t device;          /* CUDA device ordinal */
    int32_t pricing;         /* sb200_pricing */
    int32_t engine;          /* sb200_engine */
    int32_t dual_mode;       /* sb200_dual_mode */
    int32_t dual_refresh;    /* staged engine, SB200_DUAL_UPDATE: recompute y every this many pivots inside a chunk
                                (0 = only at the start of every chunk); ignored by the persistent engines */
    int32_t reinvert_period; /* recompute B^-1 from A_B every this many pivots (0 = never) */
    int32_t trace_capacity;  /* pivots kept in the device trace ring (0 = no trace) */
    double eps;              /* utils::ALMOST_ZERO, src/Utils.h:9 (1e-7) */
    int64_t iter_limit;      /* src/Simplex.cpp:274 (1e9) */
    int64_t chunk_iters;     /* iterations per device launch/graph replay between host polls */
    void * stream;           /* cudaStream_t to run on, NULL = the library's own stream */
    /* column/row sharding of ONE LP over several GPUs; world = 1: not sharded */
    int32_t rank;
    int32_t world;
    /* Numerical hygiene off the hot path (ABI 2). The reference factorises A_B afresh in every iteration
     * (src/Simplex.cpp:305-307); this library carries B^-1, x_B and y by recurrences. Every hygiene_period
     * launches, and always before an OPTIMAL or UNBOUNDED ending is accepted, it measures
     *     r = max( |b - A_B x_B|_inf / (1 + |b|_inf),  max over hygiene_rows sampled rows i of |e_i^T B^-1 A_B - e_i^T|_inf )
     * and, when r > hygiene_tol, recomputes B^-1 from the resident A_B (Gauss-Jordan, partial pivoting) and
     * x_B = B^-1 b and lets the loop go on (an ending is then re-examined with the fresh inverse; the
     * iteration count is not advanced twice). hygiene_tol = 0 switches all of it off. */
    double hygiene_tol;      /* default 1e-9 */
    int32_t hygiene_period;  /* launches between two checks while running (default 16; 0 = only at the ending) */
    int32_t hygiene_rows;    /* rows of B^-1 sampled per check (default 8, at most 64) */
} sb200_opts;

typedef struct sb200_result {
    int32_t term;        /* sb200_term */
    int32_t reserved;
    int64_t iterations;  /* loop entries, including bound-flip iterations and the final
                            optimality-detecting one (src/Simplex.cpp:277) */
    int64_t pivots;      /* basis changes (src/Basis.cpp:48-57) */
    int64_t bound_flips; /* iterations that returned -2 (src/Simplex.cpp:160-164) */
    double objective;    /* c_B . x_B in the scaled, possibly flipped space (before
                            solution_found's shift/sign handling, src/Simplex.cpp:210-211) */
    double loop_seconds; /* device time of the loop (CUDA events), excluding load and B^-1 setup */
} sb200_result;

/* One pivot as the reference logs it (src/Basis.cpp:55-56):
 * "leaving basis: <leaving_col> (<row>), entering basis: <entering_col> (<pos>)" */
typedef struct sb200_pivot {
    int32_t leaving_col;
    int32_t row;
    int32_t entering_col;
    int32_t pos;
} sb200_pivot;

/* ---- lifecycle ------------------------------------------------------------------------- */

/* Fill *o with the reference's behaviour: first-eligible pricing, eps 1e-7, 1e9 iterations; persistent engine,
 * duals by recurrence (refreshed every launch), hygiene checks on. */
void sb200_default_opts(sb200_opts * o);

int sb200_create(sb200_ctx ** out, const sb200_opts * opts);
void sb200_destroy(sb200_ctx * ctx);
const char * sb200_last_error(const sb200_ctx * ctx); /* ctx may be NULL: last create error */
int sb200_abi_version(void);

/* ---- data in: replaces the reads of m_lp.matrix_A / vector_b / vector_c / var_ubnd that the
 *      reference does through `friend class Simplex` (src/LinearProgram.h:44) ---------------- */

/* Host -> HBM, once per phase. ub == NULL: no finite upper bound was parsed
 * (has_non_trivial_upper_bounds() false, src/Simplex.cpp:108) -> plain ratio test. Otherwise
 * ub[j] is var_ubnd[j], DBL_MAX meaning "none" (src/LinearProgram.cpp:225). */
int sb200_load(sb200_ctx * ctx, int64_t m, int64_t N, const double * A_colmajor, int64_t lda, const double * b,
               const double * c, const double * ub_or_null);

/* Same, from device pointers on ctx's device (synthetic data generated in HBM; copies D2D). */
int sb200_load_device(sb200_ctx * ctx, int64_t m, int64_t N, const double * dA_colmajor, int64_t lda,
                      const double * db, const double * dc, const double * dub_or_null);

/* Replace c only (Phase I -> Phase II keeps A resident: artificial columns are a trailing block,
 * src/InitialBasis.cpp:45-48). N_new <= N drops trailing columns. */
int sb200_set_costs(sb200_ctx * ctx, int64_t N_new, const double * c);

/* ---- the seam: replaces the body of Simplex::simplex(Basis&) ------------------------------- */

/* basic[m] / nonbasic[N-m] in and out (Basis, src/Basis.h:18-34). n_basic must equal m.
 * Returns SB200_OK for every regular ending (optimal / unbounded / iteration limit): look at
 * out->term. */
int sb200_run(sb200_ctx * ctx, int64_t * basic, int64_t n_basic, int64_t * nonbasic, int64_t n_nonbasic,
              sb200_result * out);

/* ---- read-out needed by solution_found (src/Simplex.cpp:181-245) --------------------------- */
int sb200_get_xB(sb200_ctx * ctx, double * xB /* [m] */);
int sb200_get_y(sb200_ctx * ctx, double * y /* [m] */);
int sb200_get_d(sb200_ctx * ctx, double * d /* [m], last FTRAN column */);
int sb200_get_flip_flags(sb200_ctx * ctx, uint8_t * flags /* [N], ub_substitutions */);
int sb200_get_Binv(sb200_ctx * ctx, double * Binv_rowmajor /* [m*m] */);
int sb200_get_b(sb200_ctx * ctx, double * b /* [m], after bound flips */);
/* Copies up to cap pivots of the last run, oldest first; *n_total = pivots made. */
int sb200_get_trace(sb200_ctx * ctx, sb200_pivot * out, int64_t cap, int64_t * n_total);

/* ---- single-phase entry points (each wraps exactly one kernel of the pivot; used by the
 *      parity tests and by profiling; state must have been set up by sb200_load + sb200_prepare) */
int sb200_prepare(sb200_ctx * ctx, const int64_t * basic, int64_t n_basic, const int64_t * nonbasic,
                  int64_t n_nonbasic);                  /* K0: B^-1, x_B = B^-1 b, index maps */
int sb200_step_dual(sb200_ctx * ctx);                    /* K1: y = c_B^T B^-1 */
int sb200_step_price(sb200_ctx * ctx, int64_t * pos, double * value); /* K2: pos = -1 if none */
int sb200_step_ftran(sb200_ctx * ctx, int64_t pos);      /* K3: d = B^-1 a_q + ratio candidates */
int sb200_step_ratio(sb200_ctx * ctx, int64_t pos, int64_t * row, double * theta); /* K4: -1 / -2 as the reference */
int sb200_step_update(sb200_ctx * ctx);                  /* K5: rank-1 update of B^-1, x_B, lists */

/* ---- batched mode: many independent small LPs, one CTA per LP (north_star, config 5) -------
 * All LPs share (m, N). A: [batch][N][m] column-major per LP, b: [batch][m], c: [batch][N].
 * basic/nonbasic: [batch][m] / [batch][N-m] in/out (int32). No upper bounds. */
int sb200_run_batched(sb200_ctx * ctx, int64_t batch, int64_t m, int64_t N, const double * A, const double * b,
                      const double * c, int32_t * basic, int32_t * nonbasic, double * xB /* [batch][m] out */,
                      double * objective /* [batch] out */, int32_t * term /* [batch] out */,
                      int32_t * iterations /* [batch] out */, int device_resident /* pointers are device */,
                      double * seconds /* out: device time */);

/* The same with finite upper bounds (reference src/Simplex.cpp:116-179): ub[batch][N], DBL_MAX = no bound
 * (src/LinearProgram.cpp:225). Runs the shared-memory kernel (the whole tableau of an LP is rewritten in place by a
 * substitution, src/LinearProgram.cpp:154-169). Out: flip_flags[batch][N] = ub_substitutions at the end of every LP,
 * bound_flips[batch] = iterations that moved the entering variable to its bound instead of pivoting. */
int sb200_run_batched_bounded(sb200_ctx * ctx, int64_t batch, int64_t m, int64_t N, const double * A, const double * b,
                              const double * c, const double * ub, int32_t * basic, int32_t * nonbasic, double * xB,
                              double * objective, int32_t * term, int32_t * iterations, uint8_t * flip_flags,
                              int32_t * bound_flips, int device_resident, double * seconds);

/* ---- sharded mode (config 4): peer windows for the per-pivot NVLink exchanges ----------------
 * Each rank creates a context with (rank, world), queries the size/handle of its exchange
 * window, the host exchanges the 64-byte handles (torch.distributed) and hands all of them back. */
int sb200_shard_window_handle(sb200_ctx * ctx, void * handle64 /* out, 64 bytes */);
int sb200_shard_connect(sb200_ctx * ctx, const void * handles /* world * 64 bytes */);
/* The same for ranks that live in ONE process (one context per rank, one host thread per rank calling
 * sb200_shard_run at the same time): peers[r] = rank r's context, windows are wired directly. Ranks may sit on
 * different devices (peer access is enabled) or SHARE a device, in which case every rank runs on its share of
 * the SMs -- which is how the G-independence of the pivot sequence is tested on a one-GPU box. */
int sb200_shard_connect_local(sb200_ctx * ctx, sb200_ctx * const * peers, int64_t n_peers);
/* Column-shard load: this rank holds the CYCLIC column set {rank + k*world} of the N-column tableau
 * (col0 = rank, ncols = ceil((N - rank)/world); A_cols holds them in that order, each column
 * contiguous). Cyclic rather than blocked so that structural and slack columns, hence the pricing
 * work, stay balanced over the ranks. b and c are replicated. B^-1 and x_B are row-sharded by the
 * library (blocks of round_up(ceil(m/world), 32) rows). Call order per rank:
 *   create(rank, world) -> load_shard -> window_handle -> [all-gather the handles] -> connect
 *   -> shard_prepare -> [sum diag_partial over the ranks] -> shard_finish_prepare
 *   -> [barrier] -> shard_run.
 * The two bracketed host steps are the only host-side collectives (torch.distributed plumbing);
 * every per-pivot exchange is done by the kernels over NVLink (see csrc/sb200_sharded.cuh). */
int sb200_load_shard(sb200_ctx * ctx, int64_t m, int64_t N, int64_t col0, int64_t ncols,
                     const double * A_cols_colmajor, int64_t lda, const double * b, const double * c,
                     int device_resident);
/* K0 of the sharded mode: uploads the (replicated) lists, resets the exchange window, and returns
 * in diag_partial[m] the pivot entry a_ii of every basic column this rank owns (0 elsewhere).
 * The start basis must consist of unit columns (slack / artificial crash basis). */
int sb200_shard_prepare(sb200_ctx * ctx, const int64_t * basic, int64_t n_basic, const int64_t * nonbasic,
                        int64_t n_nonbasic, double * diag_partial);
int sb200_shard_finish_prepare(sb200_ctx * ctx, const double * diag /* [m], summed over ranks */);
/* The loop. Lists are global and identical on every rank; out->objective is the partial sum over
 * this rank's rows (add the ranks up). */
int sb200_shard_run(sb200_ctx * ctx, int64_t * basic, int64_t * nonbasic, sb200_result * out);
int sb200_shard_get_xB_local(sb200_ctx * ctx, double * xB_local, int64_t * row0, int64_t * nrows);
/* Finite upper bounds on the sharded engine (after sb200_load_shard; ub[N] replicated, DBL_MAX = none; NULL switches back
 * to the plain ratio test): the bounded ratio test and both substitutions of src/Simplex.cpp:116-179 on the shards. A
 * substituted column keeps a sign flag until the exit of the launch rewrites it (as in the fused single-GPU kernel); b is
 * replicated and follows on every rank (the column of a substituted LEAVING variable is pushed by its owner, a third,
 * rare exchange). sb200_get_flip_flags / sb200_get_b read the replicated flags and right-hand side. */
int sb200_shard_set_bounds(sb200_ctx * ctx, const double * ub_or_null);
/* Phase I -> Phase II on the sharded engine (src/InitialBasis.cpp:85-121, src/Simplex.cpp:35-41): the artificial
 * columns are the trailing ids, so sb200_shard_set_costs installs the Phase-II costs and drops them (N_new <= N; the
 * cyclic shards stay where they are), and sb200_shard_continue starts the next sb200_shard_run from the basis the
 * previous one ended on -- B^-1 and x_B are CARRIED OVER (row-sharded as they are; nothing is re-inverted), the
 * non-basic list is the caller's (Phase I's final list without the artificial ids), counters and exchange windows
 * restart. Collective like sb200_shard_prepare: every rank calls both with the same arguments before anyone runs. */
int sb200_shard_set_costs(sb200_ctx * ctx, int64_t N_new, const double * c);
int sb200_shard_continue(sb200_ctx * ctx, const int64_t * basic, int64_t n_basic, const int64_t * nonbasic,
                         int64_t n_nonbasic);
/* ANY start basis, and re-inversion, on the sharded engine. B^-1 is row-sharded, but inverting a basis is a one-off on
 * m x m doubles: every rank receives the whole basis matrix (each basic column from the rank that owns it, gathered by
 * the host), runs the single-GPU Gauss-Jordan on it -- same kernels, same bits, hence the same inverse on every rank and
 * on one GPU -- and keeps its rows. Call order (collective, same arguments on every rank):
 *   sb200_shard_prepare_general(lists)            instead of shard_prepare / finish_prepare (no unit-basis requirement)
 *   sb200_shard_get_basic_columns -> [host: assemble AB, column k = A[:, basic[k]], from all ranks]
 *   sb200_shard_invert(AB)                        B^-1 rows and x_B = B^-1 b of this rank; then sb200_shard_run.
 * Between two runs the last two calls alone rebuild B^-1 and x_B of the current basis (re-inversion). */
int sb200_shard_prepare_general(sb200_ctx * ctx, const int64_t * basic, int64_t n_basic, const int64_t * nonbasic,
                                int64_t n_nonbasic);
/* cols[k][m] = the k-th basic column THIS rank owns, positions[k] = its basis position; cols == NULL: count only. */
int sb200_shard_get_basic_columns(sb200_ctx * ctx, double * cols, int64_t * positions, int64_t cap, int64_t * n_owned);
int sb200_shard_invert(sb200_ctx * ctx, const double * AB_colmajor /* [m][m], leading dimension m */);
/* A sharded run that stopped at its iteration limit goes on (same counters, same exchange sequence numbers) up to
 * new_iter_limit: sb200_shard_run again, no prepare in between. With sb200_shard_invert between the two runs this is the
 * periodic / residual-triggered re-inversion of the sharded mode (simplex_b200/dist.py: reinvert_period, hygiene_tol). */
int sb200_shard_resume(sb200_ctx * ctx, int64_t new_iter_limit);
/* Hygiene on the sharded engine: partial[i] = sum over the basic columns THIS rank owns of A[i][col] * xB_full[k]
 * (xB_full = the whole x_B, gathered from the ranks). Summed over the ranks and subtracted from b it is the primal
 * residual sb200_primal_residual measures on one GPU. */
int sb200_shard_residual_partial(sb200_ctx * ctx, const double * xB_full /* [m] */, double * partial /* [m] out */);

/* ---- numerical hygiene off the hot path. The reference needs none of this: it
 *      factorises the basis from scratch in every iteration (src/Simplex.cpp:305). With
 *      sb200_opts.reinvert_period = P > 0 the loop recomputes B^-1 from the resident A_B (Gauss-Jordan,
 *      partial pivoting) and x_B = B^-1 b every P pivots, between two device launches. */
/* max_i | b_i - (A_B x_B)_i | of the basis left by the last run / prepare. */
int sb200_primal_residual(sb200_ctx * ctx, double * max_abs);
/* max over n_rows sampled rows i (evenly spread, first one = row `first`) of | e_i^T B^-1 A_B - e_i^T |_inf. */
int sb200_inverse_residual(sb200_ctx * ctx, int64_t n_rows, int64_t first, double * max_abs);
/* Hygiene checks made / re-inversions they triggered during the last sb200_run, and the largest r seen. */
int sb200_hygiene_stats(const sb200_ctx * ctx, int64_t * checks, int64_t * refreshes, double * worst_r);
/* Re-inversions done by this context since creation. */
int64_t sb200_reinversion_count(const sb200_ctx * ctx);

/* Out-of-bounds-write check of the library's own: every device buffer of a context lies between two 256-byte
 * guard zones filled with a pattern at allocation. Returns how many buffers there are and how many of them have
 * a damaged zone. (compute-sanitizer is the better tool where it may run; tests call this after every context.) */
int sb200_debug_check_guards(sb200_ctx * ctx, int64_t * n_buffers, int64_t * n_damaged);

/* ---- measurement helpers ---------------------------------------------------------------- */
/* ALGORITHMIC bytes of one pivot, SURVEY.md §8d: 8*(m*N_nb + 4*m^2). */
double sb200_algorithmic_bytes_per_pivot(int64_t m, int64_t n_nonbasic);
/* Number of kernel launches issued by the library on this context since creation. */
int64_t sb200_launch_count(const sb200_ctx * ctx);
/* 1 if the last sb200_run / sb200_prepare started from the very basis the previous run on this context had ended
 * on (no sb200_load in between) and therefore CARRIED B^-1 and x_B OVER instead of inverting the basis again: the
 * Phase II call after a Phase I with A resident (sb200_set_costs). The reference factorises that basis afresh
 * (src/Simplex.cpp:305); here the hygiene gate watches the drift. SB200_NO_CARRY=1 in the environment disables it. */
int sb200_inverse_carried(const sb200_ctx * ctx);
/* Which loop kernel the last sb200_run used. SB200_ENGINE_PERSISTENT picks, in this order: the
 * shared-memory-resident kernel (dual update, and the tableau fits the SMs' shared memory: B^-1 rows +
 * dense columns on chip), the fused HBM-streaming kernel (dual update), else the 4-phase persistent
 * kernel (SB200_DUAL_RECOMPUTE). All three take LPs with or without finite upper bounds. */
typedef enum sb200_engine_used_t {
    SB200_ENGINE_USED_NONE = 0,
    SB200_ENGINE_USED_STAGED = 1,
    SB200_ENGINE_USED_PERSISTENT = 2,
    SB200_ENGINE_USED_FUSED = 3,
    SB200_ENGINE_USED_RESIDENT = 4
} sb200_engine_used_t;
int sb200_engine_used(const sb200_ctx * ctx);
/* Persistent engine: SM-clock cycles CTA 0 spent per phase during the last run, summed over
 * iterations: [0] price [1] barrier [2] ftran+ratio [3] barrier [4] rank-1 update [5] barrier
 * [6] dual reduce + barrier + restage [7] iterations counted. (The reference's only profiling aid
 * is a wall-clock stopwatch around the whole solve, src/Utils.h:11-25.) */
int sb200_get_phase_cycles(sb200_ctx * ctx, int64_t * out8);

#ifdef __cplusplus
}
#endif
#endif /* SIMPLEX_B200_H */

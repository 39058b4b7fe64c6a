/* TEST INFRASTRUCTURE — CPU restatement of the reference's hot path. NOT product code.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this. See simplex_oracle.c for the reference file:line each function follows. */
#ifndef SB200_SIMPLEX_ORACLE_H
#define SB200_SIMPLEX_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORACLE_RUNNING = 0, ORACLE_OPTIMAL = 1, ORACLE_UNBOUNDED = 2, ORACLE_ITER_LIMIT = 3 };
enum { ORACLE_RULE_FIRST = 0, ORACLE_RULE_MOST_NEG = 1 };

typedef struct oracle_result {
    int32_t term;
    int32_t pad;
    int64_t iterations;
    int64_t pivots;
    int64_t flips;
    double objective; /* c_B . x at exit, before solution_found's shifts */
} oracle_result;

/* Simplex::simplex(Basis&) restated with a fresh dense LU every iteration, exactly as the
 * reference does. A (column-major, lda = m), b, c are MUTATED by bound flips like the
 * reference's matrix_A/vector_b/vector_c; ub == NULL means has_non_trivial_upper_bounds()==false.
 * trace: 4 ints per pivot (leaving_col,row,entering_col,pos), up to trace_cap pivots. */
int oracle_simplex_lu(int64_t m, int64_t N, double * A, double * b, double * c, const double * ub, int64_t * basic,
                      int64_t * nonbasic, int rule, double eps, int64_t iter_limit, double * x_out, double * y_out,
                      uint8_t * flip_out, int32_t * trace, int64_t trace_cap, oracle_result * out);

/* Same mathematics with an explicit inverse kept by product-form updates (the algorithm the
 * GPU path runs), for sizes where an O(m^3) iteration is too slow. Parallel over columns with
 * OpenMP when compiled with -fopenmp. Validated against oracle_simplex_lu and oracle/_ref. */
int oracle_simplex_pfi(int64_t m, int64_t N, double * A, double * b, double * c, const double * ub, int64_t * basic,
                       int64_t * nonbasic, int rule, double eps, int64_t iter_limit, double * x_out, double * y_out,
                       uint8_t * flip_out, int32_t * trace, int64_t trace_cap, oracle_result * out);

/* LinearProgram::rescale_matrix (LinearProgram.cpp:171-195) on a column-major m x N tableau. */
void oracle_rescale(int64_t m, int64_t N, double * A, double * b, double * c);

/* LinearProgram::upper_bound_substitution (LinearProgram.cpp:154-169). */
void oracle_upper_bound_substitution(int64_t m, double * A, double * b, double * c, uint8_t * flip, int64_t col,
                                     double u);

/* Synthetic LP G(m,n,n_eq,seed) of SURVEY.md §8d (std::mt19937_64 restated).
 * c[n], a_rowmajor[m*n], rhs[m]. */
void oracle_synth(int64_t m, int64_t n, int64_t n_eq, uint64_t seed, double * c, double * a_rowmajor, double * rhs);

/* Dense standard-form tableau of G as the reference flow builds it: structurals, then one slack
 * per '<=' row in row order (LinearProgram.cpp:110-141), then (phase1 != 0) one +1 artificial per
 * '=' row (InitialBasis.cpp:23-55), rescaled (LinearProgram.cpp:171-195); Phase-I costs are 1 on
 * the artificials (InitialBasis.cpp:73-78). Returns N. A must hold m*(n+m) doubles. The crash
 * basis goes to basic[m] / nonbasic[N-m]. */
int64_t oracle_synth_tableau(int64_t m, int64_t n, int64_t n_eq, uint64_t seed, int phase1, double * A, double * b,
                             double * c, int64_t * basic, int64_t * nonbasic);

#ifdef __cplusplus
}
#endif
#endif

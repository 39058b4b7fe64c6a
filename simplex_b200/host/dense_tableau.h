// Dense standard-form tableau helpers of the host side (product code, C++17).
//
// The reference produces the inputs of its hot loop in LinearProgram::initialize_tableau
// (reference src/LinearProgram.cpp:64-97) followed by rescale_matrix (:171-195). This header
// offers the same two steps on plain column-major arrays so that large LPs (the synthetic
// benchmark configurations of BASELINE.json) never have to go through the string-keyed model.
#pragma once
#include <cstdint>
#include <vector>

namespace simplex_b200
{
    // Row then column scaling exactly as the reference: a row is divided by its max |a| only
    // when that max exceeds 1 (LinearProgram.cpp:179); every column is divided by its max |a|
    // together with c_j (:186-194). A is column-major m x N with leading dimension m.
    void rescale_tableau(int64_t m, int64_t N, double * A, double * b, double * c);

    // Synthetic dense feasible bounded LP G(m, n, n_eq, seed):
    //   minimise c.x, rows 0..n_eq-1 are equalities, the others '<=' ; all data > 0.
    struct SynthLP
    {
        int64_t m = 0, n = 0, n_eq = 0;
        std::vector<double> c;    // n
        std::vector<double> a;    // m*n row-major
        std::vector<double> rhs;  // m
    };
    SynthLP make_synth(int64_t m, int64_t n, int64_t n_eq, uint64_t seed);

    // Number of columns of the standard-form tableau of G: structurals + one slack per '<=' row
    // (+ one artificial per '=' row in Phase I).
    inline int64_t synth_num_cols(int64_t m, int64_t n, int64_t n_eq, bool phase1)
    {
        return n + (m - n_eq) + (phase1 ? n_eq : 0);
    }

    // Fills A (m x N column-major, or only the ncols columns col0, col0+col_stride, ... when sharded), b, c and the
    // crash basis the reference's InitialBasis would build (src/InitialBasis.cpp:23-55, 9-18).
    // basic/nonbasic may be null. Returns N.
    int64_t synth_tableau(int64_t m, int64_t n, int64_t n_eq, uint64_t seed, bool phase1, double * A, double * b,
                          double * c, int64_t * basic, int64_t * nonbasic, int64_t col0 = 0, int64_t ncols = -1,
                          int64_t col_stride = 1);
}  // namespace simplex_b200

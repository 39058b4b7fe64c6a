// Dense fast path of the host side: the same two-phase flow as
// Simplex::solve() / InitialBasis::get_basis() (reference src/Simplex.cpp:22-45,
// src/InitialBasis.cpp:23-137) for problems given as plain arrays, so that large LPs (the benchmark
// configurations) never pass through the string-keyed model (std::map nodes of a 1024 x 2048 LP
// already cost seconds; the 8192 x 262144 one would need > 200 GB, src/LinearProgram.h:24,80).
//
//     min c.x   s.t.  row_i . x  (<=, >=, =)  rhs_i,   0 <= x <= ub
//
// Rows are taken in the given order (= the reference's label order when labels sort like indices),
// variable ids are: structurals 0..n-1, then one slack per inequality row in row order, then one
// artificial per row without a usable slack ('=' rows and rows with a negative right-hand side).
// Everything the device sees (scaling, slack/artificial signs, Phase-I costs, crash basis, list
// hand-off between the phases) follows the reference line by line; only the containers differ.
#pragma once

#include "solver.h"

#include <cstdint>
#include <string>
#include <vector>

namespace simplex
{
    struct DenseProblem
    {
        Index m = 0, n = 0;
        std::vector<double> a;      // m x n, row-major
        std::vector<char> type;     // m entries: '<', '>' or '='
        std::vector<double> rhs;    // m
        std::vector<double> c;      // n (minimise)
        std::vector<double> ub;     // empty = no finite upper bound; else n entries, DBL_MAX = none
    };

    struct DenseSolution
    {
        Status status = Status::not_solved;
        double objective = 0.0;            // in the scaled space, as the reference reports it
        std::vector<double> x_scaled;      // n structural values as the reference reports them
        std::vector<double> x;             // n structural values of the original problem
        std::vector<LoopStats> calls;      // Phase I (if any) and Phase II
        double seconds_host = 0.0;         // wall time of the whole call
    };

    // Throws std::string / const char * like Simplex::solve(). There is no CPU loop: needs a GPU.
    DenseSolution solve_dense(const DenseProblem & problem, const DeviceOptions & options);
}  // namespace simplex

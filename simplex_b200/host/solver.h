// Phase control of the two-phase bounded revised simplex on the host, with the per-iteration
// loop on the GPU (product code).
//
// `simplex::Simplex` and `simplex::InitialBasis` keep the reference's public surface
// (reference src/Simplex.h:17-47, src/InitialBasis.h:16-37): construct on a LinearProgram, call
// solve(), then write("sol_test.xml"). The seam `int Simplex::simplex(Basis &)` is where the
// reference runs its loop on the CPU (src/Simplex.cpp:255-353); here it marshals the tableau
// into the C ABI of libsimplex_b200 (include/simplex_b200.h) and the loop runs as CUDA kernels.
// There is no CPU loop in this library: without a CUDA device simplex() throws.
#pragma once

#include "basis.h"
#include "model.h"

#include <cstdint>
#include <map>
#include <string>
#include <vector>

struct sb200_ctx;

namespace simplex
{
    enum class Pricing { first_eligible = 0, most_negative = 1 };  // src/Simplex.cpp:247-251
    enum class Status { not_solved, optimal, unbounded, infeasible };

    struct DeviceOptions
    {
        int device = 0;
        Pricing pricing = Pricing::first_eligible;  // what the reference ships with (:249)
        int engine = 0;                             // sb200_engine: persistent
        int dual_mode = 1;                          // sb200_dual_mode: update (refreshed every launch)
        std::int64_t chunk_iters = 256;
        int trace_capacity = 0;                     // > 0: keep that many pivots for the debug log
    };

    // Statistics of one call of the seam.
    struct LoopStats
    {
        Index M = 0, N = 0;
        int term = 0;  // sb200_term
        std::int64_t iterations = 0, pivots = 0, bound_flips = 0;
        double loop_seconds = 0.0;  // device time of the loop
        double objective = 0.0;     // c_B . x_B in the scaled space
        bool reused_resident_tableau = false;
    };

    class Simplex
    {
    public:
        explicit Simplex(LinearProgram & lp) : m_lp(lp) {}
        ~Simplex();
        Simplex(const Simplex &) = delete;
        Simplex & operator=(const Simplex &) = delete;

        // ---- the reference's surface
        void solve();
        void write(const std::string & filename) const;
        int simplex(Basis & arg_basis);

        // ---- additions (the reference has no getters, src/Simplex.h:31-33)
        DeviceOptions & options() { return m_opts; }
        Status status() const { return m_status; }
        double objective_value() const { return m_objective_value; }
        const std::map<std::string, double> & solution() const { return m_solution; }
        const std::vector<LoopStats> & calls() const { return m_calls; }

        // Host half of the optimal exit (reference Simplex::solution_found, src/Simplex.cpp:181-245):
        // `flags` = the device's ub_substitutions at exit. Mirrors those flips onto the host copy
        // of the tableau, then undoes them the way the reference does while it reads the solution.
        void absorb_optimum(std::vector<double> x_B, const std::vector<std::uint8_t> & flags, const Basis & basis);

    private:
        LinearProgram & m_lp;
        DeviceOptions m_opts;
        std::map<std::string, double> m_solution;
        double m_objective_value = 0.0;
        Status m_status = Status::not_solved;
        std::vector<LoopStats> m_calls;

        // device side
        sb200_ctx * m_ctx = nullptr;
        Index m_loaded_rows = 0, m_loaded_cols = 0;
        bool m_loaded_with_bounds = false;
        bool m_device_tableau_dirty = true;  // bound flips rewrote A/b on the device
        std::vector<std::uint64_t> m_column_prints;  // fingerprints of what is resident
        std::uint64_t m_rhs_print = 0;
        std::vector<double> m_loaded_ub;

        void solution_found(std::vector<double> vector_bx, const std::vector<double> & vector_c_B, const Basis & basis);
        bool resident_tableau_matches(const Tableau & t, const std::vector<double> & ub, std::vector<std::uint64_t> & prints,
                                      std::uint64_t & rhs_print) const;
        void ensure_context();
        [[noreturn]] void device_failure(const char * what) const;
    };

    class InitialBasis
    {
    public:
        InitialBasis(LinearProgram & lp, Simplex & simplex) : m_lp(lp), m_simplex(simplex) {}

        Basis get_basis();

        // The two host halves of get_basis() around the Phase-I call of the seam:
        // open: slacks, artificials, crash basis; when artificials were needed also the Phase-I
        //       tableau, costs and the ascending non-basic list. Returns the number of artificials.
        // close: drops the artificial columns again (throws when one is still basic).
        size_t open_phase_one(Basis & basis);
        void close_phase_one(Basis & basis);

    private:
        LinearProgram & m_lp;
        Simplex & m_simplex;
        std::map<std::string, Index> m_artificial_of_row;

        size_t add_artificial_variables_for_initial_basis(Basis & basis);
        void fill_non_basic(Basis & basis) const;
    };
}  // namespace simplex

#include "solver.h"

#include "log.h"
#include "simplex_b200.h"

#include <chrono>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <limits>

namespace simplex
{
    namespace
    {
        const std::string ARTIFICIAL_PREFIX = "_ARTIFICIAL_";

        // 64-bit fingerprint of a run of doubles (multiply-xorshift over the bit patterns). Used to
        // recognise that the tableau of this call is the one already resident in HBM.
        std::uint64_t fingerprint(const double * v, Index n)
        {
            std::uint64_t h = 0x9e3779b97f4a7c15ull ^ static_cast<std::uint64_t>(n);
            for (Index i = 0; i < n; ++i)
            {
                std::uint64_t w;
                std::memcpy(&w, v + i, sizeof w);
                h = (h ^ w) * 0xff51afd7ed558ccdull;
                h ^= h >> 32;
            }
            return h;
        }
    }  // namespace

    // ====================================================================== Simplex
    Simplex::~Simplex()
    {
        if (m_ctx) sb200_destroy(m_ctx);
    }

    void Simplex::device_failure(const char * what) const
    {
        throw std::string(what) + ": " + sb200_last_error(m_ctx);
    }

    void Simplex::ensure_context()
    {
        if (m_ctx) return;
        sb200_opts o;
        sb200_default_opts(&o);
        o.device = m_opts.device;
        o.pricing = static_cast<int>(m_opts.pricing);
        o.engine = m_opts.engine;
        o.dual_mode = m_opts.dual_mode;
        o.chunk_iters = m_opts.chunk_iters;
        o.trace_capacity = m_opts.trace_capacity;
        if (sb200_create(&m_ctx, &o) != SB200_OK)
        {
            m_ctx = nullptr;
            throw std::string("GPU solver unavailable: ") + sb200_last_error(nullptr);
        }
    }

    bool Simplex::resident_tableau_matches(const Tableau & t, const std::vector<double> & ub,
                                           std::vector<std::uint64_t> & prints, std::uint64_t & rhs_print) const
    {
        const bool with_bounds = !ub.empty();
        prints.resize(static_cast<size_t>(t.cols));
        for (Index j = 0; j < t.cols; ++j) prints[static_cast<size_t>(j)] = fingerprint(t.column(j), t.rows);
        rhs_print = fingerprint(t.b.data(), t.rows);
        if (!m_ctx || m_device_tableau_dirty || with_bounds != m_loaded_with_bounds) return false;
        if (t.rows != m_loaded_rows || t.cols > m_loaded_cols || rhs_print != m_rhs_print) return false;
        if (with_bounds && std::memcmp(ub.data(), m_loaded_ub.data(), ub.size() * sizeof(double)) != 0) return false;
        for (Index j = 0; j < t.cols; ++j)
            if (prints[static_cast<size_t>(j)] != m_column_prints[static_cast<size_t>(j)]) return false;
        return true;
    }

    // The seam (reference src/Simplex.cpp:255-353). Everything between "gather A_B" and
    // "basis.update" of the reference loop happens inside sb200_run on the device.
    int Simplex::simplex(Basis & basis)
    {
        const Index M = static_cast<Index>(m_lp.get_num_rows());
        const Index N = static_cast<Index>(m_lp.get_num_vars());
        if (static_cast<Index>(basis.get_basic_columns().size()) != M)
            throw std::string("Basis size must be equal to M=" + std::to_string(M));

        Tableau & t = m_lp.tableau();
        if (t.rows != M || t.cols != N) throw std::string("simplex(): initialize_tableau() has not been called for this model");
        m_lp.flip_flags().assign(static_cast<size_t>(N), false);  // :271

        ensure_context();
        const bool bounded = m_lp.has_non_trivial_upper_bounds();
        LoopStats stats;
        stats.M = M;
        stats.N = N;

        // Phase I -> Phase II: the Phase-II tableau is the Phase-I one without its trailing block of
        // artificial columns, so A and b can stay in HBM and only the costs are replaced (unless a
        // bound flip of Phase I rewrote A and b on the device).
        std::vector<std::uint64_t> prints;
        std::uint64_t rhs_print = 0;
        std::vector<double> ub;
        if (bounded) ub = m_lp.dense_upper_bounds();
        if (resident_tableau_matches(t, ub, prints, rhs_print))
        {
            if (sb200_set_costs(m_ctx, N, t.c.data()) != SB200_OK) device_failure("sb200_set_costs");
            stats.reused_resident_tableau = true;
            m_column_prints.resize(static_cast<size_t>(N));
            m_loaded_cols = N;
        }
        else
        {
            if (sb200_load(m_ctx, M, N, t.A.data(), M, t.b.data(), t.c.data(), bounded ? ub.data() : nullptr) != SB200_OK)
                device_failure("sb200_load");
            m_loaded_rows = M;
            m_loaded_cols = N;
            m_loaded_with_bounds = bounded;
            m_loaded_ub = ub;
            m_column_prints.swap(prints);
            m_rhs_print = rhs_print;
            m_device_tableau_dirty = false;
        }

        sb200_result res;
        std::vector<Index> & basic = basis.basic_columns();
        std::vector<Index> & non_basic = basis.non_basic_columns();
        const int rc = sb200_run(m_ctx, basic.data(), static_cast<int64_t>(basic.size()), non_basic.data(),
                                 static_cast<int64_t>(non_basic.size()), &res);
        if (rc == SB200_ERR_BASIS_SIZE) throw std::string("Basis size must be equal to M=" + std::to_string(M));
        if (rc != SB200_OK) device_failure("sb200_run");
        stats.term = res.term;
        stats.iterations = res.iterations;
        stats.pivots = res.pivots;
        stats.bound_flips = res.bound_flips;
        stats.loop_seconds = res.loop_seconds;
        stats.objective = res.objective;
        m_calls.push_back(stats);
        if (res.bound_flips > 0) m_device_tableau_dirty = true;

        SB_LOG(debug) << "SIMPLEX iterations: " << res.iterations << " (pivots " << res.pivots << ", bound flips "
                      << res.bound_flips << ") in " << res.loop_seconds << " s on the device";
        if (log::level() >= log::debug && m_opts.trace_capacity > 0)
        {
            int64_t total = 0;
            std::vector<sb200_pivot> trace(static_cast<size_t>(m_opts.trace_capacity));
            if (sb200_get_trace(m_ctx, trace.data(), m_opts.trace_capacity, &total) == SB200_OK)
            {
                const int64_t shown = std::min<int64_t>(total, m_opts.trace_capacity);
                for (int64_t k = 0; k < shown; ++k)
                    SB_LOG(debug) << "leaving basis: " << trace[k].leaving_col << " (" << trace[k].row
                                  << "), entering basis: " << trace[k].entering_col << " (" << trace[k].pos << ")";
            }
        }

        if (res.term == SB200_TERM_OPTIMAL)
        {
            std::vector<double> x_B(static_cast<size_t>(M));
            std::vector<std::uint8_t> flags(static_cast<size_t>(N), 0);
            if (sb200_get_xB(m_ctx, x_B.data()) != SB200_OK) device_failure("sb200_get_xB");
            if (bounded && sb200_get_flip_flags(m_ctx, flags.data()) != SB200_OK) device_failure("sb200_get_flip_flags");
            absorb_optimum(std::move(x_B), flags, basis);
        }
        else if (res.term == SB200_TERM_UNBOUNDED)
        {
            SB_LOG(debug) << "LP IS UNBOUNDED.\n";
            m_status = Status::unbounded;
            if (bounded)
            {
                // the reference leaves its tableau flipped on this path (:334-338): mirror it
                std::vector<std::uint8_t> flags(static_cast<size_t>(N), 0);
                if (sb200_get_flip_flags(m_ctx, flags.data()) != SB200_OK) device_failure("sb200_get_flip_flags");
                const std::vector<double> ub = m_lp.dense_upper_bounds();
                for (Index j = 0; j < N; ++j)
                    if (flags[static_cast<size_t>(j)]) m_lp.upper_bound_substitution(j, ub[static_cast<size_t>(j)]);
                if (sb200_get_b(m_ctx, t.b.data()) != SB200_OK) device_failure("sb200_get_b");
            }
        }
        else
        {
            throw "MAXIMUM NUMBER OF SIMPLEX ITERATIONS REACHED";  // :347-350
        }
        return 0;
    }

    void Simplex::absorb_optimum(std::vector<double> x_B, const std::vector<std::uint8_t> & flags, const Basis & basis)
    {
        // Bring the host copy of the tableau to the state the reference's would have at this
        // point: every column the device flipped (and its cost) negated, flag set. The right-hand
        // side is not mirrored: nothing reads b after the loop (it is rebuilt by the next
        // initialize_tableau) and solution_found() undoes the flips below.
        Tableau & t = m_lp.tableau();
        const std::vector<double> saved_b = t.b;
        for (size_t j = 0; j < flags.size(); ++j)
            if (flags[j]) m_lp.upper_bound_substitution(static_cast<Index>(j), 0.0);
        std::vector<double> c_B;
        c_B.reserve(basis.get_basic_columns().size());
        for (Index col : basis.get_basic_columns()) c_B.push_back(t.c[static_cast<size_t>(col)]);
        solution_found(std::move(x_B), c_B, basis);
        t.b = saved_b;
    }

    // Reference src/Simplex.cpp:181-245, including its quirks: values are reported in the
    // column-scaled space, shifts are added unscaled, a basic variable that is still flagged as
    // flipped is un-flipped in the tableau but reported as x' (SURVEY.md App. B 3, 5, 13).
    void Simplex::solution_found(std::vector<double> vector_bx, const std::vector<double> & vector_c_B, const Basis & basis)
    {
        double cost = 0.0;
        m_solution.clear();
        std::vector<bool> & flipped = m_lp.flip_flags();
        const auto & upper = m_lp.upper_bounds();
        const auto & shifts = m_lp.shifts();
        const auto bound_of = [&upper](Index id) {
            const auto it = upper.find(id);
            return it == upper.end() ? 0.0 : it->second;
        };

        const std::vector<Index> & basic = basis.get_basic_columns();
        for (size_t i = 0; i < basic.size(); ++i)
        {
            const Index id = basic[i];
            const std::string & name = m_lp.name_of(id);
            if (static_cast<size_t>(id) < flipped.size() && flipped[static_cast<size_t>(id)])
            {
                SB_LOG(debug) << "REVERSING UB-subst of " << name;
                m_lp.upper_bound_substitution(id, bound_of(id));
            }
            const auto shift = shifts.find(id);
            if (shift != shifts.end())
            {
                SB_LOG(debug) << "SHIFTING: basis var " << name << " by " << shift->second;
                vector_bx[i] += shift->second;
            }
            SB_LOG(debug) << "var (basis) " << name << " = " << vector_bx[i];
            m_solution[name] = vector_bx[i];
            cost += vector_c_B[i] * vector_bx[i];
        }

        for (Index id : basis.get_non_basic_columns())
        {
            double q = 0.0;
            const std::string & name = m_lp.name_of(id);
            if (static_cast<size_t>(id) < flipped.size() && flipped[static_cast<size_t>(id)])
            {
                const double ub = bound_of(id);
                m_lp.upper_bound_substitution(id, ub);
                q += ub;
                SB_LOG(debug) << "REVERSING UB-subst of " << name << " ub= " << ub;
            }
            const auto shift = shifts.find(id);
            if (shift != shifts.end())
            {
                SB_LOG(debug) << "SHIFTING: non-basis var " << name << " by " << shift->second;
                q += shift->second;
            }
            SB_LOG(debug) << "var (non-basis) " << name << " = " << q;
            m_solution[name] = q;
            cost += m_lp.tableau().c[static_cast<size_t>(id)] * q;
        }
        m_objective_value = (m_lp.sense() == 'm') ? cost : -cost;
        m_status = Status::optimal;
        SB_LOG(debug) << "OPTIMAL X FOUND.\nValue = " << m_objective_value;
    }

    void Simplex::solve()
    {
        const auto t0 = std::chrono::steady_clock::now();
        InitialBasis crash(m_lp, *this);
        Basis basis = crash.get_basis();

        if (m_objective_value > 0.0)  // Phase-I optimum; no tolerance (src/Simplex.cpp:29)
        {
            SB_LOG(debug) << "LP is infeasible. Aborting.";
            m_status = Status::infeasible;
        }
        else
        {
            m_lp.initialize_tableau();
            SB_LOG(debug) << "*** PHASE II ***";
            simplex(basis);
        }
        const auto us = std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0);
        SB_LOG(info) << "Solver: elapsed " << us.count() / 1000000 << " s (" << us.count() / 1000 << " ms, " << us.count()
                     << " us)";
    }

    // Same text as the reference writes (src/Simplex.cpp:356-373): fixed notation, 6 decimals,
    // variables in name order.
    void Simplex::write(const std::string & filename) const
    {
        SB_LOG(debug) << "Writing solution...\n";
        std::ofstream out(filename);
        out << std::fixed;
        out << "<?xml version = \"1.0\" encoding=\"UTF-8\" standalone=\"yes\"?>" << '\n';
        out << "<TestSolution>\n<header objectiveValue=\"" << m_objective_value << "\"/>" << '\n';
        for (const auto & [name, value] : m_solution)
            out << "<variable name=\"" << name << "\" value=\"" << value << "\"/>" << '\n';
        out << "</TestSolution>" << '\n';
    }

    // ====================================================================== InitialBasis
    size_t InitialBasis::add_artificial_variables_for_initial_basis(Basis & basis)
    {
        // Row by row in label order: the row's slack is basic when its right-hand side is
        // non-negative, otherwise (no slack, or negative rhs) a +-1 artificial column is appended
        // (reference src/InitialBasis.cpp:23-55).
        size_t count = 0;
        for (auto & [label, row] : m_lp.rows())
        {
            if (row.type != '=') throw "All constraints must be equality at this point.";
            const auto slack = m_lp.slack_of_row().find(label);
            if (slack != m_lp.slack_of_row().end() && row.rhs >= 0.0)
            {
                basis.insert(slack->second);
                continue;
            }
            const std::string name = ARTIFICIAL_PREFIX + std::to_string(count++);
            row.add_term(name, row.rhs >= 0.0 ? 1.0 : -1.0);
            const Index id = m_lp.add_variable(name);
            basis.insert(id);
            m_artificial_of_row[label] = id;
            SB_LOG(debug) << "added artificial variable: " << name;
        }
        SB_LOG(debug) << "added " << count << " artificial variables";
        return count;
    }

    void InitialBasis::fill_non_basic(Basis & basis) const
    {
        // every id that is not basic, ascending (src/InitialBasis.cpp:9-18), in O(N)
        const size_t n = m_lp.get_num_vars();
        std::vector<char> is_basic(n, 0);
        for (Index col : basis.get_basic_columns())
            if (col >= 0 && static_cast<size_t>(col) < n) is_basic[static_cast<size_t>(col)] = 1;
        for (size_t id = 0; id < n; ++id)
            if (!is_basic[id]) basis.insert_to_non_basic(static_cast<Index>(id));
    }

    size_t InitialBasis::open_phase_one(Basis & basis)
    {
        m_lp.add_slack_variables_for_inequality_constraints();
        const size_t n_art = add_artificial_variables_for_initial_basis(basis);
        if (n_art == 0)
        {
            SB_LOG(debug) << "Using all slacks for initial basis, skipping Phase I.";
            fill_non_basic(basis);
            return 0;
        }
        SB_LOG(debug) << "*** PHASE I ***";
        m_lp.initialize_tableau();
        std::vector<double> & c = m_lp.tableau().c;  // minimise the sum of the artificials
        std::fill(c.begin(), c.end(), 0.0);
        for (size_t k = 0; k < n_art; ++k) c[static_cast<size_t>(m_lp.get_variable_id(ARTIFICIAL_PREFIX + std::to_string(k)))] = 1.0;
        fill_non_basic(basis);
        return n_art;
    }

    void InitialBasis::close_phase_one(Basis & basis)
    {
        for (auto & [label, row] : m_lp.rows())
        {
            const auto art = m_artificial_of_row.find(label);
            if (art == m_artificial_of_row.end()) continue;
            const Index id = art->second;
            if (basis.contains(id))
            {
                // the reference gives up here with `throw const char *` (src/InitialBasis.cpp:97-114)
                SB_LOG(info) << "artificial variable " << m_lp.name_of(id) << " in the basis.";
                throw "artificial variable in the basis.";
            }
            const std::string name = m_lp.name_of(id);
            row.remove_term(name);
            m_lp.remove_variable(id);
            m_lp.forget_name(name);
            basis.remove(id);
            SB_LOG(debug) << "removed artificial variable: " << name;
        }
        m_lp.tableau().zero();
    }

    Basis InitialBasis::get_basis()
    {
        Basis basis;
        if (open_phase_one(basis) > 0)
        {
            m_simplex.simplex(basis);
            close_phase_one(basis);
        }
        return basis;
    }
}  // namespace simplex

#include "dense_solver.h"

#include "log.h"
#include "simplex_b200.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <limits>

namespace simplex
{
    namespace
    {
        constexpr double NO_BOUND = std::numeric_limits<double>::max();  // src/LinearProgram.cpp:225

        struct Context
        {
            sb200_ctx * ctx = nullptr;
            ~Context()
            {
                if (ctx) sb200_destroy(ctx);
            }
        };

        [[noreturn]] void device_failure(sb200_ctx * ctx, const char * what)
        {
            throw std::string(what) + ": " + sb200_last_error(ctx);
        }

        // Standard-form rows after the slack step of the reference (src/LinearProgram.cpp:110-142):
        // '>=' rows multiplied by -1, one +1 slack per inequality row.
        struct StandardForm
        {
            Index m = 0, n = 0, n_slack = 0, n_art = 0;
            std::vector<double> sign;         // +1 / -1 per row (the negation of '>=' rows)
            std::vector<double> rhs;          // after the negation
            std::vector<Index> slack_of_row;  // -1: equality row
            std::vector<Index> art_of_row;    // -1: the slack is basic
            std::vector<double> art_coeff;    // +1 / -1 (src/InitialBasis.cpp:44)
        };

        StandardForm standard_form(const DenseProblem & p)
        {
            StandardForm f;
            f.m = p.m;
            f.n = p.n;
            f.sign.assign(static_cast<size_t>(p.m), 1.0);
            f.rhs = p.rhs;
            f.slack_of_row.assign(static_cast<size_t>(p.m), -1);
            f.art_of_row.assign(static_cast<size_t>(p.m), -1);
            f.art_coeff.assign(static_cast<size_t>(p.m), 0.0);
            for (Index i = 0; i < p.m; ++i)
            {
                const char t = p.type[static_cast<size_t>(i)];
                if (t == '=') continue;
                if (t == '>')
                {
                    f.sign[static_cast<size_t>(i)] = -1.0;
                    f.rhs[static_cast<size_t>(i)] = -f.rhs[static_cast<size_t>(i)];
                }
                f.slack_of_row[static_cast<size_t>(i)] = p.n + f.n_slack++;
            }
            for (Index i = 0; i < p.m; ++i)
            {
                const bool slack_usable = f.slack_of_row[static_cast<size_t>(i)] >= 0 && f.rhs[static_cast<size_t>(i)] >= 0.0;
                if (slack_usable) continue;
                f.art_of_row[static_cast<size_t>(i)] = p.n + f.n_slack + f.n_art++;
                f.art_coeff[static_cast<size_t>(i)] = f.rhs[static_cast<size_t>(i)] >= 0.0 ? 1.0 : -1.0;
            }
            return f;
        }

        // Column-major tableau of the first N columns (N = n + n_slack [+ n_art]) with the reference's
        // scaling (src/LinearProgram.cpp:171-195); col_scale gets the divisor of every column.
        void build_tableau(const DenseProblem & p, const StandardForm & f, bool with_artificials, std::vector<double> & A,
                           std::vector<double> & b, std::vector<double> & c, std::vector<double> & col_scale)
        {
            const Index m = f.m, n = f.n;
            const Index N = n + f.n_slack + (with_artificials ? f.n_art : 0);
            A.assign(static_cast<size_t>(m) * static_cast<size_t>(N), 0.0);
            b = f.rhs;
            c.assign(static_cast<size_t>(N), 0.0);
            std::vector<double> row_max(static_cast<size_t>(m), 0.0);
            for (Index i = 0; i < m; ++i)
            {
                const double sg = f.sign[static_cast<size_t>(i)];
                const double * row = p.a.data() + static_cast<size_t>(i) * static_cast<size_t>(n);
                double top = 0.0;
                for (Index j = 0; j < n; ++j)
                {
                    // |a| < 1e-7 never enters a row (Constraint::add_term, src/LinearProgram.cpp:17-20)
                    const double v = is_float_zero(row[j]) ? 0.0 : sg * row[j];
                    A[static_cast<size_t>(j) * m + i] = v;
                    top = std::max(top, std::fabs(v));
                }
                if (f.slack_of_row[static_cast<size_t>(i)] >= 0)
                {
                    A[static_cast<size_t>(f.slack_of_row[static_cast<size_t>(i)]) * m + i] = 1.0;
                    top = std::max(top, 1.0);
                }
                if (with_artificials && f.art_of_row[static_cast<size_t>(i)] >= 0)
                {
                    A[static_cast<size_t>(f.art_of_row[static_cast<size_t>(i)]) * m + i] = f.art_coeff[static_cast<size_t>(i)];
                    top = std::max(top, 1.0);
                }
                row_max[static_cast<size_t>(i)] = top;
            }
            for (Index j = 0; j < n; ++j) c[static_cast<size_t>(j)] = p.c[static_cast<size_t>(j)];
            for (Index i = 0; i < m; ++i)
                if (row_max[static_cast<size_t>(i)] > 1.0) b[static_cast<size_t>(i)] /= row_max[static_cast<size_t>(i)];
            col_scale.assign(static_cast<size_t>(N), 0.0);
            for (Index j = 0; j < N; ++j)
            {
                double * col = A.data() + static_cast<size_t>(j) * m;
                double top = 0.0;
                for (Index i = 0; i < m; ++i)
                {
                    if (row_max[static_cast<size_t>(i)] > 1.0) col[i] /= row_max[static_cast<size_t>(i)];
                    top = std::max(top, std::fabs(col[i]));
                }
                col_scale[static_cast<size_t>(j)] = top;
                for (Index i = 0; i < m; ++i) col[i] /= top;  // unguarded, like the reference
                c[static_cast<size_t>(j)] /= top;
            }
        }

        LoopStats run_phase(sb200_ctx * ctx, Index M, Index N, std::vector<Index> & basic, std::vector<Index> & non_basic,
                            bool resident)
        {
            sb200_result res;
            const int rc = sb200_run(ctx, basic.data(), static_cast<int64_t>(basic.size()), non_basic.data(),
                                     static_cast<int64_t>(non_basic.size()), &res);
            if (rc != SB200_OK) device_failure(ctx, "sb200_run");
            LoopStats s;
            s.M = M;
            s.N = N;
            s.term = res.term;
            s.iterations = res.iterations;
            s.pivots = res.pivots;
            s.bound_flips = res.bound_flips;
            s.loop_seconds = res.loop_seconds;
            s.objective = res.objective;
            s.reused_resident_tableau = resident;
            if (res.term == SB200_TERM_ITER_LIMIT) throw "MAXIMUM NUMBER OF SIMPLEX ITERATIONS REACHED";
            return s;
        }
    }  // namespace

    DenseSolution solve_dense(const DenseProblem & p, const DeviceOptions & options)
    {
        const auto t0 = std::chrono::steady_clock::now();
        if (p.m <= 0 || p.n <= 0 || static_cast<Index>(p.a.size()) != p.m * p.n || static_cast<Index>(p.type.size()) != p.m ||
            static_cast<Index>(p.rhs.size()) != p.m || static_cast<Index>(p.c.size()) != p.n ||
            (!p.ub.empty() && static_cast<Index>(p.ub.size()) != p.n))
            throw std::string("solve_dense: inconsistent array sizes");

        DenseSolution out;
        const auto stamp = [&out, t0]() {
            out.seconds_host = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        };
        const StandardForm f = standard_form(p);
        const Index m = p.m, n = p.n;
        const Index N2 = n + f.n_slack, N1 = N2 + f.n_art;
        const bool bounded = !p.ub.empty();
        std::vector<double> ub;
        if (bounded)
        {
            ub.assign(static_cast<size_t>(N1), NO_BOUND);
            std::copy(p.ub.begin(), p.ub.end(), ub.begin());
        }

        Context dev;
        sb200_opts o;
        sb200_default_opts(&o);
        o.device = options.device;
        o.pricing = static_cast<int>(options.pricing);
        o.engine = options.engine;
        o.dual_mode = options.dual_mode;
        o.chunk_iters = options.chunk_iters;
        if (sb200_create(&dev.ctx, &o) != SB200_OK) throw std::string("GPU solver unavailable: ") + sb200_last_error(nullptr);

        // crash basis: per row its slack or its artificial (src/InitialBasis.cpp:23-55), the rest ascending (:9-18)
        std::vector<Index> basic(static_cast<size_t>(m)), non_basic;
        std::vector<char> is_basic(static_cast<size_t>(N1), 0);
        for (Index i = 0; i < m; ++i)
        {
            const Index col = f.art_of_row[static_cast<size_t>(i)] >= 0 ? f.art_of_row[static_cast<size_t>(i)]
                                                                         : f.slack_of_row[static_cast<size_t>(i)];
            basic[static_cast<size_t>(i)] = col;
            is_basic[static_cast<size_t>(col)] = 1;
        }

        std::vector<double> A, b, c, col_scale;
        bool tableau_resident = false;
        if (f.n_art > 0)
        {
            SB_LOG(debug) << "*** PHASE I ***";
            build_tableau(p, f, true, A, b, c, col_scale);
            std::fill(c.begin(), c.end(), 0.0);  // minimise the sum of the artificials (:73-78)
            for (Index i = 0; i < m; ++i)
                if (f.art_of_row[static_cast<size_t>(i)] >= 0) c[static_cast<size_t>(f.art_of_row[static_cast<size_t>(i)])] = 1.0;
            for (Index j = 0; j < N1; ++j)
                if (!is_basic[static_cast<size_t>(j)]) non_basic.push_back(j);
            if (sb200_load(dev.ctx, m, N1, A.data(), m, b.data(), c.data(), bounded ? ub.data() : nullptr) != SB200_OK)
                device_failure(dev.ctx, "sb200_load");
            const LoopStats s1 = run_phase(dev.ctx, m, N1, basic, non_basic, false);
            out.calls.push_back(s1);
            // Phase-I optimum > 0 (no tolerance, src/Simplex.cpp:29) or an artificial still basic (:97-114)
            for (Index col : basic)
                if (col >= N2) throw "artificial variable in the basis.";
            if (s1.term == SB200_TERM_OPTIMAL && s1.objective > 0.0)
            {
                out.status = Status::infeasible;
                stamp();
                return out;
            }
            non_basic.erase(std::remove_if(non_basic.begin(), non_basic.end(), [N2](Index col) { return col >= N2; }),
                            non_basic.end());
            tableau_resident = s1.bound_flips == 0;
        }
        else
        {
            for (Index j = 0; j < N2; ++j)
                if (!is_basic[static_cast<size_t>(j)]) non_basic.push_back(j);
        }

        SB_LOG(debug) << "*** PHASE II ***";
        if (tableau_resident)
        {
            // same A and b without the trailing artificial block: only the costs are replaced
            std::vector<double> c2(static_cast<size_t>(N2), 0.0);
            for (Index j = 0; j < n; ++j) c2[static_cast<size_t>(j)] = p.c[static_cast<size_t>(j)] / col_scale[static_cast<size_t>(j)];
            c.swap(c2);
            if (sb200_set_costs(dev.ctx, N2, c.data()) != SB200_OK) device_failure(dev.ctx, "sb200_set_costs");
        }
        else
        {
            build_tableau(p, f, false, A, b, c, col_scale);
            if (sb200_load(dev.ctx, m, N2, A.data(), m, b.data(), c.data(), bounded ? ub.data() : nullptr) != SB200_OK)
                device_failure(dev.ctx, "sb200_load");
        }
        const LoopStats s2 = run_phase(dev.ctx, m, N2, basic, non_basic, tableau_resident);
        out.calls.push_back(s2);
        if (s2.term == SB200_TERM_UNBOUNDED)
        {
            out.status = Status::unbounded;
            stamp();
            return out;
        }

        // read-out as Simplex::solution_found (src/Simplex.cpp:181-245), without names and shifts
        std::vector<double> x_B(static_cast<size_t>(m));
        std::vector<uint8_t> flags(static_cast<size_t>(N2), 0);
        if (sb200_get_xB(dev.ctx, x_B.data()) != SB200_OK) device_failure(dev.ctx, "sb200_get_xB");
        if (bounded && sb200_get_flip_flags(dev.ctx, flags.data()) != SB200_OK) device_failure(dev.ctx, "sb200_get_flip_flags");
        std::vector<double> value(static_cast<size_t>(N2), 0.0);
        double cost = 0.0;
        for (Index i = 0; i < m; ++i)
        {
            const Index col = basic[static_cast<size_t>(i)];
            const double cB = flags[static_cast<size_t>(col)] ? -c[static_cast<size_t>(col)] : c[static_cast<size_t>(col)];
            value[static_cast<size_t>(col)] = x_B[static_cast<size_t>(i)];
            cost += cB * x_B[static_cast<size_t>(i)];
        }
        for (Index col : non_basic)
        {
            const double q = flags[static_cast<size_t>(col)] ? ub[static_cast<size_t>(col)] : 0.0;
            value[static_cast<size_t>(col)] = q;
            cost += c[static_cast<size_t>(col)] * q;
        }
        out.status = Status::optimal;
        out.objective = cost;
        out.x_scaled.assign(value.begin(), value.begin() + n);
        out.x.resize(static_cast<size_t>(n));
        for (Index j = 0; j < n; ++j) out.x[static_cast<size_t>(j)] = out.x_scaled[static_cast<size_t>(j)] / col_scale[static_cast<size_t>(j)];
        stamp();
        return out;
    }
}  // namespace simplex

#include "presolve.h"

#include "log.h"

#include <cstdlib>
#include <vector>

namespace simplex
{
    namespace
    {
        Presolve::InfeasibleHandler & handler_slot()
        {
            static Presolve::InfeasibleHandler slot;
            return slot;
        }
    }  // namespace

    void Presolve::set_infeasible_handler(InfeasibleHandler handler) { handler_slot() = std::move(handler); }

    void Presolve::infeasible(const std::string & reason)
    {
        SB_LOG(debug) << reason;
        if (handler_slot())
        {
            handler_slot()(reason);
            return;
        }
        std::cout.flush();
        std::exit(0);  // the reference ends the process with status 0 here
    }

    void Presolve::substitute_fixed_variable(Index, double) { throw "Presolve: not implemented."; }

    // Activity bounds of every row from the variable bounds: L <= a.x <= U. A '<=' row with
    // U <= rhs (or '>=' with L >= rhs) can never bind and is removed; a row whose activity range
    // misses the right-hand side proves infeasibility (reference src/Presolve.cpp:85-127).
    void Presolve::drop_rows_that_cannot_bind() const
    {
        auto & lower = m_lp.lower_bounds();
        auto & upper = m_lp.upper_bounds();
        std::vector<std::string> never_binding;
        for (const auto & [label, row] : m_lp.rows())
        {
            double hi = 0.0, lo = 0.0;
            for (const auto & [var_name, a] : row.name_to_coeff)
            {
                const Index id = m_lp.ids_by_name().at(var_name);
                const double ub = upper[id];
                const double lb = lower[id];
                if (a > 0.0)
                {
                    hi += a * ub;
                    lo += a * lb;
                }
                else if (a < 0.0)
                {
                    hi += a * lb;
                    lo += a * ub;
                }
            }
            SB_LOG(debug) << "Presolve: applying reduction to constraint:" << label << " L=" << lo << " U=" << hi;
            const bool le = row.type == '<', ge = row.type == '>', eq = row.type == '=';
            if ((le && hi <= row.rhs) || (ge && lo >= row.rhs))
            {
                SB_LOG(debug) << "  constraint is redundant";
                never_binding.push_back(label);
            }
            if (((le || eq) && lo > row.rhs) || ((ge || eq) && hi < row.rhs)) infeasible("  no feasible solution");
        }
        for (const std::string & label : never_binding) m_lp.rows().erase(label);
    }

    // x' = x - lb, x' >= 0 for every variable with a non-zero lower bound: the upper bound moves
    // with it, the shift is remembered for the read-out, every right-hand side absorbs lb * a
    // (reference src/Presolve.cpp:26-83).
    void Presolve::shift_lower_bounds_to_zero()
    {
        for (auto & [var_id, lb_slot] : m_lp.lower_bounds())
        {
            if (is_float_zero(lb_slot)) continue;
            const double lb = lb_slot;
            const auto ub_it = m_lp.upper_bounds().find(var_id);
            if (ub_it != m_lp.upper_bounds().end())
            {
                const double ub = ub_it->second;
                if (lb > ub)
                    infeasible("Presolve: problem infeasible, lb>ub for variable: " + std::to_string(var_id));
                else if (is_float_zero(ub - lb))
                    substitute_fixed_variable(var_id, lb);
                else
                    ub_it->second = ub - lb;
            }
            m_lp.shifts()[var_id] = lb;
            lb_slot = 0.0;

            const std::string var_name = m_lp.name_of(var_id);
            const auto cj = m_lp.objective().find(var_id);
            if (cj != m_lp.objective().end())
            {
                SB_LOG(debug) << "Presolve: applying shift " << lb << " to obj.fun. variable: " << var_name;
                m_lp.objective_shift() += cj->second * lb;
            }
            for (auto & [label, row] : m_lp.rows())
            {
                const auto a = row.get_coefficient(var_name);
                if (!a.has_value()) continue;
                SB_LOG(debug) << "Presolve: applying shift " << lb << " to constraint " << label << " variable: " << var_name;
                row.rhs -= lb * a.value();
            }
        }
    }

    void Presolve::run()
    {
        if (m_reductions_enabled) drop_rows_that_cannot_bind();
        shift_lower_bounds_to_zero();
    }
}  // namespace simplex

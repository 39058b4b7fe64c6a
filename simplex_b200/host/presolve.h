// Presolve of the text model (reference src/Presolve.h:18-39, src/Presolve.cpp): row reductions
// from the variable bounds, then elimination of non-zero lower bounds by the shift x' = x - lb.
#pragma once

#include "model.h"

#include <functional>
#include <string>

namespace simplex
{
    class Presolve
    {
    public:
        Presolve(LinearProgram & lp, bool reductions_enabled = true) : m_lp(lp), m_reductions_enabled(reductions_enabled) {}

        void run();

        // What happens when presolve proves the LP infeasible. Default = the reference's behaviour:
        // log the reason and std::exit(0) (src/Presolve.cpp:40-41, :118-119). A host application
        // can install a handler that throws instead.
        using InfeasibleHandler = std::function<void(const std::string & reason)>;
        static void set_infeasible_handler(InfeasibleHandler handler);

    private:
        static void infeasible(const std::string & reason);
        void drop_rows_that_cannot_bind() const;                     // reference: apply_reductions
        void shift_lower_bounds_to_zero();                           // reference: eliminate_lbs
        void substitute_fixed_variable(Index var_id, double value);  // reference: fix_variable (not implemented there either)

        LinearProgram & m_lp;
        bool m_reductions_enabled;
    };
}  // namespace simplex

#include "model.h"

#include "log.h"

#include <algorithm>
#include <cmath>
#include <sstream>

namespace simplex
{
    // ------------------------------------------------------------------ Constraint
    Constraint & Constraint::add_term(std::string_view name, double a)
    {
        if (is_float_zero(a)) return *this;
        name_to_coeff[std::string(name)] += a;  // value-initialised to 0.0 when new
        return *this;
    }

    bool Constraint::has_variable(std::string_view name) const
    {
        return name_to_coeff.find(std::string(name)) != name_to_coeff.end();
    }

    std::optional<double> Constraint::get_coefficient(std::string_view name) const
    {
        const auto it = name_to_coeff.find(std::string(name));
        if (it == name_to_coeff.end()) return std::nullopt;
        return it->second;
    }

    Constraint & Constraint::remove_term(std::string_view name)
    {
        name_to_coeff.erase(std::string(name));
        return *this;
    }

    void Constraint::negate_sides()
    {
        rhs = -rhs;
        for (auto & term : name_to_coeff) term.second = -term.second;
        if (type == '<')
            type = '>';
        else if (type == '>')
            type = '<';
    }

    // ------------------------------------------------------------------ Tableau
    void Tableau::reset(Index m, Index n)
    {
        rows = m;
        cols = n;
        A.assign(static_cast<size_t>(m) * static_cast<size_t>(n), 0.0);
        b.assign(static_cast<size_t>(m), 0.0);
        c.assign(static_cast<size_t>(n), 0.0);
    }

    void Tableau::zero()
    {
        std::fill(A.begin(), A.end(), 0.0);
        std::fill(b.begin(), b.end(), 0.0);
        std::fill(c.begin(), c.end(), 0.0);
    }

    // ------------------------------------------------------------------ LinearProgram: model
    void LinearProgram::add_constraint(const std::string & name, Constraint && constraint)
    {
        m_rows.emplace(name, std::move(constraint));  // an existing label wins
    }

    const Constraint & LinearProgram::get_constraint(const std::string & name) const
    {
        const auto it = m_rows.find(name);
        if (it == m_rows.end()) throw std::string("Constraint not found: " + name);
        return it->second;
    }

    Index LinearProgram::add_variable(const std::string & var_name, double coeff)
    {
        const Index id = m_next_id++;
        m_name_to_id[var_name] = id;
        m_id_to_name[id] = var_name;
        m_objective[id] = coeff;
        m_lower[id] = 0.0;
        m_upper[id] = std::numeric_limits<double>::max();
        return id;
    }

    void LinearProgram::remove_variable(Index var_id)
    {
        const auto it = m_id_to_name.find(var_id);
        if (it == m_id_to_name.end()) return;
        m_objective.erase(var_id);
        m_lower.erase(var_id);
        m_upper.erase(var_id);
        m_name_to_id.erase(it->second);
        m_id_to_name.erase(it);
    }

    bool LinearProgram::has_variable(const std::string & var_name) const
    {
        return m_name_to_id.find(var_name) != m_name_to_id.end();
    }

    Index LinearProgram::get_variable_id(const std::string & var_name) const
    {
        const auto it = m_name_to_id.find(var_name);
        if (it == m_name_to_id.end()) throw std::string("Variable not found: " + var_name);
        return it->second;
    }

    const std::string & LinearProgram::name_of(Index var_id) const
    {
        static const std::string none;
        const auto it = m_id_to_name.find(var_id);
        return it == m_id_to_name.end() ? none : it->second;
    }

    void LinearProgram::set_lower_bound(Index var_id, double value) { m_lower[var_id] = value; }

    void LinearProgram::set_upper_bound(Index var_id, double value)
    {
        m_upper[var_id] = value;
        m_has_ubs = true;  // from now on the loop uses the bounded ratio test (src/Simplex.cpp:108)
    }

    std::vector<double> LinearProgram::dense_upper_bounds() const
    {
        std::vector<double> ub(get_num_vars(), std::numeric_limits<double>::max());
        for (const auto & [id, u] : m_upper)
            if (id >= 0 && static_cast<size_t>(id) < ub.size()) ub[static_cast<size_t>(id)] = u;
        return ub;
    }

    size_t LinearProgram::add_slack_variables_for_inequality_constraints()
    {
        // One slack per inequality row, in row (label) order; a '>=' row is first multiplied by
        // -1 so that every slack enters with +1 (reference src/LinearProgram.cpp:110-142).
        size_t added = 0;
        for (auto & [label, row] : m_rows)
        {
            if (row.type == '=') continue;
            const std::string slack_name = "__SLACK" + std::to_string(added++);
            if (row.type == '>') row.negate_sides();  // now a '<' row
            if (row.type == '<') row.add_term(slack_name);
            row.type = '=';
            m_slack_of_row[label] = add_variable(slack_name);
            SB_LOG(debug) << "added slack variable: " << slack_name << " (" << m_slack_of_row[label] << ")";
        }
        SB_LOG(debug) << "added " << added << " slack variables";
        return added;
    }

    // ------------------------------------------------------------------ LinearProgram: tableau
    void LinearProgram::initialize_tableau()
    {
        const Index M = static_cast<Index>(get_num_rows());
        const Index N = static_cast<Index>(get_num_vars());
        SB_LOG(debug) << "problem size: N=" << N << ", M=" << M;
        m_tab.reset(M, N);

        Index i = 0;
        for (const auto & [label, row] : m_rows)
        {
            (void) label;
            for (const auto & [var_name, a] : row.name_to_coeff)
            {
                // the reference indexes with operator[] (src/LinearProgram.cpp:84), i.e. an
                // unknown name silently becomes a new map entry with id 0
                const Index j = m_name_to_id[var_name];
                m_tab.at(i, j) = a;
            }
            m_tab.b[static_cast<size_t>(i)] = row.rhs;
            ++i;
        }
        for (const auto & [id, coeff] : m_objective)
            m_tab.c[static_cast<size_t>(id)] = (m_sense == 'm') ? coeff : -coeff;

        rescale_matrix();
    }

    void LinearProgram::rescale_matrix()
    {
        // Row pass: divide a row (and its rhs) by max|a| only when that exceeds 1; column pass:
        // divide EVERY column and its cost by max|a| — unguarded, a zero column turns into NaNs
        // exactly as in the reference (src/LinearProgram.cpp:171-195).
        const Index M = m_tab.rows, N = m_tab.cols;
        m_row_scale.assign(static_cast<size_t>(M), 0.0);
        for (Index j = 0; j < N; ++j)
        {
            const double * col = m_tab.column(j);
            for (Index i = 0; i < M; ++i)
                m_row_scale[static_cast<size_t>(i)] = std::max(m_row_scale[static_cast<size_t>(i)], std::fabs(col[i]));
        }
        for (Index i = 0; i < M; ++i)
            if (m_row_scale[static_cast<size_t>(i)] > 1.0) m_tab.b[static_cast<size_t>(i)] /= m_row_scale[static_cast<size_t>(i)];

        m_col_scale.assign(static_cast<size_t>(N), 0.0);
        for (Index j = 0; j < N; ++j)
        {
            double * col = m_tab.column(j);
            double top = 0.0;
            for (Index i = 0; i < M; ++i)
            {
                const double r = m_row_scale[static_cast<size_t>(i)];
                if (r > 1.0) col[i] /= r;
                top = std::max(top, std::fabs(col[i]));
            }
            m_col_scale[static_cast<size_t>(j)] = top;
            for (Index i = 0; i < M; ++i) col[i] /= top;
            m_tab.c[static_cast<size_t>(j)] /= top;
        }
    }

    void LinearProgram::upper_bound_substitution(Index var_id, double ub)
    {
        // x_s -> u_s - x_s on the host copy: c_s and column s change sign, b -= a_s * u_s
        // (reference src/LinearProgram.cpp:154-169). The device applies the same rule to its own
        // resident copy during the loop (csrc/sb200_phases.cuh, k_select).
        const size_t s = static_cast<size_t>(var_id);
        m_tab.c[s] = -m_tab.c[s];
        double * col = m_tab.column(var_id);
        for (Index i = 0; i < m_tab.rows; ++i)
        {
            const double a = col[i];
            m_tab.b[static_cast<size_t>(i)] = m_tab.b[static_cast<size_t>(i)] - a * ub;
            col[i] = -a;
        }
        if (m_flipped.size() <= s) m_flipped.resize(s + 1, false);
        m_flipped[s] = !m_flipped[s];
    }

    void LinearProgram::print_tableau() const
    {
        std::ostringstream os;
        os << "A=\n";
        for (Index i = 0; i < m_tab.rows; ++i)
        {
            for (Index j = 0; j < m_tab.cols; ++j) os << (j ? " " : "") << m_tab.at(i, j);
            os << '\n';
        }
        os << "b=\n";
        for (double v : m_tab.b) os << v << '\n';
        os << "c=\n";
        for (double v : m_tab.c) os << v << '\n';
        SB_LOG(debug) << os.str();
    }
}  // namespace simplex

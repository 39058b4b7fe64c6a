// Host-side text model of a linear program (product code, C++17, no Eigen).
//
// Mirrors the public surface of the reference's `simplex::Constraint` / `simplex::LinearProgram`
// (reference src/LinearProgram.h:20-107) so code written against the reference compiles against
// this header: same namespace, same member names, same argument meaning, same exception types
// (the reference throws `std::string`, src/LinearProgram.cpp:209,254). What differs is behind the
// surface: the dense tableau is a plain column-major `std::vector<double>` that is handed to the
// GPU library through the C ABI (include/simplex_b200.h) — there is no linear algebra on the host.
//
// Ordering rules that decide the layout of the tableau (and therefore every pivot sequence):
//   * rows are the constraints in lexicographic LABEL order (reference keeps them in a
//     std::map<std::string, Constraint>, src/LinearProgram.h:80, and emits them in map order,
//     src/LinearProgram.cpp:79-88);
//   * columns are variable ids, handed out in order of creation (src/LinearProgram.cpp:217-226);
//   * a second constraint with an already-used label is dropped (std::map::emplace keeps the
//     first, src/LinearProgram.cpp:197-200).
#pragma once

#include <cstdint>
#include <limits>
#include <map>
#include <optional>
#include <string>
#include <string_view>
#include <vector>

namespace simplex
{
    using Index = std::int64_t;  // the reference's Eigen::Index (ptrdiff_t)

    constexpr double ALMOST_ZERO = 1e-7;  // reference src/Utils.h:9
    inline bool is_float_zero(double x) { return (x < 0.0 ? -x : x) < ALMOST_ZERO; }

    struct Constraint
    {
        char type;   // '<', '>' or '='
        double rhs;
        std::map<std::string, double> name_to_coeff;  // variable name -> a_ij, ordered by name

        Constraint(char type_ = '<', double rhs_ = 0.0) : type(type_), rhs(rhs_) {}

        // |a| < 1e-7 is dropped, a repeated name accumulates (reference src/LinearProgram.cpp:15-32).
        Constraint & add_term(std::string_view name, double a = 1.0);
        bool has_variable(std::string_view name) const;
        std::optional<double> get_coefficient(std::string_view name) const;
        Constraint & remove_term(std::string_view name);
        void negate_sides();  // both sides times -1, '<' <-> '>' (src/LinearProgram.cpp:52-60)
    };

    // Dense tableau in the layout the device expects: A column-major, ld == rows.
    struct Tableau
    {
        Index rows = 0, cols = 0;
        std::vector<double> A, b, c;
        double & at(Index i, Index j) { return A[static_cast<size_t>(j) * rows + i]; }
        double at(Index i, Index j) const { return A[static_cast<size_t>(j) * rows + i]; }
        double * column(Index j) { return A.data() + static_cast<size_t>(j) * rows; }
        const double * column(Index j) const { return A.data() + static_cast<size_t>(j) * rows; }
        void reset(Index m, Index n);  // resize and zero
        void zero();
    };

    // A standard-form LP:  min c^T x  s.t.  Ax = b, 0 <= x <= UB.
    class LinearProgram
    {
    public:
        // ---- the reference's public surface (src/LinearProgram.h:46-72)
        void add_constraint(const std::string & name, Constraint && constraint);
        void add_constraint(const std::string & name, const Constraint & constraint)  // addition: by copy
        {
            add_constraint(name, Constraint(constraint));
        }
        [[nodiscard]] const Constraint & get_constraint(const std::string & name) const;  // throws std::string
        Index add_variable(const std::string & var_name, double coeff = 0.0);
        void remove_variable(Index var_id);
        bool has_variable(const std::string & var_name) const;
        Index get_variable_id(const std::string & var_name) const;  // throws std::string
        void set_sense(const char s) { m_sense = s; }
        void set_objective_label(const std::string & label) { m_objective_label = label; }
        void set_lower_bound(Index var_id, double value);
        void set_upper_bound(Index var_id, double value);

        void initialize_tableau();  // constraints + objective -> dense A, b, c, then rescale_matrix()
        void print_tableau() const;
        size_t add_slack_variables_for_inequality_constraints();
        size_t get_num_vars() const { return m_name_to_id.size(); }
        size_t get_num_rows() const { return m_rows.size(); }
        void upper_bound_substitution(Index var_id, double ub);  // on the HOST copy of the tableau
        bool has_non_trivial_upper_bounds() const { return m_has_ubs; }
        void rescale_matrix();

        // ---- additions (the reference exposes these only to its friend classes)
        char sense() const { return m_sense; }
        const std::string & objective_label() const { return m_objective_label; }
        Tableau & tableau() { return m_tab; }
        const Tableau & tableau() const { return m_tab; }
        std::map<std::string, Constraint> & rows() { return m_rows; }
        const std::map<std::string, Constraint> & rows() const { return m_rows; }
        const std::map<Index, double> & objective() const { return m_objective; }
        std::map<Index, double> & lower_bounds() { return m_lower; }
        std::map<Index, double> & upper_bounds() { return m_upper; }
        const std::map<Index, double> & upper_bounds() const { return m_upper; }
        const std::map<std::string, Index> & slack_of_row() const { return m_slack_of_row; }
        const std::map<std::string, Index> & ids_by_name() const { return m_name_to_id; }
        const std::string & name_of(Index var_id) const;  // "" when unknown
        void forget_name(const std::string & var_name) { m_name_to_id.erase(var_name); }
        std::map<Index, double> & shifts() { return m_shift; }
        const std::map<Index, double> & shifts() const { return m_shift; }
        double & objective_shift() { return m_objective_shift; }
        std::vector<bool> & flip_flags() { return m_flipped; }
        const std::vector<double> & row_scaling() const { return m_row_scale; }
        const std::vector<double> & column_scaling() const { return m_col_scale; }
        // Upper bounds as the dense array the device wants: ub[j] = bound of variable id j,
        // DBL_MAX = none (src/LinearProgram.cpp:225). Sized to get_num_vars().
        std::vector<double> dense_upper_bounds() const;

    private:
        std::map<std::string, Constraint> m_rows;
        std::map<Index, double> m_objective;  // id -> c_j as written in the file
        std::map<Index, double> m_lower;      // zeroed by Presolve::eliminate_lbs
        std::map<Index, double> m_upper;
        std::string m_objective_label;
        char m_sense = 'm';  // 'M' maximise, 'm' minimise
        Tableau m_tab;
        Index m_next_id = 0;
        std::map<std::string, Index> m_name_to_id;
        std::map<Index, std::string> m_id_to_name;
        std::vector<bool> m_flipped;
        std::vector<double> m_col_scale, m_row_scale;
        std::map<std::string, Index> m_slack_of_row;
        double m_objective_shift = 0.0;
        std::map<Index, double> m_shift;
        bool m_has_ubs = false;
    };
}  // namespace simplex

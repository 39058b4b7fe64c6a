// Unit tests of the host model API (no GPU needed). The first two cases follow the reference's
// own gtest cases one to one (reference tests/UnitTests.cpp:12-34 ConstraintTests1 and :36-56
// LinearProgramTests1) without gtest, which is not available offline; the rest cover Basis and
// Presolve through the same public surface.
#include "../basis.h"
#include "../log.h"
#include "../lp_text.h"
#include "../model.h"
#include "../presolve.h"

#include <cstdio>
#include <optional>
#include <string>

static int g_failed = 0;
#define CHECK(cond)                                                        \
    do                                                                     \
    {                                                                      \
        if (!(cond))                                                       \
        {                                                                  \
            std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);  \
            ++g_failed;                                                    \
        }                                                                  \
    } while (0)

static void constraint_tests_1()
{
    simplex::Constraint constraint('>', 9);
    constraint.add_term("x1");
    constraint.add_term("y23", -3.21);
    constraint.add_term("new_var", 0.0);
    constraint.add_term("new_var1", -1.0);

    CHECK(constraint.rhs == 9);
    CHECK(constraint.type == '>');
    CHECK(constraint.has_variable("x1"));
    CHECK(constraint.has_variable("y23"));
    CHECK(!constraint.has_variable("new_var"));
    CHECK(constraint.has_variable("new_var1"));
    CHECK(constraint.get_coefficient("x1") == 1.0);
    CHECK(constraint.get_coefficient("new_var") == std::nullopt);
    CHECK(constraint.get_coefficient("y23") == -3.21);

    constraint.remove_term("y23");
    CHECK(!constraint.has_variable("y23"));
    CHECK(constraint.get_coefficient("y23") == std::nullopt);
}

static void linear_program_tests_1()
{
    simplex::LinearProgram lp;
    lp.add_variable("x1");
    lp.add_variable("x2", 1.0);
    lp.add_variable("x3", -2.0);
    auto c1 = simplex::Constraint('<', 5).add_term("x1").add_term("x2", 2.0);
    lp.add_constraint("C1", std::move(c1));
    lp.initialize_tableau();
    lp.print_tableau();

    bool threw_string = false;
    try
    {
        (void) lp.get_constraint("C2");
    }
    catch (const std::string &)
    {
        threw_string = true;
    }
    CHECK(threw_string);
    const auto & test_c1 = lp.get_constraint("C1");
    CHECK(test_c1.has_variable("x2"));
    CHECK(test_c1.get_coefficient("x2").value() == 2.0);
}

static void model_details()
{
    simplex::LinearProgram lp;
    CHECK(lp.add_variable("a", 1.0) == 0);
    CHECK(lp.add_variable("b", 2.0) == 1);
    CHECK(lp.get_num_vars() == 2);
    CHECK(lp.has_variable("a") && !lp.has_variable("zz"));
    CHECK(lp.get_variable_id("b") == 1);
    bool threw = false;
    try
    {
        lp.get_variable_id("zz");
    }
    catch (const std::string & msg)
    {
        threw = msg == "Variable not found: zz";
    }
    CHECK(threw);
    CHECK(!lp.has_non_trivial_upper_bounds());
    lp.set_upper_bound(0, 4.0);
    CHECK(lp.has_non_trivial_upper_bounds());

    // rows come out in label order, a repeated label keeps the first row; '>=' rows are negated
    lp.add_constraint("r2", simplex::Constraint('>', 2).add_term("a").add_term("b", 4.0));
    lp.add_constraint("r1", simplex::Constraint('<', 8).add_term("a", 2.0).add_term("b"));
    lp.add_constraint("r1", simplex::Constraint('<', 99).add_term("a"));
    CHECK(lp.get_num_rows() == 2);
    CHECK(lp.get_constraint("r1").rhs == 8);
    CHECK(lp.add_slack_variables_for_inequality_constraints() == 2);
    CHECK(lp.get_variable_id("__SLACK0") == 2 && lp.get_variable_id("__SLACK1") == 3);
    CHECK(lp.get_constraint("r2").type == '=' && lp.get_constraint("r2").rhs == -2);
    lp.set_sense('M');
    lp.initialize_tableau();
    const simplex::Tableau & t = lp.tableau();
    CHECK(t.rows == 2 && t.cols == 4);
    // row r1 = (2, 1 | 1, 0) / 2 then columns / their max; row r2 = (-1, -4 | 0, 1) / 4
    CHECK(t.b[0] == 4.0 && t.b[1] == -0.5);
    CHECK(t.at(0, 0) == 1.0 && t.at(1, 0) == -0.25);
    CHECK(t.at(0, 1) == 0.5 && t.at(1, 1) == -1.0);
    CHECK(t.at(0, 2) == 1.0 && t.at(1, 3) == 1.0);
    CHECK(t.c[0] == -1.0 && t.c[1] == -2.0);  // maximise: costs negated, then / column max
    // x -> u - x on column 1 with u = 3
    lp.upper_bound_substitution(1, 3.0);
    CHECK(t.c[1] == 2.0 && t.at(0, 1) == -0.5 && t.at(1, 1) == 1.0);
    CHECK(t.b[0] == 4.0 - 0.5 * 3.0 && t.b[1] == -0.5 + 3.0);
    CHECK(lp.flip_flags()[1]);
    lp.remove_variable(3);
    CHECK(!lp.has_variable("__SLACK1") && lp.get_num_vars() == 3);
}

static void basis_tests()
{
    simplex::Basis basis;
    for (simplex::Index j : {0, 1, 2, 3, 4}) basis.insert_to_non_basic(j);
    basis.insert(3);
    basis.insert(1);
    CHECK(basis.get_basic_columns() == (std::vector<simplex::Index>{3, 1}));
    CHECK(basis.get_non_basic_columns() == (std::vector<simplex::Index>{0, 2, 4}));
    CHECK(basis.contains(1) && !basis.contains(0));
    basis.update(2, 0);  // non-basic position 2 (col 4) enters, row 0 (col 3) leaves: swapped in place
    CHECK(basis.get_basic_columns() == (std::vector<simplex::Index>{4, 1}));
    CHECK(basis.get_non_basic_columns() == (std::vector<simplex::Index>{0, 2, 3}));
    CHECK(basis.show() == "4 1 ");
    basis.remove(2);
    CHECK(basis.get_non_basic_columns() == (std::vector<simplex::Index>{0, 3}));
    basis.remove(4);
    CHECK(basis.get_basic_columns() == (std::vector<simplex::Index>{1}));
    basis.clear();
    CHECK(basis.get_basic_columns().empty() && basis.get_non_basic_columns().empty());
}

static void presolve_tests()
{
    simplex::LinearProgram lp;
    const bool saw_end = simplex::run_parser_lp(
        "Maximize\n obj: x + y\nSubject To\n big: x + y <= 100\n r: x - y <= 3\nBounds\n 1 <= x <= 4\n y <= 5\nEnd\n", lp);
    CHECK(saw_end);
    CHECK(lp.get_num_rows() == 2);
    simplex::Presolve pre(lp);
    pre.run();
    CHECK(lp.get_num_rows() == 1);                 // "big" can never bind: U = 9 <= 100
    CHECK(lp.shifts().at(0) == 1.0);               // x' = x - 1
    CHECK(lp.upper_bounds().at(0) == 3.0);         // bound moves with the shift
    CHECK(lp.get_constraint("r").rhs == 2.0);      // 3 - 1 * 1
    CHECK(lp.objective_shift() == 1.0);

    std::string reason;
    simplex::Presolve::set_infeasible_handler([&reason](const std::string & why) { reason = why; });
    simplex::LinearProgram bad;
    simplex::run_parser_lp("Maximize\n obj: x\nSubject To\n r: x >= 10\nBounds\n x <= 4\nEnd\n", bad);
    simplex::Presolve(bad).run();
    CHECK(reason.find("no feasible solution") != std::string::npos);
    simplex::Presolve::set_infeasible_handler(nullptr);
}

int main()
{
    simplex::log::set_level(simplex::log::off);
    constraint_tests_1();
    linear_program_tests_1();
    model_details();
    basis_tests();
    presolve_tests();
    if (g_failed == 0) std::printf("ALL PASSED\n");
    return g_failed == 0 ? 0 : 1;
}

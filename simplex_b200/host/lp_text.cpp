// Reader of the CPLEX-LP subset the reference accepts (product code).
//
// Behavioural specification = reference src/Parser.cpp (read as a black box; restated here as a
// cursor over "squeezed" lines plus small term/relation splitters). The dialect, including the
// parts that are unusual for an .lp reader and that decide the tableau layout:
//   * a comment starts at '/', '\' or '#' and runs to the end of its line (Parser.cpp:333-351);
//   * ALL white space inside a line is removed before anything else (:30-33, :50), so keywords
//     are matched on the squeezed, lower-cased text with ':' removed: "subjectto", "bounds",
//     "end" (:61-66, :110-114);
//   * line 1 decides the sense by its first three letters (:94-112); the objective is every
//     following line up to "Subject To", joined; an optional "label:" prefix is ignored (:295-297);
//   * terms are split in front of every '+' or '-' (:134-150) — an exponent sign splits too;
//     the coefficient of a term is what precedes its first alphabetic character (:157-166);
//   * variable ids follow first appearance: objective terms left to right, then constraint terms
//     (:168-175, :211-215); naming a variable twice in the objective is an error (:169-172);
//   * a constraint is the lines joined up to and including the first one that holds '<', '=' or
//     '>' (:72-92); its relation is the FIRST of those characters (:180-181), its right-hand
//     side starts at the first digit or '-' after it (:187-191);
//   * a row without "label:" gets the label "" — and since labels are unique keys, every
//     unlabelled row after the first is dropped, although its variables are still created
//     (:125-133, LinearProgram.cpp:197-200);
//   * a bounds line is "lo <= x <= hi", "x <= hi" or "lo <= x" (also with '<' or '='); a lower
//     bound is stored unless |lo| < 1e-7, an upper bound unless it is +inf (:220-268).
// Errors: the reference throws std::string for "domain" errors (caught by its main, which prints
// "Exception: ...") and `const char *` / std::logic_error for malformed text (uncaught there:
// the process aborts). The same exception types leave this reader.
#include "lp_text.h"

#include "log.h"

#include <cctype>
#include <chrono>
#include <cstdlib>
#include <limits>
#include <stdexcept>
#include <utility>
#include <vector>

namespace
{
    using simplex::LinearProgram;

    bool is_alpha(char ch) { return std::isalpha(static_cast<unsigned char>(ch)) != 0; }
    bool is_digit(char ch) { return std::isdigit(static_cast<unsigned char>(ch)) != 0; }

    // Leading-number conversion with std::stod's contract: longest valid prefix, otherwise throw.
    double leading_number(const std::string & text)
    {
        const char * begin = text.c_str();
        char * stop = nullptr;
        const double v = std::strtod(begin, &stop);
        if (stop == begin) throw std::invalid_argument("stod");
        return v;
    }

    // Text with every comment (from '/', '\\' or '#' to the end of that line) cut out.
    std::string without_comments(std::string_view text)
    {
        std::string kept;
        kept.reserve(text.size());
        size_t from = 0;
        while (from < text.size())
        {
            const size_t mark = text.find_first_of("/\\#", from);
            if (mark == std::string_view::npos)
            {
                kept.append(text.substr(from));
                break;
            }
            kept.append(text.substr(from, mark - from));
            const size_t eol = text.find('\n', mark);
            if (eol == std::string_view::npos) break;  // comment runs to the end of the input
            from = eol + 1;                            // the newline goes with the comment
        }
        return kept;
    }

    // Hands out the non-empty lines of a text with all white space squeezed out of them.
    class SqueezedLines
    {
    public:
        explicit SqueezedLines(const std::string & text) : m_text(text) {}
        bool exhausted() const { return m_at >= m_text.size(); }

        // Next line that is non-empty after squeezing; "" once the text is exhausted.
        std::string next()
        {
            std::string line;
            while (line.empty() && !exhausted())
            {
                while (!exhausted())
                {
                    const char ch = m_text[m_at++];
                    if (ch == '\n') break;
                    if (!std::isspace(static_cast<unsigned char>(ch))) line.push_back(ch);
                }
            }
            return line;
        }

    private:
        const std::string & m_text;
        size_t m_at = 0;
    };

    // Lower-cased line with ':' removed: the form in which section keywords are recognised.
    std::string keyword_form(const std::string & line)
    {
        std::string k;
        for (char ch : line)
            if (ch != ':') k.push_back(static_cast<char>(std::tolower(static_cast<unsigned char>(ch))));
        return k;
    }

    // "+3x-y+2.5z" -> {"+3x", "-y", "+2.5z"}; always at least one (possibly empty) piece.
    std::vector<std::string> split_terms(const std::string & expr)
    {
        std::vector<std::string> pieces(1);
        for (char ch : expr)
        {
            if ((ch == '+' || ch == '-') && !pieces.back().empty()) pieces.emplace_back();
            pieces.back().push_back(ch);
        }
        return pieces;
    }

    struct Term
    {
        double coeff;
        std::string name;
    };

    // Coefficient = text in front of the first letter ("" and a lone sign mean +-1).
    Term read_term(const std::string & piece)
    {
        size_t k = 0;
        while (k < piece.size() && !is_alpha(piece[k])) ++k;
        Term t;
        t.name = piece.substr(k);
        const bool lone_sign = k == 1 && (piece[0] == '+' || piece[0] == '-');
        if (k == 0 || lone_sign)
            t.coeff = (!piece.empty() && piece[0] == '-') ? -1.0 : 1.0;
        else
            t.coeff = leading_number(piece.substr(0, k));
        return t;
    }

    // "label:expr" -> (label, expr); no ':' -> ("", line).
    std::pair<std::string, std::string> split_label(const std::string & line)
    {
        const size_t colon = line.find(':');
        if (colon == std::string::npos) return {std::string(), line};
        return {line.substr(0, colon), line.substr(colon + 1)};
    }

    void read_sense(const std::string & line, LinearProgram & lp)
    {
        if (line.size() < 3) throw "Expected keyword: Minimize or Maximize";
        std::string head = line.substr(0, 3);
        for (char & ch : head) ch = static_cast<char>(std::tolower(static_cast<unsigned char>(ch)));
        if (head == "min") lp.set_sense('m');
        if (head == "max") lp.set_sense('M');
    }

    void read_objective(const std::string & expr, LinearProgram & lp)
    {
        for (const std::string & piece : split_terms(expr))
        {
            bool named = false;
            for (char ch : piece) named = named || is_alpha(ch);
            if (!named) throw std::string("Invalid token in objective function: " + piece);
            const Term t = read_term(piece);
            if (lp.has_variable(t.name)) throw std::string("Variable already defined in objective function: " + t.name);
            lp.add_variable(t.name, t.coeff);
        }
    }

    void read_constraint(const std::string & label, const std::string & expr, LinearProgram & lp)
    {
        const size_t rel = expr.find_first_of("<=>");
        if (rel == std::string::npos) throw std::out_of_range("constraint without a relation: " + expr);
        size_t value_at = rel + 1;  // right-hand side: first digit or '-' after the relation sign
        for (;; ++value_at)
        {
            if (value_at >= expr.size()) throw std::out_of_range("constraint without a right-hand side: " + expr);
            if (is_digit(expr[value_at]) || expr[value_at] == '-') break;
        }
        simplex::Constraint row(expr[rel], leading_number(expr.substr(value_at)));
        for (const std::string & piece : split_terms(expr.substr(0, rel)))
        {
            const Term t = read_term(piece);
            if (!lp.has_variable(t.name)) lp.add_variable(t.name);
            row.add_term(t.name, t.coeff);
        }
        lp.add_constraint(label, std::move(row));
    }

    void read_bound(const std::string & line, LinearProgram & lp)
    {
        constexpr double inf = std::numeric_limits<double>::infinity();
        const size_t first = line.find_first_of("<=");
        const size_t last = line.find_last_of("<=");
        const std::string left = line.substr(0, first);  // first == npos: the whole line
        const std::string right = line.substr(last + 1);
        double lo = 0.0, hi = inf;
        std::string name;
        if (last - first <= 1)
        {
            // one relation ("<", "=", "<=" ...): the side that holds a letter is the variable
            bool left_is_name = false;
            for (char ch : left) left_is_name = left_is_name || is_alpha(ch);
            if (left_is_name)
            {
                hi = leading_number(right);
                name = left;
            }
            else
            {
                lo = leading_number(left);
                name = right;
            }
        }
        else
        {
            // "lo <= name <= hi": the name sits between the two two-character relations
            name = line.substr(first + 2, last - first - 3);
            lo = leading_number(left);
            hi = leading_number(right);
        }
        const simplex::Index id = lp.get_variable_id(name);
        if (!simplex::is_float_zero(lo)) lp.set_lower_bound(id, lo);
        if (hi < inf) lp.set_upper_bound(id, hi);
    }
}  // namespace

namespace simplex
{
    bool run_parser_lp(std::string_view input, LinearProgram & lp)
    {
        const auto t0 = std::chrono::steady_clock::now();
        const std::string text = without_comments(input);
        SqueezedLines lines(text);
        bool complete = false;

        enum class Section { sense, objective, rows, bounds } section = Section::sense;
        while (!lines.exhausted() && !complete)
        {
            switch (section)
            {
                case Section::sense:
                    read_sense(lines.next(), lp);
                    section = Section::objective;
                    break;

                case Section::objective:
                {
                    std::string joined;
                    while (!lines.exhausted())
                    {
                        const std::string line = lines.next();
                        const std::string key = keyword_form(line);
                        if (key == "subjectto" || key == "end") break;
                        joined += line;
                    }
                    read_objective(split_label(joined).second, lp);
                    section = Section::rows;
                    break;
                }

                case Section::rows:
                {
                    std::string joined, key;
                    while (!lines.exhausted())
                    {
                        const std::string line = lines.next();
                        key = keyword_form(line);
                        if (key == "bounds" || key == "end") break;
                        key.clear();
                        joined += line;
                        if (line.find_first_of("<=>") != std::string::npos) break;
                    }
                    if (key == "bounds")
                        section = Section::bounds;
                    else if (key == "end")
                        complete = true;
                    else
                    {
                        const auto [label, expr] = split_label(joined);
                        read_constraint(label, expr, lp);
                    }
                    break;
                }

                case Section::bounds:
                {
                    const std::string line = lines.next();
                    if (keyword_form(line) == "end")
                        complete = true;
                    else
                        read_bound(line, lp);
                    break;
                }
            }
        }
        const auto us = std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0);
        SB_LOG(info) << "Parser: elapsed " << us.count() << " us";
        return complete;
    }
}  // namespace simplex

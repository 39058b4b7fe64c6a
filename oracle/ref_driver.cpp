// TEST INFRASTRUCTURE — not product code.
//
// Driver around the UNMODIFIED reference sources (/root/reference/src/*.cpp, compiled where
// they lie by oracle/Makefile against oracle/eigen_shim). It exists because the reference has
// no getters (Simplex.h:31-33 keeps the solution private and write() prints 6 decimals,
// Simplex.cpp:364) and no way to feed synthetic LPs except its model API.
//
// What it does:
//   lp <file.lp>                       same sequence as main.cpp:52-61 (parse, presolve, solve)
//   synth <m> <n> <n_eq> <seed>        G(m,n,n_eq,seed) through
//                                      LinearProgram::add_variable/add_constraint, then presolve+solve
//   seam <m> <n> <n_eq> <seed> <K>     fill the dense tableau directly and call the seam
//                                      Simplex::simplex(Basis&) (Simplex.h:25); stop after K
//                                      iterations and report seconds/iteration (CPU baseline)
// Options: --out FILE (JSON), --dump-tableau, --via-solve (call Simplex::solve() itself instead
// of the replicated 6-line body, used to check both give the same trace).
//
// Everything is read out by tapping the Logger's stdout stream (Logger.h:81-82): the
// reference prints "SIMPLEX iteration", "Basis:", "leaving basis:" for every iteration
// (Simplex.cpp:278-279, Basis.cpp:55-56), and by reading privates (the access hack below
// touches only this translation unit; the reference objects are compiled unchanged).
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <random>
#include <set>
#include <sstream>
#include <streambuf>
#include <string>
#include <unistd.h>
#include <vector>

#include "Eigen/Dense"

#define private public
#include "Parser.h"
#include "Presolve.h"
#include "Simplex.h"
#include "InitialBasis.h"
#include "Logger.h"
#undef private

namespace
{
    using Clock = std::chrono::steady_clock;

    struct Pivot
    {
        long leaving_col, row, entering_col, pos;
    };

    struct CallRecord
    {
        long iterations = 0;
        std::vector<Pivot> pivots;
        std::vector<long> first_basis, last_basis;
        bool unbounded = false;
        bool optimal = false;
        // entry dump
        long M = 0, N = 0;
        std::vector<double> A, b, c, ub;
        std::vector<long> basic_in, nonbasic_in, basic_out, nonbasic_out;
    };

    struct State
    {
        simplex::LinearProgram * lp = nullptr;
        std::vector<CallRecord> calls;
        bool dump_tableau = false;
        bool saw_infeasible_log = false;
        // timing mode
        long stop_after = -1;
        std::vector<double> iter_t;
        std::string out_path;
        bool finished = false;
        std::string status = "incomplete";
        std::string exception_text;
    } g;

    void dump_entry(CallRecord & r)
    {
        auto & lp = *g.lp;
        r.M = static_cast<long>(lp.get_num_rows());
        r.N = static_cast<long>(lp.get_num_vars());
        if (!g.dump_tableau) return;
        r.A.assign(lp.matrix_A.data(), lp.matrix_A.data() + lp.matrix_A.rows() * lp.matrix_A.cols());
        r.b.assign(lp.vector_b.data(), lp.vector_b.data() + lp.vector_b.size());
        r.c.assign(lp.vector_c.data(), lp.vector_c.data() + lp.vector_c.size());
        r.ub.clear();
        for (long j = 0; j < r.N; ++j)
        {
            auto it = lp.var_ubnd.find(j);
            r.ub.push_back(it == lp.var_ubnd.end() ? -1.0 : it->second);
        }
    }

    std::vector<long> parse_longs(const std::string & s)
    {
        std::vector<long> v;
        std::istringstream is(s);
        long x;
        while (is >> x) v.push_back(x);
        return v;
    }

    void write_json_and_exit_if_timing();

    void on_line(const std::string & raw)
    {
        // strip "[time][sev] "
        std::string line = raw;
        if (!line.empty() && line[0] == '[')
        {
            auto p = line.find("] ");
            if (p != std::string::npos) line = line.substr(p + 2);
        }
        if (line.rfind("SIMPLEX iteration: ", 0) == 0)
        {
            const long k = std::strtol(line.c_str() + 19, nullptr, 10);
            if (k == 1)
            {
                g.calls.emplace_back();
                dump_entry(g.calls.back());
            }
            if (!g.calls.empty()) g.calls.back().iterations = k;
            if (g.stop_after >= 0)
            {
                g.iter_t.push_back(std::chrono::duration<double>(Clock::now().time_since_epoch()).count());
                if (static_cast<long>(g.iter_t.size()) > g.stop_after) write_json_and_exit_if_timing();
            }
        }
        else if (line.rfind("Basis: ", 0) == 0)
        {
            if (g.calls.empty() || g.stop_after >= 0) return;
            auto v = parse_longs(line.substr(7));
            if (g.calls.back().first_basis.empty()) g.calls.back().first_basis = v;
            g.calls.back().last_basis = v;
        }
        else if (line.rfind("leaving basis: ", 0) == 0)
        {
            Pivot p{};
            if (std::sscanf(line.c_str(), "leaving basis: %ld (%ld), entering basis: %ld (%ld)", &p.leaving_col, &p.row,
                            &p.entering_col, &p.pos) == 4 &&
                !g.calls.empty())
            {
                g.calls.back().pivots.push_back(p);
            }
        }
        else if (line.rfind("LP IS UNBOUNDED", 0) == 0)
        {
            if (!g.calls.empty()) g.calls.back().unbounded = true;
        }
        else if (line.rfind("OPTIMAL X FOUND", 0) == 0)
        {
            if (!g.calls.empty()) g.calls.back().optimal = true;
        }
        else if (line.rfind("LP is infeasible", 0) == 0)
        {
            g.saw_infeasible_log = true;
        }
    }

    class LineTap : public std::streambuf
    {
    protected:
        int overflow(int ch) override
        {
            if (ch == traits_type::eof()) return 0;
            put(static_cast<char>(ch));
            return ch;
        }
        std::streamsize xsputn(const char * s, std::streamsize n) override
        {
            for (std::streamsize i = 0; i < n; ++i) put(s[i]);
            return n;
        }

    public:
        void flush_line()
        {
            if (!m_cur.empty()) on_line(m_cur);
            m_cur.clear();
        }

    private:
        std::string m_cur;
        void put(char c)
        {
            if (c == '\n')
            {
                if (!m_cur.empty()) on_line(m_cur);
                m_cur.clear();
            }
            else
            {
                m_cur += c;
            }
        }
    };

    LineTap g_tap;

    // ---------------------------------------------------------------- synthetic LP G(m,n,n_eq,seed)
    struct SynthLP
    {
        long m, n, n_eq;
        std::vector<double> c;    // n
        std::vector<double> a;    // row-major m*n
        std::vector<double> rhs;  // m
    };

    SynthLP make_synth(long m, long n, long n_eq, unsigned long long seed)
    {
        SynthLP s;
        s.m = m;
        s.n = n;
        s.n_eq = n_eq;
        std::mt19937_64 rng(seed);
        auto U = [&rng]() { return static_cast<double>(rng() >> 11) * (1.0 / 9007199254740992.0); };
        s.c.resize(n);
        std::vector<double> x0(n);
        for (long j = 0; j < n; ++j) s.c[j] = -(0.5 + U());
        for (long j = 0; j < n; ++j) x0[j] = U();
        s.a.resize(static_cast<size_t>(m) * n);
        s.rhs.resize(m);
        for (long i = 0; i < m; ++i)
        {
            double * row = s.a.data() + static_cast<size_t>(i) * n;
            double acc = 0.0;
            for (long j = 0; j < n; ++j)
            {
                row[j] = 0.05 + 0.95 * U();
                const double t = row[j] * x0[j];
                acc = acc + t;
            }
            if (i >= n_eq)
            {
                const double f = 1.0 + 0.5 * U();
                acc = f * acc;
            }
            s.rhs[i] = acc;
        }
        return s;
    }

    void name7(char * buf, char prefix, long k) { std::snprintf(buf, 24, "%c%07ld", prefix, k); }

    void build_model_via_api(const SynthLP & s, simplex::LinearProgram & lp)
    {
        char nm[24];
        lp.set_sense('m');
        for (long j = 0; j < s.n; ++j)
        {
            name7(nm, 'x', j);
            lp.add_variable(nm, s.c[j]);
        }
        for (long i = 0; i < s.m; ++i)
        {
            simplex::Constraint con(i < s.n_eq ? '=' : '<', s.rhs[i]);
            for (long j = 0; j < s.n; ++j)
            {
                name7(nm, 'x', j);
                con.add_term(nm, s.a[static_cast<size_t>(i) * s.n + j]);
            }
            name7(nm, 'R', i);
            lp.add_constraint(nm, std::move(con));
        }
    }

    // ---------------------------------------------------------------- JSON
    void jnum(std::ostream & o, double v)
    {
        char buf[40];
        std::snprintf(buf, sizeof buf, "%.17g", v);
        o << buf;
    }
    template <class T>
    void jarr_l(std::ostream & o, const std::vector<T> & v)
    {
        o << "[";
        for (size_t i = 0; i < v.size(); ++i) o << (i ? "," : "") << v[i];
        o << "]";
    }
    void jarr_d(std::ostream & o, const std::vector<double> & v)
    {
        o << "[";
        for (size_t i = 0; i < v.size(); ++i)
        {
            if (i) o << ",";
            jnum(o, v[i]);
        }
        o << "]";
    }
    std::string jstr(const std::string & s)
    {
        std::string r = "\"";
        for (char c : s)
        {
            if (c == '"' || c == '\\')
            {
                r += '\\';
                r += c;
            }
            else if (c == '\n')
                r += "\\n";
            else
                r += c;
        }
        return r + "\"";
    }

    simplex::Simplex * g_sx = nullptr;

    void write_json(std::ostream & o)
    {
        o << "{\n \"status\": " << jstr(g.status) << ",\n";
        if (!g.exception_text.empty()) o << " \"exception\": " << jstr(g.exception_text) << ",\n";
        if (g_sx)
        {
            o << " \"objective\": ";
            jnum(o, g_sx->m_objective_value);
            o << ",\n \"solution\": {";
            bool first = true;
            for (const auto & [k, v] : g_sx->m_solution)
            {
                o << (first ? "" : ",") << jstr(k) << ":";
                jnum(o, v);
                first = false;
            }
            o << "},\n";
        }
        if (g.lp)
        {
            o << " \"sense\": " << jstr(std::string(1, g.lp->sense)) << ",\n";
            o << " \"var_names\": [";
            bool first = true;
            for (const auto & [id, nm] : g.lp->variable_id_to_name)
            {
                o << (first ? "" : ",") << jstr(nm);
                first = false;
            }
            o << "],\n \"var_shifts\": {";
            first = true;
            for (const auto & [id, sh] : g.lp->var_shifts)
            {
                o << (first ? "" : ",") << "\"" << id << "\":";
                jnum(o, sh);
                first = false;
            }
            o << "},\n \"row_labels\": [";
            first = true;
            for (const auto & [lab, con] : g.lp->constraints)
            {
                o << (first ? "" : ",") << jstr(lab);
                first = false;
            }
            o << "],\n";
        }
        o << " \"calls\": [";
        for (size_t k = 0; k < g.calls.size(); ++k)
        {
            const auto & r = g.calls[k];
            o << (k ? ",\n  {" : "\n  {");
            o << "\"M\":" << r.M << ",\"N\":" << r.N << ",\"iterations\":" << r.iterations
              << ",\"optimal\":" << (r.optimal ? "true" : "false") << ",\"unbounded\":" << (r.unbounded ? "true" : "false");
            o << ",\"pivots\":[";
            for (size_t i = 0; i < r.pivots.size(); ++i)
            {
                const auto & p = r.pivots[i];
                o << (i ? "," : "") << "[" << p.leaving_col << "," << p.row << "," << p.entering_col << "," << p.pos << "]";
            }
            o << "],\"first_basis\":";
            jarr_l(o, r.first_basis);
            o << ",\"last_basis\":";
            jarr_l(o, r.last_basis);
            if (!r.basic_in.empty() || !r.nonbasic_in.empty())
            {
                o << ",\"basic_in\":";
                jarr_l(o, r.basic_in);
                o << ",\"nonbasic_in\":";
                jarr_l(o, r.nonbasic_in);
                o << ",\"basic_out\":";
                jarr_l(o, r.basic_out);
                o << ",\"nonbasic_out\":";
                jarr_l(o, r.nonbasic_out);
            }
            if (!r.A.empty())
            {
                o << ",\"A\":";
                jarr_d(o, r.A);
                o << ",\"b\":";
                jarr_d(o, r.b);
                o << ",\"c\":";
                jarr_d(o, r.c);
                o << ",\"ub\":";
                jarr_d(o, r.ub);
            }
            o << "}";
        }
        o << "\n ]";
        if (g.stop_after >= 0)
        {
            o << ",\n \"iter_timestamps\": ";
            jarr_d(o, g.iter_t);
        }
        o << "\n}\n";
    }

    void emit(bool flush_tap = true)
    {
        if (g.finished) return;
        g.finished = true;
        if (flush_tap) g_tap.flush_line();
        if (g.out_path.empty())
        {
            write_json(std::cerr);
        }
        else
        {
            std::ofstream f(g.out_path);
            write_json(f);
        }
    }

    void write_json_and_exit_if_timing()
    {
        g.status = "timing_stop";
        g_sx = nullptr;  // solution maps are not meaningful mid-run
        emit(false);     // called from inside the tap: do not re-enter it
        std::fflush(nullptr);
        _exit(0);
    }

    void at_exit_hook()
    {
        // Presolve calls std::exit(0) on trivially infeasible input (Presolve.cpp:44,119).
        if (!g.finished)
        {
            if (g.status == "incomplete") g.status = "exit_called";
            emit();
        }
    }

    std::vector<long> to_long(const std::vector<Eigen::Index> & v) { return std::vector<long>(v.begin(), v.end()); }

    // The body of Simplex::solve() (Simplex.cpp:22-45) replayed call by call so that the
    // Basis object handed to each seam call can be recorded. --via-solve uses solve() itself.
    void run_solve(simplex::LinearProgram & lp, simplex::Simplex & sx, bool via_solve)
    {
        if (via_solve)
        {
            sx.solve();
            return;
        }
        simplex::InitialBasis basis_initializer(lp, sx);
        simplex::Basis init_basis = basis_initializer.get_basis();
        if (sx.m_objective_value > 0.0)
        {
            g.saw_infeasible_log = true;
        }
        else
        {
            lp.initialize_tableau();
            const auto bi = to_long(init_basis.get_basic_columns());
            const auto ni = to_long(init_basis.get_non_basic_columns());
            const size_t ncalls_before = g.calls.size();
            sx.simplex(init_basis);
            g_tap.flush_line();
            if (g.calls.size() > ncalls_before)
            {
                auto & r = g.calls.back();
                r.basic_in = bi;
                r.nonbasic_in = ni;
                r.basic_out = to_long(init_basis.get_basic_columns());
                r.nonbasic_out = to_long(init_basis.get_non_basic_columns());
            }
        }
    }

    int usage()
    {
        std::fprintf(stderr,
                     "usage: ref_driver lp FILE | synth M N NEQ SEED | seam M N NEQ SEED K  [--out F] [--dump-tableau] "
                     "[--via-solve]\n");
        return 2;
    }
}  // namespace

int main(int argc, char ** argv)
{
    std::vector<std::string> pos;
    bool via_solve = false;
    for (int i = 1; i < argc; ++i)
    {
        std::string a = argv[i];
        if (a == "--out" && i + 1 < argc)
            g.out_path = argv[++i];
        else if (a == "--dump-tableau")
            g.dump_tableau = true;
        else if (a == "--via-solve")
            via_solve = true;
        else
            pos.push_back(a);
    }
    if (pos.empty()) return usage();

    // Same stream formatting as main.cpp:17-18 (only affects the human-readable log).
    std::cout.precision(4);
    std::cout.setf(std::ios_base::showpoint);
    std::streambuf * old = std::cout.rdbuf(&g_tap);
    std::atexit(at_exit_hook);

    simplex::LinearProgram lp;
    g.lp = &lp;
    simplex::Simplex sx(lp);

    try
    {
        if (pos[0] == "lp" && pos.size() == 2)
        {
            std::ifstream fin(pos[1]);
            if (!fin)
            {
                std::cout.rdbuf(old);
                std::fprintf(stderr, "cannot open %s\n", pos[1].c_str());
                g.finished = true;
                return 1;
            }
            std::stringstream ss;
            ss << fin.rdbuf();
            std::string buffer = ss.str();
            simplex::run_parser_lp(buffer, lp);
            simplex::Presolve pre(lp);
            pre.run();
            g_sx = &sx;
            run_solve(lp, sx, via_solve);
        }
        else if (pos[0] == "synth" && pos.size() == 5)
        {
            const auto s = make_synth(std::atol(pos[1].c_str()), std::atol(pos[2].c_str()), std::atol(pos[3].c_str()),
                                      std::strtoull(pos[4].c_str(), nullptr, 10));
            build_model_via_api(s, lp);
            simplex::Presolve pre(lp);
            pre.run();
            g_sx = &sx;
            run_solve(lp, sx, via_solve);
        }
        else if (pos[0] == "seam" && pos.size() == 6)
        {
            // Dense fast path: fill matrix_A/vector_b/vector_c as initialize_tableau would
            // (LinearProgram.cpp:64-97: rows in label order = index order here, slack per '<='
            // row appended in row order, LinearProgram.cpp:110-141) and call the seam directly.
            // Only meaningful for n_eq == 0 (all-slack start; no Phase I needed).
            const long m = std::atol(pos[1].c_str()), n = std::atol(pos[2].c_str()), n_eq = std::atol(pos[3].c_str());
            if (n_eq != 0)
            {
                std::fprintf(stderr, "seam mode supports n_eq == 0 only\n");
                return 2;
            }
            g.stop_after = std::atol(pos[5].c_str());
            // G(m,n,0,seed) generated row block by row block straight into matrix_A (same draw order as
            // make_synth: costs, x0, then per row n coefficients and the slack factor), so that the widest
            // BASELINE shape (m=8192, n=262144: 17.7 GB of matrix_A) needs no second copy of the coefficients.
            std::mt19937_64 rng(std::strtoull(pos[4].c_str(), nullptr, 10));
            auto U = [&rng]() { return static_cast<double>(rng() >> 11) * (1.0 / 9007199254740992.0); };
            std::vector<double> cost(n), x0(n);
            for (long j = 0; j < n; ++j) cost[j] = -(0.5 + U());
            for (long j = 0; j < n; ++j) x0[j] = U();
            char nm[24];
            for (long j = 0; j < n; ++j)
            {
                name7(nm, 'x', j);
                lp.add_variable(nm, cost[j]);
            }
            for (long i = 0; i < m; ++i)
            {
                name7(nm, 'R', i);
                lp.add_constraint(nm, simplex::Constraint('=', 0.0));
                std::string sl = "__SLACK" + std::to_string(i);
                lp.add_variable(sl);
            }
            const long N = n + m;
            lp.vector_b.resize(m);
            lp.vector_c.resize(N);
            lp.matrix_A.resize(m, N);
            lp.vector_c.setZero();
            lp.matrix_A.setZero();
            constexpr long RB = 64;
            std::vector<double> blk(static_cast<size_t>(RB) * n);
            double * Adata = lp.matrix_A.data();  // column-major, leading dimension m (LinearProgram.cpp:101)
            for (long i0 = 0; i0 < m; i0 += RB)
            {
                const long nb = std::min(RB, m - i0);
                for (long r = 0; r < nb; ++r)
                {
                    double * row = blk.data() + static_cast<size_t>(r) * n;
                    double acc = 0.0;
                    for (long j = 0; j < n; ++j)
                    {
                        row[j] = 0.05 + 0.95 * U();
                        const double t = row[j] * x0[j];
                        acc = acc + t;
                    }
                    const double f = 1.0 + 0.5 * U();
                    lp.vector_b[i0 + r] = f * acc;
                }
                for (long j = 0; j < n; ++j)
                    for (long r = 0; r < nb; ++r) Adata[static_cast<size_t>(j) * m + i0 + r] = blk[static_cast<size_t>(r) * n + j];
            }
            for (long i = 0; i < m; ++i) lp.matrix_A(i, n + i) = 1.0;
            for (long j = 0; j < n; ++j) lp.vector_c[j] = cost[j];
            lp.rescale_matrix();
            simplex::Basis basis;
            for (long i = 0; i < m; ++i) basis.insert(n + i);
            for (long j = 0; j < n; ++j) basis.insert_to_non_basic(j);
            sx.simplex(basis);
        }
        else
        {
            std::cout.rdbuf(old);
            g.finished = true;
            return usage();
        }
        g_tap.flush_line();
        bool any_unbounded = false;
        for (const auto & r : g.calls) any_unbounded = any_unbounded || r.unbounded;
        if (g.saw_infeasible_log)
            g.status = "infeasible";
        else if (any_unbounded)
            g.status = "unbounded";
        else if (!g.calls.empty() && g.calls.back().optimal)
            g.status = "optimal";
        else
            g.status = "unknown";
    }
    catch (const std::string & msg)
    {
        g.status = "exception_string";
        g.exception_text = msg;
    }
    catch (const char * msg)
    {
        // main.cpp:63 does not catch these: the stock binary aborts (SIGABRT).
        g.status = "exception_cstr";
        g.exception_text = "(const char* thrown; the stock CLI aborts here)";
        (void)msg;  // may dangle (InitialBasis.cpp:113 throws c_str() of a temporary)
    }
    std::cout.rdbuf(old);
    emit();
    return 0;
}

// C entry points of the host library (product code): the text model, presolve and phase control
// behind plain C types, for bindings (simplex_b200/lp.py uses them through ctypes) and for the
// parity tests, which drive the host halves of the two-phase flow one step at a time and compare
// every tableau that would be handed to the device with the reference's.
//
// The step functions never compute a pivot: the loop exists only on the GPU (sbh_solve).
#include "dense_solver.h"
#include "dense_tableau.h"
#include "lp_text.h"
#include "presolve.h"
#include "solver.h"

#include <cstdio>
#include <cstring>
#include <memory>
#include <sstream>
#include <stdexcept>

namespace
{
    struct PresolveStop
    {
        std::string reason;
    };

    struct Session
    {
        simplex::LinearProgram lp;
        std::unique_ptr<simplex::Simplex> solver;
        std::unique_ptr<simplex::InitialBasis> crash;
        simplex::Basis basis;
        std::string error;
        bool presolve_infeasible = false;
        bool saw_end = false;
    };

    template <class F>
    int guarded(Session * s, F && body)
    {
        try
        {
            body();
            return 0;
        }
        catch (const std::string & msg)
        {
            s->error = "Exception: " + msg;  // what the reference's main prints
            return 1;
        }
        catch (const char * msg)
        {
            s->error = std::string("fatal: ") + msg;  // uncaught in the reference: process aborts
            return 2;
        }
        catch (const PresolveStop & stop)
        {
            s->error = stop.reason;
            s->presolve_infeasible = true;
            return 3;
        }
        catch (const std::exception & e)
        {
            s->error = std::string("fatal: ") + e.what();
            return 2;
        }
    }
}  // namespace

extern "C"
{
    void * sbh_open(void)
    {
        auto * s = new Session();
        s->solver = std::make_unique<simplex::Simplex>(s->lp);
        s->crash = std::make_unique<simplex::InitialBasis>(s->lp, *s->solver);
        return s;
    }

    void sbh_close(void * h) { delete static_cast<Session *>(h); }

    const char * sbh_error(void * h) { return static_cast<Session *>(h)->error.c_str(); }

    // Parse the text of an .lp file and (optionally) presolve. 0 ok, 1 "Exception: ...", 2 fatal
    // (the reference aborts), 3 presolve proved infeasibility (the reference exits with status 0).
    int sbh_read(void * h, const char * text, int presolve)
    {
        Session * s = static_cast<Session *>(h);
        return guarded(s, [&] {
            s->saw_end = simplex::run_parser_lp(text, s->lp);
            if (presolve)
            {
                simplex::Presolve::set_infeasible_handler([](const std::string & why) { throw PresolveStop{why}; });
                simplex::Presolve pre(s->lp);
                try
                {
                    pre.run();
                }
                catch (...)
                {
                    simplex::Presolve::set_infeasible_handler(nullptr);
                    throw;
                }
                simplex::Presolve::set_infeasible_handler(nullptr);
            }
        });
    }

    // G(m, n, n_eq, seed) of the benchmark configurations entered through the
    // model API the way a user of the reference would (add_variable / add_constraint), then presolve.
    int sbh_build_synth(void * h, int64_t m, int64_t n, int64_t n_eq, uint64_t seed)
    {
        Session * s = static_cast<Session *>(h);
        return guarded(s, [&] {
            const simplex_b200::SynthLP g = simplex_b200::make_synth(m, n, n_eq, seed);
            char name[24];
            std::vector<std::string> var_names(static_cast<size_t>(n));
            for (int64_t j = 0; j < n; ++j)
            {
                std::snprintf(name, sizeof name, "x%07lld", static_cast<long long>(j));
                var_names[static_cast<size_t>(j)] = name;
                s->lp.add_variable(name, g.c[static_cast<size_t>(j)]);
            }
            s->lp.set_sense('m');
            for (int64_t i = 0; i < m; ++i)
            {
                simplex::Constraint row(i < n_eq ? '=' : '<', g.rhs[static_cast<size_t>(i)]);
                const double * a = g.a.data() + static_cast<size_t>(i) * static_cast<size_t>(n);
                for (int64_t j = 0; j < n; ++j) row.add_term(var_names[static_cast<size_t>(j)], a[j]);
                std::snprintf(name, sizeof name, "R%07lld", static_cast<long long>(i));
                s->lp.add_constraint(name, std::move(row));
            }
            simplex::Presolve pre(s->lp);
            pre.run();
        });
    }

    int sbh_saw_end(void * h) { return static_cast<Session *>(h)->saw_end ? 1 : 0; }

    void sbh_set_options(void * h, int device, int pricing, int engine, int dual_mode, int64_t chunk_iters)
    {
        simplex::DeviceOptions & o = static_cast<Session *>(h)->solver->options();
        o.device = device;
        o.pricing = static_cast<simplex::Pricing>(pricing);
        o.engine = engine;
        o.dual_mode = dual_mode;
        if (chunk_iters > 0) o.chunk_iters = chunk_iters;
    }

    // ---- the host halves of Simplex::solve(), one at a time
    int sbh_open_phase_one(void * h, int64_t * n_artificial)
    {
        Session * s = static_cast<Session *>(h);
        return guarded(s, [&] { *n_artificial = static_cast<int64_t>(s->crash->open_phase_one(s->basis)); });
    }

    int sbh_close_phase_one(void * h)
    {
        Session * s = static_cast<Session *>(h);
        return guarded(s, [&] { s->crash->close_phase_one(s->basis); });
    }

    int sbh_initialize_tableau(void * h)
    {
        Session * s = static_cast<Session *>(h);
        return guarded(s, [&] { s->lp.initialize_tableau(); });
    }

    void sbh_dims(void * h, int64_t * rows, int64_t * cols, int64_t * n_basic, int64_t * n_non_basic, int * bounded)
    {
        Session * s = static_cast<Session *>(h);
        *rows = s->lp.tableau().rows;
        *cols = s->lp.tableau().cols;
        *n_basic = static_cast<int64_t>(s->basis.get_basic_columns().size());
        *n_non_basic = static_cast<int64_t>(s->basis.get_non_basic_columns().size());
        *bounded = s->lp.has_non_trivial_upper_bounds() ? 1 : 0;
    }

    // A column-major [cols][rows], b [rows], c [cols], ub [cols] (DBL_MAX = none)
    void sbh_copy_tableau(void * h, double * A, double * b, double * c, double * ub)
    {
        Session * s = static_cast<Session *>(h);
        const simplex::Tableau & t = s->lp.tableau();
        if (A) std::memcpy(A, t.A.data(), t.A.size() * sizeof(double));
        if (b) std::memcpy(b, t.b.data(), t.b.size() * sizeof(double));
        if (c) std::memcpy(c, t.c.data(), t.c.size() * sizeof(double));
        if (ub)
        {
            const std::vector<double> u = s->lp.dense_upper_bounds();
            std::memcpy(ub, u.data(), std::min<size_t>(u.size(), static_cast<size_t>(t.cols)) * sizeof(double));
        }
    }

    void sbh_copy_basis(void * h, int64_t * basic, int64_t * non_basic)
    {
        Session * s = static_cast<Session *>(h);
        std::memcpy(basic, s->basis.get_basic_columns().data(), s->basis.get_basic_columns().size() * sizeof(int64_t));
        std::memcpy(non_basic, s->basis.get_non_basic_columns().data(),
                    s->basis.get_non_basic_columns().size() * sizeof(int64_t));
    }

    void sbh_set_basis(void * h, const int64_t * basic, int64_t n_basic, const int64_t * non_basic, int64_t n_non_basic)
    {
        Session * s = static_cast<Session *>(h);
        s->basis.basic_columns().assign(basic, basic + n_basic);
        s->basis.non_basic_columns().assign(non_basic, non_basic + n_non_basic);
    }

    // Host half of the optimal exit of the loop: x_B [rows] and the flip flags [cols] as the device
    // returns them (sb200_get_xB / sb200_get_flip_flags).
    int sbh_absorb_optimum(void * h, const double * x_B, const uint8_t * flags)
    {
        Session * s = static_cast<Session *>(h);
        return guarded(s, [&] {
            const simplex::Tableau & t = s->lp.tableau();
            std::vector<double> x(x_B, x_B + t.rows);
            std::vector<uint8_t> f(static_cast<size_t>(t.cols), 0);
            if (flags) f.assign(flags, flags + t.cols);
            s->lp.flip_flags().assign(static_cast<size_t>(t.cols), false);
            s->solver->absorb_optimum(std::move(x), f, s->basis);
        });
    }

    // ---- the whole flow on the GPU: InitialBasis (Phase I) + Phase II, as Simplex::solve()
    int sbh_solve(void * h)
    {
        Session * s = static_cast<Session *>(h);
        return guarded(s, [&] { s->solver->solve(); });
    }

    int sbh_status(void * h) { return static_cast<int>(static_cast<Session *>(h)->solver->status()); }
    double sbh_objective(void * h) { return static_cast<Session *>(h)->solver->objective_value(); }
    char sbh_sense(void * h) { return static_cast<Session *>(h)->lp.sense(); }

    int sbh_write(void * h, const char * path)
    {
        Session * s = static_cast<Session *>(h);
        return guarded(s, [&] { s->solver->write(path); });
    }

    // Text tables "name\tvalue\n" with 17 significant digits; returns the size needed.
    static int64_t table_out(const std::string & text, char * out, int64_t cap)
    {
        if (out && cap > 0)
        {
            const size_t n = std::min<size_t>(text.size(), static_cast<size_t>(cap - 1));
            std::memcpy(out, text.data(), n);
            out[n] = '\0';
        }
        return static_cast<int64_t>(text.size()) + 1;
    }

    int64_t sbh_solution_table(void * h, char * out, int64_t cap)
    {
        std::ostringstream os;
        os.precision(17);
        for (const auto & [name, value] : static_cast<Session *>(h)->solver->solution()) os << name << '\t' << value << '\n';
        return table_out(os.str(), out, cap);
    }

    // "id\tname\n" for every live variable, ascending id; then "#rows\n" and one row label per line
    int64_t sbh_names_table(void * h, char * out, int64_t cap)
    {
        Session * s = static_cast<Session *>(h);
        std::map<simplex::Index, std::string> by_id;
        for (const auto & [name, id] : s->lp.ids_by_name()) by_id[id] = name;
        std::ostringstream os;
        for (const auto & [id, name] : by_id) os << id << '\t' << name << '\n';
        os << "#rows\n";
        for (const auto & [label, row] : s->lp.rows())
        {
            (void) row;
            os << label << '\n';
        }
        return table_out(os.str(), out, cap);
    }

    // "id\tshift\n"
    int64_t sbh_shifts_table(void * h, char * out, int64_t cap)
    {
        std::ostringstream os;
        os.precision(17);
        for (const auto & [id, v] : static_cast<Session *>(h)->lp.shifts()) os << id << '\t' << v << '\n';
        return table_out(os.str(), out, cap);
    }

    // Dense fast path (host/dense_solver.h): the two-phase flow on plain arrays. a: m x n row-major;
    // types: m chars '<' '>' '='; ub: n entries or NULL. Outputs: x_scaled / x (n each, may be NULL),
    // calls_out[k*8 + ...] as sbh_calls (up to 2 calls). Returns 0 ok, 1 "Exception: ...", 2 fatal;
    // the message goes to err (cap bytes).
    int sbh_dense_solve(int64_t m, int64_t n, const double * a, const char * types, const double * rhs, const double * c,
                        const double * ub, int device, int pricing, int engine, int dual_mode, int64_t chunk_iters,
                        int * status, double * objective, double * x_scaled, double * x, int64_t * calls_out,
                        int64_t * n_calls, double * seconds_host, char * err, int64_t cap)
    {
        Session scratch;  // only for the error text
        simplex::DenseSolution sol;
        const int rc = guarded(&scratch, [&] {
            simplex::DenseProblem p;
            p.m = m;
            p.n = n;
            p.a.assign(a, a + m * n);
            p.type.assign(types, types + m);
            p.rhs.assign(rhs, rhs + m);
            p.c.assign(c, c + n);
            if (ub) p.ub.assign(ub, ub + n);
            simplex::DeviceOptions o;
            o.device = device;
            o.pricing = static_cast<simplex::Pricing>(pricing);
            o.engine = engine;
            o.dual_mode = dual_mode;
            if (chunk_iters > 0) o.chunk_iters = chunk_iters;
            sol = simplex::solve_dense(p, o);
        });
        if (err && cap > 0)
        {
            const size_t k = std::min<size_t>(scratch.error.size(), static_cast<size_t>(cap - 1));
            std::memcpy(err, scratch.error.data(), k);
            err[k] = '\0';
        }
        *status = static_cast<int>(sol.status);
        *objective = sol.objective;
        *seconds_host = sol.seconds_host;
        if (x_scaled && !sol.x_scaled.empty()) std::memcpy(x_scaled, sol.x_scaled.data(), sol.x_scaled.size() * sizeof(double));
        if (x && !sol.x.empty()) std::memcpy(x, sol.x.data(), sol.x.size() * sizeof(double));
        *n_calls = static_cast<int64_t>(sol.calls.size());
        for (size_t k = 0; k < sol.calls.size() && k < 2; ++k)
        {
            int64_t * r = calls_out + k * 8;
            r[0] = sol.calls[k].M;
            r[1] = sol.calls[k].N;
            r[2] = sol.calls[k].term;
            r[3] = sol.calls[k].iterations;
            r[4] = sol.calls[k].pivots;
            r[5] = sol.calls[k].bound_flips;
            r[6] = static_cast<int64_t>(sol.calls[k].loop_seconds * 1e6);
            r[7] = sol.calls[k].reused_resident_tableau ? 1 : 0;
        }
        return rc;
    }

    // per-call statistics of the seam: out[k*8 + {0..7}] = M, N, term, iterations, pivots, flips,
    // loop microseconds, resident-tableau reuse; returns the number of calls made
    int64_t sbh_calls(void * h, int64_t * out, int64_t cap_calls)
    {
        const auto & calls = static_cast<Session *>(h)->solver->calls();
        for (size_t k = 0; k < calls.size() && static_cast<int64_t>(k) < cap_calls; ++k)
        {
            int64_t * r = out + k * 8;
            r[0] = calls[k].M;
            r[1] = calls[k].N;
            r[2] = calls[k].term;
            r[3] = calls[k].iterations;
            r[4] = calls[k].pivots;
            r[5] = calls[k].bound_flips;
            r[6] = static_cast<int64_t>(calls[k].loop_seconds * 1e6);
            r[7] = calls[k].reused_resident_tableau ? 1 : 0;
        }
        return static_cast<int64_t>(calls.size());
    }
}

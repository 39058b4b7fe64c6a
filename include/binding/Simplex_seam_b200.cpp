// The reference-side binding of simplex_b200: the body of
//
//     int simplex::Simplex::simplex(Basis & basis)         (reference src/Simplex.h:25, src/Simplex.cpp:255-353)
//
// as a maintainer of maciejdrwal/simplex would write it to run the loop on the GPU. This file is compiled INTO
// the reference: oracle/Makefile (target _ref/simplex_ref_gpu) takes a throw-away copy of the reference's
// Simplex.cpp, removes the body of Simplex::simplex from it, adds this translation unit and links
// libsimplex_b200.so; every other line of the reference (parser, presolve, InitialBasis, solve, solution_found,
// write, main) is the reference's own. The GPU tests run that binary on the reference's fixtures and compare its
// sol_test.xml with what the stock binary writes (tests/test_gpu_host_flow.py).
//
// It uses nothing but the reference's headers and include/simplex_b200.h (plain C).
#include "Simplex.h"

#include "Logger.h"
#include "simplex_b200.h"

#include <cfloat>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace simplex
{
    int Simplex::simplex(Basis & basis)
    {
        const auto M = m_lp.get_num_rows();
        const auto N = m_lp.get_num_vars();
        if (basis.get_basic_columns().size() != M)  // src/Simplex.cpp:260-263
        {
            throw "Basis size must be equal to M=" + std::to_string(M);
        }

        sb200_opts opts;
        sb200_default_opts(&opts);  // first-eligible pricing (the active rule, :249), eps 1e-7, 1e9 iterations
        if (const char * rule = std::getenv("SIMPLEX_B200_PRICING"))  // "dantzig": the rule kept commented at :250
            if (std::strcmp(rule, "dantzig") == 0) opts.pricing = SB200_PRICE_MOST_NEGATIVE;
        sb200_ctx * ctx = nullptr;
        if (sb200_create(&ctx, &opts) != SB200_OK) throw std::string(sb200_last_error(nullptr));
        struct Closer
        {
            sb200_ctx * c;
            ~Closer() { sb200_destroy(c); }
        } closer{ ctx };

        // var_ubnd as a dense array, only if a finite bound was parsed (:108); DBL_MAX = none (LinearProgram.cpp:225)
        std::vector<double> ub;
        if (m_lp.has_non_trivial_upper_bounds())
        {
            ub.assign(N, DBL_MAX);
            for (const auto & [id, u] : m_lp.var_ubnd) ub[static_cast<size_t>(id)] = u;
        }
        m_lp.ub_substitutions.assign(N, false);  // :271

        // Eigen::MatrixXd is column-major and contiguous: lda = M (LinearProgram.cpp:101)
        if (sb200_load(ctx, static_cast<int64_t>(M), static_cast<int64_t>(N), m_lp.matrix_A.data(), static_cast<int64_t>(M),
                       m_lp.vector_b.data(), m_lp.vector_c.data(), ub.empty() ? nullptr : ub.data()) != SB200_OK)
            throw std::string(sb200_last_error(ctx));

        std::vector<int64_t> basic(basis.get_basic_columns().begin(), basis.get_basic_columns().end());
        std::vector<int64_t> nonbasic(basis.get_non_basic_columns().begin(), basis.get_non_basic_columns().end());
        sb200_result res;
        if (sb200_run(ctx, basic.data(), static_cast<int64_t>(basic.size()), nonbasic.data(),
                      static_cast<int64_t>(nonbasic.size()), &res) != SB200_OK)
            throw std::string(sb200_last_error(ctx));

        // the final lists, in order, through Basis' public interface (Basis.cpp:13-46)
        basis.clear();
        for (auto id : basic) basis.insert(static_cast<Eigen::Index>(id));
        for (auto id : nonbasic) basis.insert_to_non_basic(static_cast<Eigen::Index>(id));

        if (res.term == SB200_TERM_ITER_LIMIT) throw "MAXIMUM NUMBER OF SIMPLEX ITERATIONS REACHED";  // :347-350
        if (res.term == SB200_TERM_UNBOUNDED)                                                           // :334-338
        {
            LOG(debug) << "LP IS UNBOUNDED.\n";
            return 0;
        }

        // optimal (:319-323): hand solution_found (:181-245) what the reference's loop hands it
        Eigen::VectorXd x(M), c_B(M);
        if (sb200_get_xB(ctx, x.data()) != SB200_OK) throw std::string(sb200_last_error(ctx));
        std::vector<uint8_t> flags(N, 0);
        if (!ub.empty() && sb200_get_flip_flags(ctx, flags.data()) != SB200_OK) throw std::string(sb200_last_error(ctx));
        // The reference's loop leaves matrix_A / vector_b / vector_c with the substitutions applied and their flags
        // set; solution_found reverses them. The host copies were never touched by the device loop: apply them now
        // (LinearProgram.cpp:154-169 toggles the flag itself), so that the reversal restores the originals.
        for (size_t j = 0; j < N; ++j)
            if (flags[j]) m_lp.upper_bound_substitution(static_cast<Eigen::Index>(j), m_lp.var_ubnd[static_cast<Eigen::Index>(j)]);
        for (size_t i = 0; i < M; ++i) c_B[i] = m_lp.vector_c[basic[i]];  // flipped columns carry the negated cost (:157)
        solution_found(x, c_B, basis);
        return 0;
    }
}  // namespace simplex

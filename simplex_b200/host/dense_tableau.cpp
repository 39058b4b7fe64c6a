#include "dense_tableau.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <random>

namespace simplex_b200
{
    void rescale_tableau(int64_t m, int64_t N, double * A, double * b, double * c)
    {
        std::vector<double> rmax(static_cast<size_t>(m), 0.0);
        for (int64_t j = 0; j < N; ++j)
        {
            const double * col = A + j * m;
            for (int64_t i = 0; i < m; ++i) rmax[i] = std::max(rmax[i], std::fabs(col[i]));
        }
        for (int64_t i = 0; i < m; ++i)
        {
            if (rmax[i] > 1.0) b[i] /= rmax[i];
        }
        for (int64_t j = 0; j < N; ++j)
        {
            double * col = A + j * m;
            double cmax = 0.0;
            for (int64_t i = 0; i < m; ++i)
            {
                if (rmax[i] > 1.0) col[i] /= rmax[i];
                cmax = std::max(cmax, std::fabs(col[i]));
            }
            for (int64_t i = 0; i < m; ++i) col[i] /= cmax;  // unguarded like the reference (:192)
            c[j] /= cmax;
        }
    }

    SynthLP make_synth(int64_t m, int64_t n, int64_t n_eq, uint64_t seed)
    {
        SynthLP s;
        s.m = m;
        s.n = n;
        s.n_eq = n_eq;
        std::mt19937_64 rng(seed);
        auto U = [&rng]() { return static_cast<double>(rng() >> 11) * 0x1.0p-53; };
        s.c.resize(static_cast<size_t>(n));
        std::vector<double> x0(static_cast<size_t>(n));
        for (auto & v : s.c) v = -(0.5 + U());
        for (auto & v : x0) v = U();
        s.a.resize(static_cast<size_t>(m) * n);
        s.rhs.resize(static_cast<size_t>(m));
        for (int64_t i = 0; i < m; ++i)
        {
            double * row = s.a.data() + static_cast<size_t>(i) * n;
            double sum = 0.0;
            for (int64_t j = 0; j < n; ++j)
            {
                row[j] = 0.05 + 0.95 * U();
                const double prod = row[j] * x0[j];
                sum = sum + prod;
            }
            s.rhs[i] = (i >= n_eq) ? (1.0 + 0.5 * U()) * sum : sum;
        }
        return s;
    }

    int64_t synth_tableau(int64_t m, int64_t n, int64_t n_eq, uint64_t seed, bool phase1, double * A, double * b,
                          double * c, int64_t * basic, int64_t * nonbasic, int64_t col0, int64_t ncols)
    {
        const int64_t n_slack = m - n_eq;
        const int64_t N = synth_num_cols(m, n, n_eq, phase1);
        if (ncols < 0) ncols = N - col0;
        const int64_t col1 = col0 + ncols;
        const SynthLP s = make_synth(m, n, n_eq, seed);

        // In G every row maximum is below 1 (a_ij < 1, slack/artificial entries are 1), so the
        // reference's row pass leaves rows alone; columns are scaled by their own maximum. The
        // general routine is still used on the full tableau when it is built in one piece.
        std::memset(A, 0, sizeof(double) * static_cast<size_t>(m) * ncols);
        std::vector<double> cfull(static_cast<size_t>(N), 0.0);
        for (int64_t j = 0; j < n; ++j) cfull[j] = s.c[j];
        for (int64_t i = 0; i < m; ++i) b[i] = s.rhs[i];
        for (int64_t j = std::max<int64_t>(col0, 0); j < std::min(col1, n); ++j)
        {
            double * col = A + (j - col0) * m;
            for (int64_t i = 0; i < m; ++i) col[i] = s.a[static_cast<size_t>(i) * n + j];
        }
        for (int64_t i = n_eq; i < m; ++i)
        {
            const int64_t j = n + (i - n_eq);
            if (j >= col0 && j < col1) A[(j - col0) * m + i] = 1.0;
        }
        if (phase1)
        {
            for (int64_t i = 0; i < n_eq; ++i)
            {
                const int64_t j = n + n_slack + i;
                if (j >= col0 && j < col1) A[(j - col0) * m + i] = 1.0;  // rhs >= 0: +1 (InitialBasis.cpp:44)
            }
        }
        if (col0 == 0 && ncols == N)
        {
            rescale_tableau(m, N, A, b, cfull.data());
        }
        else
        {
            // sharded build: column scaling is local to a column; row maxima of G are < 1 by construction
            for (int64_t j = col0; j < col1; ++j)
            {
                double * col = A + (j - col0) * m;
                double cmax = 0.0;
                for (int64_t i = 0; i < m; ++i) cmax = std::max(cmax, std::fabs(col[i]));
                for (int64_t i = 0; i < m; ++i) col[i] /= cmax;
            }
            // c needs every column maximum, including the columns of other shards
            for (int64_t j = 0; j < n; ++j)
            {
                double cmax = 0.0;
                for (int64_t i = 0; i < m; ++i) cmax = std::max(cmax, std::fabs(s.a[static_cast<size_t>(i) * n + j]));
                cfull[j] /= cmax;
            }
        }
        if (phase1)
        {
            std::fill(cfull.begin(), cfull.end(), 0.0);  // InitialBasis.cpp:73-78
            for (int64_t i = 0; i < n_eq; ++i) cfull[n + n_slack + i] = 1.0;
        }
        std::memcpy(c, cfull.data(), sizeof(double) * static_cast<size_t>(N));
        if (basic && nonbasic && (phase1 || n_eq == 0))
        {
            for (int64_t i = 0; i < m; ++i) basic[i] = (i < n_eq) ? (n + n_slack + i) : (n + (i - n_eq));
            for (int64_t j = 0; j < n; ++j) nonbasic[j] = j;
        }
        return N;
    }
}  // namespace simplex_b200

extern "C" {
int64_t sbh_synth_num_cols(int64_t m, int64_t n, int64_t n_eq, int phase1)
{
    return simplex_b200::synth_num_cols(m, n, n_eq, phase1 != 0);
}
int64_t sbh_synth_tableau(int64_t m, int64_t n, int64_t n_eq, uint64_t seed, int phase1, double * A, double * b,
                          double * c, int64_t * basic, int64_t * nonbasic, int64_t col0, int64_t ncols)
{
    return simplex_b200::synth_tableau(m, n, n_eq, seed, phase1 != 0, A, b, c, basic, nonbasic, col0, ncols);
}
void sbh_rescale_tableau(int64_t m, int64_t N, double * A, double * b, double * c)
{
    simplex_b200::rescale_tableau(m, N, A, b, c);
}
}

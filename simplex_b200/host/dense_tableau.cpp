#include "dense_tableau.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <random>
#include <thread>

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
                          double * c, int64_t * basic, int64_t * nonbasic, int64_t col0, int64_t ncols,
                          int64_t col_stride)
    {
        const int64_t n_slack = m - n_eq;
        const int64_t N = synth_num_cols(m, n, n_eq, phase1);
        if (col_stride < 1) col_stride = 1;
        if (ncols < 0) ncols = (N - col0 + col_stride - 1) / col_stride;
        // local column k is global column col0 + k*col_stride
        auto local_of = [&](int64_t j) -> int64_t {
            if (j < col0 || (j - col0) % col_stride != 0) return -1;
            const int64_t k = (j - col0) / col_stride;
            return k < ncols ? k : -1;
        };

        // Streaming build: the m x n structural block is never held as a whole (config 4 would need
        // 16 GiB per process). The random stream is consumed in the order of make_synth (row by row),
        // ROWS_PER_TILE rows are buffered and transposed tile-wise into the local column block.
        // In G every entry is below 1 (slack/artificial entries are exactly 1), so the reference's
        // row pass (LinearProgram.cpp:179, only rows whose maximum exceeds 1) leaves every row alone;
        // columns are scaled by their own maximum (LinearProgram.cpp:186-194).
        std::memset(A, 0, sizeof(double) * static_cast<size_t>(m) * ncols);
        std::vector<double> cfull(static_cast<size_t>(N), 0.0);
        std::vector<double> colmax(static_cast<size_t>(n), 0.0);
        {
            std::mt19937_64 rng(seed);
            auto U = [&rng]() { return static_cast<double>(rng() >> 11) * 0x1.0p-53; };
            std::vector<double> x0(static_cast<size_t>(n));
            for (int64_t j = 0; j < n; ++j) cfull[j] = -(0.5 + U());
            for (auto & v : x0) v = U();
            constexpr int64_t ROWS_PER_TILE = 64;
            // local structural columns: k = 0..nl-1 with col0 + k*col_stride < n
            const int64_t nl = std::min<int64_t>(ncols, col0 < n ? (n - col0 + col_stride - 1) / col_stride : 0);
            std::vector<double> tile(static_cast<size_t>(ROWS_PER_TILE) * std::max<int64_t>(nl, 1));
            std::vector<double> rowbuf(static_cast<size_t>(n));
            for (int64_t i0 = 0; i0 < m; i0 += ROWS_PER_TILE)
            {
                const int64_t ib = std::min(ROWS_PER_TILE, m - i0);
                for (int64_t r = 0; r < ib; ++r)
                {
                    const int64_t i = i0 + r;
                    double sum = 0.0;
                    for (int64_t j = 0; j < n; ++j)
                    {
                        const double a = 0.05 + 0.95 * U();
                        rowbuf[j] = a;
                        const double prod = a * x0[j];
                        sum = sum + prod;
                        if (a > colmax[j]) colmax[j] = a;
                    }
                    b[i] = (i >= n_eq) ? (1.0 + 0.5 * U()) * sum : sum;
                    double * trow = tile.data() + r * nl;
                    if (col_stride == 1)
                    {
                        if (nl > 0) std::memcpy(trow, rowbuf.data() + col0, sizeof(double) * nl);
                    }
                    else
                    {
                        for (int64_t k = 0; k < nl; ++k) trow[k] = rowbuf[col0 + k * col_stride];
                    }
                }
                for (int64_t k = 0; k < nl; ++k)
                {
                    double * col = A + k * m + i0;
                    for (int64_t r = 0; r < ib; ++r) col[r] = tile[r * nl + k];
                }
            }
        }
        for (int64_t i = n_eq; i < m; ++i)
        {
            const int64_t k = local_of(n + (i - n_eq));
            if (k >= 0) A[k * m + i] = 1.0;
        }
        if (phase1)
        {
            for (int64_t i = 0; i < n_eq; ++i)
            {
                const int64_t k = local_of(n + n_slack + i);
                if (k >= 0) A[k * m + i] = 1.0;  // rhs >= 0: +1 (InitialBasis.cpp:44)
            }
        }
        // column scaling: local structural columns by their maximum (slack / artificial columns have
        // maximum 1 and stay as they are); c needs every column's maximum, also those of other shards
        for (int64_t k = 0; k < ncols; ++k)
        {
            const int64_t j = col0 + k * col_stride;
            if (j >= n) break;
            double * col = A + k * m;
            const double cmax = colmax[j];
            for (int64_t i = 0; i < m; ++i) col[i] /= cmax;
        }
        for (int64_t j = 0; j < n; ++j) cfull[j] /= colmax[j];
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
                          double * c, int64_t * basic, int64_t * nonbasic, int64_t col0, int64_t ncols,
                          int64_t col_stride)
{
    return simplex_b200::synth_tableau(m, n, n_eq, seed, phase1 != 0, A, b, c, basic, nonbasic, col0, ncols,
                                       col_stride);
}
// batch of independent LPs G(m, n, 0, seed0 + k), k = 0..batch-1, in the layout of
// sb200_run_batched: A [batch][N][m], b [batch][m], c [batch][N], int32 crash lists. Threads split
// the batch (the generator of one LP is sequential by construction).
void sbh_synth_batch(int64_t batch, int64_t m, int64_t n, uint64_t seed0, double * A, double * b, double * c,
                     int32_t * basic, int32_t * nonbasic, int threads)
{
    const int64_t N = n + m;
    if (threads < 1) threads = static_cast<int>(std::max(1u, std::thread::hardware_concurrency()));
    threads = static_cast<int>(std::min<int64_t>(threads, batch));
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t)
    {
        pool.emplace_back([=]() {
            std::vector<int64_t> bas(static_cast<size_t>(m)), non(static_cast<size_t>(n));
            for (int64_t k = t; k < batch; k += threads)
            {
                simplex_b200::synth_tableau(m, n, 0, seed0 + static_cast<uint64_t>(k), false, A + k * N * m, b + k * m,
                                            c + k * N, bas.data(), non.data(), 0, -1, 1);
                for (int64_t i = 0; i < m; ++i) basic[k * m + i] = static_cast<int32_t>(bas[static_cast<size_t>(i)]);
                for (int64_t j = 0; j < n; ++j) nonbasic[k * n + j] = static_cast<int32_t>(non[static_cast<size_t>(j)]);
            }
        });
    }
    for (auto & th : pool) th.join();
}
// raw G(m, n, n_eq, seed): c[n], a[m*n] row-major, rhs[m] (rows 0..n_eq-1 '=', the others '<=')
void sbh_synth_raw(int64_t m, int64_t n, int64_t n_eq, uint64_t seed, double * c, double * a, double * rhs)
{
    const simplex_b200::SynthLP g = simplex_b200::make_synth(m, n, n_eq, seed);
    std::memcpy(c, g.c.data(), g.c.size() * sizeof(double));
    std::memcpy(a, g.a.data(), g.a.size() * sizeof(double));
    std::memcpy(rhs, g.rhs.data(), g.rhs.size() * sizeof(double));
}
void sbh_rescale_tableau(int64_t m, int64_t N, double * A, double * b, double * c)
{
    simplex_b200::rescale_tableau(m, N, A, b, c);
}
}

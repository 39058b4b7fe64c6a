/* TEST INFRASTRUCTURE — CPU restatement of the reference's per-iteration hot path.
 *
 * This file is the parity checker for the CUDA path. It is NOT product code and nothing in
 * simplex_b200/ may link, load or call it: only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py do.
 *
 * Parity is PINNED: tests/test_oracle_golden.py checks this restatement against
 * tests/golden/*.json, which were produced by the UNMODIFIED reference sources compiled
 * against oracle/eigen_shim (oracle/_ref/ref_driver_{stock,dantzig}; script
 * tests/golden/make_golden.py): pivot traces, iteration counts, status and objectives of all
 * 14 reference fixtures and of the synthetic table of SURVEY.md App. A.3.
 *
 * Every function cites the reference lines (under /root/reference/src) it follows.
 */
#include "simplex_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ dense LU, partial pivoting
 * Stands for Eigen::PartialPivLU<Ref<MatrixXd>> (Simplex.cpp:305): in place, P*A = L*U, the
 * pivot of step k is the FIRST row >= k with the largest |a_ik|. */
static void lu_factor(int64_t n, double * a, int64_t * piv)
{
    for (int64_t k = 0; k < n; ++k)
    {
        double * colk = a + k * n;
        int64_t p = k;
        double best = fabs(colk[k]);
        for (int64_t i = k + 1; i < n; ++i)
        {
            const double v = fabs(colk[i]);
            if (v > best)
            {
                best = v;
                p = i;
            }
        }
        piv[k] = p;
        if (p != k)
        {
            for (int64_t j = 0; j < n; ++j)
            {
                const double t = a[k + j * n];
                a[k + j * n] = a[p + j * n];
                a[p + j * n] = t;
            }
        }
        const double pv = colk[k];
        if (pv != 0.0)
        {
            const double inv = 1.0 / pv;
            for (int64_t i = k + 1; i < n; ++i) colk[i] *= inv;
        }
        for (int64_t j = k + 1; j < n; ++j)
        {
            double * colj = a + j * n;
            const double ukj = colj[k];
            if (ukj != 0.0)
            {
                for (int64_t i = k + 1; i < n; ++i) colj[i] -= colk[i] * ukj;
            }
        }
    }
}

/* lu.solve(rhs): Simplex.cpp:306, :331 */
static void lu_solve(int64_t n, const double * a, const int64_t * piv, const double * rhs, double * x)
{
    memcpy(x, rhs, (size_t)n * sizeof(double));
    for (int64_t k = 0; k < n; ++k)
    {
        const int64_t p = piv[k];
        if (p != k)
        {
            const double t = x[k];
            x[k] = x[p];
            x[p] = t;
        }
    }
    for (int64_t k = 0; k < n; ++k)
    {
        const double xk = x[k];
        if (xk != 0.0)
        {
            const double * col = a + k * n;
            for (int64_t i = k + 1; i < n; ++i) x[i] -= col[i] * xk;
        }
    }
    for (int64_t k = n - 1; k >= 0; --k)
    {
        const double * col = a + k * n;
        x[k] /= col[k];
        const double xk = x[k];
        if (xk != 0.0)
        {
            for (int64_t i = 0; i < k; ++i) x[i] -= col[i] * xk;
        }
    }
}

/* lu.transpose().solve(rhs): Simplex.cpp:307 */
static void lu_solve_transposed(int64_t n, const double * a, const int64_t * piv, const double * rhs, double * y)
{
    memcpy(y, rhs, (size_t)n * sizeof(double));
    for (int64_t k = 0; k < n; ++k)
    {
        const double * col = a + k * n;
        double acc = y[k];
        for (int64_t i = 0; i < k; ++i) acc -= col[i] * y[i];
        y[k] = acc / col[k];
    }
    for (int64_t k = n - 1; k >= 0; --k)
    {
        const double * col = a + k * n;
        double acc = y[k];
        for (int64_t i = k + 1; i < n; ++i) acc -= col[i] * y[i];
        y[k] = acc;
    }
    for (int64_t k = n - 1; k >= 0; --k)
    {
        const int64_t p = piv[k];
        if (p != k)
        {
            const double t = y[k];
            y[k] = y[p];
            y[p] = t;
        }
    }
}

/* ------------------------------------------------------------------ selectors */

/* select_entering_variable_Bland, Simplex.cpp:49-61: first POSITION with s < -eps */
static int64_t enter_first(int64_t nnb, const double * s, double eps)
{
    for (int64_t i = 0; i < nnb; ++i)
        if (s[i] < -eps) return i;
    return -1;
}

/* select_entering_variable_most_neg, Simplex.cpp:87-102: strict '<' keeps the first minimum */
static int64_t enter_most_neg(int64_t nnb, const double * s, double eps)
{
    int64_t min_i = -1;
    double min_value = DBL_MAX;
    for (int64_t i = 0; i < nnb; ++i)
    {
        if (s[i] < -eps && s[i] < min_value)
        {
            min_i = i;
            min_value = s[i];
        }
    }
    return min_i;
}

/* select_leaving_variable_Bland, Simplex.cpp:66-84 */
static int64_t leave_bland(int64_t m, const double * x, const double * d, double eps)
{
    int64_t min_i = -1;
    double min_lbd = DBL_MAX;
    for (int64_t i = 0; i < m; ++i)
    {
        if (d[i] > eps)
        {
            const double lbd = x[i] / d[i];
            if (lbd < min_lbd)
            {
                min_i = i;
                min_lbd = lbd;
            }
        }
    }
    return min_i;
}

/* LinearProgram::upper_bound_substitution, LinearProgram.cpp:154-169 */
void oracle_upper_bound_substitution(int64_t m, double * A, double * b, double * c, uint8_t * flip, int64_t col,
                                     double u)
{
    c[col] = -c[col];
    for (int64_t i = 0; i < m; ++i)
    {
        const double a = A[i + col * m];
        b[i] = b[i] - a * u;
        A[i + col * m] = -a;
    }
    flip[col] = (uint8_t)!flip[col];
}

/* select_leaving_variable_SUB, Simplex.cpp:116-179. Returns row, -1 (unbounded) or -2 (the
 * entering variable was substituted, no pivot). */
static int64_t leave_sub(int64_t m, double * A, double * b, double * c, const double * ub, uint8_t * flip,
                         const double * x, const double * d, const int64_t * basic, const int64_t * nonbasic,
                         int64_t entering_index, double eps)
{
    int need_substitution = 0;
    int64_t min_i = -1;
    double min_theta = DBL_MAX;
    for (int64_t i = 0; i < m; ++i)
    {
        if (d[i] > eps)
        {
            const double t = x[i] / d[i];
            if (t < min_theta)
            {
                min_i = i;
                min_theta = t;
            }
        }
    }
    for (int64_t i = 0; i < m; ++i)
    {
        if (d[i] < -eps)
        {
            const double u = ub[basic[i]];
            const double t = (u - x[i]) / (-d[i]);
            if (t < min_theta)
            {
                min_i = i;
                min_theta = t;
                need_substitution = 1;
            }
        }
    }
    const double ub_s = ub[nonbasic[entering_index]];
    if (ub_s < min_theta)
    {
        oracle_upper_bound_substitution(m, A, b, c, flip, nonbasic[entering_index], ub_s);
        return -2;
    }
    if (need_substitution)
    {
        const int64_t k = basic[min_i];
        oracle_upper_bound_substitution(m, A, b, c, flip, k, ub[k]);
    }
    return min_i;
}

static void record_pivot(int32_t * trace, int64_t cap, int64_t k, int64_t leaving_col, int64_t row, int64_t entering_col,
                         int64_t pos)
{
    if (trace && k < cap)
    {
        trace[4 * k + 0] = (int32_t)leaving_col;
        trace[4 * k + 1] = (int32_t)row;
        trace[4 * k + 2] = (int32_t)entering_col;
        trace[4 * k + 3] = (int32_t)pos;
    }
}

/* ------------------------------------------------------------------ Simplex::simplex, Simplex.cpp:255-353 */
int oracle_simplex_lu(int64_t m, int64_t N, double * A, double * b, double * c, const double * ub, int64_t * basic,
                      int64_t * nonbasic, int rule, double eps, int64_t iter_limit, double * x_out, double * y_out,
                      uint8_t * flip_out, int32_t * trace, int64_t trace_cap, oracle_result * out)
{
    const int64_t nnb = N - m;
    double * AB = (double *)malloc((size_t)(m * m) * sizeof(double));
    int64_t * piv = (int64_t *)malloc((size_t)m * sizeof(int64_t));
    double * cB = (double *)malloc((size_t)m * sizeof(double));
    double * x = (double *)malloc((size_t)m * sizeof(double));
    double * y = (double *)malloc((size_t)m * sizeof(double));
    double * d = (double *)malloc((size_t)m * sizeof(double));
    double * s = (double *)malloc((size_t)(nnb > 0 ? nnb : 1) * sizeof(double));
    uint8_t * flip = (uint8_t *)calloc((size_t)N, 1); /* :271 */
    if (!AB || !piv || !cB || !x || !y || !d || !s || !flip) return -1;

    int64_t iterations = 0, pivots = 0, flips = 0;
    int term = ORACLE_RUNNING;
    while (iterations < iter_limit) /* :275 */
    {
        iterations++;
        /* :285-297 gather A_B, c_B (A_N, c_N are used in place) */
        for (int64_t j = 0; j < m; ++j)
        {
            memcpy(AB + j * m, A + basic[j] * m, (size_t)m * sizeof(double));
            cB[j] = c[basic[j]];
        }
        lu_factor(m, AB, piv);                   /* :305 */
        lu_solve(m, AB, piv, b, x);              /* :306 */
        lu_solve_transposed(m, AB, piv, cB, y);  /* :307 */
        for (int64_t k = 0; k < nnb; ++k)        /* :313 */
        {
            const double * col = A + nonbasic[k] * m;
            double acc = 0.0;
            for (int64_t i = 0; i < m; ++i) acc += col[i] * y[i];
            s[k] = c[nonbasic[k]] - acc;
        }
        const int64_t e = (rule == ORACLE_RULE_MOST_NEG) ? enter_most_neg(nnb, s, eps) : enter_first(nnb, s, eps); /* :318 */
        if (e == -1)
        {
            term = ORACLE_OPTIMAL; /* :319-323 */
            break;
        }
        lu_solve(m, AB, piv, A + nonbasic[e] * m, d); /* :326, :331 */
        int64_t l;
        if (ub) /* :108 */
            l = leave_sub(m, A, b, c, ub, flip, x, d, basic, nonbasic, e, eps);
        else
            l = leave_bland(m, x, d, eps);
        if (l == -1)
        {
            term = ORACLE_UNBOUNDED; /* :334-338 */
            break;
        }
        if (l != -2)
        {
            /* Basis::update, Basis.cpp:48-53 */
            const int64_t leaving_col = basic[l], entering_col = nonbasic[e];
            basic[l] = entering_col;
            nonbasic[e] = leaving_col;
            record_pivot(trace, trace_cap, pivots, leaving_col, l, entering_col, e);
            pivots++;
        }
        else
        {
            flips++;
        }
    }
    if (term == ORACLE_RUNNING || iterations == iter_limit) term = ORACLE_ITER_LIMIT; /* :347-350 */

    double obj = 0.0;
    for (int64_t i = 0; i < m; ++i) obj += cB[i] * x[i]; /* :210-211 before shifts */
    if (x_out) memcpy(x_out, x, (size_t)m * sizeof(double));
    if (y_out) memcpy(y_out, y, (size_t)m * sizeof(double));
    if (flip_out) memcpy(flip_out, flip, (size_t)N);
    if (out)
    {
        out->term = term;
        out->pad = 0;
        out->iterations = iterations;
        out->pivots = pivots;
        out->flips = flips;
        out->objective = obj;
    }
    free(AB);
    free(piv);
    free(cB);
    free(x);
    free(y);
    free(d);
    free(s);
    free(flip);
    return 0;
}

/* ------------------------------------------------------------------ explicit-inverse variant
 * Identical control flow; lu.solve(v) is replaced by Binv*v and lu.transpose().solve(v) by
 * Binv^T*v, with Binv = A_B^-1 carried by product-form updates (and by the sign change of row p
 * when the basic variable of row p is substituted at its upper bound). */
int oracle_simplex_pfi(int64_t m, int64_t N, double * A, double * b, double * c, const double * ub, int64_t * basic,
                       int64_t * nonbasic, int rule, double eps, int64_t iter_limit, double * x_out, double * y_out,
                       uint8_t * flip_out, int32_t * trace, int64_t trace_cap, oracle_result * out)
{
    const int64_t nnb = N - m;
    double * Bi = (double *)malloc((size_t)(m * m) * sizeof(double)); /* row-major */
    double * cB = (double *)malloc((size_t)m * sizeof(double));
    double * x = (double *)malloc((size_t)m * sizeof(double));
    double * y = (double *)malloc((size_t)m * sizeof(double));
    double * d = (double *)malloc((size_t)m * sizeof(double));
    double * pr = (double *)malloc((size_t)m * sizeof(double));
    double * s = (double *)malloc((size_t)(nnb > 0 ? nnb : 1) * sizeof(double));
    uint8_t * flip = (uint8_t *)calloc((size_t)N, 1);
    if (!Bi || !cB || !x || !y || !d || !pr || !s || !flip) return -1;

    /* initial inverse through one LU of the start basis */
    {
        double * AB = (double *)malloc((size_t)(m * m) * sizeof(double));
        int64_t * piv = (int64_t *)malloc((size_t)m * sizeof(int64_t));
        double * e = (double *)calloc((size_t)m, sizeof(double));
        double * col = (double *)malloc((size_t)m * sizeof(double));
        if (!AB || !piv || !e || !col) return -1;
        for (int64_t j = 0; j < m; ++j) memcpy(AB + j * m, A + basic[j] * m, (size_t)m * sizeof(double));
        lu_factor(m, AB, piv);
        for (int64_t j = 0; j < m; ++j)
        {
            e[j] = 1.0;
            lu_solve(m, AB, piv, e, col);
            e[j] = 0.0;
            for (int64_t i = 0; i < m; ++i) Bi[i * m + j] = col[i];
        }
        free(AB);
        free(piv);
        free(e);
        free(col);
    }

    int64_t iterations = 0, pivots = 0, flips = 0;
    int term = ORACLE_RUNNING;
    while (iterations < iter_limit)
    {
        iterations++;
        for (int64_t j = 0; j < m; ++j) cB[j] = c[basic[j]];
        /* x = Binv b ; y = Binv^T c_B */
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < m; ++i)
        {
            const double * row = Bi + i * m;
            double acc = 0.0;
            for (int64_t j = 0; j < m; ++j) acc += row[j] * b[j];
            x[i] = acc;
        }
#pragma omp parallel for schedule(static)
        for (int64_t j0 = 0; j0 < m; j0 += 512)
        {
            const int64_t j1 = j0 + 512 < m ? j0 + 512 : m;
            for (int64_t j = j0; j < j1; ++j) y[j] = 0.0;
            for (int64_t i = 0; i < m; ++i)
            {
                const double cb = cB[i];
                const double * row = Bi + i * m;
                for (int64_t j = j0; j < j1; ++j) y[j] += cb * row[j];
            }
        }
#pragma omp parallel for schedule(static)
        for (int64_t k = 0; k < nnb; ++k)
        {
            const double * col = A + nonbasic[k] * m;
            double acc = 0.0;
            for (int64_t i = 0; i < m; ++i) acc += col[i] * y[i];
            s[k] = c[nonbasic[k]] - acc;
        }
        const int64_t e = (rule == ORACLE_RULE_MOST_NEG) ? enter_most_neg(nnb, s, eps) : enter_first(nnb, s, eps);
        if (e == -1)
        {
            term = ORACLE_OPTIMAL;
            break;
        }
        const double * aq = A + nonbasic[e] * m;
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < m; ++i)
        {
            const double * row = Bi + i * m;
            double acc = 0.0;
            for (int64_t j = 0; j < m; ++j) acc += row[j] * aq[j];
            d[i] = acc;
        }
        int64_t l;
        int64_t k_before = -1;
        uint8_t flag_before = 0;
        if (ub)
        {
            /* detect a substitution of a BASIC variable (row l of Binv changes sign) */
            l = leave_sub(m, A, b, c, ub, flip, x, d, basic, nonbasic, e, eps);
            (void)k_before;
            (void)flag_before;
        }
        else
        {
            l = leave_bland(m, x, d, eps);
        }
        if (l == -1)
        {
            term = ORACLE_UNBOUNDED;
            break;
        }
        if (l == -2)
        {
            flips++;
            continue; /* Binv unchanged; x is recomputed from the shifted b next iteration */
        }
        double dl = d[l];
        if (ub && dl < -eps)
        {
            /* leave_sub substituted basic[l] (Simplex.cpp:165-176): its column changed sign */
            for (int64_t j = 0; j < m; ++j) Bi[l * m + j] = -Bi[l * m + j];
            dl = -dl;
            d[l] = dl;
        }
        for (int64_t j = 0; j < m; ++j) pr[j] = Bi[l * m + j] / dl;
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < m; ++i)
        {
            double * row = Bi + i * m;
            if (i == l)
            {
                for (int64_t j = 0; j < m; ++j) row[j] = pr[j];
            }
            else
            {
                const double di = d[i];
                if (di != 0.0)
                    for (int64_t j = 0; j < m; ++j) row[j] -= di * pr[j];
            }
        }
        const int64_t leaving_col = basic[l], entering_col = nonbasic[e];
        basic[l] = entering_col;
        nonbasic[e] = leaving_col;
        record_pivot(trace, trace_cap, pivots, leaving_col, l, entering_col, e);
        pivots++;
    }
    if (term == ORACLE_RUNNING || iterations == iter_limit) term = ORACLE_ITER_LIMIT;

    double obj = 0.0;
    for (int64_t i = 0; i < m; ++i) obj += cB[i] * x[i];
    if (x_out) memcpy(x_out, x, (size_t)m * sizeof(double));
    if (y_out) memcpy(y_out, y, (size_t)m * sizeof(double));
    if (flip_out) memcpy(flip_out, flip, (size_t)N);
    if (out)
    {
        out->term = term;
        out->pad = 0;
        out->iterations = iterations;
        out->pivots = pivots;
        out->flips = flips;
        out->objective = obj;
    }
    free(Bi);
    free(cB);
    free(x);
    free(y);
    free(d);
    free(pr);
    free(s);
    free(flip);
    return 0;
}

/* ------------------------------------------------------------------ LinearProgram::rescale_matrix, LinearProgram.cpp:171-195 */
void oracle_rescale(int64_t m, int64_t N, double * A, double * b, double * c)
{
    for (int64_t i = 0; i < m; ++i) /* rows: only when max > 1 (:179) */
    {
        double mx = fabs(A[i]);
        for (int64_t j = 1; j < N; ++j)
        {
            const double v = fabs(A[i + j * m]);
            if (v > mx) mx = v;
        }
        if (mx > 1.0)
        {
            for (int64_t j = 0; j < N; ++j) A[i + j * m] /= mx;
            b[i] /= mx;
        }
    }
    for (int64_t j = 0; j < N; ++j) /* columns: always (:186-194), no zero guard (:192) */
    {
        double * col = A + j * m;
        double mx = fabs(col[0]);
        for (int64_t i = 1; i < m; ++i)
        {
            const double v = fabs(col[i]);
            if (v > mx) mx = v;
        }
        for (int64_t i = 0; i < m; ++i) col[i] /= mx;
        c[j] /= mx;
    }
}

/* ------------------------------------------------------------------ std::mt19937_64 restated */
typedef struct
{
    uint64_t mt[312];
    int idx;
} mt64;

static void mt64_seed(mt64 * g, uint64_t seed)
{
    g->mt[0] = seed;
    for (int i = 1; i < 312; ++i) g->mt[i] = 6364136223846793005ULL * (g->mt[i - 1] ^ (g->mt[i - 1] >> 62)) + (uint64_t)i;
    g->idx = 312;
}

static uint64_t mt64_next(mt64 * g)
{
    if (g->idx >= 312)
    {
        const uint64_t UM = 0xFFFFFFFF80000000ULL, LM = 0x7FFFFFFFULL, MA = 0xB5026F5AA96619E9ULL;
        for (int i = 0; i < 312; ++i)
        {
            const uint64_t x = (g->mt[i] & UM) | (g->mt[(i + 1) % 312] & LM);
            g->mt[i] = g->mt[(i + 156) % 312] ^ (x >> 1) ^ ((x & 1ULL) ? MA : 0ULL);
        }
        g->idx = 0;
    }
    uint64_t x = g->mt[g->idx++];
    x ^= (x >> 29) & 0x5555555555555555ULL;
    x ^= (x << 17) & 0x71D67FFFEDA60000ULL;
    x ^= (x << 37) & 0xFFF7EEE000000000ULL;
    x ^= (x >> 43);
    return x;
}

static double mt64_u(mt64 * g) { return (double)(mt64_next(g) >> 11) * (1.0 / 9007199254740992.0); }

/* G(m,n,n_eq,seed); same draw order as oracle/ref_driver.cpp::make_synth */
void oracle_synth(int64_t m, int64_t n, int64_t n_eq, uint64_t seed, double * c, double * a, double * rhs)
{
    mt64 g;
    mt64_seed(&g, seed);
    double * x0 = (double *)malloc((size_t)n * sizeof(double));
    for (int64_t j = 0; j < n; ++j) c[j] = -(0.5 + mt64_u(&g));
    for (int64_t j = 0; j < n; ++j) x0[j] = mt64_u(&g);
    for (int64_t i = 0; i < m; ++i)
    {
        double * row = a + i * n;
        double acc = 0.0;
        for (int64_t j = 0; j < n; ++j)
        {
            row[j] = 0.05 + 0.95 * mt64_u(&g);
            const double t = row[j] * x0[j];
            acc = acc + t;
        }
        if (i >= n_eq)
        {
            const double f = 1.0 + 0.5 * mt64_u(&g);
            acc = f * acc;
        }
        rhs[i] = acc;
    }
    free(x0);
}

int64_t oracle_synth_tableau(int64_t m, int64_t n, int64_t n_eq, uint64_t seed, int phase1, double * A, double * b,
                             double * c, int64_t * basic, int64_t * nonbasic)
{
    const int64_t n_slack = m - n_eq;
    const int64_t n_art = phase1 ? n_eq : 0;
    const int64_t N = n + n_slack + n_art;
    double * cs = (double *)malloc((size_t)n * sizeof(double));
    double * a = (double *)malloc((size_t)(m * n) * sizeof(double));
    oracle_synth(m, n, n_eq, seed, cs, a, b);
    memset(A, 0, (size_t)(m * N) * sizeof(double));
    memset(c, 0, (size_t)N * sizeof(double));
    for (int64_t i = 0; i < m; ++i)
        for (int64_t j = 0; j < n; ++j) A[i + j * m] = a[i * n + j];
    /* slack ids in row order (LinearProgram.cpp:114-131), artificial ids after them (InitialBasis.cpp:26-49) */
    for (int64_t i = n_eq; i < m; ++i) A[i + (n + (i - n_eq)) * m] = 1.0;
    for (int64_t i = 0; i < n_art; ++i) A[i + (n + n_slack + i) * m] = 1.0; /* rhs >= 0 -> +1 (:44) */
    for (int64_t j = 0; j < n; ++j) c[j] = cs[j];                           /* sense 'm' (:93) */
    oracle_rescale(m, N, A, b, c);
    if (phase1)
    {
        for (int64_t j = 0; j < N; ++j) c[j] = 0.0; /* InitialBasis.cpp:73-78 */
        for (int64_t i = 0; i < n_art; ++i) c[n + n_slack + i] = 1.0;
    }
    if (basic && nonbasic && (phase1 || n_eq == 0))
    {
        /* crash basis in row order (InitialBasis.cpp:36-51), the rest ascending (:9-18) */
        for (int64_t i = 0; i < m; ++i) basic[i] = (i < n_eq) ? (n + n_slack + i) : (n + (i - n_eq));
        int64_t k = 0;
        for (int64_t j = 0; j < n; ++j) nonbasic[k++] = j;
    }
    free(cs);
    free(a);
    return N;
}

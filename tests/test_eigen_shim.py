"""The Eigen stand-in the reference is compiled against here (oracle/eigen_shim/Eigen/Dense, test infrastructure):
its packed, register-blocked GEMM must give, element by element, the bits of the plain loop it replaced (every
element of C receives its products one at a time in k order), because the committed golden traces were produced
with that loop and regenerate bit-identically with the blocked kernel. CPU only; needs g++."""
import os
import shutil
import subprocess
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SRC = textwrap.dedent(r"""
    #include <Eigen/Dense>
    #include <cstdio>
    #include <cstdlib>
    #include <cstring>
    #include <vector>
    // the plain loop (256-row blocks, four columns at a time, C updated in memory for every k), built per ISA like
    // the stand-in's kernel so that both see the same contraction of a*b into the subtraction
    __attribute__((target_clones("avx512f", "avx2,fma", "default"))) void
    plain(long mr, long nr, long kb, const double * A, long lda, const double * B, long ldb, double * C, long ldc)
    {
        for (long i0 = 0; i0 < mr; i0 += 256)
        {
            const long ib = (mr - i0 < 256) ? mr - i0 : 256;
            for (long j = 0; j < nr; ++j)
            {
                double * c0 = C + i0 + j * ldc;
                for (long k = 0; k < kb; ++k)
                {
                    const double * a = A + i0 + k * lda;
                    const double b0 = B[k + j * ldb];
                    for (long i = 0; i < ib; ++i) c0[i] -= a[i] * b0;
                }
            }
        }
    }
    int main()
    {
        const long shapes[][3] = {{300, 77, 64}, {31, 5, 64}, {256, 6, 17}, {33, 13, 64}, {513, 100, 64}, {64, 64, 1}, {7, 3, 5},
                                  {4032, 50, 64}};
        srand(7);
        for (const auto & sh : shapes)
        {
            const long mr = sh[0], nr = sh[1], kb = sh[2], lda = mr + 3, ldb = kb + 1, ldc = mr + 5;
            std::vector<double> A(lda * kb), B(ldb * nr), C(ldc * nr), R;
            for (auto & x : A) x = rand() / (double)RAND_MAX - 0.5;
            for (auto & x : B) x = (rand() % 5 == 0) ? 0.0 : rand() / (double)RAND_MAX - 0.5;
            for (auto & x : C) x = rand() / (double)RAND_MAX - 0.5;
            R = C;
            plain(mr, nr, kb, A.data(), lda, B.data(), ldb, R.data(), ldc);
            Eigen::shim_detail::gemm_minus(mr, nr, kb, A.data(), lda, B.data(), ldb, C.data(), ldc);
            if (std::memcmp(C.data(), R.data(), sizeof(double) * C.size()) != 0)
            {
                std::printf("MISMATCH %ld %ld %ld\n", mr, nr, kb);
                return 1;
            }
        }
        std::printf("OK\n");
        return 0;
    }
""")


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_stand_in_gemm_is_the_plain_loop_bit_for_bit(tmp_path):
    src = tmp_path / "t.cpp"
    src.write_text(SRC)
    exe = tmp_path / "t"
    subprocess.run(["g++", "-std=c++17", "-O3", "-fno-fast-math", "-I", os.path.join(ROOT, "oracle", "eigen_shim"),
                    str(src), "-o", str(exe)], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip() == "OK", r.stdout + r.stderr

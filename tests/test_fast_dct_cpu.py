"""Host build of the in-register fast DCT-II (mp3_b200/csrc/fast_dct.h) against the definition."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SRC = r'''
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include "fast_dct.h"
template <int N> double check() {
    double worst = 0;
    for (int trial = 0; trial < 400; trial++) {
        float x[N]; double xd[N], ref[N], peak = 1e-30;
        for (int k = 0; k < N; k++) { x[k] = (float)((rand() / (double)RAND_MAX - 0.5) * 2); xd[k] = x[k]; }
        for (int n = 0; n < N; n++) {
            double s = 0;
            for (int k = 0; k < N; k++) s += xd[k] * cos(M_PI * n * (2 * k + 1) / (2.0 * N));
            ref[n] = s; if (fabs(s) > peak) peak = fabs(s);
        }
        L3Dct2<N>::run(x);
        for (int n = 0; n < N; n++) { double e = fabs(x[n] - ref[n]) / peak; if (e > worst) worst = e; }
    }
    return worst;
}
int main() {
    srand(7);
    double e2 = check<2>(), e8 = check<8>(), e32 = check<32>();
    printf("%g %g %g\n", e2, e8, e32);
    return (e2 < 1e-6 && e8 < 1e-6 && e32 < 2e-6) ? 0 : 1;
}
'''


SRC_IMDCT = r'''
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include "fast_imdct.h"
int main() {
    double worst = 0; srand(3);
    for (int t = 0; t < 2000; t++) {
        float X[18], Z[18]; double xd[18], peak = 1e-30, ref[18];
        for (int k = 0; k < 18; k++) { X[k] = (float)((rand() / (double)RAND_MAX - 0.5) * 2); xd[k] = X[k]; }
        for (int n = 0; n < 18; n++) {
            double s = 0;
            for (int k = 0; k < 18; k++) s += xd[k] * cos(M_PI / 72 * (2 * n + 1) * (2 * k + 1));
            ref[n] = s; if (fabs(s) > peak) peak = fabs(s);
        }
        l3_dct4_18(X, Z);
        for (int n = 0; n < 18; n++) { double e = fabs(Z[n] - ref[n]) / peak; if (e > worst) worst = e; }
    }
    printf("%g\\n", worst);
    return worst < 2e-6 ? 0 : 1;
}
'''


def test_fast_dct4_18_matches_definition(tmp_path):
    src = tmp_path / "i.cpp"
    src.write_text(SRC_IMDCT)
    exe = tmp_path / "i"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "mp3_b200", "csrc"), "-o", str(exe), str(src)])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout


def test_fast_dct_matches_definition(tmp_path):
    src = tmp_path / "t.cpp"
    src.write_text(SRC)
    exe = tmp_path / "t"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "mp3_b200", "csrc"), "-o", str(exe), str(src)])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout

// CPU check of csrc/m3s_fast_transforms.cuh: the functions the FP32 hybrid kernel runs per thread, compiled for the host, against the
// direct (reference) formulas in double.  Prints the worst absolute error of each transform over random inputs of unit scale.
//   g++ -O2 -std=c++17 -o fast_transforms_check fast_transforms_check.cpp && ./fast_transforms_check
#include <stdio.h>
#include <stdlib.h>
#include "../../mp3-steganography-lib_b200/csrc/m3s_fast_transforms.cuh"

static double rnd() { return 2.0 * rand() / RAND_MAX - 1.0; }

template <int I> struct PickAll {
    template <typename R> static void run(const R (&c)[18], R *x) { x[I] = imdct36_pick<I, R>(c); PickAll<I + 1>::run(c, x); }
};
template <> struct PickAll<36> { template <typename R> static void run(const R (&)[18], R *) {} };

int main()
{
    const double PI = 3.141592653589793;
    m3s_fast_const_build(h_fast);
    m3s_fast_const_build(h_fast_d);
    srand(7);
    double e_imdct = 0, e_mat = 0, e_mat_direct = 0;
    for (int it = 0; it < 20000; it++) {
        float X[18], c[18], x[36];
        for (int k = 0; k < 18; k++) X[k] = (float)rnd();
        dct4_18<float>(X, c);
        PickAll<0>::run(c, x);
        for (int i = 0; i < 36; i++) {   // Frame.py:119-133
            double ref = 0;
            for (int k = 0; k < 18; k++) ref += (double)X[k] * cos(PI / 72.0 * (2 * i + 1 + 18) * (2 * k + 1));
            double d = fabs(ref - x[i]);
            if (d > e_imdct) e_imdct = d;
        }
        float S[32], D[32];
        for (int j = 0; j < 32; j++) D[j] = S[j] = (float)rnd();
        dct2_lee<32, float>(D);
        for (int i = 0; i < 64; i++) {   // Frame.py:81-87
            double ref = 0;
            float direct = 0.f;
            for (int j = 0; j < 32; j++) {
                ref += (double)S[j] * cos((16.0 + i) * (2.0 * j + 1.0) * (PI / 64.0));
                direct = fmaf(S[j], (float)cos((16.0 + i) * (2.0 * j + 1.0) * (PI / 64.0)), direct);
            }
            float v;
            if (i < 16) v = D[16 + i];
            else if (i == 16) v = 0.f;
            else if (i <= 48) v = -D[48 - i];
            else v = -D[i - 48];
            double d = fabs(ref - v);
            if (d > e_mat) e_mat = d;
            d = fabs(ref - direct);
            if (d > e_mat_direct) e_mat_direct = d;
        }
    }
    double e_imdct_d = 0, e_mat_d = 0;   // the float64 instantiation (M3S_DEC_EXACT) against long double direct sums
    for (int it = 0; it < 5000; it++) {
        double X[18], c[18], x[36];
        for (int k = 0; k < 18; k++) X[k] = rnd();
        dct4_18<double>(X, c);
        PickAll<0>::run(c, x);
        for (int i = 0; i < 36; i++) {
            long double ref = 0;
            for (int k = 0; k < 18; k++) ref += (long double)X[k] * cosl(3.14159265358979323846264338327950288L / 72.0L * (2 * i + 1 + 18) * (2 * k + 1));
            double d = fabs((double)(ref - x[i]));
            if (d > e_imdct_d) e_imdct_d = d;
        }
        double S[32], D[32];
        for (int j = 0; j < 32; j++) D[j] = S[j] = rnd();
        dct2_lee<32, double>(D);
        for (int i = 0; i < 64; i++) {
            long double ref = 0;
            for (int j = 0; j < 32; j++) ref += (long double)S[j] * cosl((16.0L + i) * (2.0L * j + 1.0L) * (3.14159265358979323846264338327950288L / 64.0L));
            double v = i < 16 ? D[16 + i] : (i == 16 ? 0.0 : (i <= 48 ? -D[48 - i] : -D[i - 48]));
            double d = fabs((double)(ref - v));
            if (d > e_mat_d) e_mat_d = d;
        }
    }
    printf("float64: imdct36 max abs error %.3g, matrixing %.3g\n", e_imdct_d, e_mat_d);
    if (!(e_imdct_d < 1e-13 && e_mat_d < 1e-13)) return 1;
    printf("imdct36 via dct4_18: max abs error %.3g\n", e_imdct);
    printf("matrixing via dct2_lee<32>: max abs error %.3g (direct float32 form: %.3g)\n", e_mat, e_mat_direct);
    return (e_imdct < 2e-5 && e_mat < 2e-5) ? 0 : 1;   // unit-scale inputs: outputs reach +-10, i.e. ~1e-6 relative
}

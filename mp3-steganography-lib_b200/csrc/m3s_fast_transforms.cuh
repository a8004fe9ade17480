// m3s_fast_transforms.cuh -- fast forms of the two big transforms of the decoder's FP32 instantiation, computed entirely in one
// thread's registers with compile-time indexing (every twiddle is an immediate constant-bank operand on the device):
//
//   dct2_lee<32>   the 64 x 32 synthesis matrixing (Frame.py:81-87: V[i] = sum_j cos((16 + i)(2 j + 1) pi / 64) S[j]) is a 32-point
//                  DCT-II D[m] = sum_j S[j] cos(pi (2 j + 1) m / 64) read through sign / index symmetries
//                  (V[i] = D[16 + i] for i < 16, V[16] = 0, V[i] = -D[48 - i] for 17 <= i <= 48, V[i] = -D[i - 48] above);
//                  Lee's recursion: 80 multiplications + 209 additions instead of 512 multiply-adds
//   imdct36_fast   the 36-point IMDCT (Frame.py:119-133: x[i] = sum_k X[k] cos(pi / 72 (2 i + 19)(2 k + 1))) is an 18-point DCT-IV
//                  c[n] = sum_k X[k] cos(pi (2 n + 1)(2 k + 1) / 72) read through x[i] = c[i + 9] (i < 9), -c[26 - i] (9 <= i <= 26),
//                  -c[i - 27] (i >= 27); the DCT-IV comes from a DCT-II of X[k] 2 cos((2 k + 1) pi / 72) by the running
//                  difference c[n] = d[n] - c[n - 1], the DCT-II(18) from two 9-point DCT-IIs: ~170 operations instead of 324
//
// The float64 instantiation (M3S_DEC_EXACT) keeps the direct forms and their summation order: its goldens are sample-exact.
// Host + device code: tests/model/fast_transforms_check.cpp runs the same functions on the CPU against the direct formulas.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define M3S_HD __host__ __device__ __forceinline__
#else
#define M3S_HD inline
#endif

template <typename R>
struct M3sFastConstT {
    R lee[32];      // 1 / (2 cos((2 k + 1) pi / (2 N))): N = 32 at [0, 16), N = 16 at [16, 24), N = 8 at [24, 28), N = 4 at [28, 30), N = 2 at [30]
    R pre18[18];    // 2 cos((2 k + 1) pi / 72)
    R pre9[12];     // 2 cos((2 k + 1) pi / 36), k < 9
    R c9e[5][4];    // cos(pi (2 k + 1)(2 j) / 18), j < 5, k < 4
    R c9o[4][4];    // cos(pi (2 k + 1)(2 j + 1) / 18), j < 4, k < 4
};
typedef M3sFastConstT<float> M3sFastConst;
typedef M3sFastConstT<double> M3sFastConstD;

template <typename R>
inline void m3s_fast_const_build(M3sFastConstT<R> &c)
{
    const double PI = 3.141592653589793;
    int o = 0;
    for (int n = 32; n >= 2; n >>= 1)
        for (int k = 0; k < n / 2; k++) c.lee[o++] = (R)(1.0 / (2.0 * cos((2 * k + 1) * PI / (2.0 * n))));
    c.lee[31] = (R)0;
    for (int k = 0; k < 18; k++) c.pre18[k] = (R)(2.0 * cos((2 * k + 1) * PI / 72.0));
    for (int k = 0; k < 12; k++) c.pre9[k] = k < 9 ? (R)(2.0 * cos((2 * k + 1) * PI / 36.0)) : (R)0;
    for (int j = 0; j < 5; j++)
        for (int k = 0; k < 4; k++) c.c9e[j][k] = (R)cos(PI * (2 * k + 1) * (2 * j) / 18.0);
    for (int j = 0; j < 4; j++)
        for (int k = 0; k < 4; k++) c.c9o[j][k] = (R)cos(PI * (2 * k + 1) * (2 * j + 1) / 18.0);
}

#if defined(__CUDACC__)
__constant__ M3sFastConst c_fast;
__constant__ M3sFastConstD c_fast_d;
#endif
static M3sFastConst h_fast;
static M3sFastConstD h_fast_d;
// the constants of precision R: the device's constant-memory copy or the host's (the host check runs the same templates)
template <typename R> struct M3sFC;
template <> struct M3sFC<float> {
    static M3S_HD const M3sFastConst &get()
    {
#if defined(__CUDA_ARCH__)
        return c_fast;
#else
        return h_fast;
#endif
    }
};
template <> struct M3sFC<double> {
    static M3S_HD const M3sFastConstD &get()
    {
#if defined(__CUDA_ARCH__)
        return c_fast_d;
#else
        return h_fast_d;
#endif
    }
};
M3S_HD float m3s_fma(float a, float b, float c) { return fmaf(a, b, c); }
M3S_HD double m3s_fma(double a, double b, double c) { return fma(a, b, c); }

template <int N> struct M3sLeeOff;
template <> struct M3sLeeOff<32> { static constexpr int v = 0; };
template <> struct M3sLeeOff<16> { static constexpr int v = 16; };
template <> struct M3sLeeOff<8> { static constexpr int v = 24; };
template <> struct M3sLeeOff<4> { static constexpr int v = 28; };
template <> struct M3sLeeOff<2> { static constexpr int v = 30; };

// in place: x[k] -> X[m] = sum_k x[k] cos(pi (2 k + 1) m / (2 N))   (Lee 1984: even outputs from the folded sums, odd outputs from the
// folded differences scaled by 1 / (2 cos) through X[2 m + 1] = B[m] + B[m + 1])
template <int N, typename R>
M3S_HD void dct2_lee(R (&x)[N])
{
    if constexpr (N == 1) {
        return;
    } else {
        R a[N / 2], b[N / 2];
#pragma unroll
        for (int k = 0; k < N / 2; k++) {
            a[k] = x[k] + x[N - 1 - k];
            b[k] = (x[k] - x[N - 1 - k]) * M3sFC<R>::get().lee[M3sLeeOff<N>::v + k];
        }
        dct2_lee<N / 2, R>(a);
        dct2_lee<N / 2, R>(b);
#pragma unroll
        for (int m = 0; m < N / 2; m++) {
            x[2 * m] = a[m];
            x[2 * m + 1] = m + 1 < N / 2 ? b[m] + b[m + 1] : b[m];
        }
    }
}

// Z[m] = sum_{k<9} z[k] cos(pi (2 k + 1) m / 18): folded sums feed the even outputs (5 x 4 multiply-adds), folded differences the odd ones (4 x 4)
template <typename R>
M3S_HD void dct2_9(const R (&z)[9], R (&Z)[9])
{
    const M3sFastConstT<R> &FC = M3sFC<R>::get();
    const R p0 = z[0] + z[8], p1 = z[1] + z[7], p2 = z[2] + z[6], p3 = z[3] + z[5], p4 = z[4];
    const R q0 = z[0] - z[8], q1 = z[1] - z[7], q2 = z[2] - z[6], q3 = z[3] - z[5];
#pragma unroll
    for (int j = 0; j < 5; j++) {
        R acc = (j & 1) ? -p4 : p4;
        acc = m3s_fma(p0, FC.c9e[j][0], acc);
        acc = m3s_fma(p1, FC.c9e[j][1], acc);
        acc = m3s_fma(p2, FC.c9e[j][2], acc);
        acc = m3s_fma(p3, FC.c9e[j][3], acc);
        Z[2 * j] = acc;
    }
#pragma unroll
    for (int j = 0; j < 4; j++) {
        R acc = q0 * FC.c9o[j][0];
        acc = m3s_fma(q1, FC.c9o[j][1], acc);
        acc = m3s_fma(q2, FC.c9o[j][2], acc);
        acc = m3s_fma(q3, FC.c9o[j][3], acc);
        Z[2 * j + 1] = acc;
    }
}

// c[n] = sum_{k<18} X[k] cos(pi (2 n + 1)(2 k + 1) / 72)
template <typename R>
M3S_HD void dct4_18(const R (&X)[18], R (&c)[18])
{
    const M3sFastConstT<R> &FC = M3sFC<R>::get();
    R s[9], r[9];
#pragma unroll
    for (int k = 0; k < 9; k++) {
        const R ya = X[k] * FC.pre18[k], yb = X[17 - k] * FC.pre18[17 - k];
        s[k] = ya + yb;
        r[k] = (ya - yb) * FC.pre9[k];
    }
    R E[9], O[9];
    dct2_9<R>(s, E);   // d[2 m]
    dct2_9<R>(r, O);   // e[m] + e[m - 1], e = d[2 m + 1]
    R e = (R)0.5 * O[0];
    R cn = (R)0.5 * E[0];
    c[0] = cn;
    cn = e - cn;
    c[1] = cn;
#pragma unroll
    for (int m = 1; m < 9; m++) {
        cn = E[m] - cn;
        c[2 * m] = cn;
        e = O[m] - e;
        cn = e - cn;
        c[2 * m + 1] = cn;
    }
}

// the 36 IMDCT outputs from the 18 DCT-IV values: x[i] = imdct36_pick(c, i) (compile-time i)
template <int I, typename R>
M3S_HD R imdct36_pick(const R (&c)[18])
{
    if constexpr (I < 9) return c[I + 9];
    else if constexpr (I <= 26) return -c[26 - I];
    else return -c[I - 27];
}

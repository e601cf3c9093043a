// sin and cos of a float64 phase of up to ~5e7 rad, for the bit-faithful mixer (precise.cu).
//
// The reference mixes in complex64: every sample is multiplied by (float)cos(phi), (float)sin(phi) of a float64 phase
// that grows to w * chunk ~ 2.6e7 rad inside a chunk (processing.py:289-297).  CUDA's sincos() takes its Payne-Hanek
// path for arguments that large (~130 instructions per sample: the mixer was issue-bound at 83 %,
// profiles/r02_ncu_full_k_mix_exact_raw.csv).  For |x| < 2^25 * pi/2 a Cody-Waite reduction is exact enough and three
// times cheaper: k = rint(x * 2/pi) < 2^25, pi/2 = P1 + P1t with P1 the leading 33 bits -- x - k*P1 is a multiple of
// 2^-32 below 2 in magnitude, so ONE fma computes it exactly -- and k * P1t carries an absolute error below 2e-19.
// The reduced argument is kept as head + tail and evaluated with the fdlibm kernels (k_sin.c / k_cos.c, error < 1 ulp).
// tools/sincos_check.cu compares against sincos() over 2^28 random arguments: same float32 roundings, |diff| <= 2.3e-16.
#pragma once

namespace iq2a {

__device__ __forceinline__ void sincos_cw(double x, double* sn, double* cs) {
    if (!(fabs(x) < 5.0e7)) {          // outside the range the one-step reduction is exact for (also NaN / inf)
        sincos(x, sn, cs);
        return;
    }
    const double fn = rint(x * 6.36619772367581382433e-01);            // 2/pi
    const int n = (int)fn;
    const double r = fma(-fn, 1.57079632673412561417e+00, x);          // exact (see above)
    const double w = fn * 6.07710050650619224932e-11;                  // k * (pi/2 - P1)
    const double y = r - w;
    const double yt = (r - y) - w;                                     // tail of the reduced argument
    const double z = y * y;
    // __kernel_sin(y, yt, 1)
    const double v = z * y;
    const double rs = fma(z, fma(z, fma(z, fma(z, 1.58969099521155010221e-10, -2.50507602534068634195e-08),
                                        2.75573137070700676789e-06), -1.98412698298579493134e-04), 8.33333333332248946124e-03);
    const double ks = y - ((z * (0.5 * yt - v * rs) - yt) - v * -1.66666666666666324348e-01);
    // __kernel_cos(y, yt), the |x| < 0.3 form with the correction term kept for the whole range (error < 2 ulp)
    const double rc = z * fma(z, fma(z, fma(z, fma(z, fma(z, -1.13596475577881948265e-11, 2.08757232129817482790e-09),
                                               -2.75573143513906633035e-07), 2.48015872894767294178e-05),
                                     -1.38888888888741095749e-03), 4.16666666666666019037e-02);
    const double hz = 0.5 * z;
    const double a = 1.0 - hz;
    const double kc = a + (((1.0 - a) - hz) + (z * rc - y * yt));
    const double s = (n & 1) ? kc : ks, c = (n & 1) ? ks : kc;
    *sn = (n & 2) ? -s : s;
    *cs = ((n + 1) & 2) ? -c : c;
}

}  // namespace iq2a

/*
 * oracle/seq_f32.c -- TEST INFRASTRUCTURE ONLY (CPU oracle, never shipped, never
 * on the product path).
 *
 * Plain-C restatement of the two per-sample Python loops on the reference's
 * demodulation path, so that the oracle can run them at more than a few
 * thousand samples per second while keeping the reference's exact float32
 * rounding sequence (NumPy >= 2 weak-scalar promotion: np.float32 <op> Python
 * float stays float32; Python float <op> Python float is float64).
 *
 *   dc_block_f32  follows  src/iq_to_audio/decoders/common.py:16-30
 *   agc_f32       follows  src/iq_to_audio/decoders/ssb.py:67-80
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off: no FMA contraction,
 * the reference evaluates every product and sum as a separately rounded op).
 */
#include <math.h>
#include <stddef.h>

/* y[n] = x[n] - x[n-1] + r*y[n-1]; state carried as doubles (Python floats).
 *
 * Rounding sequence of the reference (common.py:23-27), per sample:
 *   t1 = sample - x_prev        -> float32 subtract
 *   t2 = r * y_prev             -> first sample of a call: x_prev, y_prev are
 *                                  Python floats, so r*y_prev is a float64
 *                                  product that is then rounded to float32 when
 *                                  added to the float32 t1 (common.py:20-21,24);
 *                                  later samples: y_prev is np.float32 and r is
 *                                  a weak Python scalar -> float32 multiply by
 *                                  (float)r
 *   y  = t1 + t2                -> float32 add
 * The carried state is float(x_prev), float(y_prev) (common.py:28-29), i.e.
 * exactly representable float32 values, so "sample - x_prev" is a float32
 * subtract on the first sample as well.
 */
void dc_block_f32(const float *x, float *y, size_t n, double r,
                  double *x_prev_io, double *y_prev_io)
{
    if (n == 0) return;
    const float rf = (float)r;
    /* first sample: float64 product, rounded once to float32 */
    float xp = (float)(*x_prev_io);
    float t1 = x[0] - xp;
    float t2 = (float)(r * (*y_prev_io));
    float yp = t1 + t2;
    y[0] = yp;
    xp = x[0];
    for (size_t i = 1; i < n; ++i) {
        float s = x[i];
        float a = s - xp;
        float b = rf * yp;
        float v = a + b;
        y[i] = v;
        xp = s;
        yp = v;
    }
    *x_prev_io = (double)xp;
    *y_prev_io = (double)yp;
}

/* AGC, ssb.py:67-80.  gain restarts at 1.0 on every call (ssb.py:72).
 *
 *   magnitude = abs(sample)                      float32
 *   if magnitude > 1e-6:                         float32 vs weak Python float:
 *                                                compared as float32 (1e-6f)
 *       desired = target / magnitude             weak/float32 -> float32 divide
 *       gain += decay * (desired - gain)         float32 throughout; on the very
 *                                                first update "gain" is the Python
 *                                                float 1.0 (exact in float32)
 *   out = sample * gain                          float32 (gain==1.0 exactly while
 *                                                it is still a Python float)
 */
void agc_f32(const float *x, float *y, size_t n, double target, double decay)
{
    const float tf = (float)target;
    const float df = (float)decay;
    const float floor_f = (float)1e-6;
    float gain = 1.0f;
    for (size_t i = 0; i < n; ++i) {
        float s = x[i];
        float mag = fabsf(s);
        if (mag > floor_f) {
            float desired = tf / mag;
            float diff = desired - gain;
            float step = df * diff;
            gain = gain + step;
        }
        y[i] = s * gain;
    }
}

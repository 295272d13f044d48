/*
 * orc_math.h — transcendental backend of the CPU oracle (TEST INFRASTRUCTURE).
 *
 * Default: the bit-reproducible functions of pirip_b200/csrc/exact_math.h, so the
 * CUDA path can be compared with the oracle float for float (same IEEE op
 * sequence on both machines).  With -DORC_USE_LIBM the oracle calls glibc's
 * cosf/sinf/atan2f/expf/logf/log10f exactly as upstream codec2 does; the CPU tests
 * tests/test_exact_math.py check the two builds agree (every function within 1 ulp
 * of glibc; hard bits / nin / f_est identical and floats within 1e-4 on noisy
 * captures), so the exact-math substitution is itself pinned.
 */
#ifndef ORC_MATH_H
#define ORC_MATH_H
#include <math.h>
#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif
#ifdef ORC_USE_LIBM
#define ORC_MATH_BACKEND 1
static inline float om_cosf(float x) { return cosf(x); }
static inline float om_sinf(float x) { return sinf(x); }
static inline float om_atan2f(float y, float x) { return atan2f(y, x); }
static inline double om_exp(double x) { return exp(x); }
static inline double om_log(double x) { return log(x); }
static inline float om_log10f(float x) { return log10f(x); }
#else
#define ORC_MATH_BACKEND 0
#include "../pirip_b200/csrc/exact_math.h"
static inline float om_cosf(float x) { return em_cosf(x); }
static inline float om_sinf(float x) { return em_sinf(x); }
static inline float om_atan2f(float y, float x) { return em_atan2f(y, x); }
static inline double om_exp(double x) { return em_exp(x); }
static inline double om_log(double x) { return em_log(x); }
static inline float om_log10f(float x) { return em_log10f(x); }
#endif
#endif

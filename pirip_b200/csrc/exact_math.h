/*
 * exact_math.h — bit-reproducible elementary functions for host (gcc/g++) and
 * device (nvcc, sm_100a) code.
 *
 * Why this exists: the parity contract of this repo is "the CUDA path produces
 * the same bytes AND the same floats as the CPU oracle".  IEEE-754 guarantees
 * identical results for + - * / sqrt and float<->double<->int conversions on
 * both machines, but NOT for libm/libdevice transcendentals (glibc cosf and
 * CUDA cosf differ in the last ulp).  Every transcendental the FSK receive
 * chain needs (codec2 fsk.c uses cosf/sinf/atan2f/log10f, mpdecode_core.c uses
 * exp/log) is therefore restated here with nothing but IEEE double arithmetic,
 * floor() and integer bit operations, evaluated in a fixed order.
 *
 * Build rules that make the results identical on CPU and GPU:
 *   gcc/g++ : -O2 -ffp-contract=off, never -ffast-math
 *   nvcc    : -fmad=false (no FMA contraction), default -prec-div/-prec-sqrt
 *
 * Accuracy: all functions evaluate in double to ~1e-16 relative and round once
 * to float, so they agree with a correctly-rounded libm in all but ~1e-8 of
 * arguments (tests/test_exact_math.py checks <= 1 ulp against numpy/glibc).
 */
#ifndef PIRIP_EXACT_MATH_H
#define PIRIP_EXACT_MATH_H

#include <stdint.h>
#include <math.h>
#include <string.h>

#if defined(__CUDACC__)
#define EM_FN __host__ __device__ __forceinline__
#else
#define EM_FN static inline
#endif

#define EM_PI 3.14159265358979323846
#define EM_TWO_PI 6.28318530717958647692

EM_FN uint64_t em_d2u(double d) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(d);
#else
    uint64_t u;
    memcpy(&u, &d, sizeof u);
    return u;
#endif
}

EM_FN double em_u2d(uint64_t u) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    double d;
    memcpy(&d, &u, sizeof d);
    return d;
#endif
}

/* 2^k for -1022 <= k <= 1023 */
EM_FN double em_pow2i(int k) { return em_u2d((uint64_t)(k + 1023) << 52); }

/* sin and cos of a double argument, |x| < ~1e5 */
EM_FN void em_sincos(double x, double *s, double *c) {
    /* pi/2 split so that k*HI is exact for |k| < 2^20 */
    const double PIO2_HI = 1.57079632673412561417e+00; /* 33 significant bits */
    const double PIO2_LO = 6.07710050650619224932e-11;
    const double TWO_OVER_PI = 6.36619772367581382433e-01;
    double kf = floor(x * TWO_OVER_PI + 0.5);
    double r = (x - kf * PIO2_HI) - kf * PIO2_LO;
    double z = r * r;
    /* Taylor series, |r| <= pi/4: truncation < 1e-19 */
    double ps = -2.81145725434552076320e-15;            /* -1/17! */
    ps = ps * z + 7.64716373181981647590e-13;           /*  1/15! */
    ps = ps * z - 1.60590438368216145994e-10;           /* -1/13! */
    ps = ps * z + 2.50521083854417187751e-08;           /*  1/11! */
    ps = ps * z - 2.75573192239858906526e-06;           /* -1/9!  */
    ps = ps * z + 1.98412698412698412698e-04;           /*  1/7!  */
    ps = ps * z - 8.33333333333333333333e-03;           /* -1/5!  */
    ps = ps * z + 1.66666666666666666667e-01;           /*  1/3!  */
    double sr = r - (r * z) * ps;
    double pc = 1.56192069685862264622e-16;             /*  1/18! */
    pc = pc * z - 4.77947733238738529744e-14;           /* -1/16! */
    pc = pc * z + 1.14707455977297247139e-11;           /*  1/14! */
    pc = pc * z - 2.08767569878680989792e-09;           /* -1/12! */
    pc = pc * z + 2.75573192239858906526e-07;           /*  1/10! */
    pc = pc * z - 2.48015873015873015873e-05;           /* -1/8!  */
    pc = pc * z + 1.38888888888888888889e-03;           /*  1/6!  */
    pc = pc * z - 4.16666666666666666667e-02;           /* -1/4!  */
    pc = pc * z + 0.5;                                  /*  1/2!  */
    double cr = 1.0 - z * pc;
    long long k = (long long)kf;
    switch ((int)(k & 3)) {
        case 0: *s = sr;  *c = cr;  break;
        case 1: *s = cr;  *c = -sr; break;
        case 2: *s = -sr; *c = -cr; break;
        default: *s = -cr; *c = sr; break;
    }
}

EM_FN float em_cosf(float x) {
    double s, c;
    em_sincos((double)x, &s, &c);
    return (float)c;
}

EM_FN float em_sinf(float x) {
    double s, c;
    em_sincos((double)x, &s, &c);
    return (float)s;
}

/* atan(t) for 0 <= t <= 1 */
EM_FN double em_atan01(double t) {
    const double TAN_PI_8 = 0.41421356237309504880;
    const double PI_4 = 0.78539816339744830962;
    double base = 0.0;
    if (t > TAN_PI_8) {
        t = (t - 1.0) / (t + 1.0);
        base = PI_4;
    }
    /* |t| <= tan(pi/8); odd Taylor series to t^47, truncation < 2e-20 */
    double z = t * t;
    double p = 0.0;
    for (int n = 23; n >= 1; --n) {
        double coef = 1.0 / (double)(2 * n + 1);
        p = (n & 1) ? (p * z - coef) : (p * z + coef);
    }
    /* p = -1/3 + z/5 - z^2/7 ... */
    return base + (t + (t * z) * p);
}

EM_FN double em_atan2(double y, double x) {
    double ax = fabs(x), ay = fabs(y);
    double r;
    if (ax == 0.0 && ay == 0.0) {
        r = 0.0;
    } else if (ay <= ax) {
        r = em_atan01(ay / ax);
    } else {
        r = (EM_PI / 2.0) - em_atan01(ax / ay);
    }
    if (x < 0.0) r = EM_PI - r;
    if (y < 0.0) r = -r;
    return r;
}

EM_FN float em_atan2f(float y, float x) { return (float)em_atan2((double)y, (double)x); }

/* exp(x) in double; result saturates to +inf / 0 outside [-708, 709] */
EM_FN double em_exp(double x) {
    const double LOG2E = 1.44269504088896340736;
    const double LN2_HI = 6.93147180369123816490e-01;
    const double LN2_LO = 1.90821492927058770002e-10;
    if (x != x) return x;
    if (x > 709.0) return em_u2d(0x7ff0000000000000ULL);
    if (x < -708.0) return 0.0;
    double kf = floor(x * LOG2E + 0.5);
    double r = (x - kf * LN2_HI) - kf * LN2_LO;
    /* Taylor to r^13, |r| <= 0.347: truncation < 5e-18 */
    double p = 1.60590438368216145994e-10;  /* 1/13! */
    p = p * r + 2.08767569878680989792e-09; /* 1/12! */
    p = p * r + 2.50521083854417187751e-08; /* 1/11! */
    p = p * r + 2.75573192239858906526e-07; /* 1/10! */
    p = p * r + 2.75573192239858906526e-06; /* 1/9!  */
    p = p * r + 2.48015873015873015873e-05; /* 1/8!  */
    p = p * r + 1.98412698412698412698e-04; /* 1/7!  */
    p = p * r + 1.38888888888888888889e-03; /* 1/6!  */
    p = p * r + 8.33333333333333333333e-03; /* 1/5!  */
    p = p * r + 4.16666666666666666667e-02; /* 1/4!  */
    p = p * r + 1.66666666666666666667e-01; /* 1/3!  */
    p = p * r + 0.5;
    p = p * r + 1.0;
    p = p * r + 1.0;
    return p * em_pow2i((int)kf);
}

/* natural log of a positive normal double */
EM_FN double em_log(double x) {
    const double LN2_HI = 6.93147180369123816490e-01;
    const double LN2_LO = 1.90821492927058770002e-10;
    const double SQRT2 = 1.41421356237309504880;
    if (x != x) return x;
    if (x < 0.0) return em_u2d(0x7ff8000000000000ULL);
    if (x == 0.0) return em_u2d(0xfff0000000000000ULL);
    uint64_t u = em_d2u(x);
    if ((u >> 52) == 0x7ffULL) return x; /* +inf */
    int e = (int)((u >> 52) & 0x7ffULL) - 1023;
    double m = em_u2d((u & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL);
    if (m > SQRT2) {
        m = m * 0.5;
        e += 1;
    }
    double s = (m - 1.0) / (m + 1.0);
    double z = s * s;
    /* 2*atanh(s) = 2s(1 + z/3 + z^2/5 + ...), |s| <= 0.1716 */
    double p = 0.0;
    for (int n = 12; n >= 1; --n) p = p * z + 1.0 / (double)(2 * n + 1);
    double lm = 2.0 * (s + (s * z) * p);
    double ef = (double)e;
    return ef * LN2_HI + (lm + ef * LN2_LO);
}

EM_FN float em_expf(float x) { return (float)em_exp((double)x); }
EM_FN float em_logf(float x) { return (float)em_log((double)x); }

EM_FN double em_log10(double x) {
    const double INV_LN10 = 4.34294481903251827651e-01;
    return em_log(x) * INV_LN10;
}
EM_FN float em_log10f(float x) { return (float)em_log10((double)x); }

#endif /* PIRIP_EXACT_MATH_H */

/*
 * detmath.h -- the oracle's libm: exp, expm1, log, log1p as plain IEEE-754 binary64 operation
 * sequences (TEST INFRASTRUCTURE, like everything under oracle/).
 *
 * Why the oracle does not call the platform libm: the reference keeps an observation iff
 * `prob_ref != prob_alt` EXACTLY (src/variants/evidence/realignment/mod.rs:303), so the last bit of
 * exp / ln_1p decides discrete outputs.  Rust's f64::exp / ln_1p are the platform's libm, whose
 * last bits differ between glibc versions and between their FMA / non-FMA builds -- the reference
 * is not bit-reproducible across machines there.  The oracle therefore fixes ONE libm: the
 * classic Sun/FreeBSD msun algorithms (e_exp.c, s_expm1.c, e_log.c, s_log1p.c; < 1 ulp), restated
 * here operation by operation.  Compiled with -ffp-contract=off every operation is a single
 * correctly rounded binary64 operation, so any other implementation of the same sequences (the
 * library's device code uses __dmul_rn / __dadd_rn / __ddiv_rn) produces the same bits.
 * tests/test_detmath.py checks the accuracy against the platform libm (<= 1 ulp) and the bit
 * identity with the library's copy.
 */
#ifndef VLRO_DETMATH_H
#define VLRO_DETMATH_H

#include <math.h>
#include <stdint.h>
#include <string.h>

static inline int32_t dm_hi(double x) { uint64_t u; memcpy(&u, &x, 8); return (int32_t)(u >> 32); }
static inline uint32_t dm_lo(double x) { uint64_t u; memcpy(&u, &x, 8); return (uint32_t)u; }
static inline double dm_with_hi(double x, int32_t hi) {
    uint64_t u; memcpy(&u, &x, 8);
    u = (u & 0xffffffffull) | ((uint64_t)(uint32_t)hi << 32);
    memcpy(&x, &u, 8); return x;
}

static const double DM_LN2_HI = 6.93147180369123816490e-01, DM_LN2_LO = 1.90821492927058770002e-10,
                    DM_INVLN2 = 1.44269504088896338700e+00;

static inline double vlro_exp(double x) {
    static const double P1 = 1.66666666666666019037e-01, P2 = -2.77777777770155933842e-03,
                        P3 = 6.61375632143793436117e-05, P4 = -1.65339022054652515390e-06,
                        P5 = 4.13813679705723846039e-08;
    double hi = 0.0, lo = 0.0, c, t, y;
    int32_t k = 0, hx = dm_hi(x);
    const int xsb = (hx >> 31) & 1;
    hx &= 0x7fffffff;
    if (hx >= 0x40862E42) { /* |x| >= 709.78 */
        if (hx >= 0x7ff00000) {
            if (((hx & 0xfffff) | dm_lo(x)) != 0) return x + x; /* NaN */
            return xsb ? 0.0 : x;                                  /* exp(+-inf) */
        }
        if (x > 7.09782712893383973096e+02) return INFINITY;
        if (x < -7.45133219101941108420e+02) return 0.0;
    }
    if (hx > 0x3fd62e42) { /* |x| > 0.5 ln2 */
        if (hx < 0x3FF0A2B2) { /* |x| < 1.5 ln2 */
            hi = xsb ? x + DM_LN2_HI : x - DM_LN2_HI;
            lo = xsb ? -DM_LN2_LO : DM_LN2_LO;
            k = 1 - xsb - xsb;
        } else {
            k = (int32_t)(DM_INVLN2 * x + (xsb ? -0.5 : 0.5));
            t = (double)k;
            hi = x - t * DM_LN2_HI;
            lo = t * DM_LN2_LO;
        }
        x = hi - lo;
    } else if (hx < 0x3e300000) { /* |x| < 2^-28 */
        return 1.0 + x;
    }
    t = x * x;
    c = x - t * (P1 + t * (P2 + t * (P3 + t * (P4 + t * P5))));
    if (k == 0) return 1.0 - ((x * c) / (c - 2.0) - x);
    y = 1.0 - ((lo - (x * c) / (2.0 - c)) - hi);
    if (k >= -1021) return dm_with_hi(y, dm_hi(y) + (k << 20));
    y = dm_with_hi(y, dm_hi(y) + ((k + 1000) << 20));
    return y * 9.33263618503218878990e-302; /* 2^-1000 */
}

static inline double vlro_expm1(double x) {
    static const double Q1 = -3.33333333333331316428e-02, Q2 = 1.58730158725481460165e-03,
                        Q3 = -7.93650757867487942473e-05, Q4 = 4.00821782732936239552e-06,
                        Q5 = -2.01099218183624371326e-07;
    double y, hi, lo, c = 0.0, t, e, hxs, hfx, r1;
    int32_t k, hx = dm_hi(x);
    const int xsb = (hx >> 31) & 1;
    hx &= 0x7fffffff;
    if (hx >= 0x4043687A) { /* |x| >= 56 ln2 */
        if (hx >= 0x40862E42) {
            if (hx >= 0x7ff00000) {
                if (((hx & 0xfffff) | dm_lo(x)) != 0) return x + x;
                return xsb ? -1.0 : x;
            }
            if (x > 7.09782712893383973096e+02) return INFINITY;
        }
        if (xsb) return -1.0;
    }
    if (hx > 0x3fd62e42) {
        if (hx < 0x3FF0A2B2) {
            if (!xsb) { hi = x - DM_LN2_HI; lo = DM_LN2_LO; k = 1; }
            else { hi = x + DM_LN2_HI; lo = -DM_LN2_LO; k = -1; }
        } else {
            k = (int32_t)(DM_INVLN2 * x + (xsb ? -0.5 : 0.5));
            t = (double)k;
            hi = x - t * DM_LN2_HI;
            lo = t * DM_LN2_LO;
        }
        x = hi - lo;
        c = (hi - x) - lo;
    } else if (hx < 0x3c900000) { /* |x| < 2^-54 */
        return x;
    } else {
        k = 0;
    }
    hfx = 0.5 * x;
    hxs = x * hfx;
    r1 = 1.0 + hxs * (Q1 + hxs * (Q2 + hxs * (Q3 + hxs * (Q4 + hxs * Q5))));
    t = 3.0 - r1 * hfx;
    e = hxs * ((r1 - t) / (6.0 - x * t));
    if (k == 0) return x - (x * e - hxs);
    e = (x * (e - c) - c);
    e -= hxs;
    if (k == -1) return 0.5 * (x - e) - 0.5;
    if (k == 1) {
        if (x < -0.25) return -2.0 * (e - (x + 0.5));
        return 1.0 + 2.0 * (x - e);
    }
    if (k <= -2 || k > 56) {
        y = 1.0 - (e - x);
        if (k == 1024) y = y * 2.0 * 8.98846567431157953865e+307;
        else y = dm_with_hi(y, dm_hi(y) + (k << 20));
        return y - 1.0;
    }
    t = 1.0;
    if (k < 20) {
        t = dm_with_hi(t, 0x3ff00000 - (0x200000 >> k)); /* 1 - 2^-k */
        y = t - (e - x);
        return dm_with_hi(y, dm_hi(y) + (k << 20));
    }
    t = dm_with_hi(t, (0x3ff - k) << 20); /* 2^-k */
    y = x - (e + t);
    y += 1.0;
    return dm_with_hi(y, dm_hi(y) + (k << 20));
}

static const double DM_L1 = 6.666666666666735130e-01, DM_L2 = 3.999999999940941908e-01,
                    DM_L3 = 2.857142874366239149e-01, DM_L4 = 2.222219843214978396e-01,
                    DM_L5 = 1.818357216161805012e-01, DM_L6 = 1.531383769920937332e-01,
                    DM_L7 = 1.479819860511658591e-01;

static inline double vlro_log(double x) {
    double hfsq, f, s, z, R, w, t1, t2, dk;
    int32_t k = 0, hx = dm_hi(x), i, j;
    const uint32_t lx = dm_lo(x);
    if (hx < 0x00100000) { /* x < 2^-1022 */
        if (((hx & 0x7fffffff) | lx) == 0) return -INFINITY;
        if (hx < 0) return NAN;
        k -= 54;
        x *= 1.80143985094819840000e+16; /* 2^54 */
        hx = dm_hi(x);
    }
    if (hx >= 0x7ff00000) return x + x;
    k += (hx >> 20) - 1023;
    hx &= 0x000fffff;
    i = (hx + 0x95f64) & 0x100000;
    x = dm_with_hi(x, hx | (i ^ 0x3ff00000)); /* normalise x or x/2 */
    k += (i >> 20);
    f = x - 1.0;
    if ((0x000fffff & (2 + hx)) < 3) { /* |f| < 2^-20 */
        if (f == 0.0) {
            if (k == 0) return 0.0;
            dk = (double)k;
            return dk * DM_LN2_HI + dk * DM_LN2_LO;
        }
        R = f * f * (0.5 - 0.33333333333333333 * f);
        if (k == 0) return f - R;
        dk = (double)k;
        return dk * DM_LN2_HI - ((R - dk * DM_LN2_LO) - f);
    }
    s = f / (2.0 + f);
    dk = (double)k;
    z = s * s;
    i = hx - 0x6147a;
    w = z * z;
    j = 0x6b851 - hx;
    t1 = w * (DM_L2 + w * (DM_L4 + w * DM_L6));
    t2 = z * (DM_L1 + w * (DM_L3 + w * (DM_L5 + w * DM_L7)));
    i |= j;
    R = t2 + t1;
    if (i > 0) {
        hfsq = 0.5 * f * f;
        if (k == 0) return f - (hfsq - s * (hfsq + R));
        return dk * DM_LN2_HI - ((hfsq - (s * (hfsq + R) + dk * DM_LN2_LO)) - f);
    }
    if (k == 0) return f - s * (f - R);
    return dk * DM_LN2_HI - ((s * (f - R) - dk * DM_LN2_LO) - f);
}

static inline double vlro_log1p(double x) {
    double hfsq, f = 0.0, c = 0.0, s, z, R, u, dk;
    int32_t k = 1, hx = dm_hi(x), hu = 0, ax;
    ax = hx & 0x7fffffff;
    if (hx < 0x3FDA827A) { /* x < 0.41422 */
        if (ax >= 0x3ff00000) { /* x <= -1 */
            if (x == -1.0) return -INFINITY;
            return NAN;
        }
        if (ax < 0x3e200000) { /* |x| < 2^-29 */
            if (ax < 0x3c900000) return x; /* |x| < 2^-54 */
            return x - x * x * 0.5;
        }
        if (hx > 0 || hx <= (int32_t)0xbfd2bec3) { /* -0.2929 < x < 0.41422 */
            k = 0; f = x; hu = 1;
        }
    }
    if (hx >= 0x7ff00000) return x + x;
    if (k != 0) {
        if (hx < 0x43400000) {
            u = 1.0 + x;
            hu = dm_hi(u);
            k = (hu >> 20) - 1023;
            c = (k > 0) ? 1.0 - (u - x) : x - (u - 1.0); /* correction term */
            c /= u;
        } else {
            u = x;
            hu = dm_hi(u);
            k = (hu >> 20) - 1023;
            c = 0.0;
        }
        hu &= 0x000fffff;
        if (hu < 0x6a09e) {
            u = dm_with_hi(u, hu | 0x3ff00000); /* normalise u */
        } else {
            k += 1;
            u = dm_with_hi(u, hu | 0x3fe00000); /* normalise u/2 */
            hu = (0x00100000 - hu) >> 2;
        }
        f = u - 1.0;
    }
    hfsq = 0.5 * f * f;
    dk = (double)k;
    if (hu == 0) { /* |f| < 2^-20 */
        if (f == 0.0) {
            if (k == 0) return 0.0;
            c += dk * DM_LN2_LO;
            return dk * DM_LN2_HI + c;
        }
        R = hfsq * (1.0 - 0.66666666666666666 * f);
        if (k == 0) return f - R;
        return dk * DM_LN2_HI - ((R - (dk * DM_LN2_LO + c)) - f);
    }
    s = f / (2.0 + f);
    z = s * s;
    R = z * (DM_L1 + z * (DM_L2 + z * (DM_L3 + z * (DM_L4 + z * (DM_L5 + z * (DM_L6 + z * DM_L7))))));
    if (k == 0) return f - (hfsq - s * (hfsq + R));
    return dk * DM_LN2_HI - ((hfsq - (s * (hfsq + R) + (dk * DM_LN2_LO + c))) - f);
}

#endif

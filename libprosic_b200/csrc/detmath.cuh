// detmath.cuh -- the library's reference-arithmetic libm: exp, expm1, log, log1p as fixed sequences
// of correctly rounded binary64 operations (the classic Sun/FreeBSD msun algorithms e_exp.c,
// s_expm1.c, e_log.c, s_log1p.c; < 1 ulp), identical on the host and on the device.
//
// Used wherever a result must be reproducible bit for bit: the reference keeps an observation iff
// prob_ref != prob_alt EXACTLY (src/variants/evidence/realignment/mod.rs:303), so near ties are
// re-evaluated in the reference's literal log-space operation order with this libm
// (pairhmm_literal, pairhmm_kernel.cuh), and the constants / emission LUT that enter that path are
// built with it on the host.  Device code uses __dmul_rn / __dadd_rn / __ddiv_rn (never contracted
// into FMAs); host code is compiled without FMA contraction.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define VLR_HD __host__ __device__ __forceinline__
#else
#define VLR_HD inline
#endif

namespace vlr {
namespace dm {

VLR_HD double mul(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
VLR_HD double add(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}
VLR_HD double sub(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dsub_rn(a, b);
#else
    return a - b;
#endif
}
VLR_HD double div(double a, double b) {
#ifdef __CUDA_ARCH__
    return __ddiv_rn(a, b);
#else
    return a / b;
#endif
}
VLR_HD int32_t hi(double x) {
#ifdef __CUDA_ARCH__
    return __double2hiint(x);
#else
    uint64_t u; memcpy(&u, &x, 8); return (int32_t)(u >> 32);
#endif
}
VLR_HD uint32_t lo(double x) {
#ifdef __CUDA_ARCH__
    return (uint32_t)__double2loint(x);
#else
    uint64_t u; memcpy(&u, &x, 8); return (uint32_t)u;
#endif
}
VLR_HD double with_hi(double x, int32_t h) {
#ifdef __CUDA_ARCH__
    return __hiloint2double(h, __double2loint(x));
#else
    uint64_t u; memcpy(&u, &x, 8);
    u = (u & 0xffffffffull) | ((uint64_t)(uint32_t)h << 32);
    memcpy(&x, &u, 8); return x;
#endif
}
VLR_HD double pos_inf() { return with_hi(0.0, 0x7ff00000); }
VLR_HD double neg_inf() { return with_hi(0.0, (int32_t)0xfff00000); }
VLR_HD double quiet_nan() { return with_hi(0.0, 0x7ff80000); }

#define VLR_DM_LN2_HI 6.93147180369123816490e-01
#define VLR_DM_LN2_LO 1.90821492927058770002e-10
#define VLR_DM_INVLN2 1.44269504088896338700e+00

VLR_HD double exp(double x) {
    const double P1 = 1.66666666666666019037e-01, P2 = -2.77777777770155933842e-03,
                 P3 = 6.61375632143793436117e-05, P4 = -1.65339022054652515390e-06,
                 P5 = 4.13813679705723846039e-08;
    double h = 0.0, l = 0.0, c, t, y;
    int32_t k = 0, hx = hi(x);
    const int xsb = (hx >> 31) & 1;
    hx &= 0x7fffffff;
    if (hx >= 0x40862E42) {
        if (hx >= 0x7ff00000) {
            if (((hx & 0xfffff) | lo(x)) != 0) return add(x, x);
            return xsb ? 0.0 : x;
        }
        if (x > 7.09782712893383973096e+02) return pos_inf();
        if (x < -7.45133219101941108420e+02) return 0.0;
    }
    if (hx > 0x3fd62e42) {
        if (hx < 0x3FF0A2B2) {
            h = xsb ? add(x, VLR_DM_LN2_HI) : sub(x, VLR_DM_LN2_HI);
            l = xsb ? -VLR_DM_LN2_LO : VLR_DM_LN2_LO;
            k = 1 - xsb - xsb;
        } else {
            k = (int32_t)add(mul(VLR_DM_INVLN2, x), xsb ? -0.5 : 0.5);
            t = (double)k;
            h = sub(x, mul(t, VLR_DM_LN2_HI));
            l = mul(t, VLR_DM_LN2_LO);
        }
        x = sub(h, l);
    } else if (hx < 0x3e300000) {
        return add(1.0, x);
    }
    t = mul(x, x);
    c = sub(x, mul(t, add(P1, mul(t, add(P2, mul(t, add(P3, mul(t, add(P4, mul(t, P5))))))))));
    if (k == 0) return sub(1.0, sub(div(mul(x, c), sub(c, 2.0)), x));
    y = sub(1.0, sub(sub(l, div(mul(x, c), sub(2.0, c))), h));
    if (k >= -1021) return with_hi(y, hi(y) + (k << 20));
    y = with_hi(y, hi(y) + ((k + 1000) << 20));
    return mul(y, 9.33263618503218878990e-302);
}

VLR_HD double expm1(double x) {
    const double Q1 = -3.33333333333331316428e-02, Q2 = 1.58730158725481460165e-03,
                 Q3 = -7.93650757867487942473e-05, Q4 = 4.00821782732936239552e-06,
                 Q5 = -2.01099218183624371326e-07;
    double y, h, l, c = 0.0, t, e, hxs, hfx, r1;
    int32_t k, hx = hi(x);
    const int xsb = (hx >> 31) & 1;
    hx &= 0x7fffffff;
    if (hx >= 0x4043687A) {
        if (hx >= 0x40862E42) {
            if (hx >= 0x7ff00000) {
                if (((hx & 0xfffff) | lo(x)) != 0) return add(x, x);
                return xsb ? -1.0 : x;
            }
            if (x > 7.09782712893383973096e+02) return pos_inf();
        }
        if (xsb) return -1.0;
    }
    if (hx > 0x3fd62e42) {
        if (hx < 0x3FF0A2B2) {
            if (!xsb) { h = sub(x, VLR_DM_LN2_HI); l = VLR_DM_LN2_LO; k = 1; }
            else { h = add(x, VLR_DM_LN2_HI); l = -VLR_DM_LN2_LO; k = -1; }
        } else {
            k = (int32_t)add(mul(VLR_DM_INVLN2, x), xsb ? -0.5 : 0.5);
            t = (double)k;
            h = sub(x, mul(t, VLR_DM_LN2_HI));
            l = mul(t, VLR_DM_LN2_LO);
        }
        x = sub(h, l);
        c = sub(sub(h, x), l);
    } else if (hx < 0x3c900000) {
        return x;
    } else {
        k = 0;
    }
    hfx = mul(0.5, x);
    hxs = mul(x, hfx);
    r1 = add(1.0, mul(hxs, add(Q1, mul(hxs, add(Q2, mul(hxs, add(Q3, mul(hxs, add(Q4, mul(hxs, Q5))))))))));
    t = sub(3.0, mul(r1, hfx));
    e = mul(hxs, div(sub(r1, t), sub(6.0, mul(x, t))));
    if (k == 0) return sub(x, sub(mul(x, e), hxs));
    e = sub(mul(x, sub(e, c)), c);
    e = sub(e, hxs);
    if (k == -1) return sub(mul(0.5, sub(x, e)), 0.5);
    if (k == 1) {
        if (x < -0.25) return mul(-2.0, sub(e, add(x, 0.5)));
        return add(1.0, mul(2.0, sub(x, e)));
    }
    if (k <= -2 || k > 56) {
        y = sub(1.0, sub(e, x));
        if (k == 1024) y = mul(mul(y, 2.0), 8.98846567431157953865e+307);
        else y = with_hi(y, hi(y) + (k << 20));
        return sub(y, 1.0);
    }
    t = 1.0;
    if (k < 20) {
        t = with_hi(t, 0x3ff00000 - (0x200000 >> k));
        y = sub(t, sub(e, x));
        return with_hi(y, hi(y) + (k << 20));
    }
    t = with_hi(t, (0x3ff - k) << 20);
    y = sub(x, add(e, t));
    y = add(y, 1.0);
    return with_hi(y, hi(y) + (k << 20));
}

#define VLR_DM_L1 6.666666666666735130e-01
#define VLR_DM_L2 3.999999999940941908e-01
#define VLR_DM_L3 2.857142874366239149e-01
#define VLR_DM_L4 2.222219843214978396e-01
#define VLR_DM_L5 1.818357216161805012e-01
#define VLR_DM_L6 1.531383769920937332e-01
#define VLR_DM_L7 1.479819860511658591e-01

VLR_HD double log(double x) {
    double hfsq, f, s, z, R, w, t1, t2, dk;
    int32_t k = 0, hx = hi(x), i, j;
    const uint32_t lx = lo(x);
    if (hx < 0x00100000) {
        if (((hx & 0x7fffffff) | lx) == 0) return neg_inf();
        if (hx < 0) return quiet_nan();
        k -= 54;
        x = mul(x, 1.80143985094819840000e+16);
        hx = hi(x);
    }
    if (hx >= 0x7ff00000) return add(x, x);
    k += (hx >> 20) - 1023;
    hx &= 0x000fffff;
    i = (hx + 0x95f64) & 0x100000;
    x = with_hi(x, hx | (i ^ 0x3ff00000));
    k += (i >> 20);
    f = sub(x, 1.0);
    if ((0x000fffff & (2 + hx)) < 3) {
        if (f == 0.0) {
            if (k == 0) return 0.0;
            dk = (double)k;
            return add(mul(dk, VLR_DM_LN2_HI), mul(dk, VLR_DM_LN2_LO));
        }
        R = mul(mul(f, f), sub(0.5, mul(0.33333333333333333, f)));
        if (k == 0) return sub(f, R);
        dk = (double)k;
        return sub(mul(dk, VLR_DM_LN2_HI), sub(sub(R, mul(dk, VLR_DM_LN2_LO)), f));
    }
    s = div(f, add(2.0, f));
    dk = (double)k;
    z = mul(s, s);
    i = hx - 0x6147a;
    w = mul(z, z);
    j = 0x6b851 - hx;
    t1 = mul(w, add(VLR_DM_L2, mul(w, add(VLR_DM_L4, mul(w, VLR_DM_L6)))));
    t2 = mul(z, add(VLR_DM_L1, mul(w, add(VLR_DM_L3, mul(w, add(VLR_DM_L5, mul(w, VLR_DM_L7)))))));
    i |= j;
    R = add(t2, t1);
    if (i > 0) {
        hfsq = mul(mul(0.5, f), f);
        if (k == 0) return sub(f, sub(hfsq, mul(s, add(hfsq, R))));
        return sub(mul(dk, VLR_DM_LN2_HI), sub(sub(hfsq, add(mul(s, add(hfsq, R)), mul(dk, VLR_DM_LN2_LO))), f));
    }
    if (k == 0) return sub(f, mul(s, sub(f, R)));
    return sub(mul(dk, VLR_DM_LN2_HI), sub(sub(mul(s, sub(f, R)), mul(dk, VLR_DM_LN2_LO)), f));
}

VLR_HD double log1p(double x) {
    double hfsq, f = 0.0, c = 0.0, s, z, R, u, dk;
    int32_t k = 1, hx = hi(x), hu = 0, ax;
    ax = hx & 0x7fffffff;
    if (hx < 0x3FDA827A) {
        if (ax >= 0x3ff00000) {
            if (x == -1.0) return neg_inf();
            return quiet_nan();
        }
        if (ax < 0x3e200000) {
            if (ax < 0x3c900000) return x;
            return sub(x, mul(mul(x, x), 0.5));
        }
        if (hx > 0 || hx <= (int32_t)0xbfd2bec3) { k = 0; f = x; hu = 1; }
    }
    if (hx >= 0x7ff00000) return add(x, x);
    if (k != 0) {
        if (hx < 0x43400000) {
            u = add(1.0, x);
            hu = hi(u);
            k = (hu >> 20) - 1023;
            c = (k > 0) ? sub(1.0, sub(u, x)) : sub(x, sub(u, 1.0));
            c = div(c, u);
        } else {
            u = x;
            hu = hi(u);
            k = (hu >> 20) - 1023;
            c = 0.0;
        }
        hu &= 0x000fffff;
        if (hu < 0x6a09e) {
            u = with_hi(u, hu | 0x3ff00000);
        } else {
            k += 1;
            u = with_hi(u, hu | 0x3fe00000);
            hu = (0x00100000 - hu) >> 2;
        }
        f = sub(u, 1.0);
    }
    hfsq = mul(mul(0.5, f), f);
    dk = (double)k;
    if (hu == 0) {
        if (f == 0.0) {
            if (k == 0) return 0.0;
            c = add(c, mul(dk, VLR_DM_LN2_LO));
            return add(mul(dk, VLR_DM_LN2_HI), c);
        }
        R = mul(hfsq, sub(1.0, mul(0.66666666666666666, f)));
        if (k == 0) return sub(f, R);
        return sub(mul(dk, VLR_DM_LN2_HI), sub(sub(R, add(mul(dk, VLR_DM_LN2_LO), c)), f));
    }
    s = div(f, add(2.0, f));
    z = mul(s, s);
    R = mul(z, add(VLR_DM_L1, mul(z, add(VLR_DM_L2, mul(z, add(VLR_DM_L3, mul(z, add(VLR_DM_L4,
            mul(z, add(VLR_DM_L5, mul(z, add(VLR_DM_L6, mul(z, VLR_DM_L7)))))))))))));
    if (k == 0) return sub(f, sub(hfsq, mul(s, add(hfsq, R))));
    return sub(mul(dk, VLR_DM_LN2_HI),
               sub(sub(hfsq, add(mul(s, add(hfsq, R)), add(mul(dk, VLR_DM_LN2_LO), c))), f));
}

// bio::stats::probs on top of it, literal operation order
VLR_HD double ln_add_exp(double a, double b) {
    double p0 = a, p1 = b;
    if (p1 > p0) { const double t = p0; p0 = p1; p1 = t; }
    if (p0 == neg_inf()) return p0;
    if (p0 == pos_inf()) return p0;
    if (p1 == neg_inf()) return p0;
    return add(p0, log1p(exp(sub(p1, p0))));
}
VLR_HD double ln_one_minus_exp(double p) {
    if (p < -0.693) return log1p(-exp(p));
    return log(-expm1(p));
}

}  // namespace dm
}  // namespace vlr

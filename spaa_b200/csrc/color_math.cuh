// Per-pixel sRGB -> Lab and the reference's CIEDE2000 *variant*, forward and hand-derived reverse mode.
//
// Semantics follow /root/reference/src/python/perc_al/differential_color_functions.py:12-180
// (see SURVEY.md App. A3): sRGB threshold 0.0405, 4-digit matrix, f(0)=0, neutral handling, T with 39 deg,
// non-positive squares -> 0.  The reference blends both branches of each piecewise function with 0/1 float
// masks; these functions SELECT the active branch, which is value- and gradient-identical whenever both
// branches are finite (the reference yields NaN when the unselected power branch is NaN, e.g. XYZ < 0).
//
// Templated on the real type so the same code is compiled for the device (float) and, in
// tests/hostsim, for the host in float and double to validate the derivatives.
#pragma once
#include "common.cuh"
#include <cmath>

namespace spaa {
namespace color {

template <typename R> struct K {
    static constexpr R deg = R(180.0 / 3.14159265358979323846);
    static constexpr R rad = R(3.14159265358979323846 / 180.0);
};

template <typename R> SPAA_HD R rpow(R x, R p) { return pow(x, p); }
#if defined(__CUDACC__)
template <> SPAA_HD float rpow<float>(float x, float p) { return powf(x, p); }
#endif
// Integer powers by multiplication (at least as accurate as pow(); the reference's `x ** 7.`, :135-143,163) and the cube
// root by cbrt (1 ulp): a precise powf costs ~150 instructions, and the fused loss kernel would otherwise issue 17 of them
// per pixel and be instruction-bound at 3% of its HBM roofline.
template <typename R> SPAA_HD void pow67(R x, R& x6, R& x7) { const R x2 = x * x, x3 = x2 * x; x6 = x3 * x3; x7 = x6 * x; }
// Division used in the REVERSE pass only: MUFU.RCP-based (2 ulp) on the device instead of the ~12-instruction IEEE
// sequence with its slow-path branch -- the fused loss kernel has ~35 of them per pixel.  Forward values keep IEEE
// division (they are compared with the reference at 1e-5); gradients carry a relative tolerance of 1e-4.
template <typename R> SPAA_HD R fdiv(R a, R b) {
#if defined(__CUDA_ARCH__)
    if constexpr (sizeof(R) == 4) return (R)__fdividef((float)a, (float)b);
    else return a / b;
#else
    return a / b;
#endif
}
template <typename R> SPAA_HD R rcbrt(R x) { return cbrt(x); }               // float argument -> cbrtf on host and device
template <typename R> SPAA_HD void rsincos(R x, R& sn, R& cs) {
#if defined(__CUDA_ARCH__)
    if constexpr (sizeof(R) == 4) sincosf((float)x, (float*)&sn, (float*)&cs);       // one range reduction for both
    else sincos((double)x, (double*)&sn, (double*)&cs);
#else
    sn = sin(x); cs = cos(x);
#endif
}

// Math policy.  F = false: IEEE division / square root and the accurate powf / cbrtf / sincosf / expf -- the 1e-5 parity arithmetic.
// F = true (device, float only; chosen by the 16-bit tensor-core modes, whose PCNet output already carries ~3e-4 of rounding): MUFU-based
// approximations (relative error ~1e-6): ~1 200 -> ~500 instructions per pixel for the fused loss, which is bound by instruction issue.
template <typename R, bool F> struct M {
    static SPAA_HD R div(R a, R b) { return a / b; }
    static SPAA_HD R sqrt_(R x) { return sqrt(x); }
    static SPAA_HD R pow24(R y) { return rpow(y, R(2.4)); }
    static SPAA_HD R cbrt_(R x) { return rcbrt(x); }
    static SPAA_HD R exp_(R x) { return exp(x); }
    static SPAA_HD void sincos_(R x, R& sn, R& cs) { rsincos(x, sn, cs); }
    static SPAA_HD R cos_(R x) { return cos(x); }
    static SPAA_HD R sin_(R x) { return sin(x); }
};
#if defined(__CUDACC__)
template <> struct M<float, true> {
    static SPAA_HD float div(float a, float b) {
#if defined(__CUDA_ARCH__)
        return __fdividef(a, b);
#else
        return a / b;
#endif
    }
    static SPAA_HD float sqrt_(float x) {
#if defined(__CUDA_ARCH__)
        float r;
        asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
        return r;
#else
        return sqrtf(x);
#endif
    }
    static SPAA_HD float pow24(float y) {
#if defined(__CUDA_ARCH__)
        return __powf(y, 2.4f);
#else
        return powf(y, 2.4f);
#endif
    }
    static SPAA_HD float cbrt_(float x) {          // x > 0.008856 here
#if defined(__CUDA_ARCH__)
        return exp2f(__log2f(x) * (1.0f / 3.0f));
#else
        return cbrtf(x);
#endif
    }
    static SPAA_HD float exp_(float x) {
#if defined(__CUDA_ARCH__)
        return __expf(x);
#else
        return expf(x);
#endif
    }
    static SPAA_HD void sincos_(float x, float& sn, float& cs) {
#if defined(__CUDA_ARCH__)
        __sincosf(x, &sn, &cs);
#else
        sn = sinf(x); cs = cosf(x);
#endif
    }
    static SPAA_HD float cos_(float x) {
#if defined(__CUDA_ARCH__)
        return __cosf(x);
#else
        return cosf(x);
#endif
    }
    static SPAA_HD float sin_(float x) {
#if defined(__CUDA_ARCH__)
        return __sinf(x);
#else
        return sinf(x);
#endif
    }
};
#endif

// ---- sRGB channel -> 100 * linear, value and derivative  (differential_color_functions.py:16-20) -------
// y^1.4 of the derivative is y^2.4 / y: one powf serves both.
template <typename R, bool F = false> SPAA_HD void srgb_lin100_vg(R c, R& v, R& g) {
    if (c > R(0.0405)) {
        const R y = F ? (c + R(0.055)) * R(1.0 / 1.055) : (c + R(0.055)) / R(1.055);
        const R p = M<R, F>::pow24(y);
        v = R(100) * p;
        g = R(100.0 * 2.4 / 1.055) * fdiv(p, y);
    } else {
        v = R(100) * (c / R(12.92));
        g = R(100) / R(12.92);
    }
}
template <typename R> SPAA_HD R srgb_lin100(R c) {
    return c > R(0.0405) ? R(100) * rpow((c + R(0.055)) / R(1.055), R(2.4)) : R(100) * (c / R(12.92));
}

// ---- Lab f()  (:27-36), value and derivative: exact zero -> 0 with zero slope -------------------------
template <typename R, bool F = false> SPAA_HD void lab_f_vg(R t, R& f, R& g) {
    if (t == R(0)) { f = R(0); g = R(0); return; }
    if (t > R(0.008856)) { f = M<R, F>::cbrt_(t); g = fdiv(R(1.0 / 3.0), f * f); }      // d t^(1/3) = t^(-2/3) / 3
    else { f = R(7.787) * t + R(16.0 / 116.0); g = R(7.787); }
}
template <typename R> SPAA_HD R lab_f(R t) {
    if (t == R(0)) return R(0);
    return t > R(0.008856) ? rcbrt(t) : R(7.787) * t + R(16.0 / 116.0);
}

template <typename R> struct White {
    static constexpr R xn = R(95.0489), yn = R(100.0), zn = R(108.8840);
};

// rgb2lab_diff (:39-64) for one pixel
template <typename R> SPAA_HD void rgb_to_lab(R r, R g, R b, R& L, R& A, R& B) {
    const R lr = srgb_lin100(r), lg = srgb_lin100(g), lb = srgb_lin100(b);
    const R X = R(0.4124) * lr + R(0.3576) * lg + R(0.1805) * lb;
    const R Y = R(0.2126) * lr + R(0.7152) * lg + R(0.0722) * lb;
    const R Z = R(0.0193) * lr + R(0.1192) * lg + R(0.9504) * lb;
    const R fx = lab_f(X / White<R>::xn), fy = lab_f(Y / White<R>::yn), fz = lab_f(Z / White<R>::zn);
    L = R(116) * fy - R(16);
    A = R(500) * (fx - fy);
    B = R(200) * (fy - fz);
}

// Forward that keeps the six local derivatives the reverse pass needs (no transcendental is evaluated twice).
template <typename R> struct LabJac { R gr, gg, gb, jx, jy, jz; };     // d lin/d c per channel; f'(t)/white per axis
template <typename R, bool F = false> SPAA_HD void rgb_to_lab_jac(R r, R g, R b, R& L, R& A, R& B, LabJac<R>& J) {
    R lr, lg, lb;
    srgb_lin100_vg<R, F>(r, lr, J.gr); srgb_lin100_vg<R, F>(g, lg, J.gg); srgb_lin100_vg<R, F>(b, lb, J.gb);
    const R X = R(0.4124) * lr + R(0.3576) * lg + R(0.1805) * lb;
    const R Y = R(0.2126) * lr + R(0.7152) * lg + R(0.0722) * lb;
    const R Z = R(0.0193) * lr + R(0.1192) * lg + R(0.9504) * lb;
    R fx, fy, fz;
    if (F) {
        lab_f_vg<R, F>(X * R(1.0 / 95.0489), fx, J.jx); lab_f_vg<R, F>(Y * R(1.0 / 100.0), fy, J.jy); lab_f_vg<R, F>(Z * R(1.0 / 108.8840), fz, J.jz);
    } else {
        lab_f_vg(X / White<R>::xn, fx, J.jx); lab_f_vg(Y / White<R>::yn, fy, J.jy); lab_f_vg(Z / White<R>::zn, fz, J.jz);
    }
    J.jx = J.jx * R(1.0 / 95.0489); J.jy = J.jy * R(1.0 / 100.0); J.jz = J.jz * R(1.0 / 108.8840);
    L = R(116) * fy - R(16);
    A = R(500) * (fx - fy);
    B = R(200) * (fy - fz);
}
template <typename R> SPAA_HD void lab_jac_bwd(const LabJac<R>& J, R dL, R dA, R dB, R& dr, R& dg, R& db) {
    const R dX = (R(500) * dA) * J.jx;
    const R dY = (R(116) * dL - R(500) * dA + R(200) * dB) * J.jy;
    const R dZ = (-R(200) * dB) * J.jz;
    dr = (R(0.4124) * dX + R(0.2126) * dY + R(0.0193) * dZ) * J.gr;
    dg = (R(0.3576) * dX + R(0.7152) * dY + R(0.1192) * dZ) * J.gg;
    db = (R(0.1805) * dX + R(0.0722) * dY + R(0.9504) * dZ) * J.gb;
}

// reverse mode of rgb_to_lab: (dL,dA,dB) -> (dr,dg,db)
template <typename R> SPAA_HD void rgb_to_lab_bwd(R r, R g, R b, R dL, R dA, R dB, R& dr, R& dg, R& db) {
    R L, A, B;
    LabJac<R> J;
    rgb_to_lab_jac(r, g, b, L, A, B, J);
    lab_jac_bwd(J, dL, dA, dB, dr, dg, db);
}

// hue in degrees, [0,360)  (:73-81).  After the neutral nudge the arguments are never both zero.
template <typename R> SPAA_HD R hue_deg(R y, R x) {
    R h = K<R>::deg * atan2(y, x);
    return h < R(0) ? h + R(360) : h;
}

// ---- ciede2000_diff (:109-180), forward; optionally reverse mode -------------------------------------
// g1[3], g2[3] receive d(dE)/d(L1,A1,B1) and d(dE)/d(L2,A2,B2) (multiply by the cotangent outside).
template <typename R, bool WithGrad, bool F = false>
SPAA_HD R de2000(R L1, R A1, R B1, R L2, R A2, R B2, R* g1, R* g2) {
    const R P25_7 = R(6103515625.0);  // 25^7
    const bool n1 = (A1 == R(0)) && (B1 == R(0));
    const bool n2 = (A2 == R(0)) && (B2 == R(0));
    if (n1) B1 += R(0.0001);
    if (n2) B2 += R(0.0001);
    const R C1 = M<R, F>::sqrt_(A1 * A1 + B1 * B1);
    const R C2 = M<R, F>::sqrt_(A2 * A2 + B2 * B2);
    const R cbar = (C1 + C2) * R(0.5);
    R c6, c7;
    pow67(cbar, c6, c7);
    const R u = M<R, F>::div(c7, c7 + P25_7);
    const R su = M<R, F>::sqrt_(u);
    const R G = R(0.5) * (R(1) - su);
    const R a1p = (R(1) + G) * A1, a2p = (R(1) + G) * A2;
    const R c1p = M<R, F>::sqrt_(a1p * a1p + B1 * B1);
    const R c2p = M<R, F>::sqrt_(a2p * a2p + B2 * B2);
    const R h1p = n1 ? R(0) : hue_deg(B1, a1p);
    const R h2p = n2 ? R(0) : hue_deg(B2, a2p);
    const bool nz = (C1 * C2) != R(0);
    const bool on = !(n1 || n2);
    const R dLp = L2 - L1;
    const R dCp = c2p - c1p;
    const R dh = h2p - h1p;
    R dhp = R(0);
    if (nz) dhp = (fabs(dh) <= R(180)) ? dh : (dh > R(180) ? dh - R(360) : dh + R(360));
    const R sq12 = M<R, F>::sqrt_(c1p * c2p);
    const R half_ang = K<R>::rad * dhp * R(0.5);
    R sn, cs_half;
    M<R, F>::sincos_(half_ang, sn, cs_half);
    const R dHp = on ? R(2) * sq12 * sn : R(0);
    const R Lbar = (L1 + L2) * R(0.5);
    const R cpbar = (c1p + c2p) * R(0.5);
    const R hs = h1p + h2p;
    R hbar = R(0);
    if (nz) {
        const bool near = fabs(dh) <= R(180);
        const bool lt360 = fabs(hs) < R(360);
        hbar = (near ? hs : (lt360 ? hs + R(360) : hs - R(360))) * R(0.5);
    }
    const R ang1 = K<R>::rad * (hbar - R(39)), ang2 = K<R>::rad * (R(2) * hbar);
    const R ang3 = K<R>::rad * (R(3) * hbar + R(6)), ang4 = K<R>::rad * (R(4) * hbar - R(63));
    R s1, k1, s2, k2, s3, k3, s4, k4;
    if (WithGrad) { M<R, F>::sincos_(ang1, s1, k1); M<R, F>::sincos_(ang2, s2, k2); M<R, F>::sincos_(ang3, s3, k3); M<R, F>::sincos_(ang4, s4, k4); }
    else { k1 = M<R, F>::cos_(ang1); k2 = M<R, F>::cos_(ang2); k3 = M<R, F>::cos_(ang3); k4 = M<R, F>::cos_(ang4); s1 = s2 = s3 = s4 = R(0); }
    const R T = R(1) - R(0.17) * k1 + R(0.24) * k2 + R(0.32) * k3 - R(0.2) * k4;
    const R hq = (hbar - R(275)) / R(25);
    const R ex = M<R, F>::exp_(-(hq * hq));
    const R dtheta = R(30) * ex;
    R cp6, cp7;
    pow67(cpbar, cp6, cp7);
    const R v = M<R, F>::div(cp7, cp7 + P25_7);
    const R rC = M<R, F>::sqrt_(v);
    const R Lm = Lbar - R(50);
    const R q = Lm * Lm;
    const R sq20 = M<R, F>::sqrt_(R(20) + q);
    const R sL = R(1) + M<R, F>::div(R(0.015) * q, sq20);
    const R sC = R(1) + R(0.045) * cpbar;
    const R sH = R(1) + R(0.015) * cpbar * T;
    const R ang5 = K<R>::rad * (R(2) * dtheta);
    R s5, k5;
    if (WithGrad) M<R, F>::sincos_(ang5, s5, k5); else { s5 = M<R, F>::sin_(ang5); k5 = R(0); }
    const R rT = R(-2) * rC * s5;
    const R tl = M<R, F>::div(dLp, sL), tc = M<R, F>::div(dCp, sC), th = M<R, F>::div(dHp, sH);
    const R sq = on ? (tl * tl + tc * tc + th * th + rT * tc * th) : (tl * tl);
    const bool pos = sq > R(0);
    const R res = pos ? M<R, F>::sqrt_(sq) : R(0);
    if (!WithGrad) return res;

    // ------------------------------- reverse mode ---------------------------------------------------
    R gL1 = 0, gA1 = 0, gB1 = 0, gL2 = 0, gA2 = 0, gB2 = 0;
    if (pos) {
        const R dsq = fdiv(R(0.5), res);
        const R isL = fdiv(R(1), sL), isC = fdiv(R(1), sC), isH = fdiv(R(1), sH);
        // sq = tl^2 + on*(tc^2 + th^2 + rT*tc*th)
        const R d_tl = dsq * R(2) * tl;
        R d_tc = 0, d_th = 0, d_rT = 0;
        if (on) {
            d_tc = dsq * (R(2) * tc + rT * th);
            d_th = dsq * (R(2) * th + rT * tc);
            d_rT = dsq * tc * th;
        }
        // tl = dLp/sL ; tc = dCp/sC ; th = dHp/sH
        const R d_dLp = d_tl * isL;
        R d_sL = -d_tl * tl * isL;
        const R d_dCp = d_tc * isC;
        R d_sC = -d_tc * tc * isC;
        const R d_dHp = d_th * isH;
        R d_sH = -d_th * th * isH;
        // rT = -2 rC sin(rad*2*dtheta)
        const R d_rC = d_rT * R(-2) * s5;
        const R d_dtheta = d_rT * R(-2) * rC * k5 * K<R>::rad * R(2);
        // sH = 1 + 0.015 cpbar T ; sC = 1 + 0.045 cpbar
        R d_cpbar = d_sH * R(0.015) * T + d_sC * R(0.045);
        const R d_T = d_sH * R(0.015) * cpbar;
        // sL = 1 + 0.015 q / sqrt(20+q), q = (Lbar-50)^2
        const R isq20 = fdiv(R(1), sq20);
        const R d_q = d_sL * R(0.015) * (isq20 - R(0.5) * q * isq20 * isq20 * isq20);      // 1/(sq20*(20+q)) = 1/sq20^3
        const R d_Lbar = d_q * R(2) * Lm;
        // rC = sqrt(v), v = cp7/(cp7+25^7), cp7 = cpbar^7
        {
            const R d_v = fdiv(d_rC * R(0.5), rC);
            const R den = cp7 + P25_7;
            const R d_cp7 = fdiv(d_v * P25_7, den * den);
            d_cpbar += d_cp7 * R(7) * cp6;
        }
        // dtheta = 30 exp(-hq^2), hq = (hbar-275)/25
        R d_hbar = d_dtheta * R(30) * ex * (R(-2) * hq) * R(1.0 / 25.0);
        // T
        d_hbar += d_T * K<R>::rad * (R(0.17) * s1 - R(0.24) * R(2) * s2 - R(0.32) * R(3) * s3 + R(0.2) * R(4) * s4);
        // hbar = 0.5*(h1p+h2p (+-360)) when nz
        R d_h1p = 0, d_h2p = 0;
        if (nz) { d_h1p += R(0.5) * d_hbar; d_h2p += R(0.5) * d_hbar; }
        // cpbar = (c1p+c2p)/2 ; Lbar = (L1+L2)/2
        R d_c1p = R(0.5) * d_cpbar, d_c2p = R(0.5) * d_cpbar;
        gL1 += R(0.5) * d_Lbar; gL2 += R(0.5) * d_Lbar;
        // dHp = on * 2 sqrt(c1p c2p) sin(rad*dhp/2)
        if (on) {
            const R d_sq12 = d_dHp * R(2) * sn;
            const R d_dhp = d_dHp * R(2) * sq12 * cs_half * K<R>::rad * R(0.5);
            const R d_prod = fdiv(d_sq12 * R(0.5), sq12);
            d_c1p += d_prod * c2p;
            d_c2p += d_prod * c1p;
            if (nz) { d_h2p += d_dhp; d_h1p -= d_dhp; }
        }
        // dCp = c2p - c1p ; dLp = L2 - L1
        d_c2p += d_dCp; d_c1p -= d_dCp;
        gL2 += d_dLp; gL1 -= d_dLp;
        // h1p = hue(B1, a1p) unless neutral: d atan2(y,x): dy = x/(x^2+y^2), dx = -y/(x^2+y^2); degrees
        R d_a1p = 0, d_a2p = 0;
        if (!n1) {
            const R k = fdiv(d_h1p * K<R>::deg, a1p * a1p + B1 * B1);
            gB1 += k * a1p;
            d_a1p += k * (-B1);
        }
        if (!n2) {
            const R k = fdiv(d_h2p * K<R>::deg, a2p * a2p + B2 * B2);
            gB2 += k * a2p;
            d_a2p += k * (-B2);
        }
        // c1p = sqrt(a1p^2 + B1^2)
        {
            const R k1 = fdiv(d_c1p, c1p), k2 = fdiv(d_c2p, c2p);
            d_a1p += k1 * a1p; gB1 += k1 * B1;
            d_a2p += k2 * a2p; gB2 += k2 * B2;
        }
        // a1p = (1+G) A1
        gA1 += d_a1p * (R(1) + G); gA2 += d_a2p * (R(1) + G);
        const R d_G = d_a1p * A1 + d_a2p * A2;
        // G = 0.5 (1 - sqrt(u)), u = c7/(c7+25^7), c7 = cbar^7, cbar = (C1+C2)/2
        {
            const R d_su = R(-0.5) * d_G;
            const R d_u = fdiv(d_su * R(0.5), su);
            const R den = c7 + P25_7;
            const R d_c7 = fdiv(d_u * P25_7, den * den);
            const R d_cbar = d_c7 * R(7) * c6;
            const R d_C = R(0.5) * d_cbar;
            const R k1 = fdiv(d_C, C1), k2 = fdiv(d_C, C2);
            gA1 += k1 * A1; gB1 += k1 * B1;
            gA2 += k2 * A2; gB2 += k2 * B2;
        }
    }
    g1[0] = gL1; g1[1] = gA1; g1[2] = gB1;
    g2[0] = gL2; g2[1] = gA2; g2[2] = gB2;
    return res;
}

}  // namespace color
}  // namespace spaa

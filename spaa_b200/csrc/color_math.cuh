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

// ---- sRGB channel -> 100 * linear  (differential_color_functions.py:16-20) ----------------------------
template <typename R> SPAA_HD R srgb_lin100(R c) {
    return c > R(0.0405) ? R(100) * rpow((c + R(0.055)) / R(1.055), R(2.4)) : R(100) * (c / R(12.92));
}
template <typename R> SPAA_HD R srgb_lin100_grad(R c) {
    return c > R(0.0405) ? R(100) * (R(2.4) * rpow((c + R(0.055)) / R(1.055), R(1.4))) / R(1.055) : R(100) / R(12.92);
}

// ---- Lab f()  (:27-36): exact zero -> 0 with zero slope ----------------------------------------------
template <typename R> SPAA_HD R lab_f(R t) {
    if (t == R(0)) return R(0);
    return t > R(0.008856) ? rpow(t, R(1.0 / 3.0)) : R(7.787) * t + R(16.0 / 116.0);
}
template <typename R> SPAA_HD R lab_f_grad(R t) {
    if (t == R(0)) return R(0);
    return t > R(0.008856) ? R(1.0 / 3.0) * rpow(t, R(1.0 / 3.0 - 1.0)) : R(7.787);
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

// reverse mode of rgb_to_lab: (dL,dA,dB) -> (dr,dg,db)
template <typename R> SPAA_HD void rgb_to_lab_bwd(R r, R g, R b, R dL, R dA, R dB, R& dr, R& dg, R& db) {
    const R lr = srgb_lin100(r), lg = srgb_lin100(g), lb = srgb_lin100(b);
    const R X = R(0.4124) * lr + R(0.3576) * lg + R(0.1805) * lb;
    const R Y = R(0.2126) * lr + R(0.7152) * lg + R(0.0722) * lb;
    const R Z = R(0.0193) * lr + R(0.1192) * lg + R(0.9504) * lb;
    const R dfx = R(500) * dA;
    const R dfy = R(116) * dL - R(500) * dA + R(200) * dB;
    const R dfz = -R(200) * dB;
    const R dX = dfx * lab_f_grad(X / White<R>::xn) / White<R>::xn;
    const R dY = dfy * lab_f_grad(Y / White<R>::yn) / White<R>::yn;
    const R dZ = dfz * lab_f_grad(Z / White<R>::zn) / White<R>::zn;
    dr = (R(0.4124) * dX + R(0.2126) * dY + R(0.0193) * dZ) * srgb_lin100_grad(r);
    dg = (R(0.3576) * dX + R(0.7152) * dY + R(0.1192) * dZ) * srgb_lin100_grad(g);
    db = (R(0.1805) * dX + R(0.0722) * dY + R(0.9504) * dZ) * srgb_lin100_grad(b);
}

// hue in degrees, [0,360)  (:73-81).  After the neutral nudge the arguments are never both zero.
template <typename R> SPAA_HD R hue_deg(R y, R x) {
    R h = K<R>::deg * atan2(y, x);
    return h < R(0) ? h + R(360) : h;
}

// ---- ciede2000_diff (:109-180), forward; optionally reverse mode -------------------------------------
// g1[3], g2[3] receive d(dE)/d(L1,A1,B1) and d(dE)/d(L2,A2,B2) (multiply by the cotangent outside).
template <typename R, bool WithGrad>
SPAA_HD R de2000(R L1, R A1, R B1, R L2, R A2, R B2, R* g1, R* g2) {
    const R P25_7 = R(6103515625.0);  // 25^7
    const bool n1 = (A1 == R(0)) && (B1 == R(0));
    const bool n2 = (A2 == R(0)) && (B2 == R(0));
    if (n1) B1 += R(0.0001);
    if (n2) B2 += R(0.0001);
    const R C1 = sqrt(A1 * A1 + B1 * B1);
    const R C2 = sqrt(A2 * A2 + B2 * B2);
    const R cbar = (C1 + C2) * R(0.5);
    const R c7 = rpow(cbar, R(7));
    const R u = c7 / (c7 + P25_7);
    const R su = sqrt(u);
    const R G = R(0.5) * (R(1) - su);
    const R a1p = (R(1) + G) * A1, a2p = (R(1) + G) * A2;
    const R c1p = sqrt(a1p * a1p + B1 * B1);
    const R c2p = sqrt(a2p * a2p + B2 * B2);
    const R h1p = n1 ? R(0) : hue_deg(B1, a1p);
    const R h2p = n2 ? R(0) : hue_deg(B2, a2p);
    const bool nz = (C1 * C2) != R(0);
    const bool on = !(n1 || n2);
    const R dLp = L2 - L1;
    const R dCp = c2p - c1p;
    const R dh = h2p - h1p;
    R dhp = R(0);
    if (nz) dhp = (fabs(dh) <= R(180)) ? dh : (dh > R(180) ? dh - R(360) : dh + R(360));
    const R sq12 = sqrt(c1p * c2p);
    const R half_ang = K<R>::rad * dhp * R(0.5);
    const R sn = sin(half_ang);
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
    const R T = R(1) - R(0.17) * cos(ang1) + R(0.24) * cos(ang2) + R(0.32) * cos(ang3) - R(0.2) * cos(ang4);
    const R hq = (hbar - R(275)) / R(25);
    const R ex = exp(-(hq * hq));
    const R dtheta = R(30) * ex;
    const R cp7 = rpow(cpbar, R(7));
    const R v = cp7 / (cp7 + P25_7);
    const R rC = sqrt(v);
    const R Lm = Lbar - R(50);
    const R q = Lm * Lm;
    const R sq20 = sqrt(R(20) + q);
    const R sL = R(1) + (R(0.015) * q) / sq20;
    const R sC = R(1) + R(0.045) * cpbar;
    const R sH = R(1) + R(0.015) * cpbar * T;
    const R ang5 = K<R>::rad * (R(2) * dtheta);
    const R s5 = sin(ang5);
    const R rT = R(-2) * rC * s5;
    const R tl = dLp / sL, tc = dCp / sC, th = dHp / sH;
    const R sq = on ? (tl * tl + tc * tc + th * th + rT * tc * th) : (tl * tl);
    const bool pos = sq > R(0);
    const R res = pos ? sqrt(sq) : R(0);
    if (!WithGrad) return res;

    // ------------------------------- reverse mode ---------------------------------------------------
    R gL1 = 0, gA1 = 0, gB1 = 0, gL2 = 0, gA2 = 0, gB2 = 0;
    if (pos) {
        const R dsq = R(0.5) / res;
        // sq = tl^2 + on*(tc^2 + th^2 + rT*tc*th)
        const R d_tl = dsq * R(2) * tl;
        R d_tc = 0, d_th = 0, d_rT = 0;
        if (on) {
            d_tc = dsq * (R(2) * tc + rT * th);
            d_th = dsq * (R(2) * th + rT * tc);
            d_rT = dsq * tc * th;
        }
        // tl = dLp/sL ; tc = dCp/sC ; th = dHp/sH
        const R d_dLp = d_tl / sL;
        R d_sL = -d_tl * tl / sL;
        const R d_dCp = d_tc / sC;
        R d_sC = -d_tc * tc / sC;
        const R d_dHp = d_th / sH;
        R d_sH = -d_th * th / sH;
        // rT = -2 rC sin(rad*2*dtheta)
        const R d_rC = d_rT * R(-2) * s5;
        const R d_dtheta = d_rT * R(-2) * rC * cos(ang5) * K<R>::rad * R(2);
        // sH = 1 + 0.015 cpbar T ; sC = 1 + 0.045 cpbar
        R d_cpbar = d_sH * R(0.015) * T + d_sC * R(0.045);
        const R d_T = d_sH * R(0.015) * cpbar;
        // sL = 1 + 0.015 q / sqrt(20+q), q = (Lbar-50)^2
        const R d_q = d_sL * R(0.015) * (R(1) / sq20 - R(0.5) * q / (sq20 * (R(20) + q)));
        const R d_Lbar = d_q * R(2) * Lm;
        // rC = sqrt(v), v = cp7/(cp7+25^7), cp7 = cpbar^7
        {
            const R d_v = d_rC * R(0.5) / rC;
            const R den = cp7 + P25_7;
            const R d_cp7 = d_v * P25_7 / (den * den);
            d_cpbar += d_cp7 * R(7) * rpow(cpbar, R(6));
        }
        // dtheta = 30 exp(-hq^2), hq = (hbar-275)/25
        R d_hbar = d_dtheta * R(30) * ex * (R(-2) * hq) / R(25);
        // T
        d_hbar += d_T * K<R>::rad * (R(0.17) * sin(ang1) - R(0.24) * R(2) * sin(ang2) - R(0.32) * R(3) * sin(ang3) +
                                      R(0.2) * R(4) * sin(ang4));
        // hbar = 0.5*(h1p+h2p (+-360)) when nz
        R d_h1p = 0, d_h2p = 0;
        if (nz) { d_h1p += R(0.5) * d_hbar; d_h2p += R(0.5) * d_hbar; }
        // cpbar = (c1p+c2p)/2 ; Lbar = (L1+L2)/2
        R d_c1p = R(0.5) * d_cpbar, d_c2p = R(0.5) * d_cpbar;
        gL1 += R(0.5) * d_Lbar; gL2 += R(0.5) * d_Lbar;
        // dHp = on * 2 sqrt(c1p c2p) sin(rad*dhp/2)
        if (on) {
            const R d_sq12 = d_dHp * R(2) * sn;
            const R d_dhp = d_dHp * R(2) * sq12 * cos(half_ang) * K<R>::rad * R(0.5);
            const R d_prod = d_sq12 * R(0.5) / sq12;
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
            const R den = a1p * a1p + B1 * B1;
            gB1 += d_h1p * K<R>::deg * a1p / den;
            d_a1p += d_h1p * K<R>::deg * (-B1) / den;
        }
        if (!n2) {
            const R den = a2p * a2p + B2 * B2;
            gB2 += d_h2p * K<R>::deg * a2p / den;
            d_a2p += d_h2p * K<R>::deg * (-B2) / den;
        }
        // c1p = sqrt(a1p^2 + B1^2)
        d_a1p += d_c1p * a1p / c1p; gB1 += d_c1p * B1 / c1p;
        d_a2p += d_c2p * a2p / c2p; gB2 += d_c2p * B2 / c2p;
        // a1p = (1+G) A1
        gA1 += d_a1p * (R(1) + G); gA2 += d_a2p * (R(1) + G);
        const R d_G = d_a1p * A1 + d_a2p * A2;
        // G = 0.5 (1 - sqrt(u)), u = c7/(c7+25^7), c7 = cbar^7, cbar = (C1+C2)/2
        {
            const R d_su = R(-0.5) * d_G;
            const R d_u = d_su * R(0.5) / su;
            const R den = c7 + P25_7;
            const R d_c7 = d_u * P25_7 / (den * den);
            const R d_cbar = d_c7 * R(7) * rpow(cbar, R(6));
            const R d_C = R(0.5) * d_cbar;
            gA1 += d_C * A1 / C1; gB1 += d_C * B1 / C1;
            gA2 += d_C * A2 / C2; gB2 += d_C * B2 / C2;
        }
    }
    g1[0] = gL1; g1[1] = gA1; g1[2] = gB1;
    g2[0] = gL2; g2[1] = gA2; g2[2] = gB2;
    return res;
}

}  // namespace color
}  // namespace spaa

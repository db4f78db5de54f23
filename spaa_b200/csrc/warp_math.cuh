// Per-pixel math of WarpingNet's sampling grid: affine grid (F.affine_grid, align_corners=True) sampled by a
// thin-plate-spline grid (pytorch_tps.tps_grid) -- /root/reference/src/python/models.py:163-172 and
// pytorch_tps.py:54-106 -- fused analytically: the affine "image" is never materialised.
// Also the bilinear tap set-up shared by the grid_sample kernels (zeros padding, align_corners=True).
#pragma once
#include "common.cuh"
#include <cmath>

namespace spaa {
namespace warp {

// torch.linspace(lo, hi, n)[i] as ATen evaluates it (symmetric around the midpoint)
template <typename R> SPAA_HD R linspace_at(R lo, R hi, int n, int i) {
    if (n <= 1) return lo;
    const R step = (hi - lo) / R(n - 1);
    return (i < n / 2) ? lo + step * R(i) : hi - step * R(n - 1 - i);
}

template <typename R> struct Taps {
    int x0, y0;        // north-west integer corner
    R wx1, wy1;        // weights of x0+1 / y0+1 ; (1-wx1),(1-wy1) belong to x0 / y0
    bool vx0, vx1, vy0, vy1;
};

// grid coordinate in [-1,1] -> taps on an H x W image (ATen grid_sampler_2d, bilinear, zeros, align_corners)
template <typename R> SPAA_HD Taps<R> make_taps(R gx, R gy, int H, int W) {
    Taps<R> t;
    const R ix = (gx + R(1)) * R(0.5) * R(W - 1);
    const R iy = (gy + R(1)) * R(0.5) * R(H - 1);
    const R fx = floor(ix), fy = floor(iy);
    t.x0 = (int)fx; t.y0 = (int)fy;
    t.wx1 = ix - fx; t.wy1 = iy - fy;
    t.vx0 = t.x0 >= 0 && t.x0 < W; t.vx1 = t.x0 + 1 >= 0 && t.x0 + 1 < W;
    t.vy0 = t.y0 >= 0 && t.y0 < H; t.vy1 = t.y0 + 1 >= 0 && t.y0 + 1 < H;
    return t;
}

// TPS radial basis as the reference evaluates it (pytorch_tps.py:61-62): D^2 * log(D + 1e-6)
template <typename R> SPAA_HD R tps_u(R dx, R dy) {
    const R D = sqrt(dx * dx + dy * dy);
    return (D * D) * log(D + R(1e-6));
}

// TPS sampling grid value at output pixel (y,x) of an H x W grid.  theta: [(T+2)][2] reduced form
// (rows 0..T-2 = w_1..w_{T-1}, rows T-1..T+1 = affine part), ctrl: [T][2] (x,y) in [0,1].
template <typename R>
SPAA_HD void tps_point(const R* theta, const R* ctrl, int T, int H, int W, int y, int x, R& gx, R& gy) {
    const R px = linspace_at<R>(R(0), R(1), W, x), py = linspace_at<R>(R(0), R(1), H, y);
    const R* a = theta + (T - 1) * 2;
    R zx = a[0] + a[2] * px + a[4] * py;
    R zy = a[1] + a[3] * px + a[5] * py;
    const R u0 = tps_u(px - ctrl[0], py - ctrl[1]);
    R sx = 0, sy = 0;
    for (int t = 1; t < T; ++t) {
        const R u = tps_u(px - ctrl[2 * t], py - ctrl[2 * t + 1]) - u0;   // w_0 = -sum(w)  (pytorch_tps.py:66-69)
        sx += u * theta[2 * (t - 1)];
        sy += u * theta[2 * (t - 1) + 1];
    }
    gx = (px + (zx + sx)) * R(2) - R(1);
    gy = (py + (zy + sy)) * R(2) - R(1);
}

// value of the (virtual) affine-grid image channel c at integer pixel (v,u) of an Hin x Win image
template <typename R> SPAA_HD R affine_img(const R* aff, int c, int Hin, int Win, int v, int u) {
    const R xn = linspace_at<R>(R(-1), R(1), Win, u), yn = linspace_at<R>(R(-1), R(1), Hin, v);
    return aff[3 * c] * xn + aff[3 * c + 1] * yn + aff[3 * c + 2];
}

// coarse grid = grid_sample(affine_grid(aff; Hin x Win), tps_grid(theta; H x W))   (models.py:168-172)
template <typename R>
SPAA_HD void coarse_grid_point(const R* aff, const R* theta, const R* ctrl, int T, int Hin, int Win, int H, int W,
                               int y, int x, R& ox, R& oy) {
    R gx, gy;
    tps_point(theta, ctrl, T, H, W, y, x, gx, gy);
    const Taps<R> t = make_taps(gx, gy, Hin, Win);
    const R w00 = (R(1) - t.wx1) * (R(1) - t.wy1), w01 = t.wx1 * (R(1) - t.wy1);
    const R w10 = (R(1) - t.wx1) * t.wy1, w11 = t.wx1 * t.wy1;
    R o[2];
    for (int c = 0; c < 2; ++c) {
        R s = 0;
        if (t.vy0 && t.vx0) s += affine_img(aff, c, Hin, Win, t.y0, t.x0) * w00;
        if (t.vy0 && t.vx1) s += affine_img(aff, c, Hin, Win, t.y0, t.x0 + 1) * w01;
        if (t.vy1 && t.vx0) s += affine_img(aff, c, Hin, Win, t.y0 + 1, t.x0) * w10;
        if (t.vy1 && t.vx1) s += affine_img(aff, c, Hin, Win, t.y0 + 1, t.x0 + 1) * w11;
        o[c] = s;
    }
    ox = o[0]; oy = o[1];
}

// reverse mode at one pixel: given d(out) returns the local contributions
//   daff[6]  : gradient wrt the 2x3 affine matrix (row-major)
//   dzx,dzy  : gradient wrt the TPS offset z (before *2-1), to be contracted with [1,px,py] and (U_t-U_0)
template <typename R>
SPAA_HD void coarse_grid_point_bwd_local(const R* aff, const R* theta, const R* ctrl, int T, int Hin, int Win, int H,
                                         int W, int y, int x, R dox, R doy, R* daff, R& dzx, R& dzy) {
    R gx, gy;
    tps_point(theta, ctrl, T, H, W, y, x, gx, gy);
    const Taps<R> t = make_taps(gx, gy, Hin, Win);
    const R wx0 = R(1) - t.wx1, wy0 = R(1) - t.wy1;
    const R dout[2] = {dox, doy};
    R gix = 0, giy = 0;
    for (int i = 0; i < 6; ++i) daff[i] = 0;
    for (int c = 0; c < 2; ++c) {
        for (int k = 0; k < 4; ++k) {
            const int dy = k >> 1, dx = k & 1;
            const bool ok = (dy ? t.vy1 : t.vy0) && (dx ? t.vx1 : t.vx0);
            if (!ok) continue;
            const int v = t.y0 + dy, u = t.x0 + dx;
            const R wgt = (dx ? t.wx1 : wx0) * (dy ? t.wy1 : wy0);
            const R xn = linspace_at<R>(R(-1), R(1), Win, u), yn = linspace_at<R>(R(-1), R(1), Hin, v);
            const R val = aff[3 * c] * xn + aff[3 * c + 1] * yn + aff[3 * c + 2];
            daff[3 * c] += wgt * xn * dout[c];
            daff[3 * c + 1] += wgt * yn * dout[c];
            daff[3 * c + 2] += wgt * dout[c];
            gix += val * (dx ? R(1) : R(-1)) * (dy ? t.wy1 : wy0) * dout[c];
            giy += val * (dy ? R(1) : R(-1)) * (dx ? t.wx1 : wx0) * dout[c];
        }
    }
    // ix = (gx+1)/2*(Win-1), gx = (px+z)*2-1  ->  d ix / d z = (Win-1)
    dzx = gix * R(0.5) * R(Win - 1) * R(2);
    dzy = giy * R(0.5) * R(Hin - 1) * R(2);
}

// host/test convenience: accumulate the full parameter gradients for one pixel
template <typename R>
SPAA_HD void coarse_grid_point_bwd(const R* aff, const R* theta, const R* ctrl, int T, int Hin, int Win, int H, int W,
                                   int y, int x, R dox, R doy, R* daff_acc, R* dtheta_acc) {
    R da[6], dzx, dzy;
    coarse_grid_point_bwd_local(aff, theta, ctrl, T, Hin, Win, H, W, y, x, dox, doy, da, dzx, dzy);
    for (int i = 0; i < 6; ++i) daff_acc[i] += da[i];
    const R px = linspace_at<R>(R(0), R(1), W, x), py = linspace_at<R>(R(0), R(1), H, y);
    R* a = dtheta_acc + (T - 1) * 2;
    a[0] += dzx; a[1] += dzy; a[2] += dzx * px; a[3] += dzy * px; a[4] += dzx * py; a[5] += dzy * py;
    const R u0 = tps_u(px - ctrl[0], py - ctrl[1]);
    for (int t = 1; t < T; ++t) {
        const R u = tps_u(px - ctrl[2 * t], py - ctrl[2 * t + 1]) - u0;
        dtheta_acc[2 * (t - 1)] += u * dzx;
        dtheta_acc[2 * (t - 1) + 1] += u * dzy;
    }
}

}  // namespace warp
}  // namespace spaa

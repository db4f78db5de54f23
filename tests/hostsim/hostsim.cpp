// TEST-ONLY host build of the __host__ __device__ math headers in spaa_b200/csrc (no CUDA needed).
// Lets the CPU test-suite validate the per-pixel device math (and hand-derived derivatives) against the
// oracle before any GPU time is spent.  Never loaded by the product package.
#include <cmath>
#include <cstdint>
using std::sqrt; using std::pow; using std::atan2; using std::fabs; using std::sin; using std::cos; using std::exp;
using std::log; using std::floor; using std::cbrt;
#include "../../spaa_b200/csrc/color_math.cuh"
#include "../../spaa_b200/csrc/warp_math.cuh"

using namespace spaa;

template <typename R>
static void lab_fwd_t(const R* rgb, R* lab, int64_t n) {       // planar [3][n]
    for (int64_t i = 0; i < n; ++i) color::rgb_to_lab(rgb[i], rgb[n + i], rgb[2 * n + i], lab[i], lab[n + i], lab[2 * n + i]);
}
template <typename R>
static void lab_bwd_t(const R* rgb, const R* dlab, R* drgb, int64_t n) {
    for (int64_t i = 0; i < n; ++i)
        color::rgb_to_lab_bwd(rgb[i], rgb[n + i], rgb[2 * n + i], dlab[i], dlab[n + i], dlab[2 * n + i], drgb[i], drgb[n + i], drgb[2 * n + i]);
}
template <typename R>
static void de_t(const R* l1, const R* l2, R* de, R* g1, R* g2, int64_t n) {
    for (int64_t i = 0; i < n; ++i) {
        R a[3], b[3];
        de[i] = color::de2000<R, true>(l1[i], l1[n + i], l1[2 * n + i], l2[i], l2[n + i], l2[2 * n + i], a, b);
        for (int c = 0; c < 3; ++c) { g1[c * n + i] = a[c]; g2[c * n + i] = b[c]; }
    }
}

extern "C" {
void hs_lab_fwd_f32(const float* rgb, float* lab, int64_t n) { lab_fwd_t(rgb, lab, n); }
void hs_lab_fwd_f64(const double* rgb, double* lab, int64_t n) { lab_fwd_t(rgb, lab, n); }
void hs_lab_bwd_f32(const float* rgb, const float* d, float* o, int64_t n) { lab_bwd_t(rgb, d, o, n); }
void hs_lab_bwd_f64(const double* rgb, const double* d, double* o, int64_t n) { lab_bwd_t(rgb, d, o, n); }
void hs_de_f32(const float* a, const float* b, float* de, float* g1, float* g2, int64_t n) { de_t(a, b, de, g1, g2, n); }
void hs_de_f64(const double* a, const double* b, double* de, double* g1, double* g2, int64_t n) { de_t(a, b, de, g1, g2, n); }

// warp math ------------------------------------------------------------------------------------------
void hs_coarse_grid_f32(const float* aff, const float* theta, const float* ctrl, int T, int Hin, int Win, int H, int W, float* out) {
    for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) {
        float gx, gy; warp::coarse_grid_point<float>(aff, theta, ctrl, T, Hin, Win, H, W, y, x, gx, gy);
        out[(y * W + x) * 2] = gx; out[(y * W + x) * 2 + 1] = gy;
    }
}
void hs_coarse_grid_f64(const double* aff, const double* theta, const double* ctrl, int T, int Hin, int Win, int H, int W, double* out) {
    for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) {
        double gx, gy; warp::coarse_grid_point<double>(aff, theta, ctrl, T, Hin, Win, H, W, y, x, gx, gy);
        out[(y * W + x) * 2] = gx; out[(y * W + x) * 2 + 1] = gy;
    }
}
void hs_coarse_grid_bwd_f64(const double* aff, const double* theta, const double* ctrl, int T, int Hin, int Win, int H, int W,
                            const double* dgrid, double* daff, double* dtheta) {
    for (int i = 0; i < 6; ++i) daff[i] = 0;
    for (int i = 0; i < (T + 2) * 2; ++i) dtheta[i] = 0;
    for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x)
        warp::coarse_grid_point_bwd<double>(aff, theta, ctrl, T, Hin, Win, H, W, y, x, dgrid[(y * W + x) * 2], dgrid[(y * W + x) * 2 + 1], daff, dtheta);
}
}

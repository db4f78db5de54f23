// WarpingNet kernels: fused affine o TPS sampling-grid generation (fwd/bwd) and bilinear grid_sample
// (fwd, bwd-input scatter, bwd-grid).  All HBM/L2-bound gather/scatter work; see DESIGN.md section 4.
#include "common.cuh"
#include "warp_math.cuh"
#include "../../include/spaa_b200.h"

using namespace spaa;

namespace {

constexpr int kThreads = 256;
constexpr int kMaxT = 64;          // control points held in shared memory (reference uses 6x6 = 36)

struct GridParams {
    int T, Hin, Win, H, W;
};

// Backward: every thread owns pixels, accumulates the 6 + 2(T+2) parameter gradients privately in a
// round-robin over shared-memory reductions, block partials go to the workspace and the last block sums them.
template <int TT>       // accumulator capacity: control points <= TT (36 = the reference's 6 x 6 grid: ~110 registers, two blocks per SM; kMaxT otherwise)
__global__ void __launch_bounds__(kThreads, TT <= 36 ? 2 : 1) coarse_grid_bwd_kernel(const float* __restrict__ aff, GridParams gp, const float* __restrict__ theta, const float* __restrict__ ctrl,
                                                                    const float* __restrict__ dgrid, float* __restrict__ daff,
                                                                    float* __restrict__ dtheta, float* __restrict__ partial,
                                                                    unsigned* __restrict__ counter) {
    __shared__ float s_theta[(kMaxT + 2) * 2];
    __shared__ float s_ctrl[kMaxT * 2];
    __shared__ float s_acc[6 + (kMaxT + 2) * 2];
    __shared__ float s_aff[6];
    __shared__ bool is_last;
    const int nout = 6 + (gp.T + 2) * 2;
    const bool has_aff = aff != nullptr;
    if (threadIdx.x < 6) s_aff[threadIdx.x] = has_aff ? aff[threadIdx.x] : 0.f;
    for (int i = threadIdx.x; i < (gp.T + 2) * 2; i += blockDim.x) s_theta[i] = theta[i];
    for (int i = threadIdx.x; i < gp.T * 2; i += blockDim.x) s_ctrl[i] = ctrl[i];
    for (int i = threadIdx.x; i < nout; i += blockDim.x) s_acc[i] = 0.f;
    __syncthreads();
    const int HW = gp.H * gp.W;
    const int lane = threadIdx.x & 31;
    // Every thread keeps its 6 + 2 (T + 2) partial sums in REGISTERS over the pixels it owns and the block reduces them once at the end.  (The first
    // version reduced every quantity across the warp for every 32 pixels -- 82 shuffle trees and as many shared-memory atomics per iteration: 68 us for
    // 76 800 pixels, latency bound, 2 % of a training step.)
    float da_acc[6] = {0, 0, 0, 0, 0, 0};
    float ax[TT + 2], ay[TT + 2];            // theta gradient: rows 0..T-2 = TPS weights, rows T-1..T+1 = its affine part [1, x, y]
#pragma unroll
    for (int t = 0; t < TT + 2; ++t) { ax[t] = 0.f; ay[t] = 0.f; }
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += gridDim.x * blockDim.x) {
        float da[6] = {0, 0, 0, 0, 0, 0}, dzx = 0.f, dzy = 0.f;
        const int y = p / gp.W, x = p - y * gp.W;
        const float gox = __ldg(dgrid + p), goy = __ldg(dgrid + HW + p);
        const float px = warp::linspace_at<float>(0.f, 1.f, gp.W, x);
        const float py = warp::linspace_at<float>(0.f, 1.f, gp.H, y);
        if (has_aff) {
            warp::coarse_grid_point_bwd_local<float>(s_aff, s_theta, s_ctrl, gp.T, gp.Hin, gp.Win, gp.H, gp.W, y, x, gox, goy, da, dzx, dzy);
#pragma unroll
            for (int i = 0; i < 6; ++i) da_acc[i] += da[i];
        } else {  // grid = (p + z)*2 - 1
            dzx = 2.f * gox; dzy = 2.f * goy;
        }
        const float u0 = warp::tps_u<float>(px - s_ctrl[0], py - s_ctrl[1]);
#pragma unroll
        for (int t = 1; t < TT; ++t) {
            if (t < gp.T) {
                const float u = warp::tps_u<float>(px - s_ctrl[2 * t], py - s_ctrl[2 * t + 1]) - u0;
                ax[t - 1] += u * dzx; ay[t - 1] += u * dzy;
            }
        }
        ax[TT - 1] += dzx; ay[TT - 1] += dzy; ax[TT] += dzx * px; ay[TT] += dzy * px; ax[TT + 1] += dzx * py; ay[TT + 1] += dzy * py;
    }
    if (has_aff) {
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            const float v = warp_sum(da_acc[i]);
            if (lane == 0) atomicAdd(&s_acc[i], v);
        }
    }
    float* acc_t = s_acc + 6;
#pragma unroll
    for (int t = 0; t < TT + 2; ++t) {
        // register slot t -> row of theta: TPS weights 0..T-2 stay, the affine slots TT-1 .. TT+1 are rows T-1 .. T+1
        const int row = t < TT - 1 ? t : gp.T - 1 + (t - (TT - 1));
        if (t < TT - 1 && t >= gp.T - 1) continue;
        const float vx = warp_sum(ax[t]), vy = warp_sum(ay[t]);
        if (lane == 0) { atomicAdd(acc_t + 2 * row, vx); atomicAdd(acc_t + 2 * row + 1, vy); }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nout; i += blockDim.x) partial[(int64_t)blockIdx.x * nout + i] = s_acc[i];
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(counter, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        const volatile float* q = partial;
        for (int i = threadIdx.x; i < nout; i += blockDim.x) {
            float s = 0.f;
            for (unsigned k = 0; k < gridDim.x; ++k) s += q[(int64_t)k * nout + i];
            if (i < 6) { if (daff) daff[i] = s; }
            else dtheta[i - 6] = s;
        }
        if (threadIdx.x == 0) *counter = 0u;
    }
}

constexpr int kGridBwdBlocks = 296;        // two blocks per SM (coarse_grid_bwd_kernel<36>)

__global__ void __launch_bounds__(kThreads) grid_finish_fwd_kernel(const float* __restrict__ coarse, const float* __restrict__ refine,
                                                                    float* __restrict__ fine, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float v = __ldg(coarse + i);
        if (refine) v = __ldg(refine + i) + v;
        fine[i] = fminf(fmaxf(v, -1.f), 1.f);
    }
}

__global__ void __launch_bounds__(kThreads) grid_finish_bwd_kernel(const float* __restrict__ coarse, const float* __restrict__ refine,
                                                                    const float* __restrict__ dfine, float* __restrict__ dsum, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float v = __ldg(coarse + i);
        if (refine) v = __ldg(refine + i) + v;
        dsum[i] = (v >= -1.f && v <= 1.f) ? __ldg(dfine + i) : 0.f;   // torch.clamp backward: inclusive bounds
    }
}

SPAA_D float clamp01_if(float v, int on) { return on ? fminf(fmaxf(v, 0.f), 1.f) : v; }

// one thread per output pixel, loops over channels (C is 2 or 3 here)
__global__ void __launch_bounds__(kThreads) grid_sample_fwd_kernel(const float* __restrict__ img, int C, int Hi, int Wi,
                                                                    const float* __restrict__ grid, int64_t grid_bs, int H, int W, int clamp01,
                                                                    const float* __restrict__ mask, float* __restrict__ out,
                                                                    const float* __restrict__ rough, int64_t rough_bs, float* __restrict__ out2,
                                                                    int64_t out2_bs) {
    const int b = blockIdx.y;
    const int HW = H * W;
    const int64_t HWi = (int64_t)Hi * Wi;
    const float* g = grid + (int64_t)b * grid_bs;
    const float* ib = img + (int64_t)b * C * HWi;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += gridDim.x * blockDim.x) {
        const warp::Taps<float> t = warp::make_taps<float>(__ldg(g + p), __ldg(g + HW + p), Hi, Wi);
        const float wx0 = 1.f - t.wx1, wy0 = 1.f - t.wy1;
        const float w00 = wx0 * wy0, w01 = t.wx1 * wy0, w10 = wx0 * t.wy1, w11 = t.wx1 * t.wy1;
        const float m = mask ? __ldg(mask + p) : 1.f;
        const int64_t o00 = (int64_t)t.y0 * Wi + t.x0;
        for (int c = 0; c < C; ++c) {
            const float* ic = ib + c * HWi;
            float v = 0.f;
            if (t.vy0 && t.vx0) v += clamp01_if(__ldg(ic + o00), clamp01) * w00;
            if (t.vy0 && t.vx1) v += clamp01_if(__ldg(ic + o00 + 1), clamp01) * w01;
            if (t.vy1 && t.vx0) v += clamp01_if(__ldg(ic + o00 + Wi), clamp01) * w10;
            if (t.vy1 && t.vx1) v += clamp01_if(__ldg(ic + o00 + Wi + 1), clamp01) * w11;
            if (mask) v *= m;
            out[((int64_t)b * C + c) * HW + p] = v;
            if (out2) out2[(int64_t)b * out2_bs + (int64_t)c * HW + p] = v * __ldg(rough + (int64_t)b * rough_bs + (int64_t)c * HW + p);
        }
    }
}

__global__ void __launch_bounds__(kThreads) grid_sample_bwd_input_kernel(const float* __restrict__ dout, const float* __restrict__ dout2,
                                                                          int64_t dout2_bs, const float* __restrict__ rough, int64_t rough_bs,
                                                                          const float* __restrict__ mask, const float* __restrict__ grid,
                                                                          int64_t grid_bs, int C, int Hi, int Wi, int H, int W,
                                                                          float* __restrict__ dimg) {
    const int b = blockIdx.y;
    const int HW = H * W;
    const int64_t HWi = (int64_t)Hi * Wi;
    const float* g = grid + (int64_t)b * grid_bs;
    float* db = dimg + (int64_t)b * C * HWi;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += gridDim.x * blockDim.x) {
        const float m = mask ? __ldg(mask + p) : 1.f;
        if (mask && m == 0.f) continue;
        const warp::Taps<float> t = warp::make_taps<float>(__ldg(g + p), __ldg(g + HW + p), Hi, Wi);
        const float wx0 = 1.f - t.wx1, wy0 = 1.f - t.wy1;
        const float w00 = wx0 * wy0, w01 = t.wx1 * wy0, w10 = wx0 * t.wy1, w11 = t.wx1 * t.wy1;
        const int64_t o00 = (int64_t)t.y0 * Wi + t.x0;
        for (int c = 0; c < C; ++c) {
            float d = __ldg(dout + ((int64_t)b * C + c) * HW + p);
            if (dout2) d += __ldg(dout2 + (int64_t)b * dout2_bs + (int64_t)c * HW + p) * __ldg(rough + (int64_t)b * rough_bs + (int64_t)c * HW + p);
            d *= m;
            float* dc = db + c * HWi;
            if (t.vy0 && t.vx0) atomicAdd(dc + o00, d * w00);
            if (t.vy0 && t.vx1) atomicAdd(dc + o00 + 1, d * w01);
            if (t.vy1 && t.vx0) atomicAdd(dc + o00 + Wi, d * w10);
            if (t.vy1 && t.vx1) atomicAdd(dc + o00 + Wi + 1, d * w11);
        }
    }
}

// dgrid for a grid shared by the whole batch (grid_bs == 0 -> sum over b) or per-sample grids.
__global__ void __launch_bounds__(kThreads) grid_sample_bwd_grid_kernel(const float* __restrict__ dout, const float* __restrict__ dout2,
                                                                         int64_t dout2_bs, const float* __restrict__ rough, int64_t rough_bs,
                                                                         const float* __restrict__ mask, const float* __restrict__ img, int clamp01,
                                                                         const float* __restrict__ grid, int64_t grid_bs, int B, int C, int Hi,
                                                                         int Wi, int H, int W, float* __restrict__ dgrid, int bchunk) {
    const int HW = H * W;
    const int64_t HWi = (int64_t)Hi * Wi;
    // blockIdx.y: per-sample grids -> the sample; one grid shared by the batch -> a chunk of `bchunk` samples whose contribution is added
    // to the (zero-filled) result atomically.  (One thread looping over all 24 training samples left the GPU a quarter full: 137 us.)
    const int bo = grid_bs == 0 ? 0 : blockIdx.y;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += gridDim.x * blockDim.x) {
        const float* g = grid + (int64_t)bo * grid_bs;
        const warp::Taps<float> t = warp::make_taps<float>(__ldg(g + p), __ldg(g + HW + p), Hi, Wi);
        const float wx0 = 1.f - t.wx1, wy0 = 1.f - t.wy1;
        const float m = mask ? __ldg(mask + p) : 1.f;
        const int64_t o00 = (int64_t)t.y0 * Wi + t.x0;
        float gix = 0.f, giy = 0.f;
        const int b_lo = grid_bs == 0 ? (int)blockIdx.y * bchunk : bo, b_hi = grid_bs == 0 ? min(B, b_lo + bchunk) : bo + 1;
        for (int b = b_lo; b < b_hi; ++b) {
            const float* ib = img + (int64_t)b * C * HWi;
            for (int c = 0; c < C; ++c) {
                float d = __ldg(dout + ((int64_t)b * C + c) * HW + p);
                if (dout2) d += __ldg(dout2 + (int64_t)b * dout2_bs + (int64_t)c * HW + p) * __ldg(rough + (int64_t)b * rough_bs + (int64_t)c * HW + p);
                d *= m;
                const float* ic = ib + c * HWi;
                const float v00 = (t.vy0 && t.vx0) ? clamp01_if(__ldg(ic + o00), clamp01) : 0.f;
                const float v01 = (t.vy0 && t.vx1) ? clamp01_if(__ldg(ic + o00 + 1), clamp01) : 0.f;
                const float v10 = (t.vy1 && t.vx0) ? clamp01_if(__ldg(ic + o00 + Wi), clamp01) : 0.f;
                const float v11 = (t.vy1 && t.vx1) ? clamp01_if(__ldg(ic + o00 + Wi + 1), clamp01) : 0.f;
                gix += d * ((v01 - v00) * wy0 + (v11 - v10) * t.wy1);
                giy += d * ((v10 - v00) * wx0 + (v11 - v01) * t.wx1);
            }
        }
        float* dg = dgrid + (int64_t)bo * 2 * HW;
        if (grid_bs == 0 && bchunk < B) {
            atomicAdd(dg + p, gix * 0.5f * (float)(Wi - 1));
            atomicAdd(dg + HW + p, giy * 0.5f * (float)(Hi - 1));
        } else {
            dg[p] = gix * 0.5f * (float)(Wi - 1);
            dg[HW + p] = giy * 0.5f * (float)(Hi - 1);
        }
    }
}

// 3-channel warp that emits the padded 16-channel 16-bit NHWC tensor the tensor-core convolutions read:
// [x0,x1,x2, s0,s1,s2, x0*s0,x1*s1,x2*s2, 0 x 7] per pixel (x = bilinear(clamp(img)) * mask, s = surface image).
template <bool F16>
__global__ void __launch_bounds__(kThreads) grid_sample_fwd_packed_kernel(const float* __restrict__ img, int Hi, int Wi, const float* __restrict__ grid,
                                                                           int64_t grid_bs, int H, int W, int clamp01, const float* __restrict__ mask,
                                                                           const float* __restrict__ rough, int64_t rough_bs, uint4* __restrict__ out16) {
    const int b = blockIdx.y;
    const int HW = H * W;
    const int64_t HWi = (int64_t)Hi * Wi;
    const float* g = grid + (int64_t)b * grid_bs;
    const float* ib = img + (int64_t)b * 3 * HWi;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += gridDim.x * blockDim.x) {
        const warp::Taps<float> t = warp::make_taps<float>(__ldg(g + p), __ldg(g + HW + p), Hi, Wi);
        const float wx0 = 1.f - t.wx1, wy0 = 1.f - t.wy1;
        const float w00 = wx0 * wy0, w01 = t.wx1 * wy0, w10 = wx0 * t.wy1, w11 = t.wx1 * t.wy1;
        const float m = mask ? __ldg(mask + p) : 1.f;
        const int64_t o00 = (int64_t)t.y0 * Wi + t.x0;
        float x[3], sv[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float* ic = ib + c * HWi;
            float v = 0.f;
            if (t.vy0 && t.vx0) v += clamp01_if(__ldg(ic + o00), clamp01) * w00;
            if (t.vy0 && t.vx1) v += clamp01_if(__ldg(ic + o00 + 1), clamp01) * w01;
            if (t.vy1 && t.vx0) v += clamp01_if(__ldg(ic + o00 + Wi), clamp01) * w10;
            if (t.vy1 && t.vx1) v += clamp01_if(__ldg(ic + o00 + Wi + 1), clamp01) * w11;
            if (mask) v *= m;
            x[c] = v;
            sv[c] = rough ? __ldg(rough + (int64_t)b * rough_bs + (int64_t)c * HW + p) : 0.f;
        }
        auto pk = [](float a, float c) -> uint32_t {
            if constexpr (F16) { const __half2 h = __floats2half2_rn(a, c); return *reinterpret_cast<const uint32_t*>(&h); }
            else { const __nv_bfloat162 h = __floats2bfloat162_rn(a, c); return *reinterpret_cast<const uint32_t*>(&h); }
        };
        uint4 lo, hi;
        lo.x = pk(x[0], x[1]); lo.y = pk(x[2], sv[0]); lo.z = pk(sv[1], sv[2]); lo.w = pk(x[0] * sv[0], x[1] * sv[1]);
        hi.x = pk(x[2] * sv[2], 0.f); hi.y = 0u; hi.z = 0u; hi.w = 0u;
        uint4* o = out16 + ((int64_t)b * HW + p) * 2;
        o[0] = lo; o[1] = hi;
    }
}

// ---- gather form of the warp's adjoint for a FIXED grid (the attack loops: the model is frozen) -------------------------
// warp_taps_kernel lists, per output pixel p and bilinear tap k, the input pixel q it reads and its weight (mask folded in; q = HWi marks an
// unused tap).  Sorted by q once per attack (host side: a stable sort, so the summation order is fixed) this is the CSR matrix of the adjoint:
// grid_sample_bwd_gather_kernel then gives every input pixel the sum of its contributions -- no atomics, no zero-fill of dimg, a deterministic
// result -- and the per-sample squared norm of the (clamp-masked) gradient that the normalised step needs comes out of the same pass.
__global__ void __launch_bounds__(kThreads) warp_taps_kernel(const float* __restrict__ grid, const float* __restrict__ mask, int Hi, int Wi, int H, int W,
                                                             int* __restrict__ ent_q, float* __restrict__ ent_w) {
    const int HW = H * W, HWi = Hi * Wi;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += gridDim.x * blockDim.x) {
        const warp::Taps<float> t = warp::make_taps<float>(__ldg(grid + p), __ldg(grid + HW + p), Hi, Wi);
        const float wx0 = 1.f - t.wx1, wy0 = 1.f - t.wy1;
        const float m = mask ? __ldg(mask + p) : 1.f;
        const bool on = !(mask && m == 0.f);
        const int o00 = t.y0 * Wi + t.x0;
        const bool v[4] = {t.vy0 && t.vx0, t.vy0 && t.vx1, t.vy1 && t.vx0, t.vy1 && t.vx1};
        const int q[4] = {o00, o00 + 1, o00 + Wi, o00 + Wi + 1};
        const float w[4] = {wx0 * wy0, t.wx1 * wy0, wx0 * t.wy1, t.wx1 * t.wy1};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const bool ok = on && v[k];
            ent_q[k * HW + p] = ok ? q[k] : HWi;
            ent_w[k * HW + p] = ok ? w[k] : 0.f;
        }
    }
}

__global__ void __launch_bounds__(kThreads) grid_sample_bwd_gather_kernel(const float* __restrict__ dout, const float* __restrict__ dout2, int64_t dout2_bs,
                                                                          const float* __restrict__ rough, int64_t rough_bs,
                                                                          const int* __restrict__ row_ptr, const int* __restrict__ ent_p,
                                                                          const float* __restrict__ ent_w, const float* __restrict__ ent_m, int C, int HWi,
                                                                          int HW, const float* __restrict__ xclamp, float lo, float hi,
                                                                          float* __restrict__ dimg, float* __restrict__ sq, float* __restrict__ partial,
                                                                          unsigned* __restrict__ counter) {
    __shared__ float red[32];
    __shared__ bool is_last;
    const int b = blockIdx.y;
    const float* db = dout + (int64_t)b * C * HW;
    const float* d2 = dout2 ? dout2 + (int64_t)b * dout2_bs : nullptr;
    const float* rb = dout2 ? rough + (int64_t)b * rough_bs : nullptr;
    float s[1] = {0.f};
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < HWi; q += gridDim.x * blockDim.x) {
        float acc[3] = {0.f, 0.f, 0.f};
        const int e0 = __ldg(row_ptr + q), e1 = __ldg(row_ptr + q + 1);
        for (int e = e0; e < e1; ++e) {
            const int p = __ldg(ent_p + e);
            const float w = __ldg(ent_w + e), m = __ldg(ent_m + e);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                if (c < C) {
                    float d = __ldg(db + (int64_t)c * HW + p);
                    if (d2) d += __ldg(d2 + (int64_t)c * HW + p) * __ldg(rb + (int64_t)c * HW + p);
                    d *= m;                                     // same factor order as the scatter kernel: ((dout + dout2 * rough) * mask) * weight
                    acc[c] = fmaf(d, w, acc[c]);
                }
            }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            if (c < C) {
                const int64_t o = ((int64_t)b * C + c) * HWi + q;
                dimg[o] = acc[c];
                float v = acc[c];
                if (xclamp) { const float x = __ldg(xclamp + o); if (!(x >= lo && x <= hi)) v = 0.f; }
                s[0] = fmaf(v, v, s[0]);
            }
        }
    }
    if (sq) row_reduce_finish<1>(s, b, partial, counter, sq, red, &is_last);
}

// Tiled form of the same gather.  The plain kernel above reads, per entry, three scattered words of dout (+ dout2, rough): 4-8 cache lines per warp
// request, 95 us at B = 32 against 75 us for the atomics path.  Here a block owns a 32 x 32 tile of INPUT pixels; the output pixels that contribute to
// it lie in one compact rectangle (the warp is smooth), which the block first copies -- coalesced rows, dout + dout2 * rough combined -- into shared
// memory; every entry then reads shared memory through a tile-local index precomputed with the CSR (ent_l).  Same entries, same order, same
// arithmetic as the plain kernel: bit-identical results.
constexpr int kGT = 32;
__global__ void __launch_bounds__(256) grid_sample_bwd_gather_tiled_kernel(const float* __restrict__ dout, const float* __restrict__ dout2, int64_t dout2_bs,
                                                                           const float* __restrict__ rough, int64_t rough_bs, const int* __restrict__ row_ptr,
                                                                           const int* __restrict__ ent_l, const float* __restrict__ ent_w,
                                                                           const float* __restrict__ ent_m, const int4* __restrict__ boxes, int C, int Hi,
                                                                           int Wi, int W, int HW, int tiles_x, const float* __restrict__ xclamp, float lo,
                                                                           float hi, float* __restrict__ dimg, float* __restrict__ sq,
                                                                           float* __restrict__ partial, unsigned* __restrict__ counter) {
    extern __shared__ float s_reg[];                       // [C][h * w] of this tile's region
    __shared__ float red[32];
    __shared__ bool is_last;
    const int tile = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    const int4 bx = __ldg(boxes + tile);                   // y0, x0, h, w of the contributing output rectangle (h = 0: no contribution)
    const int n = bx.z * bx.w;
    const float* db = dout + (int64_t)b * C * HW;
    const float* d2 = dout2 ? dout2 + (int64_t)b * dout2_bs : nullptr;
    const float* rb = dout2 ? rough + (int64_t)b * rough_bs : nullptr;
    for (int idx = tid; idx < n; idx += 256) {
        const int ry = idx / bx.w, rx = idx - ry * bx.w;
        const int p = (bx.x + ry) * W + bx.y + rx;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            if (c < C) {
                float d = __ldg(db + (int64_t)c * HW + p);
                if (d2) d += __ldg(d2 + (int64_t)c * HW + p) * __ldg(rb + (int64_t)c * HW + p);
                s_reg[c * n + idx] = d;
            }
        }
    }
    __syncthreads();
    const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
    const int HWi = Hi * Wi;
    float s[1] = {0.f};
#pragma unroll
    for (int k = 0; k < (kGT * kGT) / 256; ++k) {
        const int l = tid + k * 256;
        const int qy = ty * kGT + (l >> 5), qx = tx * kGT + (l & 31);
        if (qy < Hi && qx < Wi) {
            const int q = qy * Wi + qx;
            float acc[3] = {0.f, 0.f, 0.f};
            const int e0 = __ldg(row_ptr + q), e1 = __ldg(row_ptr + q + 1);
            for (int e = e0; e < e1; ++e) {
                const int li = __ldg(ent_l + e);
                const float w = __ldg(ent_w + e), m = __ldg(ent_m + e);
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    if (c < C) {
                        float d = s_reg[c * n + li];
                        d *= m;                                 // same factor order as the plain gather and the scatter kernel
                        acc[c] = fmaf(d, w, acc[c]);
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                if (c < C) {
                    const int64_t o = ((int64_t)b * C + c) * HWi + q;
                    dimg[o] = acc[c];
                    float v = acc[c];
                    if (xclamp) { const float x = __ldg(xclamp + o); if (!(x >= lo && x <= hi)) v = 0.f; }
                    s[0] = fmaf(v, v, s[0]);
                }
            }
        }
    }
    if (sq) row_reduce_finish<1>(s, b, partial, counter, sq, red, &is_last);
}

inline int blocks_for(int64_t n, int cap_mult = 8) {
    int64_t g = (n + kThreads - 1) / kThreads;
    const int64_t cap = (int64_t)kNumSMs * cap_mult;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

extern "C" {

int spaa_tps_grid_fwd(const float* theta, const float* ctrl, int T, int H, int W, float* grid, spaa_stream_t stream) {
    return spaa_coarse_grid_fwd(nullptr, theta, ctrl, T, 0, 0, H, W, grid, stream);
}

int64_t spaa_coarse_grid_ws_bytes(int T, int H, int W) {
    (void)H; (void)W;
    return (int64_t)kGridBwdBlocks * (6 + (T + 2) * 2) * sizeof(float) + 16;
}

}  // extern "C"

namespace {

__global__ void __launch_bounds__(kThreads) coarse_grid_fwd_dev(const float* __restrict__ aff, GridParams gp, const float* __restrict__ theta,
                                                                 const float* __restrict__ ctrl, float* __restrict__ grid) {
    __shared__ float s_theta[(kMaxT + 2) * 2];
    __shared__ float s_ctrl[kMaxT * 2];
    __shared__ float s_aff[6];
    for (int i = threadIdx.x; i < (gp.T + 2) * 2; i += blockDim.x) s_theta[i] = theta[i];
    for (int i = threadIdx.x; i < gp.T * 2; i += blockDim.x) s_ctrl[i] = ctrl[i];
    if (threadIdx.x < 6) s_aff[threadIdx.x] = aff ? aff[threadIdx.x] : 0.f;
    __syncthreads();
    const int HW = gp.H * gp.W;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += gridDim.x * blockDim.x) {
        const int y = p / gp.W, x = p - y * gp.W;
        float gx, gy;
        if (aff) warp::coarse_grid_point<float>(s_aff, s_theta, s_ctrl, gp.T, gp.Hin, gp.Win, gp.H, gp.W, y, x, gx, gy);
        else warp::tps_point<float>(s_theta, s_ctrl, gp.T, gp.H, gp.W, y, x, gx, gy);
        grid[p] = gx;
        grid[HW + p] = gy;
    }
}

}  // namespace

extern "C" {

int spaa_coarse_grid_fwd(const float* affine, const float* theta, const float* ctrl, int T, int Hin, int Win, int H, int W, float* grid,
                         spaa_stream_t stream) {
    SPAA_CHECK_ARG(theta && ctrl && grid && T >= 2 && T <= kMaxT && H > 0 && W > 0, "spaa_coarse_grid_fwd: bad arguments (T<=%d)", kMaxT);
    SPAA_CHECK_ARG(!affine || (Hin > 1 && Win > 1), "spaa_coarse_grid_fwd: affine needs Hin,Win > 1");
    GridParams gp{T, Hin, Win, H, W};
    coarse_grid_fwd_dev<<<blocks_for((int64_t)H * W, 4), kThreads, 0, (cudaStream_t)stream>>>(affine, gp, theta, ctrl, grid);
    SPAA_CHECK_LAUNCH("spaa_coarse_grid_fwd");
    return SPAA_OK;
}

}  // extern "C"

extern "C" {

int spaa_coarse_grid_bwd(const float* affine, const float* theta, const float* ctrl, int T, int Hin, int Win, int H, int W,
                         const float* dgrid, float* daffine, float* dtheta, void* ws, spaa_stream_t stream) {
    SPAA_CHECK_ARG(theta && ctrl && dgrid && dtheta && ws && T >= 2 && T <= kMaxT && H > 0 && W > 0, "spaa_coarse_grid_bwd: bad arguments");
    SPAA_CHECK_ARG(!affine || (Hin > 1 && Win > 1 && daffine), "spaa_coarse_grid_bwd: affine needs Hin,Win > 1 and daffine");
    GridParams gp{T, Hin, Win, H, W};
    float* partial = (float*)ws;
    unsigned* counter = (unsigned*)(partial + (int64_t)kGridBwdBlocks * (6 + (T + 2) * 2));
    if (T <= 36) coarse_grid_bwd_kernel<36><<<kGridBwdBlocks, kThreads, 0, (cudaStream_t)stream>>>(affine, gp, theta, ctrl, dgrid, daffine, dtheta, partial, counter);
    else coarse_grid_bwd_kernel<kMaxT><<<kGridBwdBlocks, kThreads, 0, (cudaStream_t)stream>>>(affine, gp, theta, ctrl, dgrid, daffine, dtheta, partial, counter);
    SPAA_CHECK_LAUNCH("spaa_coarse_grid_bwd");
    return SPAA_OK;
}

int spaa_grid_finish_fwd(const float* coarse, const float* refine, float* fine, int64_t n, spaa_stream_t stream) {
    SPAA_CHECK_ARG(coarse && fine && n > 0, "spaa_grid_finish_fwd: bad arguments");
    grid_finish_fwd_kernel<<<blocks_for(n), kThreads, 0, (cudaStream_t)stream>>>(coarse, refine, fine, n);
    SPAA_CHECK_LAUNCH("spaa_grid_finish_fwd");
    return SPAA_OK;
}

int spaa_grid_finish_bwd(const float* coarse, const float* refine, const float* dfine, float* dsum, int64_t n, spaa_stream_t stream) {
    SPAA_CHECK_ARG(coarse && dfine && dsum && n > 0, "spaa_grid_finish_bwd: bad arguments");
    grid_finish_bwd_kernel<<<blocks_for(n), kThreads, 0, (cudaStream_t)stream>>>(coarse, refine, dfine, dsum, n);
    SPAA_CHECK_LAUNCH("spaa_grid_finish_bwd");
    return SPAA_OK;
}

int spaa_grid_sample_fwd(const float* img, int64_t B, int C, int Hi, int Wi, const float* grid, int64_t grid_bstride, int H, int W,
                         int clamp01, const float* mask, float* out, const float* rough, int64_t rough_bstride, float* out2,
                         int64_t out2_bstride, spaa_stream_t stream) {
    SPAA_CHECK_ARG(img && grid && out && B > 0 && B < 65536 && C > 0 && Hi > 1 && Wi > 1 && H > 0 && W > 0, "spaa_grid_sample_fwd: bad arguments");
    SPAA_CHECK_ARG((out2 == nullptr) == (rough == nullptr), "spaa_grid_sample_fwd: out2 and rough go together");
    dim3 g(blocks_for((int64_t)H * W, 4), (unsigned)B);
    grid_sample_fwd_kernel<<<g, kThreads, 0, (cudaStream_t)stream>>>(img, C, Hi, Wi, grid, grid_bstride, H, W, clamp01, mask, out, rough,
                                                                   rough_bstride, out2, out2_bstride);
    SPAA_CHECK_LAUNCH("spaa_grid_sample_fwd");
    return SPAA_OK;
}

int spaa_grid_sample_bwd_input(const float* dout, const float* dout2, int64_t dout2_bstride, const float* rough, int64_t rough_bstride,
                               const float* mask, const float* grid, int64_t grid_bstride, int64_t B, int C, int Hi, int Wi, int H, int W,
                               float* dimg, spaa_stream_t stream) {
    SPAA_CHECK_ARG(dout && grid && dimg && B > 0 && B < 65536 && C > 0 && Hi > 1 && Wi > 1, "spaa_grid_sample_bwd_input: bad arguments");
    SPAA_CHECK_ARG((dout2 == nullptr) || rough, "spaa_grid_sample_bwd_input: dout2 needs rough");
    dim3 g(blocks_for((int64_t)H * W, 4), (unsigned)B);
    grid_sample_bwd_input_kernel<<<g, kThreads, 0, (cudaStream_t)stream>>>(dout, dout2, dout2_bstride, rough, rough_bstride, mask, grid,
                                                                         grid_bstride, C, Hi, Wi, H, W, dimg);
    SPAA_CHECK_LAUNCH("spaa_grid_sample_bwd_input");
    return SPAA_OK;
}

int spaa_grid_sample_bwd_grid(const float* dout, const float* dout2, int64_t dout2_bstride, const float* rough, int64_t rough_bstride,
                              const float* mask, const float* img, int clamp01, const float* grid, int64_t grid_bstride, int64_t B, int C,
                              int Hi, int Wi, int H, int W, float* dgrid, spaa_stream_t stream) {
    SPAA_CHECK_ARG(dout && img && grid && dgrid && B > 0 && B < 65536 && C > 0 && Hi > 1 && Wi > 1, "spaa_grid_sample_bwd_grid: bad arguments");
    SPAA_CHECK_ARG((dout2 == nullptr) || rough, "spaa_grid_sample_bwd_grid: dout2 needs rough");
    int bchunk = (int)B;
    if (grid_bstride == 0 && B >= 8) {
        bchunk = 4;
        if (cudaMemsetAsync(dgrid, 0, (size_t)2 * H * W * sizeof(float), (cudaStream_t)stream) != cudaSuccess) {
            set_last_error("spaa_grid_sample_bwd_grid: cudaMemsetAsync failed");
            return SPAA_ERR_CUDA;
        }
    }
    dim3 g(blocks_for((int64_t)H * W, 4), grid_bstride == 0 ? (unsigned)((B + bchunk - 1) / bchunk) : (unsigned)B);
    grid_sample_bwd_grid_kernel<<<g, kThreads, 0, (cudaStream_t)stream>>>(dout, dout2, dout2_bstride, rough, rough_bstride, mask, img, clamp01,
                                                                        grid, grid_bstride, (int)B, C, Hi, Wi, H, W, dgrid, bchunk);
    SPAA_CHECK_LAUNCH("spaa_grid_sample_bwd_grid");
    return SPAA_OK;
}

int spaa_grid_sample_fwd_packed(const float* img, int64_t B, int Hi, int Wi, const float* grid, int64_t grid_bstride, int H, int W, int clamp01,
                                const float* mask, const float* rough, int64_t rough_bstride, void* out16, int dtype, spaa_stream_t stream) {
    SPAA_CHECK_ARG(img && grid && out16 && B > 0 && B < 65536 && Hi > 1 && Wi > 1 && H > 0 && W > 0 && (dtype == 1 || dtype == 2),
                   "spaa_grid_sample_fwd_packed: bad arguments");
    dim3 g(blocks_for((int64_t)H * W, 4), (unsigned)B);
    if (dtype == 2)
        grid_sample_fwd_packed_kernel<true><<<g, kThreads, 0, (cudaStream_t)stream>>>(img, Hi, Wi, grid, grid_bstride, H, W, clamp01, mask, rough, rough_bstride,
                                                                                    (uint4*)out16);
    else
        grid_sample_fwd_packed_kernel<false><<<g, kThreads, 0, (cudaStream_t)stream>>>(img, Hi, Wi, grid, grid_bstride, H, W, clamp01, mask, rough, rough_bstride,
                                                                                     (uint4*)out16);
    SPAA_CHECK_LAUNCH("spaa_grid_sample_fwd_packed");
    return SPAA_OK;
}

}  // extern "C"

extern "C" {

int spaa_warp_taps(const float* grid, const float* mask, int Hi, int Wi, int H, int W, int32_t* ent_q, float* ent_w, spaa_stream_t stream) {
    SPAA_CHECK_ARG(grid && ent_q && ent_w && Hi > 0 && Wi > 0 && H > 0 && W > 0 && (int64_t)Hi * Wi < (1ll << 30) && (int64_t)H * W < (1ll << 28),
                   "spaa_warp_taps: bad arguments");
    warp_taps_kernel<<<blocks_for((int64_t)H * W), kThreads, 0, (cudaStream_t)stream>>>(grid, mask, Hi, Wi, H, W, ent_q, ent_w);
    SPAA_CHECK_LAUNCH("spaa_warp_taps");
    return SPAA_OK;
}

int64_t spaa_grid_sample_bwd_gather_ws_bytes(int64_t B, int Hi, int Wi) {
    return row_reduce_ws_bytes(B, row_reduce_nblk((int64_t)Hi * Wi, 2048), 1);
}

int64_t spaa_grid_sample_bwd_gather_tiled_ws_bytes(int64_t B, int Hi, int Wi) {
    return row_reduce_ws_bytes(B, ((Hi + kGT - 1) / kGT) * ((Wi + kGT - 1) / kGT), 1);
}

int spaa_grid_sample_bwd_gather_tiled(const float* dout, const float* dout2, int64_t dout2_bstride, const float* rough, int64_t rough_bstride,
                                      const int32_t* row_ptr, const int32_t* ent_l, const float* ent_w, const float* ent_m, const int32_t* boxes,
                                      int max_region, int64_t B, int C, int Hi, int Wi, int H, int W, const float* x_for_clamp, float lo, float hi,
                                      float* dimg, float* sq, void* ws, spaa_stream_t stream) {
    SPAA_CHECK_ARG(dout && row_ptr && ent_l && ent_w && ent_m && boxes && dimg && B > 0 && B < 65536 && C >= 1 && C <= 3 && Hi > 0 && Wi > 0 && H > 0 && W > 0,
                   "spaa_grid_sample_bwd_gather_tiled: bad arguments");
    SPAA_CHECK_ARG((dout2 == nullptr) || rough, "spaa_grid_sample_bwd_gather_tiled: dout2 needs rough");
    SPAA_CHECK_ARG((sq == nullptr) || ws, "spaa_grid_sample_bwd_gather_tiled: sq needs the workspace");
    const size_t smem = (size_t)C * (size_t)max_region * sizeof(float);
    SPAA_CHECK_ARG(max_region > 0 && smem <= 160 * 1024, "spaa_grid_sample_bwd_gather_tiled: a tile's source region does not fit in shared memory (use the plain gather)");
    static SmemOptIn opt;
    if (smem > 48 * 1024 && !opt.ensure(grid_sample_bwd_gather_tiled_kernel, smem)) {
        set_last_error("spaa_grid_sample_bwd_gather_tiled: cannot reserve %zu bytes of shared memory", smem);
        return SPAA_ERR_CUDA;
    }
    const int tiles_x = (Wi + kGT - 1) / kGT, tiles_y = (Hi + kGT - 1) / kGT;
    float* partial = (float*)ws;
    unsigned* counter = ws ? (unsigned*)(partial + B * tiles_x * tiles_y) : nullptr;
    grid_sample_bwd_gather_tiled_kernel<<<dim3((unsigned)(tiles_x * tiles_y), (unsigned)B), 256, smem, (cudaStream_t)stream>>>(
        dout, dout2, dout2_bstride, rough, rough_bstride, row_ptr, ent_l, ent_w, ent_m, (const int4*)boxes, C, Hi, Wi, W, H * W, tiles_x, x_for_clamp, lo, hi,
        dimg, sq, partial, counter);
    SPAA_CHECK_LAUNCH("spaa_grid_sample_bwd_gather_tiled");
    return SPAA_OK;
}

int spaa_grid_sample_bwd_gather(const float* dout, const float* dout2, int64_t dout2_bstride, const float* rough, int64_t rough_bstride,
                                const int32_t* row_ptr, const int32_t* ent_p, const float* ent_w, const float* ent_m, int64_t B, int C, int Hi, int Wi,
                                int H, int W, const float* x_for_clamp, float lo, float hi, float* dimg, float* sq, void* ws, spaa_stream_t stream) {
    SPAA_CHECK_ARG(dout && row_ptr && ent_p && ent_w && ent_m && dimg && B > 0 && B < 65536 && C >= 1 && C <= 3 && Hi > 0 && Wi > 0 && H > 0 && W > 0,
                   "spaa_grid_sample_bwd_gather: bad arguments");
    SPAA_CHECK_ARG((dout2 == nullptr) || rough, "spaa_grid_sample_bwd_gather: dout2 needs rough");
    SPAA_CHECK_ARG((sq == nullptr) || ws, "spaa_grid_sample_bwd_gather: sq needs the workspace");
    const int HWi = Hi * Wi;
    const int nblk = row_reduce_nblk(HWi, 2048);
    float* partial = (float*)ws;
    unsigned* counter = ws ? (unsigned*)(partial + B * nblk) : nullptr;
    grid_sample_bwd_gather_kernel<<<dim3(nblk, (unsigned)B), kThreads, 0, (cudaStream_t)stream>>>(dout, dout2, dout2_bstride, rough, rough_bstride, row_ptr, ent_p,
                                                                                                 ent_w, ent_m, C, HWi, H * W, x_for_clamp, lo, hi, dimg, sq,
                                                                                                 partial, counter);
    SPAA_CHECK_LAUNCH("spaa_grid_sample_bwd_gather");
    return SPAA_OK;
}

}  // extern "C"

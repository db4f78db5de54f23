// Classifier pre-processing, fused: centre crop -> area resize -> ImageNet normalise, forward and backward.
// Replaces /root/reference/src/python/classifier.py:55-59 (cc(), resize() = F.interpolate(mode='area') = adaptive average
// pooling, the per-sample normalize loop :51) and the autograd graph of those ops (slice backward = zero fill + copy,
// adaptive_avg_pool2d backward with atomics, two elementwise kernels): 5 + 6 launches become 1 + 1, and the result can be
// written NHWC so the cuDNN classifier needs no layout conversion of its input.
//   window of output cell o along an axis of n_in -> n_out cells: [floor(o*n_in/n_out), ceil((o+1)*n_in/n_out))  (ATen)
#include "common.cuh"
#include "../../include/spaa_b200.h"

using namespace spaa;

namespace {

constexpr int kThreads = 256;

struct PreP {
    int B, H, W;               // image planes [B,3,H,W]
    int top, left, ch, cw;     // crop rectangle
    int oh, ow;                // network input size
    int nhwc;                  // 1: out / dout are [B,oh,ow,3]; 0: [B,3,oh,ow]; 2: space-to-depth [B,oh/2+3,ow/2+3,16] (see below)
    float mean[3], inv_std[3];
};

// Layout 2 (kS2D): the 2x2 space-to-depth fold of the network input, zero-padded by kS2DPadLo cells before and kS2DPadHi cells after each axis, 16
// channels per cell: channel (dy*2 + dx)*3 + c = pixel (2*(I - kS2DPadLo) + dy, 2*(J - kS2DPadLo) + dx), channel c; channels 12..15 are zero.
// A 7x7 stride-2 pad-3 convolution over the image is exactly a 4x4 stride-1 pad-0 convolution over this tensor (tap k of the 7 = tap 2K + d - 1 of
// the fold; classifier.S2DStem re-arranges the frozen weights), a shape for which cuDNN has tensor-core NHWC kernels: 275 vs 532 us forward + input
// gradient at B=32 (tools/stem_probe.py).
constexpr int kS2D = 2, kS2DPadLo = 2, kS2DPadHi = 1, kS2DC = 16;

// 32-bit arithmetic: fill() bounds every extent by 2^15, so o * n_in < 2^30 (64-bit integer division costs ~100 instructions and
// the first version of the backward kernel executed ~40 of them per pixel: 99 us for 49 MB of traffic)
SPAA_D int win_lo(int o, int n_in, int n_out) { return (int)(((unsigned)o * (unsigned)n_in) / (unsigned)n_out); }
SPAA_D int win_hi(int o, int n_in, int n_out) { return (int)(((unsigned)(o + 1) * (unsigned)n_in + (unsigned)n_out - 1u) / (unsigned)n_out); }

// one thread = one cell of the folded tensor (border cells included: they are written as zeros every call)
__global__ void __launch_bounds__(kThreads) preprocess_fwd_s2d_kernel(const float* __restrict__ img, float4* __restrict__ out, const PreP p) {
    const int64_t plane = (int64_t)p.H * p.W;
    const int Hs = p.oh / 2 + kS2DPadLo + kS2DPadHi, Ws = p.ow / 2 + kS2DPadLo + kS2DPadHi;
    const int b = blockIdx.y;
    for (int rc = blockIdx.x * blockDim.x + threadIdx.x; rc < Hs * Ws; rc += gridDim.x * blockDim.x) {
        const int I = rc / Ws, J = rc - I * Ws;
        float v[kS2DC];
#pragma unroll
        for (int e = 0; e < kS2DC; ++e) v[e] = 0.f;
#pragma unroll
        for (int d = 0; d < 4; ++d) {
            const int oy = 2 * (I - kS2DPadLo) + (d >> 1), ox = 2 * (J - kS2DPadLo) + (d & 1);
            if (oy < 0 || oy >= p.oh || ox < 0 || ox >= p.ow) continue;
            const int y0 = win_lo(oy, p.ch, p.oh), y1 = win_hi(oy, p.ch, p.oh);
            const int x0 = win_lo(ox, p.cw, p.ow), x1 = win_hi(ox, p.cw, p.ow);
            const float inv_area = 1.f / (float)((y1 - y0) * (x1 - x0));
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float* src = img + ((int64_t)b * 3 + c) * plane + (int64_t)p.top * p.W + p.left;
                float sum = 0.f;
                for (int y = y0; y < y1; ++y)
                    for (int x = x0; x < x1; ++x) sum += __ldg(src + (int64_t)y * p.W + x);
                v[d * 3 + c] = (sum * inv_area - p.mean[c]) * p.inv_std[c];
            }
        }
        float4* o = out + ((int64_t)b * Hs * Ws + rc) * (kS2DC / 4);
#pragma unroll
        for (int e = 0; e < kS2DC / 4; ++e) o[e] = make_float4(v[4 * e], v[4 * e + 1], v[4 * e + 2], v[4 * e + 3]);
    }
}

__global__ void __launch_bounds__(kThreads) preprocess_fwd_kernel(const float* __restrict__ img, float* __restrict__ out, const PreP p) {
    const int64_t plane = (int64_t)p.H * p.W;
    const int cells = p.oh * p.ow;
    const int b = blockIdx.y;
    for (int rc = blockIdx.x * blockDim.x + threadIdx.x; rc < cells; rc += gridDim.x * blockDim.x) {
        const int64_t i = (int64_t)b * cells + rc;
        const int oy = rc / p.ow, ox = rc - oy * p.ow;
        const int y0 = win_lo(oy, p.ch, p.oh), y1 = win_hi(oy, p.ch, p.oh);
        const int x0 = win_lo(ox, p.cw, p.ow), x1 = win_hi(ox, p.cw, p.ow);
        const float inv_area = 1.f / (float)((y1 - y0) * (x1 - x0));
        float v[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float* src = img + ((int64_t)b * 3 + c) * plane + (int64_t)p.top * p.W + p.left;
            float s = 0.f;
            for (int y = y0; y < y1; ++y)
                for (int x = x0; x < x1; ++x) s += __ldg(src + (int64_t)y * p.W + x);
            v[c] = (s * inv_area - p.mean[c]) * p.inv_std[c];
        }
        if (p.nhwc) {
            float* o = out + i * 3;
            o[0] = v[0]; o[1] = v[1]; o[2] = v[2];
        } else {
            const int64_t op = (int64_t)p.oh * p.ow;
            float* o = out + (int64_t)b * 3 * op + (int64_t)oy * p.ow + ox;
            o[0] = v[0]; o[op] = v[1]; o[2 * op] = v[2];
        }
    }
}

// gather form of the adjoint: every image pixel sums the cells whose window covers it (deterministic, no atomics); pixels
// outside the crop get an explicit zero, so the caller needs no memset.
// A block owns a kBwdTX x kBwdTY tile of the image.  Its first kBwdTX + kBwdTY threads tabulate, for each column / row of the tile, the
// covering cells (first index, count <= kMaxCover, window sizes) -- the only integer divisions of the kernel; a pixel then needs two
// shared-memory lookups and count_y * count_x loads.  (The previous version did three 32-bit divisions and a 4 x 4 candidate search per
// pixel: 56 us for 49 MB of traffic, bound by instruction issue and the dependent search.)
constexpr int kBwdTX = 32, kBwdTY = 8, kMaxCover = 4;
struct Cover { int first, count, size[kMaxCover]; };

SPAA_D Cover cover_of(int c, int n_in, int n_out) {      // cells of an axis (n_in -> n_out) whose window contains crop coordinate c
    Cover v;
    v.first = 0; v.count = 0;
#pragma unroll
    for (int a = 0; a < kMaxCover; ++a) v.size[a] = 1;
    if (c < 0 || c >= n_in) return v;
    // candidate cells: around floor(c * n_out / n_in); windows are at most ceil(n_in/n_out)+1 wide
    const int oc = (int)(((unsigned)c * (unsigned)n_out) / (unsigned)n_in);
    const int lo = oc - 1 < 0 ? 0 : oc - 1, hi = oc + 2 > n_out - 1 ? n_out - 1 : oc + 2;
    for (int o = lo; o <= hi; ++o) {
        const int w0 = win_lo(o, n_in, n_out), w1 = win_hi(o, n_in, n_out);
        if (c < w0 || c >= w1) continue;
        if (v.count == 0) v.first = o;
        if (v.count < kMaxCover) v.size[v.count] = w1 - w0;       // covering cells are consecutive
        ++v.count;
    }
    return v;
}

__global__ void __launch_bounds__(kBwdTX * kBwdTY) preprocess_bwd_kernel(const float* __restrict__ dout, float* __restrict__ dimg, const PreP p) {
    __shared__ Cover s_cx[kBwdTX], s_cy[kBwdTY];
    const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * kBwdTX + tx;
    const int x = blockIdx.x * kBwdTX + tx, y = blockIdx.y * kBwdTY + ty, b = blockIdx.z;
    if (tid < kBwdTX) s_cx[tid] = cover_of((int)blockIdx.x * kBwdTX + tid - p.left, p.cw, p.ow);
    else if (tid < kBwdTX + kBwdTY) s_cy[tid - kBwdTX] = cover_of((int)blockIdx.y * kBwdTY + (tid - kBwdTX) - p.top, p.ch, p.oh);
    __syncthreads();
    if (x >= p.W || y >= p.H) return;
    const Cover& cy = s_cy[ty];              // read in place: dynamically indexed copies would live in local memory
    const Cover& cx = s_cx[tx];
    const int64_t op = (int64_t)p.oh * p.ow;
    const int Ws = p.ow / 2 + kS2DPadLo + kS2DPadHi, Hs = p.oh / 2 + kS2DPadLo + kS2DPadHi;
    float g[3] = {0.f, 0.f, 0.f};
    for (int a = 0; a < cy.count; ++a) {
        const int oy = cy.first + a;
        for (int c = 0; c < cx.count; ++c) {
            const int ox = cx.first + c;
            const float w = 1.f / (float)(cy.size[a] * cx.size[c]);
            if (p.nhwc == kS2D) {
                const float* d = dout + (((int64_t)b * Hs + (oy >> 1) + kS2DPadLo) * Ws + (ox >> 1) + kS2DPadLo) * kS2DC + ((oy & 1) * 2 + (ox & 1)) * 3;
                g[0] += w * __ldg(d); g[1] += w * __ldg(d + 1); g[2] += w * __ldg(d + 2);
            } else if (p.nhwc) {
                const float* d = dout + (((int64_t)b * p.oh + oy) * p.ow + ox) * 3;
                g[0] += w * __ldg(d); g[1] += w * __ldg(d + 1); g[2] += w * __ldg(d + 2);
            } else {
                const float* d = dout + (int64_t)b * 3 * op + (int64_t)oy * p.ow + ox;
                g[0] += w * __ldg(d); g[1] += w * __ldg(d + op); g[2] += w * __ldg(d + 2 * op);
            }
        }
    }
    const int64_t plane = (int64_t)p.H * p.W;
    float* o = dimg + (int64_t)b * 3 * plane + (int64_t)y * p.W + x;
    o[0] = g[0] * p.inv_std[0]; o[plane] = g[1] * p.inv_std[1]; o[2 * plane] = g[2] * p.inv_std[2];
}

int fill(PreP& p, int64_t B, int H, int W, int top, int left, int ch, int cw, int oh, int ow, const float* mean, const float* stdv, int nhwc) {
    if (!(B > 0 && B < (1 << 24) && H > 0 && W > 0 && H < (1 << 15) && W < (1 << 15) && oh < (1 << 15) && ow < (1 << 15) && ch > 0 && cw > 0 && oh > 0 && ow > 0 && top >= 0 && left >= 0 && top + ch <= H && left + cw <= W && mean && stdv)) return 0;
    // the candidate search of the backward covers cells centre-1 .. centre+2: shrink by at most 3x, enlarge by at most 2x
    if (ch > 3 * oh || cw > 3 * ow || oh > 2 * ch || ow > 2 * cw) return 0;
    if (nhwc < 0 || nhwc > kS2D || (nhwc == kS2D && ((oh | ow) & 1))) return 0;      // the fold needs even extents
    p.B = (int)B; p.H = H; p.W = W; p.top = top; p.left = left; p.ch = ch; p.cw = cw; p.oh = oh; p.ow = ow; p.nhwc = nhwc;
    for (int c = 0; c < 3; ++c) { p.mean[c] = mean[c]; p.inv_std[c] = 1.f / stdv[c]; }
    return 1;
}

}  // namespace

extern "C" {

int spaa_clf_preprocess_fwd(const float* img, int64_t B, int H, int W, int top, int left, int crop_h, int crop_w, int out_h, int out_w,
                            const float* host_mean3, const float* host_std3, int nhwc, float* out, spaa_stream_t stream) {
    PreP p;
    SPAA_CHECK_ARG(img && out && fill(p, B, H, W, top, left, crop_h, crop_w, out_h, out_w, host_mean3, host_std3, nhwc), "spaa_clf_preprocess_fwd: bad arguments");
    int64_t blocks = ((int64_t)out_h * out_w + kThreads - 1) / kThreads;
    if (blocks > 2 * kNumSMs) blocks = 2 * kNumSMs;
    SPAA_CHECK_ARG(B <= 65535, "spaa_clf_preprocess_fwd: batch too large");
    if (nhwc == kS2D) {
        SPAA_CHECK_ARG(((uintptr_t)out & 15) == 0, "spaa_clf_preprocess_fwd: misaligned output");
        const int64_t cells = (int64_t)(out_h / 2 + kS2DPadLo + kS2DPadHi) * (out_w / 2 + kS2DPadLo + kS2DPadHi);
        blocks = (cells + kThreads - 1) / kThreads;
        if (blocks > 2 * kNumSMs) blocks = 2 * kNumSMs;
        preprocess_fwd_s2d_kernel<<<dim3((unsigned)blocks, (unsigned)B), kThreads, 0, (cudaStream_t)stream>>>(img, (float4*)out, p);
    } else
        preprocess_fwd_kernel<<<dim3((unsigned)blocks, (unsigned)B), kThreads, 0, (cudaStream_t)stream>>>(img, out, p);
    SPAA_CHECK_LAUNCH("spaa_clf_preprocess_fwd");
    return SPAA_OK;
}

int spaa_clf_preprocess_bwd(const float* dout, int64_t B, int H, int W, int top, int left, int crop_h, int crop_w, int out_h, int out_w,
                            const float* host_std3, int nhwc, float* dimg, spaa_stream_t stream) {
    PreP p;
    const float zero3[3] = {0.f, 0.f, 0.f};
    SPAA_CHECK_ARG(dout && dimg && fill(p, B, H, W, top, left, crop_h, crop_w, out_h, out_w, zero3, host_std3, nhwc), "spaa_clf_preprocess_bwd: bad arguments");
    SPAA_CHECK_ARG(B <= 65535, "spaa_clf_preprocess_bwd: batch too large");
    const dim3 grid((unsigned)((W + kBwdTX - 1) / kBwdTX), (unsigned)((H + kBwdTY - 1) / kBwdTY), (unsigned)B);
    preprocess_bwd_kernel<<<grid, dim3(kBwdTX, kBwdTY), 0, (cudaStream_t)stream>>>(dout, dimg, p);
    SPAA_CHECK_LAUNCH("spaa_clf_preprocess_bwd");
    return SPAA_OK;
}

}  // extern "C"

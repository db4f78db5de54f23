// Classifier glue kernels.  (1) Fused [bias +] ReLU + max-pooling (forward and adjoint) on fp32 NHWC tensors -- the two steps that follow the first convolution of
// the external classifier (torchvision ResNet stem: conv1 -> bn1 -> relu -> maxpool 3x3 s2 p1; VGG: ReLU -> MaxPool2d(2,2); Inception:
// MaxPool2d(3, 2)), classifier.py:22-33 of the reference builds those networks.  They touch the largest activations of the whole attack
// iteration (resnet18, B = 32: 103 MB in, 26 MB out) and ATen's channels_last kernels for them (max_pool_forward_nhwc 127 us,
// max_pool_backward_nhwc 229 us, plus a ReLU pass and a threshold-backward pass over the 103 MB tensor; profiles/r1_final_launches.md)
// run at a sixth of the HBM roofline.  Here: one pass each way, the arg-max kept as ONE byte per element (the tap number inside the
// window, 255 = "no gradient": the ReLU was inactive), the adjoint in gather form (every input element written exactly once: no
// zero-fill, no atomics).  max(relu(x)) == relu(max(x)) and the first maximum in row-major window order is kept, as ATen does.
// HBM-bound: algorithmic bytes = 4*(in + out) + out (index bytes) forward; 4*(in + out) + out backward.
// (2) bias_act_kernel: y = relu(x + bias[c] + residual) in one pass -- after BatchNorm folding every cuDNN convolution of the private copy
// carries a bias, which PyTorch adds with a separate (strided, non-vectorised) elementwise kernel, followed by ATen's residual add and ReLU
// kernels: 44 elementwise launches per resnet18 forward (profiles/r1_final_launches.md) become 16.
#include "common.cuh"
#include "../../include/spaa_b200.h"

using namespace spaa;

namespace {

constexpr int kThreads = 256;

struct PoolGeom { int H, W, C4, k, s, p, Ho, Wo, nchunk; };

// grid = (ceil(Wo / blockDim.y) * nchunk, Ho, N): one thread = one output pixel x 4 channels
// K: compile-time window size (2, 3: loops unrolled, the window's loads are independent and in flight together) or 0 (runtime g.k)
template <bool RELU, int K>
__global__ void __launch_bounds__(kThreads) relu_maxpool_fwd_kernel(const float4* __restrict__ x, const float4* __restrict__ bias, float4* __restrict__ y,
                                                                    uint32_t* __restrict__ idx, PoolGeom g) {
    // blockDim = (bx, 256 / bx): x = 4-channel group inside a chunk of bx groups (consecutive lanes -> consecutive 16-byte words), y = pixel of the row.
    // No per-thread division (the first version spent its time in j / C4 and the window-range divisions: 92 us for the adjoint at resnet18 B=32).
    const int chunk = blockIdx.x % g.nchunk, pblk = blockIdx.x / g.nchunk;
    const int c = chunk * blockDim.x + threadIdx.x, ow = pblk * blockDim.y + threadIdx.y;
    if (ow >= g.Wo) return;
    const int oh = blockIdx.y, n = blockIdx.z;
    const int k = K ? K : g.k;
    const int h0 = oh * g.s - g.p, w0 = ow * g.s - g.p;
    const int r0 = h0 < 0 ? -h0 : 0, q0 = w0 < 0 ? -w0 : 0;
    const float ninf = __int_as_float(0xff800000);
    float m[4] = {ninf, ninf, ninf, ninf};
    uint32_t a[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) a[e] = (uint32_t)(r0 * k + q0);
    const float4* xn = x + (int64_t)n * g.H * g.W * g.C4 + c;
#pragma unroll
    for (int r = 0; r < k; ++r) {
        const int ih = h0 + r;
#pragma unroll
        for (int q = 0; q < k; ++q) {
            const int iw = w0 + q;
            if (ih < 0 || ih >= g.H || iw < 0 || iw >= g.W) continue;
            const float4 v4 = __ldg(xn + ((int64_t)ih * g.W + iw) * g.C4);
            const float v[4] = {v4.x, v4.y, v4.z, v4.w};
            const uint32_t tap = (uint32_t)(r * k + q);
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (v[e] > m[e] || v[e] != v[e]) { m[e] = v[e]; a[e] = tap; }
        }
    }
    if (bias) {                                  // max(x) + b == max(x + b) exactly (rounding is monotonic), and the arg-max does not move
        const float4 b4 = __ldg(bias + c);
        m[0] += b4.x; m[1] += b4.y; m[2] += b4.z; m[3] += b4.w;
    }
    if (RELU) {
#pragma unroll
        for (int e = 0; e < 4; ++e)
            if (!(m[e] > 0.f) && m[e] == m[e]) { m[e] = 0.f; a[e] = 255u; }     // inactive ReLU: value 0, no gradient (NaN propagates)
    }
    const int64_t o = (((int64_t)n * g.Ho + oh) * g.Wo + ow) * g.C4 + c;
    y[o] = make_float4(m[0], m[1], m[2], m[3]);
    idx[o] = a[0] | (a[1] << 8) | (a[2] << 16) | (a[3] << 24);
}

// grid = (ceil(W / blockDim.y) * nchunk, H, N): one thread = one INPUT pixel x 4 channels; sums the (<= ceil(k/s)^2) windows that selected it.
// K, S: compile-time window / stride (3,2 and 2,2: at most ceil(K/S)^2 windows, unrolled, their loads independent) or 0,0 (runtime loops).
template <int K, int S>
__global__ void __launch_bounds__(kThreads) relu_maxpool_bwd_kernel(const float4* __restrict__ dy, const uint32_t* __restrict__ idx, float4* __restrict__ dx, PoolGeom g) {
    const int chunk = blockIdx.x % g.nchunk, pblk = blockIdx.x / g.nchunk;
    const int c = chunk * blockDim.x + threadIdx.x, iw = pblk * blockDim.y + threadIdx.y;
    if (iw >= g.W) return;
    const int ih = blockIdx.y, n = blockIdx.z;
    const int k = K ? K : g.k, s = S ? S : g.s;
    // windows o with o*s - p <= i <= o*s - p + k - 1, i.e. tap r = i + p - o*s in [0, k): o = floor((i + p) / s) - a, a = 0 .. ceil(k/s) - 1
    const int oh_top = (int)((unsigned)(ih + g.p) / (unsigned)s), ow_top = (int)((unsigned)(iw + g.p) / (unsigned)s);
    const int nw = K ? (K + S - 1) / S : (k + s - 1) / s;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const int64_t img = (int64_t)n * g.Ho * g.Wo;
#pragma unroll
    for (int a = nw - 1; a >= 0; --a) {                      // ascending window order
        const int oh = oh_top - a, r = ih + g.p - oh * s;
        if (oh < 0 || oh >= g.Ho || r >= k) continue;
#pragma unroll
        for (int b = nw - 1; b >= 0; --b) {
            const int ow = ow_top - b, q = iw + g.p - ow * s;
            if (ow < 0 || ow >= g.Wo || q >= k) continue;
            const uint32_t tap = (uint32_t)(r * k + q);
            const int64_t o = (img + (int64_t)oh * g.Wo + ow) * g.C4 + c;
            const uint32_t u = __ldg(idx + o);
            const uint32_t hit = u ^ (tap * 0x01010101u);               // a zero byte marks a channel whose arg-max is this pixel
            if (((hit - 0x01010101u) & ~hit & 0x80808080u) == 0u) continue;
            const float4 d4 = __ldg(dy + o);
            if ((hit & 0x000000ffu) == 0u) acc[0] += d4.x;
            if ((hit & 0x0000ff00u) == 0u) acc[1] += d4.y;
            if ((hit & 0x00ff0000u) == 0u) acc[2] += d4.z;
            if ((hit & 0xff000000u) == 0u) acc[3] += d4.w;
        }
    }
    dx[(((int64_t)n * g.H + ih) * g.W + iw) * g.C4 + c] = make_float4(acc[0], acc[1], acc[2], acc[3]);
}

// thread-block shape for a row of `C4` 4-channel groups: bx = the largest power of two dividing C4 (<= 64), by = kThreads / bx pixels
inline dim3 pool_block(int C4, int* nchunk) {
    int bx = 1;
    while (bx < 64 && C4 % (bx * 2) == 0) bx *= 2;
    *nchunk = C4 / bx;
    return dim3((unsigned)bx, (unsigned)(kThreads / bx), 1);
}

// y = act(x + bias[c] + res) on a dense NHWC tensor (n4 float4 elements, C4 = C / 4): the glue between two cuDNN convolutions of the
// classifier's private copy (folded-BatchNorm bias, residual add, ReLU) in one pass instead of ATen's three.
template <bool RELU>
__global__ void __launch_bounds__(kThreads) bias_act_kernel(const float4* __restrict__ x, const float4* __restrict__ bias, const float4* __restrict__ res,
                                                            float4* __restrict__ y, uint32_t n4, uint32_t C4) {
    for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < n4; i += gridDim.x * kThreads) {
        float4 v = __ldg(x + i);
        if (bias) { const float4 b = __ldg(bias + i % C4); v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w; }
        if (res) { const float4 r = __ldg(res + i); v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w; }
        if (RELU) {               // clamp_min semantics: NaN propagates
            v.x = v.x < 0.f ? 0.f : v.x; v.y = v.y < 0.f ? 0.f : v.y; v.z = v.z < 0.f ? 0.f : v.z; v.w = v.w < 0.f ? 0.f : v.w;
        }
        y[i] = v;
    }
}

bool pool_args_ok(int64_t N, int H, int W, int C, int k, int stride, int pad, int Ho, int Wo) {
    if (N < 1 || N > 65535 || H < 1 || W < 1 || C < 4 || (C & 3) || k < 1 || k > 15 || stride < 1 || pad < 0 || 2 * pad > k) return false;
    if (Ho != (H + 2 * pad - k) / stride + 1 || Wo != (W + 2 * pad - k) / stride + 1 || Ho < 1 || Wo < 1 || H > 65535 || Ho > 65535) return false;
    return (int64_t)W * (C / 4) < (int64_t)1 << 30;
}

}  // namespace

extern "C" {

int spaa_relu_maxpool_nhwc_fwd(const float* x, const float* bias, int64_t N, int H, int W, int C, int k, int stride, int pad, int Ho, int Wo, int relu,
                               float* y, uint8_t* idx, spaa_stream_t stream) {
    SPAA_CHECK_ARG(x && y && idx && pool_args_ok(N, H, W, C, k, stride, pad, Ho, Wo), "spaa_relu_maxpool_nhwc_fwd: bad arguments");
    SPAA_CHECK_ARG(((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 15) == 0 && ((uintptr_t)idx & 3) == 0 && ((uintptr_t)bias & 15) == 0,
                   "spaa_relu_maxpool_nhwc_fwd: misaligned pointer");
    PoolGeom g{H, W, C / 4, k, stride, pad, Ho, Wo, 1};
    const dim3 block = pool_block(g.C4, &g.nchunk);
    const dim3 grid((unsigned)((Wo + (int)block.y - 1) / (int)block.y * g.nchunk), (unsigned)Ho, (unsigned)N);
#define SPAA_POOL_LAUNCH(R, K_) relu_maxpool_fwd_kernel<R, K_><<<grid, block, 0, (cudaStream_t)stream>>>((const float4*)x, (const float4*)bias, (float4*)y, (uint32_t*)idx, g)
    if (relu) { if (k == 3) SPAA_POOL_LAUNCH(true, 3); else if (k == 2) SPAA_POOL_LAUNCH(true, 2); else SPAA_POOL_LAUNCH(true, 0); }
    else { if (k == 3) SPAA_POOL_LAUNCH(false, 3); else if (k == 2) SPAA_POOL_LAUNCH(false, 2); else SPAA_POOL_LAUNCH(false, 0); }
#undef SPAA_POOL_LAUNCH
    SPAA_CHECK_LAUNCH("spaa_relu_maxpool_nhwc_fwd");
    return SPAA_OK;
}

int spaa_bias_act_nhwc(const float* x, const float* bias, const float* res, int64_t n, int C, int relu, float* y, spaa_stream_t stream) {
    SPAA_CHECK_ARG(x && y && n > 0 && C >= 4 && (C & 3) == 0 && n % C == 0 && n / 4 < ((int64_t)1 << 32), "spaa_bias_act_nhwc: bad arguments");
    SPAA_CHECK_ARG((((uintptr_t)x | (uintptr_t)y | (uintptr_t)bias | (uintptr_t)res) & 15) == 0, "spaa_bias_act_nhwc: misaligned pointer");
    const uint32_t n4 = (uint32_t)(n / 4);
    int64_t blocks = ((int64_t)n4 + kThreads - 1) / kThreads;
    const int64_t cap = (int64_t)kNumSMs * 32;
    if (blocks > cap) blocks = cap;
    if (relu)
        bias_act_kernel<true><<<(unsigned)blocks, kThreads, 0, (cudaStream_t)stream>>>((const float4*)x, (const float4*)bias, (const float4*)res, (float4*)y, n4, (uint32_t)(C / 4));
    else
        bias_act_kernel<false><<<(unsigned)blocks, kThreads, 0, (cudaStream_t)stream>>>((const float4*)x, (const float4*)bias, (const float4*)res, (float4*)y, n4, (uint32_t)(C / 4));
    SPAA_CHECK_LAUNCH("spaa_bias_act_nhwc");
    return SPAA_OK;
}

int spaa_relu_maxpool_nhwc_bwd(const float* dy, const uint8_t* idx, int64_t N, int H, int W, int C, int k, int stride, int pad, int Ho, int Wo,
                               float* dx, spaa_stream_t stream) {
    SPAA_CHECK_ARG(dy && dx && idx && pool_args_ok(N, H, W, C, k, stride, pad, Ho, Wo), "spaa_relu_maxpool_nhwc_bwd: bad arguments");
    SPAA_CHECK_ARG(((uintptr_t)dy & 15) == 0 && ((uintptr_t)dx & 15) == 0 && ((uintptr_t)idx & 3) == 0, "spaa_relu_maxpool_nhwc_bwd: misaligned pointer");
    PoolGeom g{H, W, C / 4, k, stride, pad, Ho, Wo, 1};
    const dim3 block = pool_block(g.C4, &g.nchunk);
    const dim3 grid((unsigned)((W + (int)block.y - 1) / (int)block.y * g.nchunk), (unsigned)H, (unsigned)N);
#define SPAA_POOLB_LAUNCH(K_, S_) relu_maxpool_bwd_kernel<K_, S_><<<grid, block, 0, (cudaStream_t)stream>>>((const float4*)dy, (const uint32_t*)idx, (float4*)dx, g)
    if (k == 3 && stride == 2) SPAA_POOLB_LAUNCH(3, 2);
    else if (k == 2 && stride == 2) SPAA_POOLB_LAUNCH(2, 2);
    else SPAA_POOLB_LAUNCH(0, 0);
#undef SPAA_POOLB_LAUNCH
    SPAA_CHECK_LAUNCH("spaa_relu_maxpool_nhwc_bwd");
    return SPAA_OK;
}

}  // extern "C"

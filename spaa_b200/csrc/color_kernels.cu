// Colour kernels: sRGB->Lab, CIEDE2000 variant, and the fused stealth-loss forward+backward.
// HBM-bound by design (one read of each operand, one write of the gradient); see DESIGN.md section 4.
#include "common.cuh"
#include "color_math.cuh"
#include "../../include/spaa_b200.h"

using namespace spaa;

namespace {

constexpr int kThreads = 256;

// Fast: the arithmetic of color_loss_kernel<., true> (the SAME operation sequence: a pixel equal to its reference gets bit-identical Lab values,
// hence dE = 0 and a zero gradient, as with the exact arithmetic).
template <bool Fast>
__global__ void __launch_bounds__(kThreads) lab_fwd_kernel(const float* __restrict__ rgb, float* __restrict__ lab, int64_t B, int64_t HW) {
    const int64_t total = B * HW;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = i / HW, p = i - b * HW;
        const float* s = rgb + b * 3 * HW + p;
        float L, A, Bv;
        if (Fast) { color::LabJac<float> J; color::rgb_to_lab_jac<float, true>(__ldg(s), __ldg(s + HW), __ldg(s + 2 * HW), L, A, Bv, J); }
        else color::rgb_to_lab<float>(__ldg(s), __ldg(s + HW), __ldg(s + 2 * HW), L, A, Bv);
        float* d = lab + b * 3 * HW + p;
        d[0] = L; d[HW] = A; d[2 * HW] = Bv;
    }
}

__global__ void __launch_bounds__(kThreads) lab_bwd_kernel(const float* __restrict__ rgb, const float* __restrict__ dlab, float* __restrict__ drgb,
                                                           int64_t B, int64_t HW) {
    const int64_t total = B * HW;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = i / HW, p = i - b * HW;
        const int64_t o = b * 3 * HW + p;
        float dr, dg, db;
        color::rgb_to_lab_bwd<float>(__ldg(rgb + o), __ldg(rgb + o + HW), __ldg(rgb + o + 2 * HW), __ldg(dlab + o), __ldg(dlab + o + HW),
                                     __ldg(dlab + o + 2 * HW), dr, dg, db);
        drgb[o] = dr; drgb[o + HW] = dg; drgb[o + 2 * HW] = db;
    }
}

__global__ void __launch_bounds__(kThreads) de_fwd_kernel(const float* __restrict__ l1, int64_t bs1, const float* __restrict__ l2, int64_t bs2,
                                                          float* __restrict__ de, int64_t B, int64_t HW) {
    const int64_t total = B * HW;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = i / HW, p = i - b * HW;
        const float* a = l1 + b * bs1 + p;
        const float* c = l2 + b * bs2 + p;
        de[i] = color::de2000<float, false>(__ldg(a), __ldg(a + HW), __ldg(a + 2 * HW), __ldg(c), __ldg(c + HW), __ldg(c + 2 * HW), nullptr, nullptr);
    }
}

__global__ void __launch_bounds__(kThreads) de_bwd_kernel(const float* __restrict__ l1, int64_t bs1, const float* __restrict__ l2, int64_t bs2,
                                                          const float* __restrict__ cot, float* __restrict__ d1, float* __restrict__ d2, int64_t B,
                                                          int64_t HW) {
    const int64_t total = B * HW;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = i / HW, p = i - b * HW;
        const float* a = l1 + b * bs1 + p;
        const float* c = l2 + b * bs2 + p;
        float g1[3], g2[3];
        color::de2000<float, true>(__ldg(a), __ldg(a + HW), __ldg(a + 2 * HW), __ldg(c), __ldg(c + HW), __ldg(c + 2 * HW), g1, g2);
        const float w = __ldg(cot + i);
        const int64_t o = b * 3 * HW + p;
        if (d1) { d1[o] = w * g1[0]; d1[o + HW] = w * g1[1]; d1[o + 2 * HW] = w * g1[2]; }
        if (d2) { d2[o] = w * g2[0]; d2[o + HW] = w * g2[1]; d2[o + 2 * HW] = w * g2[2]; }
    }
}

__host__ __device__ inline int color_nblk(int64_t HW) {
    int64_t n = (HW + 511) / 512;
    return (int)(n < 1 ? 1 : (n > 256 ? 256 : n));
}

// Fused: Lab(cam) -> dE vs ref_lab, channel-L2 vs ref_rgb, per-sample sums, gradient wrt cam.
// grid = (nblk, B).  Deterministic: per-block partials, last block of each sample adds them in fixed order.
// Fast: MUFU-approximation arithmetic (color_math.cuh, M<float, true>) for the 16-bit tensor-core modes.
template <bool WithGrad, bool Fast>
__global__ void __launch_bounds__(kThreads) color_loss_kernel(const float* __restrict__ cam, const float* __restrict__ ref_rgb,
                                                              const float* __restrict__ ref_lab, int64_t ref_bs, int64_t HW, int cam_is_lab2,
                                                              int de_weighting, float c_de, float c_l2, float* __restrict__ stats,
                                                              float* __restrict__ grad, float* __restrict__ partial, unsigned* __restrict__ counter) {
    __shared__ float red[32];
    __shared__ bool is_last;
    const int b = blockIdx.y, nblk = gridDim.x;
    const float* cb = cam + (int64_t)b * 3 * HW;
    const float* rr = ref_rgb + (int64_t)b * ref_bs;
    const float* rl = ref_lab + (int64_t)b * ref_bs;
    float* gb = WithGrad ? grad + (int64_t)b * 3 * HW : nullptr;
    float s_de = 0.f, s_l2 = 0.f, s_de2 = 0.f;
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < HW; p += (int64_t)nblk * blockDim.x) {
        const float r = __ldg(cb + p), g = __ldg(cb + HW + p), bl = __ldg(cb + 2 * HW + p);
        const float L0 = __ldg(rl + p), A0 = __ldg(rl + HW + p), B0 = __ldg(rl + 2 * HW + p);
        float L, A, Bv;
        color::LabJac<float> J;
        if (WithGrad) color::rgb_to_lab_jac<float, Fast>(r, g, bl, L, A, Bv, J);
        else color::rgb_to_lab<float>(r, g, bl, L, A, Bv);
        float gc[3], gr[3];
        float de;
        if (cam_is_lab2) de = color::de2000<float, WithGrad, Fast>(L0, A0, B0, L, A, Bv, gr, gc);
        else de = color::de2000<float, WithGrad, Fast>(L, A, Bv, L0, A0, B0, gc, gr);
        const float dr = r - __ldg(rr + p), dg = g - __ldg(rr + HW + p), db = bl - __ldg(rr + 2 * HW + p);
        const float nrm = color::M<float, Fast>::sqrt_(dr * dr + dg * dg + db * db);
        s_de += de; s_l2 += nrm; s_de2 += de * de;
        if (WithGrad) {
            const float w = c_de * (de_weighting ? de : 1.f);
            float gr_, gg_, gb_;
            color::lab_jac_bwd<float>(J, w * gc[0], w * gc[1], w * gc[2], gr_, gg_, gb_);
            const float inv = nrm > 0.f ? color::M<float, Fast>::div(c_l2, nrm) : 0.f;   // torch.norm sub-gradient 0 at 0
            gb[p] = gr_ + inv * dr;
            gb[HW + p] = gg_ + inv * dg;
            gb[2 * HW + p] = gb_ + inv * db;
        }
    }
    s_de = block_sum(s_de, red);
    s_l2 = block_sum(s_l2, red);
    s_de2 = block_sum(s_de2, red);
    float* pp = partial + ((int64_t)b * nblk + blockIdx.x) * 4;
    if (threadIdx.x == 0) {
        pp[0] = s_de; pp[1] = s_l2; pp[2] = s_de2; pp[3] = 0.f;
        __threadfence();
        const unsigned t = atomicAdd(counter + b, 1u);
        is_last = (t == (unsigned)nblk - 1);
    }
    __syncthreads();
    if (is_last && threadIdx.x < 32) {
        __threadfence();
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
        const volatile float* q = partial + (int64_t)b * nblk * 4;
        for (int i = threadIdx.x; i < nblk; i += 32) { a0 += q[i * 4]; a1 += q[i * 4 + 1]; a2 += q[i * 4 + 2]; }
        a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2);
        if (threadIdx.x == 0) {
            stats[b * 4] = a0; stats[b * 4 + 1] = a1; stats[b * 4 + 2] = a2; stats[b * 4 + 3] = 0.f;
            counter[b] = 0u;
        }
    }
}

inline int grid_for(int64_t total) {
    int64_t g = (total + kThreads - 1) / kThreads;
    const int64_t cap = (int64_t)kNumSMs * 16;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

extern "C" {

int spaa_rgb2lab_fwd(const float* rgb, float* lab, int64_t B, int64_t HW, int fast, spaa_stream_t stream) {
    SPAA_CHECK_ARG(rgb && lab && B > 0 && HW > 0, "spaa_rgb2lab_fwd: bad arguments");
    if (fast) lab_fwd_kernel<true><<<grid_for(B * HW), kThreads, 0, (cudaStream_t)stream>>>(rgb, lab, B, HW);
    else lab_fwd_kernel<false><<<grid_for(B * HW), kThreads, 0, (cudaStream_t)stream>>>(rgb, lab, B, HW);
    SPAA_CHECK_LAUNCH("spaa_rgb2lab_fwd");
    return SPAA_OK;
}

int spaa_rgb2lab_bwd(const float* rgb, const float* dlab, float* drgb, int64_t B, int64_t HW, spaa_stream_t stream) {
    SPAA_CHECK_ARG(rgb && dlab && drgb && B > 0 && HW > 0, "spaa_rgb2lab_bwd: bad arguments");
    lab_bwd_kernel<<<grid_for(B * HW), kThreads, 0, (cudaStream_t)stream>>>(rgb, dlab, drgb, B, HW);
    SPAA_CHECK_LAUNCH("spaa_rgb2lab_bwd");
    return SPAA_OK;
}

int spaa_de2000_fwd(const float* lab1, int64_t bs1, const float* lab2, int64_t bs2, float* de, int64_t B, int64_t HW, spaa_stream_t stream) {
    SPAA_CHECK_ARG(lab1 && lab2 && de && B > 0 && HW > 0, "spaa_de2000_fwd: bad arguments");
    de_fwd_kernel<<<grid_for(B * HW), kThreads, 0, (cudaStream_t)stream>>>(lab1, bs1, lab2, bs2, de, B, HW);
    SPAA_CHECK_LAUNCH("spaa_de2000_fwd");
    return SPAA_OK;
}

int spaa_de2000_bwd(const float* lab1, int64_t bs1, const float* lab2, int64_t bs2, const float* cot, float* dlab1, float* dlab2, int64_t B,
                    int64_t HW, spaa_stream_t stream) {
    SPAA_CHECK_ARG(lab1 && lab2 && cot && (dlab1 || dlab2) && B > 0 && HW > 0, "spaa_de2000_bwd: bad arguments");
    de_bwd_kernel<<<grid_for(B * HW), kThreads, 0, (cudaStream_t)stream>>>(lab1, bs1, lab2, bs2, cot, dlab1, dlab2, B, HW);
    SPAA_CHECK_LAUNCH("spaa_de2000_bwd");
    return SPAA_OK;
}

int64_t spaa_color_loss_ws_bytes(int64_t B, int64_t HW) { return B * (int64_t)color_nblk(HW) * 4 * sizeof(float) + B * sizeof(unsigned); }

int spaa_color_loss_fwd_bwd(const float* cam, const float* ref_rgb, const float* ref_lab, int64_t ref_bstride, int64_t B, int64_t HW,
                            int cam_is_lab2, int de_weighting, float c_de, float c_l2, int fast, float* stats, float* grad, void* ws,
                            spaa_stream_t stream) {
    SPAA_CHECK_ARG(cam && ref_rgb && ref_lab && stats && ws && B > 0 && B < 65536 && HW > 0, "spaa_color_loss_fwd_bwd: bad arguments");
    const int nblk = color_nblk(HW);
    float* partial = (float*)ws;
    unsigned* counter = (unsigned*)(partial + B * nblk * 4);
    dim3 grid(nblk, (unsigned)B);
    if (grad && fast)
        color_loss_kernel<true, true><<<grid, kThreads, 0, (cudaStream_t)stream>>>(cam, ref_rgb, ref_lab, ref_bstride, HW, cam_is_lab2, de_weighting,
                                                                                   c_de, c_l2, stats, grad, partial, counter);
    else if (grad)
        color_loss_kernel<true, false><<<grid, kThreads, 0, (cudaStream_t)stream>>>(cam, ref_rgb, ref_lab, ref_bstride, HW, cam_is_lab2, de_weighting,
                                                                                    c_de, c_l2, stats, grad, partial, counter);
    else
        color_loss_kernel<false, false><<<grid, kThreads, 0, (cudaStream_t)stream>>>(cam, ref_rgb, ref_lab, ref_bstride, HW, cam_is_lab2, de_weighting,
                                                                                     c_de, c_l2, stats, nullptr, partial, counter);
    SPAA_CHECK_LAUNCH("spaa_color_loss_fwd_bwd");
    return SPAA_OK;
}

}  // extern "C"

// Training-loss and optimiser kernels.
//  * fused L1 + MSE + (1 - SSIM) forward AND backward in one pass over pred/target
//    (replaces /root/reference/src/python/train_network.py:367-392 and pytorch_ssim/__init__.py:24-61: two replicate
//    pads, five depthwise 11x11 Gaussian convolutions, ~15 elementwise kernels and their autograd graph);
//  * one-launch Adam over a flat parameter buffer with per-segment lr / weight decay
//    (replaces the three optim.Adam instances of train_network.py:253-255).
//
// SSIM tile scheme (per CTA: one 32x32 output tile of one image plane, 256 threads, everything in shared memory):
//   load pred/target on the tile + 10 px halo with replicate-clamped coordinates
//   -> separable 11-tap Gaussian of {x, y, x^2, y^2, xy} on tile + 5 px
//   -> SSIM value and its partials A = dS/dmu_x, B = dS/dE[x^2], C = dS/dE[xy] there (zero outside the image)
//   -> separable ADJOINT of (replicate pad o Gaussian) applied to A, B, C back onto the tile
//   -> grad = adj(A) + 2 x adj(B) + y adj(C), plus the L1 / L2 terms.
// HBM traffic is the algorithmic read pred + read target + write grad (halo re-reads hit L2); the kernel is bound by
// FP32 FMA / shared-memory throughput (about 180 FMA per pixel), see DESIGN.md.
#include "common.cuh"
#include "../../include/spaa_b200.h"
#include <cmath>

using namespace spaa;

namespace {

constexpr int kThreads = 256;
constexpr int T = 32, R = 5, T1 = T + 2 * R, T2 = T + 4 * R;   // 32, 42, 52
constexpr int kSmemFloats = 2 * T2 * T2 + 5 * T2 * T1 + 3 * T1 * T1;

struct Gauss {
    float g[11];    // normalised window (pytorch_ssim/__init__.py:9-12)
    float cg[11];   // prefix sums, for the adjoint of replicate padding
};

// weight with which forward position p (moment centre) receives input pixel q along one axis of length n:
// sum_k g(k) [clamp(p + k, 0, n-1) == q]
SPAA_D float adj_w(const Gauss& G, int p, int q, int n) {
    const int d = q - p;
    if (q == 0) return (p <= R) ? G.cg[R - p] : 0.f;                  // all taps k <= -p fold onto the first pixel
    if (q == n - 1) return (n - 1 - p <= R) ? G.cg[R - (n - 1 - p)] : 0.f;
    return (d >= -R && d <= R) ? G.g[d + R] : 0.f;
}

__global__ void __launch_bounds__(kThreads) ssim_l1_kernel(const float* __restrict__ pred, const float* __restrict__ target, int H, int W, Gauss G,
                                                           float c_l1, float c_l2, float c_ssim, const float* __restrict__ cot_map,
                                                           float* __restrict__ ssim_map, float* __restrict__ grad, float* __restrict__ partial,
                                                           unsigned* __restrict__ counter, float* __restrict__ sums) {
    extern __shared__ __align__(16) float smem[];
    float* sX = smem;                       // [T2][T2]
    float* sY = sX + T2 * T2;               // [T2][T2]
    float* sH = sY + T2 * T2;               // [5][T2][T1]   (later aliased by sHA [3][T1][T])
    float* sA = sH + 5 * T2 * T1;           // [3][T1][T1]
    __shared__ float red[32];
    __shared__ bool is_last;

    const int tid = threadIdx.x;
    const int plane = blockIdx.z;
    const int ty0 = blockIdx.y * T, tx0 = blockIdx.x * T;
    const float* xp = pred + (int64_t)plane * H * W;
    const float* yp = target + (int64_t)plane * H * W;

    // ---- load with replicate-clamped coordinates ------------------------------------------------------
    for (int e = tid; e < T2 * T2; e += kThreads) {
        const int i = e / T2, j = e - i * T2;
        int gy = ty0 - 2 * R + i, gx = tx0 - 2 * R + j;
        gy = gy < 0 ? 0 : (gy > H - 1 ? H - 1 : gy);
        gx = gx < 0 ? 0 : (gx > W - 1 ? W - 1 : gx);
        sX[e] = __ldg(xp + (int64_t)gy * W + gx);
        sY[e] = __ldg(yp + (int64_t)gy * W + gx);
    }
    __syncthreads();
    // ---- horizontal Gaussian of the five products: rows of R2, columns of R1 -------------------------
    for (int e = tid; e < T2 * T1; e += kThreads) {
        const int i = e / T1, j = e - i * T1;
        float m0 = 0.f, m1 = 0.f, m2 = 0.f, m3 = 0.f, m4 = 0.f;
#pragma unroll
        for (int k = 0; k < 11; ++k) {
            const float x = sX[i * T2 + j + k], y = sY[i * T2 + j + k], g = G.g[k];
            m0 = fmaf(g, x, m0); m1 = fmaf(g, y, m1);
            m2 = fmaf(g, x * x, m2); m3 = fmaf(g, y * y, m3); m4 = fmaf(g, x * y, m4);
        }
        sH[0 * T2 * T1 + e] = m0; sH[1 * T2 * T1 + e] = m1; sH[2 * T2 * T1 + e] = m2;
        sH[3 * T2 * T1 + e] = m3; sH[4 * T2 * T1 + e] = m4;
    }
    __syncthreads();
    // ---- vertical Gaussian -> moments -> SSIM and its partial derivatives on R1 ----------------------
    float s_ssim = 0.f;
    const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
    for (int e = tid; e < T1 * T1; e += kThreads) {
        const int i = e / T1, j = e - i * T1;
        const int gy = ty0 - R + i, gx = tx0 - R + j;
        float A = 0.f, Bv = 0.f, Cv = 0.f;
        if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
            float m[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int k = 0; k < 11; ++k) {
                const float g = G.g[k];
#pragma unroll
                for (int q = 0; q < 5; ++q) m[q] = fmaf(g, sH[q * T2 * T1 + (i + k) * T1 + j], m[q]);
            }
            const float mu1 = m[0], mu2 = m[1];
            const float s11 = m[2] - mu1 * mu1, s22 = m[3] - mu2 * mu2, s12 = m[4] - mu1 * mu2;
            const float n1 = 2.f * mu1 * mu2 + C1, n2 = 2.f * s12 + C2;
            const float d1 = mu1 * mu1 + mu2 * mu2 + C1, d2 = s11 + s22 + C2;
            const float inv = 1.f / (d1 * d2);
            const float S = n1 * n2 * inv;
            const bool center = i >= R && i < R + T && j >= R && j < R + T;
            if (center) {
                s_ssim += S;
                if (ssim_map) ssim_map[(int64_t)plane * H * W + (int64_t)gy * W + gx] = S;
            }
            const float cot = cot_map ? __ldg(cot_map + (int64_t)plane * H * W + (int64_t)gy * W + gx) : c_ssim;
            A = cot * (2.f * mu2 * (n2 - n1) * inv - 2.f * mu1 * S * (d2 - d1) * inv);
            Bv = cot * (-S / d2);
            Cv = cot * (2.f * n1 * inv);
        }
        sA[0 * T1 * T1 + e] = A; sA[1 * T1 * T1 + e] = Bv; sA[2 * T1 * T1 + e] = Cv;
    }
    __syncthreads();
    float s_l1 = 0.f, s_l2 = 0.f;
    if (grad) {
        // ---- adjoint, horizontal: rows of R1, columns of the tile ------------------------------------
        float* sHA = sH;    // [3][T1][T]
        for (int e = tid; e < T1 * T; e += kThreads) {
            const int i = e / T, q = e - i * T;
            const int gq = tx0 + q;
            float a0 = 0.f, a1 = 0.f, a2 = 0.f;
            if (gq < W) {
#pragma unroll
                for (int k = 0; k < 11; ++k) {          // forward centre p = gq - 5 + k  <->  R1 column q + k
                    const float w = adj_w(G, gq - R + k, gq, W);
                    const int o = i * T1 + q + k;
                    a0 = fmaf(w, sA[o], a0); a1 = fmaf(w, sA[T1 * T1 + o], a1); a2 = fmaf(w, sA[2 * T1 * T1 + o], a2);
                }
            }
            sHA[e] = a0; sHA[T1 * T + e] = a1; sHA[2 * T1 * T + e] = a2;
        }
        __syncthreads();
        // ---- adjoint, vertical + assemble the gradient -----------------------------------------------
        for (int e = tid; e < T * T; e += kThreads) {
            const int i = e / T, j = e - i * T;
            const int gy = ty0 + i, gx = tx0 + j;
            if (gy >= H || gx >= W) continue;
            float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
            for (int k = 0; k < 11; ++k) {
                const float w = adj_w(G, gy - R + k, gy, H);
                const int o = (i + k) * T + j;
                a0 = fmaf(w, sHA[o], a0); a1 = fmaf(w, sHA[T1 * T + o], a1); a2 = fmaf(w, sHA[2 * T1 * T + o], a2);
            }
            const float x = sX[(i + 2 * R) * T2 + j + 2 * R], y = sY[(i + 2 * R) * T2 + j + 2 * R];
            const float d = x - y;
            s_l1 += fabsf(d); s_l2 += d * d;
            const float sg = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
            grad[(int64_t)plane * H * W + (int64_t)gy * W + gx] = a0 + 2.f * x * a1 + y * a2 + c_l1 * sg + c_l2 * 2.f * d;
        }
    } else {
        for (int e = tid; e < T * T; e += kThreads) {
            const int i = e / T, j = e - i * T;
            if (ty0 + i >= H || tx0 + j >= W) continue;
            const float d = sX[(i + 2 * R) * T2 + j + 2 * R] - sY[(i + 2 * R) * T2 + j + 2 * R];
            s_l1 += fabsf(d); s_l2 += d * d;
        }
    }
    // ---- deterministic reduction of the three sums ----------------------------------------------------
    const int64_t nblk = (int64_t)gridDim.x * gridDim.y * gridDim.z;
    const int64_t bid = ((int64_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    s_l1 = block_sum(s_l1, red);
    s_l2 = block_sum(s_l2, red);
    s_ssim = block_sum(s_ssim, red);
    if (tid == 0) {
        partial[bid * 4] = s_l1; partial[bid * 4 + 1] = s_l2; partial[bid * 4 + 2] = s_ssim;
        __threadfence();
        const unsigned t = atomicAdd(counter, 1u);
        is_last = (t == (unsigned)(nblk - 1));
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        const volatile float* q = partial;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
        for (int64_t k = tid; k < nblk; k += kThreads) { a0 += q[k * 4]; a1 += q[k * 4 + 1]; a2 += q[k * 4 + 2]; }
        a0 = block_sum(a0, red); a1 = block_sum(a1, red); a2 = block_sum(a2, red);
        if (tid == 0) { sums[0] = a0; sums[1] = a1; sums[2] = a2; sums[3] = 0.f; *counter = 0u; }
    }
}

__global__ void __launch_bounds__(kThreads) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                                        int64_t n, const int64_t* __restrict__ seg_end, const float* __restrict__ seg_lr,
                                                        const float* __restrict__ seg_wd, int nseg, float beta1, float beta2, float eps, float bc1,
                                                        float bc2_sqrt, float grad_scale, const float* __restrict__ dyn) {
    if (dyn) { bc1 = __ldg(dyn); bc2_sqrt = __ldg(dyn + 1); }          // per-step scalars kept on the device: the launch is CUDA-graph replayable
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int s = 0;
        while (s < nseg - 1 && i >= __ldg(seg_end + s)) ++s;
        const float lr = __ldg(seg_lr + s), wd = __ldg(seg_wd + s);
        const float pv = p[i];
        float gv = __ldg(g + i) * grad_scale;
        if (wd != 0.f) gv = fmaf(wd, pv, gv);                           // L2-in-gradient weight decay (torch.optim.Adam)
        const float mv = beta1 * m[i] + (1.f - beta1) * gv;
        const float vv = beta2 * v[i] + (1.f - beta2) * gv * gv;
        m[i] = mv; v[i] = vv;
        const float denom = sqrtf(vv) / bc2_sqrt + eps;
        p[i] = pv - (lr / bc1) * (mv / denom);
    }
}

}  // namespace

extern "C" {

int64_t spaa_ssim_l1_ws_bytes(int64_t N, int H, int W) {
    const int64_t nblk = N * ((H + T - 1) / T) * ((W + T - 1) / T);
    return nblk * 4 * (int64_t)sizeof(float) + 16;
}

int spaa_ssim_l1_fwd_bwd(const float* pred, const float* target, int64_t N, int H, int W, float w_l1, float w_l2, float w_ssim, const float* cot_map,
                         float* sums, float* ssim_map, float* grad, void* ws, spaa_stream_t stream) {
    SPAA_CHECK_ARG(pred && target && sums && ws && N > 0 && N < 65536 && H > 1 && W > 1, "spaa_ssim_l1_fwd_bwd: bad arguments");
    static bool attr_set = false;
    const int smem = kSmemFloats * (int)sizeof(float);
    if (!attr_set) {
        if (cudaFuncSetAttribute(ssim_l1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) {
            set_last_error("spaa_ssim_l1_fwd_bwd: cannot reserve %d bytes of shared memory", smem);
            return SPAA_ERR_CUDA;
        }
        attr_set = true;
    }
    Gauss G;
    float sum = 0.f;
    for (int i = 0; i < 11; ++i) { G.g[i] = (float)std::exp(-(double)((i - 5) * (i - 5)) / (2.0 * 1.5 * 1.5)); sum += G.g[i]; }
    for (int i = 0; i < 11; ++i) G.g[i] = G.g[i] / sum;
    float c = 0.f;
    for (int i = 0; i < 11; ++i) { c += G.g[i]; G.cg[i] = c; }
    const double numel = (double)N * H * W;
    const int gx = (W + T - 1) / T, gy = (H + T - 1) / T;
    float* partial = (float*)ws;
    unsigned* counter = (unsigned*)(partial + (int64_t)N * gx * gy * 4);
    ssim_l1_kernel<<<dim3(gx, gy, (unsigned)N), kThreads, smem, (cudaStream_t)stream>>>(pred, target, H, W, G, (float)(w_l1 / numel), (float)(w_l2 / numel),
                                                                                        (float)(-w_ssim / numel), cot_map, ssim_map, grad, partial, counter,
                                                                                        sums);
    SPAA_CHECK_LAUNCH("spaa_ssim_l1_fwd_bwd");
    return SPAA_OK;
}

int spaa_adam_step(float* param, const float* grad, float* m, float* v, int64_t n, const int64_t* seg_end, const float* seg_lr, const float* seg_wd,
                   int nseg, float beta1, float beta2, float eps, int step, float grad_scale, spaa_stream_t stream) {
    SPAA_CHECK_ARG(param && grad && m && v && seg_end && seg_lr && seg_wd && n > 0 && nseg > 0 && step > 0, "spaa_adam_step: bad arguments");
    const float bc1 = (float)(1.0 - std::pow((double)beta1, step));
    const float bc2s = (float)std::sqrt(1.0 - std::pow((double)beta2, step));
    int64_t blocks = (n + kThreads - 1) / kThreads;
    if (blocks > 4 * kNumSMs) blocks = 4 * kNumSMs;
    adam_kernel<<<(unsigned)blocks, kThreads, 0, (cudaStream_t)stream>>>(param, grad, m, v, n, seg_end, seg_lr, seg_wd, nseg, beta1, beta2, eps, bc1, bc2s,
                                                                        grad_scale, nullptr);
    SPAA_CHECK_LAUNCH("spaa_adam_step");
    return SPAA_OK;
}

int spaa_adam_step_dev(float* param, const float* grad, float* m, float* v, int64_t n, const int64_t* seg_end, const float* seg_lr, const float* seg_wd,
                       int nseg, float beta1, float beta2, float eps, const float* bias_corr2, float grad_scale, spaa_stream_t stream) {
    SPAA_CHECK_ARG(param && grad && m && v && seg_end && seg_lr && seg_wd && bias_corr2 && n > 0 && nseg > 0, "spaa_adam_step_dev: bad arguments");
    int64_t blocks = (n + kThreads - 1) / kThreads;
    if (blocks > 4 * kNumSMs) blocks = 4 * kNumSMs;
    adam_kernel<<<(unsigned)blocks, kThreads, 0, (cudaStream_t)stream>>>(param, grad, m, v, n, seg_end, seg_lr, seg_wd, nseg, beta1, beta2, eps, 1.f, 1.f,
                                                                        grad_scale, bias_corr2);
    SPAA_CHECK_LAUNCH("spaa_adam_step_dev");
    return SPAA_OK;
}

}  // extern "C"

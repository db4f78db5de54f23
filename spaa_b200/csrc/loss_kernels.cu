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
#include "tc_ptx.cuh"
#include "../../include/spaa_b200.h"
#include <cmath>

using namespace spaa;

namespace {

constexpr int kThreads = 256;
constexpr int T = 32, R = 5, T1 = T + 2 * R, T2 = T + 4 * R;   // 32, 42, 52
constexpr int kSmemFloats = 2 * T2 * T2 + 5 * T2 * T1;      // the dS/d(moment) maps alias the input tiles (3 * T1^2 <= 2 * T2^2): 65 KB, 3 CTAs per SM
static_assert(3 * T1 * T1 <= 2 * T2 * T2, "sA must fit in the sX + sY region it aliases");

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

__global__ void __launch_bounds__(kThreads, 3) ssim_l1_kernel(const float* __restrict__ pred, const float* __restrict__ target, int H, int W, Gauss G,
                                                           float c_l1, float c_l2, float c_ssim, const float* __restrict__ cot_map,
                                                           float* __restrict__ ssim_map, float* __restrict__ grad, float* __restrict__ partial,
                                                           unsigned* __restrict__ counter, float* __restrict__ sums) {
    extern __shared__ __align__(16) float smem[];
    float* sX = smem;                       // [T2][T2]
    float* sY = sX + T2 * T2;               // [T2][T2]
    float* sH = sY + T2 * T2;               // [5][T2][T1]   (later aliased by sHA [3][T1][T])
    float* sA = sX;                         // [3][T1][T1]   (aliases sX / sY, which are dead after the horizontal pass)
    __shared__ float red[32];
    __shared__ bool is_last;

    const int tid = threadIdx.x;
    const int plane = blockIdx.z;
    const int ty0 = blockIdx.y * T, tx0 = blockIdx.x * T;
    const float* xp = pred + (int64_t)plane * H * W;
    const float* yp = target + (int64_t)plane * H * W;

    float s_ssim = 0.f, s_l1 = 0.f, s_l2 = 0.f;
    // The L1 phase of train_pcnet (iterations <= 400, train_network.py:300-303) asks for the gradient of an L1 (+ L2) loss only: with no SSIM
    // weight, cotangent map or map output the windowed statistics are not needed at all and the tile reduces to a pointwise pass
    // (sums[2] is 0 in that case; the metric calls, which want the SSIM value, pass grad == nullptr and take the full path).
    const bool want_ssim = !(grad != nullptr && c_ssim == 0.f && cot_map == nullptr && ssim_map == nullptr);
    if (!want_ssim) {
        for (int e = tid; e < T * T; e += kThreads) {
            const int i = e / T, j = e - i * T;
            const int gy = ty0 + i, gx = tx0 + j;
            if (gy >= H || gx >= W) continue;
            const float d = __ldg(xp + (int64_t)gy * W + gx) - __ldg(yp + (int64_t)gy * W + gx);
            s_l1 += fabsf(d); s_l2 += d * d;
            const float sg = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
            grad[(int64_t)plane * H * W + (int64_t)gy * W + gx] = c_l1 * sg + c_l2 * 2.f * d;
        }
    } else {
        // ---- load with replicate-clamped coordinates: every thread issues all of its ~21 element copies asynchronously, then waits once
        // (a load -> store loop exposed one global-memory latency per iteration: 25 % of the kernel's stall samples) ----------------
        for (int e = tid; e < T2 * T2; e += kThreads) {
            const int i = e / T2, j = e - i * T2;
            int gy = ty0 - 2 * R + i, gx = tx0 - 2 * R + j;
            gy = gy < 0 ? 0 : (gy > H - 1 ? H - 1 : gy);
            gx = gx < 0 ? 0 : (gx > W - 1 ? W - 1 : gx);
            tc::cp_async4(tc::smem_u32(sX + e), xp + (int64_t)gy * W + gx);
            tc::cp_async4(tc::smem_u32(sY + e), yp + (int64_t)gy * W + gx);
        }
        tc::cp_async_commit();
        tc::cp_async_wait(0);
        __syncthreads();
        // All four separable passes are REGISTER-BLOCKED: a thread produces a strip of SW consecutive outputs along the filter axis from
        // SW + 10 inputs it reads once (sliding window), instead of 11 shared-memory reads per output and map.  The first version (one
        // output per thread per pass) was bound by shared-memory bandwidth: ~300 LDS per pixel, 363 us for 24 images.
        // ---- horizontal Gaussian of the five products: rows of R2, columns of R1 -------------------------
        constexpr int SW = 11, NS = (T1 + SW - 1) / SW;        // 4 strips of 11 cover the 42 columns / rows of R1
        for (int it = tid; it < T2 * NS; it += kThreads) {
            const int i = it / NS, sidx = it - i * NS, j0 = sidx * SW;        // lanes: 8 rows x 4 strips -> conflict-free reads
            float acc[5][SW];
    #pragma unroll
            for (int q = 0; q < 5; ++q)
    #pragma unroll
                for (int o = 0; o < SW; ++o) acc[q][o] = 0.f;
    #pragma unroll
            for (int t = 0; t < SW + 10; ++t) {
                const int jj = min(j0 + t, T2 - 1);                         // columns past the halo only feed outputs that are dropped
                const float x = sX[i * T2 + jj], y = sY[i * T2 + jj];
                const float xx = x * x, yy = y * y, xy = x * y;
    #pragma unroll
                for (int o = 0; o < SW; ++o) {
                    const int k = t - o;
                    if (k >= 0 && k < 11) {
                        const float g = G.g[k];
                        acc[0][o] = fmaf(g, x, acc[0][o]); acc[1][o] = fmaf(g, y, acc[1][o]);
                        acc[2][o] = fmaf(g, xx, acc[2][o]); acc[3][o] = fmaf(g, yy, acc[3][o]); acc[4][o] = fmaf(g, xy, acc[4][o]);
                    }
                }
            }
    #pragma unroll
            for (int o = 0; o < SW; ++o) {
                if (j0 + o < T1) {
    #pragma unroll
                    for (int q = 0; q < 5; ++q) sH[q * T2 * T1 + i * T1 + j0 + o] = acc[q][o];
                }
            }
        }
        __syncthreads();
        // ---- vertical Gaussian -> moments -> SSIM and its partial derivatives on R1 ----------------------
        const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
        for (int it = tid; it < NS * T1; it += kThreads) {
            const int sidx = it / T1, j = it - sidx * T1, i0 = sidx * SW;    // lanes along a row: conflict-free
            float m[5][SW];
    #pragma unroll
            for (int q = 0; q < 5; ++q)
    #pragma unroll
                for (int o = 0; o < SW; ++o) m[q][o] = 0.f;
    #pragma unroll
            for (int t = 0; t < SW + 10; ++t) {
                const int row = min(i0 + t, T2 - 1);
                float v[5];
    #pragma unroll
                for (int q = 0; q < 5; ++q) v[q] = sH[q * T2 * T1 + row * T1 + j];
    #pragma unroll
                for (int o = 0; o < SW; ++o) {
                    const int k = t - o;
                    if (k >= 0 && k < 11) {
                        const float g = G.g[k];
    #pragma unroll
                        for (int q = 0; q < 5; ++q) m[q][o] = fmaf(g, v[q], m[q][o]);
                    }
                }
            }
            const int gx = tx0 - R + j;
    #pragma unroll
            for (int o = 0; o < SW; ++o) {
                const int i = i0 + o;
                if (i >= T1) continue;
                const int gy = ty0 - R + i;
                float A = 0.f, Bv = 0.f, Cv = 0.f;
                if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
                    const float mu1 = m[0][o], mu2 = m[1][o];
                    const float s11 = m[2][o] - mu1 * mu1, s22 = m[3][o] - mu2 * mu2, s12 = m[4][o] - mu1 * mu2;
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
                const int e = i * T1 + j;
                sA[0 * T1 * T1 + e] = A; sA[1 * T1 * T1 + e] = Bv; sA[2 * T1 * T1 + e] = Cv;
            }
        }
        __syncthreads();
        if (grad) {
            // ---- adjoint, horizontal: rows of R1, columns of the tile ------------------------------------
            // interior image columns receive the mirrored window (= the window: it is symmetric); only the first / last image column
            // collect the taps folded onto them by the replicate padding (adj_w), recomputed on a slow path
            float* sHA = sH;    // [3][T1][T]
            constexpr int AW = 8;
            for (int it = tid; it < T1 * (T / AW); it += kThreads) {
                const int i = it / (T / AW), q0 = (it - i * (T / AW)) * AW;
                float a[3][AW];
    #pragma unroll
                for (int c = 0; c < 3; ++c)
    #pragma unroll
                    for (int o = 0; o < AW; ++o) a[c][o] = 0.f;
    #pragma unroll
                for (int t = 0; t < AW + 10; ++t) {
                    const int oo = i * T1 + q0 + t;                          // q0 + t <= 24 + 17 = 41 < T1
                    const float v0 = sA[oo], v1 = sA[T1 * T1 + oo], v2 = sA[2 * T1 * T1 + oo];
    #pragma unroll
                    for (int o = 0; o < AW; ++o) {
                        const int k = t - o;
                        if (k >= 0 && k < 11) {
                            const float w = G.g[10 - k];
                            a[0][o] = fmaf(w, v0, a[0][o]); a[1][o] = fmaf(w, v1, a[1][o]); a[2][o] = fmaf(w, v2, a[2][o]);
                        }
                    }
                }
    #pragma unroll
                for (int o = 0; o < AW; ++o) {
                    const int q = q0 + o, gq = tx0 + q;
                    float a0 = a[0][o], a1 = a[1][o], a2 = a[2][o];
                    if (gq >= W) { a0 = a1 = a2 = 0.f; }
                    else if (gq == 0 || gq == W - 1) {
                        a0 = a1 = a2 = 0.f;
                        for (int k = 0; k < 11; ++k) {          // forward centre p = gq - 5 + k  <->  R1 column q + k
                            const float w = adj_w(G, gq - R + k, gq, W);
                            const int oo = i * T1 + q + k;
                            a0 = fmaf(w, sA[oo], a0); a1 = fmaf(w, sA[T1 * T1 + oo], a1); a2 = fmaf(w, sA[2 * T1 * T1 + oo], a2);
                        }
                    }
                    const int e = i * T + q;
                    sHA[e] = a0; sHA[T1 * T + e] = a1; sHA[2 * T1 * T + e] = a2;
                }
            }
            __syncthreads();
            // ---- adjoint, vertical + assemble the gradient -----------------------------------------------
            constexpr int VW = 4;
            for (int it = tid; it < (T / VW) * T; it += kThreads) {
                const int sidx = it / T, j = it - sidx * T, i0 = sidx * VW;
                const int gx = tx0 + j;
                float a[3][VW];
    #pragma unroll
                for (int c = 0; c < 3; ++c)
    #pragma unroll
                    for (int o = 0; o < VW; ++o) a[c][o] = 0.f;
    #pragma unroll
                for (int t = 0; t < VW + 10; ++t) {
                    const int oo = (i0 + t) * T + j;                         // i0 + t <= 28 + 13 = 41 < T1
                    const float v0 = sHA[oo], v1 = sHA[T1 * T + oo], v2 = sHA[2 * T1 * T + oo];
    #pragma unroll
                    for (int o = 0; o < VW; ++o) {
                        const int k = t - o;
                        if (k >= 0 && k < 11) {
                            const float w = G.g[10 - k];
                            a[0][o] = fmaf(w, v0, a[0][o]); a[1][o] = fmaf(w, v1, a[1][o]); a[2][o] = fmaf(w, v2, a[2][o]);
                        }
                    }
                }
    #pragma unroll
                for (int o = 0; o < VW; ++o) {
                    const int i = i0 + o, gy = ty0 + i;
                    if (gy >= H || gx >= W) continue;
                    float a0 = a[0][o], a1 = a[1][o], a2 = a[2][o];
                    if (gy == 0 || gy == H - 1) {
                        a0 = a1 = a2 = 0.f;
                        for (int k = 0; k < 11; ++k) {
                            const float w = adj_w(G, gy - R + k, gy, H);
                            const int oo = (i + k) * T + j;
                            a0 = fmaf(w, sHA[oo], a0); a1 = fmaf(w, sHA[T1 * T + oo], a1); a2 = fmaf(w, sHA[2 * T1 * T + oo], a2);
                        }
                    }
                    const float x = __ldg(xp + (int64_t)gy * W + gx), y = __ldg(yp + (int64_t)gy * W + gx);      // (L2 hits: the tile was just read)
                    const float d = x - y;
                    s_l1 += fabsf(d); s_l2 += d * d;
                    const float sg = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
                    grad[(int64_t)plane * H * W + (int64_t)gy * W + gx] = a0 + 2.f * x * a1 + y * a2 + c_l1 * sg + c_l2 * 2.f * d;
                }
            }
        } else {
            for (int e = tid; e < T * T; e += kThreads) {
                const int i = e / T, j = e - i * T;
                if (ty0 + i >= H || tx0 + j >= W) continue;
                const float d = __ldg(xp + (int64_t)(ty0 + i) * W + tx0 + j) - __ldg(yp + (int64_t)(ty0 + i) * W + tx0 + j);
                s_l1 += fabsf(d); s_l2 += d * d;
            }
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
    static SmemOptIn opt;
    const int smem = kSmemFloats * (int)sizeof(float);
    if (!opt.ensure(ssim_l1_kernel, (size_t)smem)) {
        set_last_error("spaa_ssim_l1_fwd_bwd: cannot reserve %d bytes of shared memory", smem);
        return SPAA_ERR_CUDA;
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

// CUDA-core fp32 "gather convolution": one kernel family for nn.Conv2d / nn.ConvTranspose2d forward and both of
// their backward-data passes (see include/spaa_b200.h, spaa_conv_desc), plus the backward-weight kernel.
// This is the exact-fp32 path (FFMA accumulate, no tensor cores) used for the 1e-5 parity mode and for the
// small-channel, HBM-bound layers; the heavy layers' bf16 path is the tcgen05 kernel in conv_tc.cu.
//
// Replaces cudnnConvolutionForward / BackwardData / BackwardFilter as reached from
// /root/reference/src/python/models.py:18-46,130-139,223-252 (F.conv2d / F.conv_transpose2d + autograd).
//
// Implicit GEMM, M = B*Hout*Wout output pixels, N = Cout, K = taps x Cin (tap-major, BK input channels per step):
//   block tile BM x BN, 256 threads, thread tile TM x TN, double-buffered shared memory, register prefetch.
#include "common.cuh"
#include "../../include/spaa_b200.h"

using namespace spaa;

namespace {

constexpr int kThreads = 256;
constexpr int BK = 8;

struct ConvP {
    spaa_conv_desc d;
    const void* in;
    const float* w;
    const float* bias;
    const void* add;
    const void* mask;
    const void* mask2;
    void* out;
    void* out2;
};

SPAA_D float apply_mask(float v, float m, int mode) {
    switch (mode) {
        case SPAA_MASK_POS: return m > 0.f ? v : 0.f;
        case SPAA_MASK_LEAKY01: return m > 0.f ? v : 0.1f * v;
        case SPAA_MASK_OPEN01: return (m > 0.f && m < 1.f) ? v : 0.f;
        default: return v;
    }
}

// local pixel index inside the block tile owned by thread column tx, register slot i
template <int BM, int TM> SPAA_D int m_local(int tx, int i) {
    if (TM == 8) return (i >> 2) * (BM / 2) + tx * 4 + (i & 3);
    return tx * TM + i;
}

template <int BM, int BN, int TM, int TN, typename InT, typename OutT>
__global__ void __launch_bounds__(kThreads) conv_gather_kernel(const ConvP p) {
    static_assert((BM / TM) * (BN / TN) == kThreads, "tile/thread mismatch");
    constexpr int NTX = BM / TM;
    constexpr int A_PER_THREAD = BM * BK / kThreads;     // 4 (BM=128) or 8 (BM=256)
    constexpr int A_KSTEP = kThreads / BM;               // k rows covered per pass: 2 or 1
    constexpr int B_PER_THREAD = (BK * BN + kThreads - 1) / kThreads;
    __shared__ __align__(16) float As[2][BK][BM];
    __shared__ __align__(16) float Bs[2][BK][BN];

    const spaa_conv_desc& d = p.d;
    const int tid = threadIdx.x;
    const int tx = tid % NTX, ty = tid / NTX;
    const int HWo = d.Hout * d.Wout;
    const int64_t M = (int64_t)d.B * HWo;
    const int64_t m0 = (int64_t)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;

    // ---- this thread's gather pixel (A loader role) -------------------------------------------------
    const int am = tid % BM;
    const int ak0 = tid / BM;
    const int64_t aP = m0 + am;
    const bool a_ok = aP < M;
    int ab = 0, aoy = 0, aox = 0;
    if (a_ok) {
        ab = (int)(aP / HWo);
        const int r = (int)(aP - (int64_t)ab * HWo);
        aoy = r / d.Wout;
        aox = r - aoy * d.Wout;
    }
    const InT* in_b = (const InT*)p.in + (int64_t)ab * d.in_bs;
    const int ci_chunks = (d.Cin + BK - 1) / BK;
    const int nsteps = d.KH * d.KW * ci_chunks;

    float a_reg[A_PER_THREAD], b_reg[B_PER_THREAD];

    auto load_step = [&](int step) {
        const int tap = step / ci_chunks;
        const int ci0 = (step - tap * ci_chunks) * BK;
        const int r = tap / d.KW, s = tap - r * d.KW;
        int iy = aoy * d.stride + r - d.pad_h, ix = aox * d.stride + s - d.pad_w;
        bool v = a_ok && iy >= 0 && ix >= 0;
        if (d.up > 1) {
            v = v && (iy % d.up == 0) && (ix % d.up == 0);
            iy /= d.up; ix /= d.up;
        }
        v = v && iy < d.Hin && ix < d.Win;
        const InT* src = in_b + ((int64_t)iy * d.Win + ix) * d.in_ps;
#pragma unroll
        for (int j = 0; j < A_PER_THREAD; ++j) {
            const int ci = ci0 + ak0 + j * A_KSTEP;
            a_reg[j] = (v && ci < d.Cin) ? ld_f(src + (int64_t)ci * d.in_cs) : 0.f;
        }
        const int wtap = d.flip ? (d.KH - 1 - r) * d.KW + (d.KW - 1 - s) : tap;
        const float* wt = p.w + (int64_t)wtap * d.w_ts;
#pragma unroll
        for (int j = 0; j < B_PER_THREAD; ++j) {
            const int e = tid + j * kThreads;
            const int kk = e / BN, nn = e - kk * BN;
            const int ci = ci0 + kk, co = n0 + nn;
            b_reg[j] = (e < BK * BN && ci < d.Cin && co < d.Cout) ? __ldg(wt + (int64_t)ci * d.w_cis + (int64_t)co * d.w_cos) : 0.f;
        }
    };
    auto store_step = [&](int buf) {
#pragma unroll
        for (int j = 0; j < A_PER_THREAD; ++j) As[buf][ak0 + j * A_KSTEP][am] = a_reg[j];
#pragma unroll
        for (int j = 0; j < B_PER_THREAD; ++j) {
            const int e = tid + j * kThreads;
            if (e < BK * BN) Bs[buf][e / BN][e % BN] = b_reg[j];
        }
    };

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    load_step(0);
    store_step(0);
    __syncthreads();
    for (int step = 0; step < nsteps; ++step) {
        const int buf = step & 1;
        if (step + 1 < nsteps) load_step(step + 1);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[TM], b[TN];
            if constexpr (TM == 8) {
                const float4 v0 = *reinterpret_cast<const float4*>(&As[buf][kk][tx * 4]);
                const float4 v1 = *reinterpret_cast<const float4*>(&As[buf][kk][BM / 2 + tx * 4]);
                a[0] = v0.x; a[1] = v0.y; a[2] = v0.z; a[3] = v0.w;
                a[4] = v1.x; a[5] = v1.y; a[6] = v1.z; a[7] = v1.w;
            } else if constexpr (TM == 4) {
                const float4 v0 = *reinterpret_cast<const float4*>(&As[buf][kk][tx * 4]);
                a[0] = v0.x; a[1] = v0.y; a[2] = v0.z; a[3] = v0.w;
            } else {
#pragma unroll
                for (int i = 0; i < TM; ++i) a[i] = As[buf][kk][tx * TM + i];
            }
            {
                const float4 v = *reinterpret_cast<const float4*>(&Bs[buf][kk][ty * TN]);
                static_assert(TN == 4, "TN"); b[0] = v.x; b[1] = v.y; b[2] = v.z; b[3] = v.w;
            }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (step + 1 < nsteps) store_step(buf ^ 1);
        __syncthreads();
    }

    // ---- epilogue -----------------------------------------------------------------------------------
    const OutT* addp = (const OutT*)p.add;
    const OutT* maskp = (const OutT*)p.mask;
    const OutT* mask2p = (const OutT*)p.mask2;
    OutT* outp = (OutT*)p.out;
    OutT* out2p = (OutT*)p.out2;
    const int ef = d.epi_flags;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int64_t P = m0 + m_local<BM, TM>(tx, i);
        if (P >= M) continue;
        const int b = (int)(P / HWo);
        const int pix = (int)(P - (int64_t)b * HWo);
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int co = n0 + ty * TN + j;
            if (co >= d.Cout) continue;
            float v = acc[i][j];
            if (p.bias) v += __ldg(p.bias + co);
            float av = 0.f;
            if (addp) av = ld_f(addp + (int64_t)b * d.add_bs + (int64_t)pix * d.add_ps + (int64_t)co * d.add_cs);
            if (addp && !(ef & SPAA_EPI_ADD_AFTER_ACT)) v += av;
            if (ef & SPAA_EPI_RELU) v = fmaxf(v, 0.f);
            if (ef & SPAA_EPI_LEAKY01) v = v > 0.f ? v : 0.1f * v;
            if (ef & SPAA_EPI_CLAMP_MAX1) v = fminf(v, 1.f);
            if (addp && (ef & SPAA_EPI_ADD_AFTER_ACT)) v += av;
            const int64_t mo = (int64_t)b * d.mask_bs + (int64_t)pix * d.mask_ps + (int64_t)co * d.mask_cs;
            if (maskp) v = apply_mask(v, ld_f(maskp + mo), d.mask_mode);
            const int64_t oo = (int64_t)b * d.out_bs + (int64_t)pix * d.out_ps + (int64_t)co * d.out_cs;
            st_f(outp + oo, v);
            if (out2p) st_f(out2p + oo, ld_f(mask2p + mo) > 0.f ? v : 0.f);
        }
    }
}

template <int BM, int BN, int TM, int TN>
int launch_cfg(const ConvP& p, cudaStream_t st) {
    const spaa_conv_desc& d = p.d;
    const int64_t M = (int64_t)d.B * d.Hout * d.Wout;
    dim3 grid((unsigned)((M + BM - 1) / BM), (unsigned)((d.Cout + BN - 1) / BN));
#define SPAA_CONV_CASE(I, O, IT, OT) \
    if (d.in_dtype == I && d.out_dtype == O) conv_gather_kernel<BM, BN, TM, TN, IT, OT><<<grid, kThreads, 0, st>>>(p)
    SPAA_CONV_CASE(0, 0, float, float);
    SPAA_CONV_CASE(0, 1, float, __nv_bfloat16);
    SPAA_CONV_CASE(0, 2, float, __half);
    SPAA_CONV_CASE(1, 0, __nv_bfloat16, float);
    SPAA_CONV_CASE(1, 1, __nv_bfloat16, __nv_bfloat16);
    SPAA_CONV_CASE(2, 0, __half, float);
    SPAA_CONV_CASE(2, 2, __half, __half);
#undef SPAA_CONV_CASE
    return 0;
}

// --------------------------------------------------------------------------------------------------------------
// backward-weight: dW[tap][ci][co] += sum_pixels in(gathered at tap)[pix][ci] * dout[pix][co]
// block = (pixel split, tap x ci-tile, co-tile); 64 x 64 output tile, 32 pixels per step; fp32 atomics on exit.
// --------------------------------------------------------------------------------------------------------------
constexpr int WB = 64;     // ci tile and co tile
constexpr int WP = 32;     // pixels per step

struct WgP {
    spaa_conv_desc d;
    const void* in;
    const void* dout;
    float* dw;
    int64_t pix_per_split;
};

template <typename InT, typename OutT>
__global__ void __launch_bounds__(kThreads) conv_bwd_weight_kernel(const WgP p) {
    __shared__ __align__(16) float As[2][WP][WB];
    __shared__ __align__(16) float Bs[2][WP][WB];
    const spaa_conv_desc& d = p.d;
    const int tid = threadIdx.x;
    const int ci_tiles = (d.Cin + WB - 1) / WB;
    const int tap = blockIdx.y / ci_tiles;
    const int ci0 = (blockIdx.y - tap * ci_tiles) * WB;
    const int co0 = blockIdx.z * WB;
    const int r = tap / d.KW, s = tap - r * d.KW;
    const int HWo = d.Hout * d.Wout;
    const int64_t M = (int64_t)d.B * HWo;
    const int64_t P0 = (int64_t)blockIdx.x * p.pix_per_split;
    const int64_t P1 = (P0 + p.pix_per_split < M) ? P0 + p.pix_per_split : M;
    const int lp = tid % WP;          // pixel slot this thread loads
    const int lc = tid / WP;          // 0..7 : channel phase
    const int tx = tid % 16, ty = tid / 16;   // compute: ci = tx*4.., co = ty*4..
    constexpr int L = WB / (kThreads / WP);   // 8 loads per operand per thread
    float a_reg[L], b_reg[L];

    auto load_step = [&](int64_t Pbase) {
        const int64_t P = Pbase + lp;
        const bool ok = P < P1;
        int b = 0, oy = 0, ox = 0;
        if (ok) {
            b = (int)(P / HWo);
            const int rem = (int)(P - (int64_t)b * HWo);
            oy = rem / d.Wout; ox = rem - oy * d.Wout;
        }
        const int iy = oy * d.stride + r - d.pad_h, ix = ox * d.stride + s - d.pad_w;
        const bool v = ok && iy >= 0 && ix >= 0 && iy < d.Hin && ix < d.Win;
        const InT* src = (const InT*)p.in + (int64_t)b * d.in_bs + ((int64_t)iy * d.Win + ix) * d.in_ps;
        const OutT* dsrc = (const OutT*)p.dout + (int64_t)b * d.out_bs + ((int64_t)oy * d.Wout + ox) * d.out_ps;
#pragma unroll
        for (int j = 0; j < L; ++j) {
            const int c = lc + j * (kThreads / WP);
            a_reg[j] = (v && ci0 + c < d.Cin) ? ld_f(src + (int64_t)(ci0 + c) * d.in_cs) : 0.f;
            b_reg[j] = (ok && co0 + c < d.Cout) ? ld_f(dsrc + (int64_t)(co0 + c) * d.out_cs) : 0.f;
        }
    };
    auto store_step = [&](int buf) {
#pragma unroll
        for (int j = 0; j < L; ++j) {
            const int c = lc + j * (kThreads / WP);
            As[buf][lp][c] = a_reg[j];
            Bs[buf][lp][c] = b_reg[j];
        }
    };
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    if (P0 >= P1) return;
    load_step(P0);
    store_step(0);
    __syncthreads();
    int buf = 0;
    for (int64_t Pb = P0; Pb < P1; Pb += WP, buf ^= 1) {
        const bool more = Pb + WP < P1;
        if (more) load_step(Pb + WP);
#pragma unroll
        for (int kk = 0; kk < WP; ++kk) {
            const float4 a = *reinterpret_cast<const float4*>(&As[buf][kk][tx * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[buf][kk][ty * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (more) store_step(buf ^ 1);
        __syncthreads();
    }
    float* dwt = p.dw + (int64_t)tap * d.w_ts;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int ci = ci0 + tx * 4 + i;
        if (ci >= d.Cin) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int co = co0 + ty * 4 + j;
            if (co < d.Cout) atomicAdd(dwt + (int64_t)ci * d.w_cis + (int64_t)co * d.w_cos, acc[i][j]);
        }
    }
}

// dbias[co] += sum over pixels of dout[pix][co]; grid = (pixel splits, Cout): one block reduction per (split, channel)
template <typename OutT>
__global__ void __launch_bounds__(kThreads) bias_grad_kernel(const spaa_conv_desc d, const OutT* __restrict__ dout, float* __restrict__ dbias,
                                                             int64_t pix_per_split) {
    __shared__ float red[32];
    const int HWo = d.Hout * d.Wout;
    const int64_t M = (int64_t)d.B * HWo;
    const int64_t P0 = (int64_t)blockIdx.x * pix_per_split;
    const int64_t P1 = (P0 + pix_per_split < M) ? P0 + pix_per_split : M;
    const int co = blockIdx.y;
    float s = 0.f;
    for (int64_t P = P0 + threadIdx.x; P < P1; P += kThreads) {
        const int b = (int)(P / HWo);
        const int pix = (int)(P - (int64_t)b * HWo);
        s += ld_f(dout + (int64_t)b * d.out_bs + (int64_t)pix * d.out_ps + (int64_t)co * d.out_cs);
    }
    s = block_sum(s, red);
    if (threadIdx.x == 0) atomicAdd(dbias + co, s);
}

// Direct convolution for the 3 -> 3 channel full-resolution layers (skipConv1.*, models.py:243-247; forward and backward-data):
// fp32 NCHW planes, stride 1, 1x1 or 3x3.  One thread per pixel computes all output channels from weights in shared memory; the
// implicit-GEMM tile above spends a 256 x 4 tile on 3 channels and runs at 5 % of the HBM roofline for these layers.
template <int K>
__global__ void __launch_bounds__(256) conv_small_kernel(const ConvP p) {
    constexpr int C = 3, KK = K * K;
    __shared__ float sw[KK * C * C];
    __shared__ float sb[C];
    const spaa_conv_desc& d = p.d;
    for (int i = threadIdx.x; i < KK * C * C; i += blockDim.x) {
        const int tap = i / (C * C), ci = (i / C) % C, co = i % C;
        const int r = tap / K, s = tap - r * K;
        const int wtap = d.flip ? (K - 1 - r) * K + (K - 1 - s) : tap;
        sw[i] = (ci < d.Cin && co < d.Cout) ? __ldg(p.w + (int64_t)wtap * d.w_ts + (int64_t)ci * d.w_cis + (int64_t)co * d.w_cos) : 0.f;
    }
    if (threadIdx.x < C) sb[threadIdx.x] = (p.bias && threadIdx.x < d.Cout) ? __ldg(p.bias + threadIdx.x) : 0.f;
    __syncthreads();
    const int H = d.Hout, W = d.Wout, HW = H * W;
    const int64_t total = (int64_t)d.B * HW;
    const float* in = (const float*)p.in;
    const float* addp = (const float*)p.add;
    const float* maskp = (const float*)p.mask;
    const float* mask2p = (const float*)p.mask2;
    float* outp = (float*)p.out;
    float* out2p = (float*)p.out2;
    const int ef = d.epi_flags;
    for (int64_t P = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; P < total; P += (int64_t)gridDim.x * blockDim.x) {
        const int b = (int)(P / HW);
        const int pix = (int)(P - (int64_t)b * HW);
        const int oy = pix / W, ox = pix - oy * W;
        float acc[C] = {sb[0], sb[1], sb[2]};
#pragma unroll
        for (int r = 0; r < K; ++r)
#pragma unroll
            for (int s = 0; s < K; ++s) {
                const int iy = oy + r - d.pad_h, ix = ox + s - d.pad_w;
                if (iy < 0 || ix < 0 || iy >= d.Hin || ix >= d.Win) continue;
                const float* src = in + (int64_t)b * d.in_bs + (int64_t)iy * d.Win + ix;
#pragma unroll
                for (int ci = 0; ci < C; ++ci) {
                    if (ci < d.Cin) {
                        const float xv = __ldg(src + (int64_t)ci * d.in_cs);
#pragma unroll
                        for (int co = 0; co < C; ++co) acc[co] = fmaf(xv, sw[((r * K + s) * C + ci) * C + co], acc[co]);
                    }
                }
            }
#pragma unroll
        for (int co = 0; co < C; ++co) {
            if (co >= d.Cout) continue;
            float v = acc[co];
            float av = 0.f;
            if (addp) av = __ldg(addp + (int64_t)b * d.add_bs + (int64_t)pix * d.add_ps + (int64_t)co * d.add_cs);
            if (addp && !(ef & SPAA_EPI_ADD_AFTER_ACT)) v += av;
            if (ef & SPAA_EPI_RELU) v = fmaxf(v, 0.f);
            if (ef & SPAA_EPI_LEAKY01) v = v > 0.f ? v : 0.1f * v;
            if (ef & SPAA_EPI_CLAMP_MAX1) v = fminf(v, 1.f);
            if (addp && (ef & SPAA_EPI_ADD_AFTER_ACT)) v += av;
            const int64_t mo = (int64_t)b * d.mask_bs + (int64_t)pix * d.mask_ps + (int64_t)co * d.mask_cs;
            if (maskp) v = apply_mask(v, __ldg(maskp + mo), d.mask_mode);
            const int64_t oo = (int64_t)b * d.out_bs + (int64_t)pix * d.out_ps + (int64_t)co * d.out_cs;
            outp[oo] = v;
            if (out2p) out2p[oo] = __ldg(mask2p + mo) > 0.f ? v : 0.f;
        }
    }
}

// Backward-weight (+ bias) for the 3 -> 3 channel full-resolution layers (skipConv1.*, models.py:243-247): fp32 NCHW planes, stride 1.
// The generic kernel above wastes a 64 x 64 tile on a 3 x 3 channel block (9 ms per layer at B=24); here every thread
// walks pixels and keeps all CO*CI*K*K partial sums (<= 81) in registers; one shuffle/shared reduction and CO*CI*K*K
// atomics per block.  HBM traffic is the algorithmic read of x and dy (neighbour taps hit L1/L2).
template <int K>
__global__ void __launch_bounds__(256) wgrad_small_kernel(const spaa_conv_desc d, const float* __restrict__ x, const float* __restrict__ dy,
                                                          float* __restrict__ dw, float* __restrict__ dbias) {
    constexpr int C = 3, KK = K * K, NACC = C * C * KK + C;
    __shared__ float red[8][NACC];
    float acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = 0.f;
    const int H = d.Hout, W = d.Wout, HW = H * W;
    const int64_t total = (int64_t)d.B * HW;
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < total; p += (int64_t)gridDim.x * blockDim.x) {
        const int b = (int)(p / HW);
        const int pix = (int)(p - (int64_t)b * HW);
        const int oy = pix / W, ox = pix - oy * W;
        float g[C];
#pragma unroll
        for (int co = 0; co < C; ++co) g[co] = co < d.Cout ? __ldg(dy + (int64_t)b * d.out_bs + (int64_t)co * d.out_cs + pix) : 0.f;
#pragma unroll
        for (int co = 0; co < C; ++co) acc[C * C * KK + co] += g[co];
#pragma unroll
        for (int r = 0; r < K; ++r)
#pragma unroll
            for (int s = 0; s < K; ++s) {
                const int iy = oy + r - d.pad_h, ix = ox + s - d.pad_w;
                const bool v = iy >= 0 && ix >= 0 && iy < d.Hin && ix < d.Win;
#pragma unroll
                for (int ci = 0; ci < C; ++ci) {
                    const float xv = (v && ci < d.Cin) ? __ldg(x + (int64_t)b * d.in_bs + (int64_t)ci * d.in_cs + (int64_t)iy * d.Win + ix) : 0.f;
#pragma unroll
                    for (int co = 0; co < C; ++co) acc[(co * C + ci) * KK + r * K + s] = fmaf(xv, g[co], acc[(co * C + ci) * KK + r * K + s]);
                }
            }
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
        const float v = warp_sum(acc[i]);
        if (lane == 0) red[wid][i] = v;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < NACC; i += blockDim.x) {
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) v += red[w][i];
        if (i < C * C * KK) {
            const int co = i / (C * KK), ci = (i / KK) % C, tap = i % KK;
            if (co < d.Cout && ci < d.Cin) atomicAdd(dw + (int64_t)tap * d.w_ts + (int64_t)ci * d.w_cis + (int64_t)co * d.w_cos, v);
        } else if (dbias && i - C * C * KK < d.Cout) {
            atomicAdd(dbias + (i - C * C * KK), v);
        }
    }
}

}  // namespace

extern "C" {

int spaa_conv_fwd(const spaa_conv_desc* d, const void* in, const float* w, const float* bias, const void* add, const void* mask,
                  const void* mask2, void* out, void* out2, spaa_stream_t stream) {
    SPAA_CHECK_ARG(d && in && w && out, "spaa_conv_fwd: null argument");
    SPAA_CHECK_ARG(d->B > 0 && d->Cin > 0 && d->Cout > 0 && d->Hin > 0 && d->Win > 0 && d->Hout > 0 && d->Wout > 0 && d->KH > 0 && d->KW > 0 &&
                       d->stride > 0 && d->up > 0,
                   "spaa_conv_fwd: bad dimensions");
    SPAA_CHECK_ARG((unsigned)d->in_dtype < 3 && (unsigned)d->out_dtype < 3 && !(d->in_dtype && d->out_dtype && d->in_dtype != d->out_dtype),
                   "spaa_conv_fwd: dtype must be 0 (fp32), 1 (bf16) or 2 (fp16); mixed bf16/fp16 is not implemented on the CUDA-core path");
    SPAA_CHECK_ARG((out2 == nullptr) == (mask2 == nullptr), "spaa_conv_fwd: out2 and mask2 go together");
    SPAA_CHECK_ARG(d->mask_mode == SPAA_MASK_NONE || mask, "spaa_conv_fwd: mask_mode needs mask");
    ConvP p{*d, in, w, bias, add, mask, mask2, out, out2};
    cudaStream_t st = (cudaStream_t)stream;
    if (d->Cin <= 3 && d->Cout <= 3 && d->in_dtype == 0 && d->out_dtype == 0 && d->stride == 1 && d->up == 1 && d->in_ps == 1 &&
        d->Hin == d->Hout && d->Win == d->Wout && d->KH == d->KW && (d->KH == 1 || d->KH == 3)) {
        const int64_t M = (int64_t)d->B * d->Hout * d->Wout;
        int64_t blocks = (M + 255) / 256;
        if (blocks > 16 * kNumSMs) blocks = 16 * kNumSMs;
        if (d->KH == 1) conv_small_kernel<1><<<(unsigned)blocks, 256, 0, st>>>(p);
        else conv_small_kernel<3><<<(unsigned)blocks, 256, 0, st>>>(p);
        SPAA_CHECK_LAUNCH("spaa_conv_fwd(small)");
        return SPAA_OK;
    }
    if (d->Cout > 32) launch_cfg<128, 64, 8, 4>(p, st);
    else if (d->Cout > 16) launch_cfg<128, 32, 4, 4>(p, st);
    else if (d->Cout > 4) launch_cfg<256, 16, 4, 4>(p, st);
    else launch_cfg<256, 4, 1, 4>(p, st);
    SPAA_CHECK_LAUNCH("spaa_conv_fwd");
    return SPAA_OK;
}

int spaa_conv_bwd_weight(const spaa_conv_desc* d, const void* in, const void* dout, float* dw, float* dbias, spaa_stream_t stream) {
    SPAA_CHECK_ARG(d && in && dout && dw, "spaa_conv_bwd_weight: null argument");
    SPAA_CHECK_ARG(d->up == 1 && d->flip == 0, "spaa_conv_bwd_weight: describe the forward gather conv (up == 1, flip == 0)");
    SPAA_CHECK_ARG((unsigned)d->in_dtype < 3 && (unsigned)d->out_dtype < 3, "spaa_conv_bwd_weight: bad dtype");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t M = (int64_t)d->B * d->Hout * d->Wout;
    if (d->Cin <= 3 && d->Cout <= 3 && d->in_dtype == 0 && d->out_dtype == 0 && d->stride == 1 && d->in_ps == 1 && d->out_ps == 1 &&
        d->Hin == d->Hout && d->Win == d->Wout && d->KH == d->KW && (d->KH == 1 || d->KH == 3)) {
        int64_t blocks = (M + 255) / 256;
        if (blocks > 4 * kNumSMs) blocks = 4 * kNumSMs;
        if (d->KH == 1) wgrad_small_kernel<1><<<(unsigned)blocks, 256, 0, st>>>(*d, (const float*)in, (const float*)dout, dw, dbias);
        else wgrad_small_kernel<3><<<(unsigned)blocks, 256, 0, st>>>(*d, (const float*)in, (const float*)dout, dw, dbias);
        SPAA_CHECK_LAUNCH("spaa_conv_bwd_weight(small)");
        return SPAA_OK;
    }
    const int ci_tiles = (d->Cin + WB - 1) / WB, co_tiles = (d->Cout + WB - 1) / WB;
    const int64_t base_blocks = (int64_t)d->KH * d->KW * ci_tiles * co_tiles;
    int64_t splits = (4 * kNumSMs + base_blocks - 1) / base_blocks;      // aim for ~4 blocks per SM in total
    const int64_t max_splits = (M + 4 * WP - 1) / (4 * WP);
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    int64_t pps = (M + splits - 1) / splits;
    pps = (pps + WP - 1) / WP * WP;
    splits = (M + pps - 1) / pps;
    WgP p{*d, in, dout, dw, pps};
    dim3 grid((unsigned)splits, (unsigned)(d->KH * d->KW * ci_tiles), (unsigned)co_tiles);
#define SPAA_WG_CASE(I, O, IT, OT) \
    if (d->in_dtype == I && d->out_dtype == O) conv_bwd_weight_kernel<IT, OT><<<grid, kThreads, 0, st>>>(p)
    SPAA_WG_CASE(0, 0, float, float);
    SPAA_WG_CASE(0, 1, float, __nv_bfloat16);
    SPAA_WG_CASE(0, 2, float, __half);
    SPAA_WG_CASE(1, 0, __nv_bfloat16, float);
    SPAA_WG_CASE(1, 1, __nv_bfloat16, __nv_bfloat16);
    SPAA_WG_CASE(1, 2, __nv_bfloat16, __half);
    SPAA_WG_CASE(2, 0, __half, float);
    SPAA_WG_CASE(2, 1, __half, __nv_bfloat16);
    SPAA_WG_CASE(2, 2, __half, __half);
#undef SPAA_WG_CASE
    SPAA_CHECK_LAUNCH("spaa_conv_bwd_weight");
    if (dbias) {
        int64_t bsplits = (4 * kNumSMs + d->Cout - 1) / d->Cout;
        if (bsplits > (M + 1023) / 1024) bsplits = (M + 1023) / 1024;
        if (bsplits < 1) bsplits = 1;
        const int64_t bpps = (M + bsplits - 1) / bsplits;
        bsplits = (M + bpps - 1) / bpps;
        const dim3 bgrid((unsigned)bsplits, (unsigned)d->Cout);
        if (d->out_dtype == 0) bias_grad_kernel<float><<<bgrid, kThreads, 0, st>>>(*d, (const float*)dout, dbias, bpps);
        else if (d->out_dtype == 1) bias_grad_kernel<__nv_bfloat16><<<bgrid, kThreads, 0, st>>>(*d, (const __nv_bfloat16*)dout, dbias, bpps);
        else bias_grad_kernel<__half><<<bgrid, kThreads, 0, st>>>(*d, (const __half*)dout, dbias, bpps);
        SPAA_CHECK_LAUNCH("spaa_conv_bwd_weight(bias)");
    }
    return SPAA_OK;
}

int spaa_channel_sum(const void* x, int dtype, int64_t B, int C, int64_t HW, int64_t bs, int64_t ps, int64_t cs, float* out, spaa_stream_t stream) {
    SPAA_CHECK_ARG(x && out && B > 0 && C > 0 && HW > 0 && HW < (1ll << 31) && (unsigned)dtype < 3, "spaa_channel_sum: bad arguments");
    spaa_conv_desc d{};
    d.B = (int)B; d.Cout = C; d.Hout = 1; d.Wout = (int)HW;
    d.out_bs = bs; d.out_ps = ps; d.out_cs = cs;
    const int64_t M = B * HW;
    int64_t splits = (4 * kNumSMs + C - 1) / C;
    if (splits > (M + 1023) / 1024) splits = (M + 1023) / 1024;
    if (splits < 1) splits = 1;
    const int64_t pps = (M + splits - 1) / splits;
    splits = (M + pps - 1) / pps;
    const dim3 grid((unsigned)splits, (unsigned)C);
    if (dtype == 0) bias_grad_kernel<float><<<grid, kThreads, 0, (cudaStream_t)stream>>>(d, (const float*)x, out, pps);
    else if (dtype == 1) bias_grad_kernel<__nv_bfloat16><<<grid, kThreads, 0, (cudaStream_t)stream>>>(d, (const __nv_bfloat16*)x, out, pps);
    else bias_grad_kernel<__half><<<grid, kThreads, 0, (cudaStream_t)stream>>>(d, (const __half*)x, out, pps);
    SPAA_CHECK_LAUNCH("spaa_channel_sum");
    return SPAA_OK;
}

}  // extern "C"

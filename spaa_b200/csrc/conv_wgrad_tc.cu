// tcgen05 backward-weight kernel for the convolution stacks (training only).
//
//   dW[tap][cx][cy] += sum_{b,oy,ox} X[b, oy*s + r - pad, ox*s + q - pad, cx] * DY[b, oy, ox, cy]
//
// (X = the operand the forward pass GATHERS through the filter taps, DY = the pointwise one; for nn.Conv2d X is the layer
// input and DY the output gradient, for nn.ConvTranspose2d the roles swap -- see ops.conv_backward_weight).  Replaces
// cudnnConvolutionBackwardFilter as reached from the autograd graph of /root/reference/src/python/models.py:18-46,223-252
// in train_network.py:304-320.
//
// GEMM view: the contraction index K is the PIXEL, which is the slow (row) index of both 16-bit NHWC operands, so both MMA
// operands are MN-major: a TMA box [pixels][64 channels] lands in shared memory as 128-byte rows = exactly the canonical
// MN-major SWIZZLE_128B layout ((T,8,m),(8,k)):((1,T,LBO),(8T,SBO)) of cute/atom/mma_traits_sm100.hpp (rows = K, 8-row
// groups SBO apart, 64-channel atoms LBO apart).  D[M = 128 X-channels][N = Cy] accumulates in TMEM over ALL pixel tiles a
// CTA owns and is added to the fp32 gradient with atomics once, at the end.
//
//   * pixel tile = 16 x 8 output pixels = 8 MMA K-steps of 16 pixels; X is fetched once per tile as a halo box (same scheme
//     as conv_halo_kernel: the view of tap (r,q) is the same smem with the start address shifted by whole rows, SBO =
//     halo_w rows; stride-2 layers use the four input-parity planes);
//   * M = 128 rows: Cx >= 128 -> 128 channels of one tap (two 64-channel atoms, LBO = box stride);
//                   Cx == 64  -> TWO taps per MMA (atom 1 = the other tap's view: LBO = distance of the two start addresses);
//                   Cx <= 32  -> the taps of one filter row per MMA: atom k = the view shifted by k pixels (LBO = one row);
//   * one launch accumulates up to 512 / Cy (tap, 128-channel) pairs; ops.py issues one launch per such set.
#include "common.cuh"
#include "tc_ptx.cuh"
#include "../../include/spaa_b200.h"
#include <cstring>
#include <cstdint>
#include <cstdlib>

using namespace spaa;
using namespace spaa::tc;

namespace {

constexpr int kThreads = 192;
constexpr int WTH = 16, WTW = 8;
constexpr int kMaxGroups = 32;     // accumulators (tap x 128-channel chunk, or tap pair) of one layer
constexpr int kMaxSets = 16;       // a CTA owns one set: accumulators that fit in TMEM together and share one X chunk

struct WgGroup {
    uint32_t a_off16;          // start of the A view inside a stage, 16 B units (plane offset + row shift)
    uint32_t a_lbo16;          // distance between the 64/32/16-channel atoms of the 128 M rows, 16 B units
    int8_t tap[8];             // filter tap of each atom (row block), -1: ignore those rows
    int32_t cx0;               // first X channel of atom 0 (Cx >= 128: both atoms are consecutive channel blocks of one tap)
};
struct WgParams {
    int32_t B, Cx, Cy, Hy, Wy;
    int32_t stride, nplanes, halo_w, halo_h, org_x, org_y;
    int32_t tiles_x, tiles_y, total_tiles;
    int32_t tile_h, ksteps;                // pixel tile = tile_h x 8 pixels = ksteps MMA K-steps of 16 pixels (16 rows, or 8 when a 16-row stage would leave no room for a second one)
    int32_t cxb, cyb;                      // channels per X / DY box (64, 32 or 16)
    int32_t x_boxes, y_boxes;              // boxes per plane of X loaded per stage; boxes of DY
    int32_t nsets;
    int16_t set_first[kMaxSets], set_count[kMaxSets], set_ch0[kMaxSets];      // accumulators / first X channel of every set
    int32_t x_plane_bytes, x_bytes, y_box_bytes, stage_bytes, tx_bytes, nstages;
    int32_t max_groups, n_cols;            // most accumulators in a set; TMEM columns per accumulator (= Cy)
    int32_t atom_rows;                     // rows of D per atom (= cxb, or 64 when Cx >= 128)
    int32_t cx_real, cx_off, cy_real;      // padded operands: real X channels sit at [cx_off, cx_off + cx_real); real DY channels [0, cy_real)
    int32_t x_f16, y_f16;
    int32_t vec4;                          // 1: dw is the scratch layout (w_ys == 1, rows 16-byte aligned): flush with red.global.add.v4.f32
    int32_t debug_noflush;                 // timing experiments only ($SPAA_WGRAD_NOFLUSH=1): skip the atomic flush (results are wrong)
    int64_t w_ts, w_xs, w_ys;
    float* dw;
    WgGroup g[kMaxGroups];
};

SPAA_D uint64_t make_mn_desc(uint32_t addr16, uint32_t lbo16, uint32_t sbo16, uint32_t layout) {
    return (uint64_t)(addr16 & 0x3FFFu) | ((uint64_t)(lbo16 & 0x3FFFu) << 16) | ((uint64_t)(sbo16 & 0x3FFFu) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)layout << 61);
}
SPAA_HD uint32_t layout_for(int ch) { return ch == 64 ? 2u : (ch == 32 ? 4u : 6u); }

__global__ void __launch_bounds__(kThreads, 3) conv_wgrad_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_y,
                                                                    const __grid_constant__ WgParams P) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* full = (uint64_t*)(smem + (size_t)P.nstages * P.stage_bytes);
    uint64_t* empty = full + 4;
    uint64_t* tdone = empty + 4;
    uint32_t* tmem_slot = (uint32_t*)(tdone + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // CTA c owns accumulator set c % nsets and every (gridDim.x / nsets)-th pixel tile: each weight receives gridDim.x / nsets
    // partial sums instead of gridDim.x (the fp32 atomics of the flush were the dominant cost with all CTAs on all sets).
    const int set = (int)blockIdx.x % P.nsets;
    const int tile0 = (int)blockIdx.x / P.nsets, tile_step = ((int)gridDim.x + P.nsets - 1 - set) / P.nsets;
    const int g0 = P.set_first[set], ngroups = P.set_count[set], x_ch0 = P.set_ch0[set];
    const uint32_t need_cols = (uint32_t)(P.max_groups * P.n_cols);
    uint32_t tmem_cols = 32;
    while (tmem_cols < need_cols) tmem_cols <<= 1;
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&map_x);
        prefetch_tmap(&map_y);
        for (int i = 0; i < 4; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
        mbar_init(tdone, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int per_img = P.tiles_x * P.tiles_y;
    const int S = P.nstages;

    if (warp == 0) {
        if (lane == 0) {
            int st = 0;
            uint32_t ph = 0;
            for (int tile = tile0; tile < P.total_tiles; tile += tile_step) {
                const int b = tile / per_img;
                const int t = tile - b * per_img;
                const int ty = t / P.tiles_x, tx = t - ty * P.tiles_x;
                mbar_wait(empty + st, ph ^ 1);
                uint8_t* dst = smem + (size_t)st * P.stage_bytes;
                mbar_expect_tx(full + st, (uint32_t)P.tx_bytes);
                const int x0 = (tx * WTW + P.org_x) * P.stride, y0 = (ty * P.tile_h + P.org_y) * P.stride;
                for (int pl = 0; pl < P.nplanes; ++pl)
                    for (int xb = 0; xb < P.x_boxes; ++xb)
                        tma_load_4d(dst + (size_t)(pl * P.x_boxes + xb) * P.x_plane_bytes, &map_x, full + st, x_ch0 + xb * P.cxb, x0 + (pl & 1), y0 + (pl >> 1), b);
                for (int yb = 0; yb < P.y_boxes; ++yb)
                    tma_load_4d(dst + P.x_bytes + (size_t)yb * P.y_box_bytes, &map_y, full + st, yb * P.cyb, tx * WTW, ty * P.tile_h, b);
                if (++st == S) { st = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            // instruction descriptor: fp32 accumulate, A / B formats, BOTH operands MN-major (bits 15, 16), N >> 3, M >> 4
            const uint32_t idesc = (1u << 4) | ((P.x_f16 ? 0u : 1u) << 7) | ((P.y_f16 ? 0u : 1u) << 10) | (1u << 15) | (1u << 16) |
                                   ((uint32_t)(P.n_cols >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint32_t rowx16 = (uint32_t)(P.cxb * 2) >> 4, rowy16 = (uint32_t)(P.cyb * 2) >> 4;
            const uint32_t a_sbo16 = (uint32_t)P.halo_w * rowx16;             // next 8 pixels of K = next tile row = halo_w rows of the halo box
            const uint32_t b_sbo16 = 8u * rowy16;
            const uint32_t b_lbo16 = (uint32_t)P.y_box_bytes >> 4;
            const uint32_t la = layout_for(P.cxb), lb = layout_for(P.cyb);
            const uint32_t smem16 = smem_u32(smem) >> 4, stage16 = (uint32_t)P.stage_bytes >> 4, xbytes16 = (uint32_t)P.x_bytes >> 4;
            // Descriptors are built ONCE (stage 0, K-step 0) and advanced by adding 16-byte-unit offsets to their 14-bit start-address field
            // (every operand lies below 228 KB, so the sum never carries out of the field).  The first version rebuilt both 64-bit descriptors
            // from the parameter block for every MMA: ~100 clk of single-thread issue per MMA, more than an N <= 64 MMA takes to execute
            // (conv6: 24 MMAs of N = 16 per tile, 146 us for 177 MB of operands).
            uint64_t a_desc[8];
#pragma unroll
            for (int g = 0; g < 8; ++g)
                a_desc[g] = g < ngroups ? make_mn_desc(smem16 + P.g[g0 + g].a_off16, P.g[g0 + g].a_lbo16, a_sbo16, la) : 0ull;
            const uint64_t b_desc = make_mn_desc(smem16 + xbytes16, b_lbo16, b_sbo16, lb);
            const uint32_t a_kstep = 2u * a_sbo16, b_kstep = 2u * b_sbo16, ncols = (uint32_t)P.n_cols;
            const int ksteps = P.ksteps;
            int st = 0;
            uint32_t ph = 0;
            uint32_t accum = 0u;
            for (int tile = tile0; tile < P.total_tiles; tile += tile_step) {
                mbar_wait(full + st, ph);
                tc_fence_after();
                const uint64_t so = (uint64_t)((uint32_t)st * stage16);
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {                             // 16 pixels (two tile rows) per MMA
                    if (ks >= ksteps) break;
                    const uint64_t bd = b_desc + so + (uint64_t)((uint32_t)ks * b_kstep);
                    const uint64_t ao = so + (uint64_t)((uint32_t)ks * a_kstep);
#pragma unroll
                    for (int g = 0; g < 8; ++g)
                        if (g < ngroups) umma_bf16(tmem_base + (uint32_t)g * ncols, a_desc[g] + ao, bd, idesc, (ks == 0) ? accum : 1u);
                }
                umma_commit(empty + st);
                accum = 1u;
                if (++st == S) { st = 0; ph ^= 1; }
            }
            umma_commit(tdone);
        }
        __syncwarp();
    } else {
        // epilogue: once per CTA.  Row of D = X channel (and tap), column = DY channel.
        const bool has_work = tile0 < P.total_tiles;
        if (has_work) {
            const int q = warp & 3;
            const int row = q * 32 + lane;
            mbar_wait(tdone, 0);
            tc_fence_after();
            for (int g = 0; g < ngroups; ++g) {
                const int blk = row / P.atom_rows;
                const int tap = P.g[g0 + g].tap[blk];
                const int cx = P.g[g0 + g].cx0 + (P.Cx >= 128 ? row : row - blk * P.atom_rows) - P.cx_off;
                const bool ok = tap >= 0 && cx >= 0 && cx < P.cx_real;
                float* dst = P.dw + (int64_t)tap * P.w_ts + (int64_t)cx * P.w_xs;
                for (int c0 = 0; c0 < P.n_cols; c0 += 16) {
                    uint32_t r[16];
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
                          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                        : "r"(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * P.n_cols + c0)));
                    tmem_wait_ld();
                    if (ok && !P.debug_noflush) {
                        if (P.vec4) {
                            // destination = the [tap][X channel][DY channel] scratch of spaa_conv_wgrad_tc_scratch: the thread's 16 columns are 64
                            // contiguous bytes -> four 16-byte reductions instead of 16 scalar ones (the scalar flush of all CTAs at the end of the
                            // kernel was 0.39 of the 0.97 ms the 17 launches of a training step took: tools/train_probe.py, $SPAA_WGRAD_NOFLUSH)
                            float* q = dst + c0;
#pragma unroll
                            for (int k = 0; k < 16; k += 4)
                                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(q + k), "f"(__uint_as_float(r[k])), "f"(__uint_as_float(r[k + 1])),
                                             "f"(__uint_as_float(r[k + 2])), "f"(__uint_as_float(r[k + 3]))
                                             : "memory");
                        } else {
#pragma unroll
                            for (int k = 0; k < 16; ++k)
                                if (c0 + k < P.cy_real) atomicAdd(dst + (int64_t)(c0 + k) * P.w_ys, __uint_as_float(r[k]));
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

// dbias[c] += sum over pixels of a dense 16-bit NHWC tensor [P][C]: 256 threads, thread = (8-channel vector, pixel phase).
template <bool F16>
__global__ void __launch_bounds__(256) channel_sum_nhwc_kernel(const uint16_t* __restrict__ x, int64_t npix, int C, float* __restrict__ out) {
    __shared__ float sacc[256 * 8];
    const int vecs = C >> 3;                       // uint4 vectors per pixel
    const int v = threadIdx.x % vecs, phase = threadIdx.x / vecs, nphase = 256 / vecs;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (phase < nphase) {
        for (int64_t p = (int64_t)blockIdx.x * nphase + phase; p < npix; p += (int64_t)gridDim.x * nphase) {
            const uint4 u = __ldg(reinterpret_cast<const uint4*>(x + p * C) + v);
            const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float2 f;
                if constexpr (F16) f = __half22float2(*reinterpret_cast<const __half2*>(&w4[e]));
                else f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w4[e]));
                acc[e * 2] += f.x; acc[e * 2 + 1] += f.y;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) sacc[threadIdx.x * 8 + k] = (phase < nphase) ? acc[k] : 0.f;
    __syncthreads();
    if (threadIdx.x < C) {                          // channel c = vector (c >> 3), element (c & 7): sum over the phases
        const int c = threadIdx.x;
        float s = 0.f;
        for (int ph = 0; ph < nphase; ++ph) s += sacc[(ph * vecs + (c >> 3)) * 8 + (c & 7)];
        atomicAdd(out + c, s);
    }
}

// the same sum for up to kMaxSumJobs tensors in ONE launch (blockIdx.y = tensor): the bias gradients of a whole backward pass
constexpr int kMaxSumJobs = 24;
struct SumJobs {
    const uint16_t* x[kMaxSumJobs];
    float* out[kMaxSumJobs];
    int64_t npix[kMaxSumJobs];
    int32_t C[kMaxSumJobs], c_real[kMaxSumJobs];
};
template <bool F16>
__global__ void __launch_bounds__(256) channel_sum_nhwc_multi_kernel(const __grid_constant__ SumJobs J) {
    __shared__ float sacc[256 * 8];
    const int job = blockIdx.y;
    const int C = J.C[job];
    const int64_t npix = J.npix[job];
    const uint16_t* __restrict__ x = J.x[job];
    const int vecs = C >> 3;
    const int v = threadIdx.x % vecs, phase = threadIdx.x / vecs, nphase = 256 / vecs;
    if ((int64_t)blockIdx.x * nphase >= npix) return;               // (whole CTA: this tensor has fewer pixel groups than the grid is wide)
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (phase < nphase) {
        for (int64_t p = (int64_t)blockIdx.x * nphase + phase; p < npix; p += (int64_t)gridDim.x * nphase) {
            const uint4 u = __ldg(reinterpret_cast<const uint4*>(x + p * C) + v);
            const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float2 f;
                if constexpr (F16) f = __half22float2(*reinterpret_cast<const __half2*>(&w4[e]));
                else f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w4[e]));
                acc[e * 2] += f.x; acc[e * 2 + 1] += f.y;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) sacc[threadIdx.x * 8 + k] = (phase < nphase) ? acc[k] : 0.f;
    __syncthreads();
    if (threadIdx.x < J.c_real[job]) {
        const int c = threadIdx.x;
        float s = 0.f;
        for (int ph = 0; ph < nphase; ++ph) s += sacc[(ph * vecs + (c >> 3)) * 8 + (c & 7)];
        atomicAdd(J.out[job] + c, s);
    }
}

// dw += scratch for up to kMaxScatterJobs layers in one launch (blockIdx.y = layer), and scratch = 0 for the next backward pass:
// scratch is [tap][real X channel][Cy] (DY channel contiguous, zero-padded beyond cy_real), dw the parameter gradient in its own layout.
constexpr int kMaxScatterJobs = 24;
struct ScatterJobs {
    float* scratch[kMaxScatterJobs];
    float* dw[kMaxScatterJobs];
    int64_t w_ts[kMaxScatterJobs], w_xs[kMaxScatterJobs], w_ys[kMaxScatterJobs];
    int32_t ntap[kMaxScatterJobs], cx_real[kMaxScatterJobs], Cy[kMaxScatterJobs], cy_real[kMaxScatterJobs];
    int32_t tiles[kMaxScatterJobs];                 // (X-channel blocks of 8) x (DY-channel blocks of 32)
};
// One CTA moves tiles of (all taps) x 8 X channels x 32 DY channels through shared memory: the scratch is read (and zeroed) along its contiguous DY
// channel, the parameter gradient is updated along ITS contiguous (X channel, tap) runs -- for the usual layout w_ts == 1, w_xs == ntap, one run of
// 8 * ntap floats per DY channel.  (A plain element-wise walk paid two 32-byte sectors per 4-byte element on one side or the other: 82 / 164 us per
// training step instead of the ~25 us this takes.)  Other layouts take the same path with scattered, still correct, updates.
constexpr int kScTX = 8, kScTY = 32;
__global__ void __launch_bounds__(256) wgrad_scatter_multi_kernel(const __grid_constant__ ScatterJobs J) {
    __shared__ float tile[kScTY][kScTX * 9 + 1];
    const int job = blockIdx.y;
    float* __restrict__ sc = J.scratch[job];
    float* __restrict__ dw = J.dw[job];
    const int ntap = J.ntap[job], cxr = J.cx_real[job], Cy = J.Cy[job], cyr = J.cy_real[job];
    const int64_t ts = J.w_ts[job], xs = J.w_xs[job], ys = J.w_ys[job];
    const int tiles_y = (cyr + kScTY - 1) / kScTY;
    const int run = kScTX * ntap;
    for (int t = blockIdx.x; t < J.tiles[job]; t += gridDim.x) {
        const int cx0 = (t / tiles_y) * kScTX, cy0 = (t % tiles_y) * kScTY;
        // (tap, 8 x 32) = 256 elements per pass and exactly one per thread: all loads of a phase are issued before anything depends on them
        const int cy_l = threadIdx.x % kScTY, cx_l = threadIdx.x / kScTY;
        const bool in = cx0 + cx_l < cxr && cy0 + cy_l < cyr;
        float v[9];
#pragma unroll
        for (int tap = 0; tap < 9; ++tap)
            v[tap] = (tap < ntap && in) ? sc[((int64_t)tap * cxr + cx0 + cx_l) * Cy + cy0 + cy_l] : 0.f;
#pragma unroll
        for (int tap = 0; tap < 9; ++tap)
            if (tap < ntap) {
                if (in) sc[((int64_t)tap * cxr + cx0 + cx_l) * Cy + cy0 + cy_l] = 0.f;
                tile[cy_l][cx_l * ntap + tap] = v[tap];
            }
        __syncthreads();
        float* dp[9];
        float old[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            const int e = i * 256 + (int)threadIdx.x;
            const int cy2 = e / run, j = e - cy2 * run;
            const int cx2 = j / ntap, tap = j - cx2 * ntap;
            const bool ok = i < ntap && e < kScTY * run && cx0 + cx2 < cxr && cy0 + cy2 < cyr;
            dp[i] = ok ? dw + ((int64_t)tap * ts + (int64_t)(cx0 + cx2) * xs + (int64_t)(cy0 + cy2) * ys) : nullptr;
            old[i] = ok ? *dp[i] : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            const int e = i * 256 + (int)threadIdx.x;
            const int cy2 = e / run, j = e - cy2 * run;
            if (dp[i]) *dp[i] = old[i] + tile[cy2][j];
        }
        __syncthreads();
    }
}

}  // namespace

extern "C" {

int spaa_conv_wgrad_tc_supported(const spaa_conv_desc* d) {
    if (!d) return 0;
    auto okc = [](int c) { return c == 16 || c == 32 || c == 64 || c == 128 || c == 256; };
    if (!(d->in_dtype == 1 || d->in_dtype == 2) || d->out_dtype != d->in_dtype) return 0;      // kind::f16 traps on mixed fp16 x bf16 operands (measured)
    if (!okc(d->Cin) || !okc(d->Cout)) return 0;
    if (d->up != 1 || d->flip != 0 || !(d->stride == 1 || d->stride == 2)) return 0;
    if (d->KH != d->KW || d->KH > 3 || d->pad_h != d->pad_w) return 0;
    // split (bf16x3 split-precision operands): x / dy point at ONE part (Cin / Cout logical channels) of tensors whose pixels hold three parts
    const int np = d->split ? 3 : 1;
    if (d->split && d->in_dtype != 1) return 0;
    if (d->in_cs != 1 || d->in_ps != np * d->Cin || (d->B > 1 && d->in_bs != (int64_t)d->Hin * d->Win * d->Cin * np)) return 0;
    if (d->out_cs != 1 || d->out_ps != np * d->Cout || (d->B > 1 && d->out_bs != (int64_t)d->Hout * d->Wout * d->Cout * np)) return 0;
    return 1;
}

static inline int wg_floordiv2(int v) { return v >= 0 ? v / 2 : -((-v + 1) / 2); }

/* d describes the FORWARD gather conv (up == 1, flip == 0): `x` = gathered operand, dense 16-bit NHWC [B,Hin,Win,Cin];
 * `dy` = pointwise operand, dense 16-bit NHWC [B,Hout,Wout,Cout]; dw fp32 (+=) addressed by d->w_ts / w_cis (X channel) / w_cos (DY channel).
 * cx_real / cx_off / cy_real: zero-padded operands (3- and 6-channel images padded to 16). */
static int wgrad_tc_impl(const spaa_conv_desc* d, const void* x, const void* dy, float* dw, int cx_real, int cx_off, int cy_real, bool scratch_layout,
                         spaa_stream_t stream) {
    SPAA_CHECK_ARG(d && x && dy && dw, "spaa_conv_wgrad_tc: null argument");
    SPAA_CHECK_ARG(spaa_conv_wgrad_tc_supported(d), "spaa_conv_wgrad_tc: unsupported shape / layout (see spaa_conv_wgrad_tc_supported)");
    SPAA_CHECK_ARG(cx_real > 0 && cx_off >= 0 && cx_off + cx_real <= d->Cin && cy_real > 0 && cy_real <= d->Cout, "spaa_conv_wgrad_tc: bad channel ranges");
    EncodeTiledFn enc = get_encode();
    if (!enc) { set_last_error("spaa_conv_wgrad_tc: cuTensorMapEncodeTiled is unavailable in this driver"); return SPAA_ERR_CUDA; }
    const int Cx = d->Cin, Cy = d->Cout;
    WgParams P;
    memset(&P, 0, sizeof(P));
    P.B = d->B; P.Cx = Cx; P.Cy = Cy; P.Hy = d->Hout; P.Wy = d->Wout;
    P.stride = d->stride; P.nplanes = d->stride * d->stride;
    P.cxb = Cx >= 64 ? 64 : Cx; P.cyb = Cy >= 64 ? 64 : Cy;
    P.y_boxes = Cy / P.cyb;
    P.n_cols = Cy;
    P.atom_rows = Cx >= 128 ? 64 : P.cxb;
    P.cx_real = cx_real; P.cx_off = cx_off; P.cy_real = cy_real;
    P.x_f16 = d->in_dtype == 2; P.y_f16 = d->out_dtype == 2;
    { static const int nf = [] { const char* e = getenv("SPAA_WGRAD_NOFLUSH"); return e ? atoi(e) : 0; }(); P.debug_noflush = nf; }
    P.w_ts = d->w_ts; P.w_xs = d->w_cis; P.w_ys = d->w_cos; P.dw = dw;
    if (scratch_layout) {           // [tap][real X channel][Cy]: the DY channel is the contiguous index
        P.w_ts = (int64_t)cx_real * Cy; P.w_xs = Cy; P.w_ys = 1; P.vec4 = 1;
        SPAA_CHECK_ARG(((uintptr_t)dw & 15) == 0, "spaa_conv_wgrad_tc_scratch: scratch must be 16-byte aligned");
    }
    // ---- taps: (plane, shift) ----
    struct TapPos { int plane, qy, qx; };
    TapPos tp[9];
    int ntaps = 0, qminx = 1 << 20, qmaxx = -(1 << 20), qminy = 1 << 20, qmaxy = -(1 << 20);
    for (int r = 0; r < d->KH; ++r)
        for (int s = 0; s < d->KW; ++s) {
            TapPos t{0, r - d->pad_h, s - d->pad_w};
            if (d->stride == 2) {
                const int ry = r - d->pad_h, rx = s - d->pad_w, py = ry & 1, px = rx & 1;
                t.plane = py * 2 + px; t.qy = wg_floordiv2(ry - py); t.qx = wg_floordiv2(rx - px);
            }
            qminx = t.qx < qminx ? t.qx : qminx; qmaxx = t.qx > qmaxx ? t.qx : qmaxx;
            qminy = t.qy < qminy ? t.qy : qminy; qmaxy = t.qy > qmaxy ? t.qy : qmaxy;
            tp[ntaps++] = t;
        }
    P.org_x = qminx; P.org_y = qminy;
    // a 16-row tile of a wide layer (conv4 / conv4_s: 46 KB of X halo + 64 KB of DY) fills shared memory with ONE stage, which serialises the
    // TMA loads and the MMAs (measured 127 us for 34 us of MMA work); 8-row tiles give three stages
    P.tile_h = WTH;
    {
        const int64_t hw16 = (int64_t)(WTW + (qmaxx - qminx)) * (WTH + (qmaxy - qminy));
        const int64_t stage16 = (int64_t)d->stride * d->stride * (Cx >= 128 ? 2 : 1) * hw16 * (Cx >= 64 ? 128 : Cx * 2) + (int64_t)128 * Cy * 2;
        if (2 * stage16 > 200 * 1024) P.tile_h = WTH / 2;
    }
    P.ksteps = P.tile_h / 2;
    P.halo_w = WTW + (qmaxx - qminx); P.halo_h = P.tile_h + (qmaxy - qminy);
    P.tiles_y = (d->Hout + P.tile_h - 1) / P.tile_h; P.tiles_x = (d->Wout + WTW - 1) / WTW;
    P.total_tiles = d->B * P.tiles_y * P.tiles_x;
    SPAA_CHECK_ARG(P.total_tiles > 0, "spaa_conv_wgrad_tc: empty problem");
    const int rowx = P.cxb * 2, rowy = P.cyb * 2;
    P.x_boxes = Cx >= 128 ? 2 : 1;                                   // one 128-channel chunk of X per launch
    P.x_plane_bytes = (P.halo_h * P.halo_w * rowx + 1023) & ~1023;
    P.x_bytes = P.nplanes * P.x_boxes * P.x_plane_bytes;
    P.y_box_bytes = P.tile_h * WTW * rowy;                            // tile_h x 8 pixels
    if (P.y_box_bytes & 1023) P.y_box_bytes = (P.y_box_bytes + 1023) & ~1023;
    P.stage_bytes = P.x_bytes + P.y_boxes * P.y_box_bytes;
    P.tx_bytes = P.nplanes * P.x_boxes * P.halo_h * P.halo_w * rowx + P.y_boxes * P.tile_h * WTW * rowy;
    auto tap_off16 = [&](int t, int xb) {
        return (uint32_t)(((tp[t].plane * P.x_boxes + xb) * P.x_plane_bytes + ((tp[t].qy - qminy) * P.halo_w + (tp[t].qx - qminx)) * rowx) >> 4);
    };
    // ---- accumulator list: (tap, 128-channel chunk) for Cx >= 128, tap pairs for Cx == 64, single taps below ----
    struct Acc { WgGroup g; int chunk; };
    Acc accs[32];
    int nacc = 0;
    if (Cx >= 128) {
        for (int ch = 0; ch < Cx / 128; ++ch)
            for (int t = 0; t < ntaps; ++t) {
                Acc& a = accs[nacc++];
                memset(&a, 0, sizeof(a));
                a.chunk = ch;
                a.g.a_off16 = tap_off16(t, 0); a.g.a_lbo16 = (uint32_t)P.x_plane_bytes >> 4;       // channels 64..127 = the next box of the same plane
                for (int k = 0; k < 8; ++k) a.g.tap[k] = (int8_t)t;
                a.g.cx0 = ch * 128;
            }
    } else if (Cx == 64) {
        for (int t = 0; t < ntaps; t += 2) {                       // two taps per MMA: atom 0 / atom 1 = the views of tap t0 / t1
            Acc& a = accs[nacc++];
            memset(&a, 0, sizeof(a));
            for (int k = 0; k < 8; ++k) a.g.tap[k] = -1;
            int t0 = t, t1 = t + 1 < ntaps ? t + 1 : -1;
            if (t1 >= 0 && tap_off16(t1, 0) < tap_off16(t0, 0)) { const int tmp = t0; t0 = t1; t1 = tmp; }
            a.g.a_off16 = tap_off16(t0, 0);
            a.g.tap[0] = (int8_t)t0;
            if (t1 >= 0) { a.g.a_lbo16 = tap_off16(t1, 0) - tap_off16(t0, 0); a.g.tap[1] = (int8_t)t1; }    // distinct taps: distinct (plane, shift)
            else a.g.a_lbo16 = (uint32_t)rowx >> 4;
        }
    } else {
        // Cx <= 32: the 128 M rows hold 4 (8) atoms of 32 (16) channels, one smem row apart (LBO = one row), i.e. atom k reads the view
        // shifted by k pixels in x: the taps (plane, qy, qx + k) of the SAME filter row ride along in the same MMA.  A 3x3 stride-1 layer
        // needs 3 accumulators instead of 9 (conv6: one accumulator set instead of two, a third of the MMAs), a stride-2 one 6.
        const int natoms = 128 / P.cxb;
        bool used[9] = {false, false, false, false, false, false, false, false, false};
        for (int t = 0; t < ntaps; ++t) {
            if (used[t]) continue;
            bool is_base = true;                                  // start runs at their smallest qx
            for (int u = 0; u < ntaps; ++u)
                if (!used[u] && u != t && tp[u].plane == tp[t].plane && tp[u].qy == tp[t].qy && tp[u].qx == tp[t].qx - 1) is_base = false;
            if (!is_base) continue;
            Acc& a = accs[nacc++];
            memset(&a, 0, sizeof(a));
            for (int k = 0; k < 8; ++k) a.g.tap[k] = -1;
            a.g.a_off16 = tap_off16(t, 0); a.g.a_lbo16 = (uint32_t)rowx >> 4; a.g.tap[0] = (int8_t)t;
            used[t] = true;
            for (int k = 1; k < natoms; ++k)
                for (int u = 0; u < ntaps; ++u)
                    if (!used[u] && tp[u].plane == tp[t].plane && tp[u].qy == tp[t].qy && tp[u].qx == tp[t].qx + k) { a.g.tap[k] = (int8_t)u; used[u] = true; }
        }
        for (int t = 0; t < ntaps; ++t) SPAA_CHECK_ARG(used[t], "spaa_conv_wgrad_tc: internal error (tap not assigned to an accumulator)");
    }
    int max_per_set = 512 / Cy;
    if (max_per_set > 8) max_per_set = 8;
    { static const int ms = [] { const char* e = getenv("SPAA_WGRAD_MAXSET"); return e ? atoi(e) : 0; }(); if (ms > 0 && ms < max_per_set) max_per_set = ms; }
    SPAA_CHECK_ARG(nacc <= kMaxGroups, "spaa_conv_wgrad_tc: too many accumulators");
    for (int i = 0; i < nacc; ++i) P.g[i] = accs[i].g;
    P.nsets = 0; P.max_groups = 0;
    for (int i = 0; i < nacc;) {
        int n = 0;
        while (i + n < nacc && accs[i + n].chunk == accs[i].chunk) ++n;          // accumulators of this X chunk
        const int k = (n + max_per_set - 1) / max_per_set;                        // sets for the chunk, evenly filled
        for (int j = 0, first = i; j < k; ++j) {
            const int cnt = (n - (first - i) + (k - j) - 1) / (k - j);
            SPAA_CHECK_ARG(P.nsets < kMaxSets, "spaa_conv_wgrad_tc: too many accumulator sets");
            P.set_first[P.nsets] = (int16_t)first; P.set_count[P.nsets] = (int16_t)cnt; P.set_ch0[P.nsets] = (int16_t)(accs[i].chunk * 128);
            if (cnt > P.max_groups) P.max_groups = cnt;
            ++P.nsets;
            first += cnt;
        }
        i += n;
    }
    // CTAs per SM: the narrow layers (conv6: 24 MMAs of N = 16 and 15 KB of operands per 128-pixel tile) are bound by the ONE thread that issues a CTA's
    // MMAs and by the bytes one CTA keeps in flight, not by the tensor pipe or TMEM -- two or three CTAs share an SM when their accumulators
    // (TMEM columns) and at least three stages each fit (measured, batch 24: conv6 106 -> 91 us, skipConv2 27.6 -> 23.5 us; the stride-2 layers
    // conv1 / conv1_s got slower, 46 -> 52 us, and keep one).  $SPAA_WGRAD_CTAS caps it (1 = the round-1 shape).
    static const int max_ctas = [] { const char* e = getenv("SPAA_WGRAD_CTAS"); return e ? atoi(e) : 3; }();
    uint32_t cols = 32;
    while (cols < (uint32_t)(P.max_groups * P.n_cols)) cols <<= 1;
    int ctas = 1;
    for (int c = max_ctas < 3 ? max_ctas : 3; c >= 2; --c)
        if (d->stride == 1 && (int)cols * c <= 512 && ((226 * 1024) / c - 2048) / P.stage_bytes >= 3) { ctas = c; break; }
    P.nstages = (ctas == 1 ? 200 * 1024 : (226 * 1024) / ctas - 2048) / P.stage_bytes;
    if (P.nstages > 4) P.nstages = 4;
    SPAA_CHECK_ARG(P.nstages >= 1, "spaa_conv_wgrad_tc: stage does not fit in shared memory");
    CUtensorMap mx, my;
    {
        const cuuint64_t cphys = (cuuint64_t)d->in_ps;             // physical channels per pixel (3 * Cx for split-precision operands)
        cuuint64_t dims[4] = {(cuuint64_t)Cx, (cuuint64_t)d->Win, (cuuint64_t)d->Hin, (cuuint64_t)d->B};
        cuuint64_t strides[3] = {cphys * 2, (cuuint64_t)d->Win * cphys * 2, (cuuint64_t)d->Hin * d->Win * cphys * 2};
        cuuint32_t box[4] = {(cuuint32_t)P.cxb, (cuuint32_t)(P.halo_w * d->stride), (cuuint32_t)(P.halo_h * d->stride), 1};
        cuuint32_t es[4] = {1, (cuuint32_t)d->stride, (cuuint32_t)d->stride, 1};
        const CUtensorMapSwizzle sw = P.cxb == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (P.cxb == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
        CUresult r = enc(&mx, P.x_f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_last_error("spaa_conv_wgrad_tc: cuTensorMapEncodeTiled(x) failed with %d", (int)r); return SPAA_ERR_CUDA; }
    }
    {
        const cuuint64_t cphys = (cuuint64_t)d->out_ps;
        cuuint64_t dims[4] = {(cuuint64_t)Cy, (cuuint64_t)d->Wout, (cuuint64_t)d->Hout, (cuuint64_t)d->B};
        cuuint64_t strides[3] = {cphys * 2, (cuuint64_t)d->Wout * cphys * 2, (cuuint64_t)d->Hout * d->Wout * cphys * 2};
        cuuint32_t box[4] = {(cuuint32_t)P.cyb, (cuuint32_t)WTW, (cuuint32_t)P.tile_h, 1};
        cuuint32_t es[4] = {1, 1, 1, 1};
        const CUtensorMapSwizzle sw = P.cyb == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (P.cyb == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
        CUresult r = enc(&my, P.y_f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(dy), dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_last_error("spaa_conv_wgrad_tc: cuTensorMapEncodeTiled(dy) failed with %d", (int)r); return SPAA_ERR_CUDA; }
    }
    const size_t smem_bytes = (size_t)P.nstages * P.stage_bytes + 128 + 1024;
    static SmemOptIn opt;
    if (smem_bytes > 220 * 1024 || !opt.ensure(conv_wgrad_tc_kernel, (size_t)(220 * 1024))) {
        set_last_error("spaa_conv_wgrad_tc: cannot reserve shared memory");
        return SPAA_ERR_CUDA;
    }
    int grid = kNumSMs * ctas;
    if ((int64_t)P.total_tiles * P.nsets < grid) grid = P.total_tiles * P.nsets;
    conv_wgrad_tc_kernel<<<grid, kThreads, smem_bytes, (cudaStream_t)stream>>>(mx, my, P);
    SPAA_CHECK_LAUNCH("spaa_conv_wgrad_tc");
    return SPAA_OK;
}

int spaa_conv_wgrad_tc(const spaa_conv_desc* d, const void* x, const void* dy, float* dw, int cx_real, int cx_off, int cy_real, spaa_stream_t stream) {
    return wgrad_tc_impl(d, x, dy, dw, cx_real, cx_off, cy_real, false, stream);
}

int64_t spaa_conv_wgrad_tc_scratch_elems(const spaa_conv_desc* d, int cx_real) {
    if (!d || cx_real <= 0) return 0;
    return (int64_t)d->KH * d->KW * cx_real * d->Cout;
}

int spaa_conv_wgrad_tc_scratch(const spaa_conv_desc* d, const void* x, const void* dy, float* scratch, int cx_real, int cx_off, int cy_real, spaa_stream_t stream) {
    return wgrad_tc_impl(d, x, dy, scratch, cx_real, cx_off, cy_real, true, stream);
}

int spaa_wgrad_scatter_multi(const float* const* scratch, float* const* dw, const int32_t* ntap, const int32_t* cx_real, const int32_t* Cy, const int32_t* cy_real,
                             const int64_t* w_ts, const int64_t* w_xs, const int64_t* w_ys, int n, spaa_stream_t stream) {
    SPAA_CHECK_ARG(scratch && dw && ntap && cx_real && Cy && cy_real && w_ts && w_xs && w_ys && n >= 1 && n <= kMaxScatterJobs, "spaa_wgrad_scatter_multi: bad arguments");
    ScatterJobs J;
    memset(&J, 0, sizeof(J));
    int64_t most = 1;
    for (int i = 0; i < n; ++i) {
        SPAA_CHECK_ARG(scratch[i] && dw[i] && ntap[i] >= 1 && cx_real[i] >= 1 && Cy[i] >= 1 && cy_real[i] >= 1 && cy_real[i] <= Cy[i], "spaa_wgrad_scatter_multi: bad job %d", i);
        SPAA_CHECK_ARG(ntap[i] <= 9, "spaa_wgrad_scatter_multi: at most 9 taps (job %d)", i);
        J.scratch[i] = const_cast<float*>(scratch[i]); J.dw[i] = dw[i];
        J.ntap[i] = ntap[i]; J.cx_real[i] = cx_real[i]; J.Cy[i] = Cy[i]; J.cy_real[i] = cy_real[i]; J.w_ts[i] = w_ts[i]; J.w_xs[i] = w_xs[i]; J.w_ys[i] = w_ys[i];
        J.tiles[i] = ((cx_real[i] + kScTX - 1) / kScTX) * ((cy_real[i] + kScTY - 1) / kScTY);
        if (J.tiles[i] > most) most = J.tiles[i];
    }
    int64_t blocks = most;
    if (blocks > 2 * kNumSMs) blocks = 2 * kNumSMs;
    wgrad_scatter_multi_kernel<<<dim3((unsigned)blocks, (unsigned)n), 256, 0, (cudaStream_t)stream>>>(J);
    SPAA_CHECK_LAUNCH("spaa_wgrad_scatter_multi");
    return SPAA_OK;
}

/* out[c] += sum over pixels of a DENSE 16-bit NHWC tensor [npix][C] (C a multiple of 8, <= 256): bias gradients. dtype 1 bf16, 2 fp16. */
int spaa_channel_sum_nhwc16(const void* x, int dtype, int64_t npix, int C, float* out, spaa_stream_t stream) {
    SPAA_CHECK_ARG(x && out && npix > 0 && C >= 8 && C <= 256 && (C & 7) == 0 && (256 % (C >> 3)) == 0 && (dtype == 1 || dtype == 2),
                   "spaa_channel_sum_nhwc16: bad arguments");
    const int nphase = 256 / (C >> 3);
    int64_t blocks = (npix + nphase * 8 - 1) / (nphase * 8);
    if (blocks > 4 * kNumSMs) blocks = 4 * kNumSMs;
    if (blocks < 1) blocks = 1;
    if (dtype == 2) channel_sum_nhwc_kernel<true><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const uint16_t*)x, npix, C, out);
    else channel_sum_nhwc_kernel<false><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const uint16_t*)x, npix, C, out);
    SPAA_CHECK_LAUNCH("spaa_channel_sum_nhwc16");
    return SPAA_OK;
}

/* n <= 24 such sums in one launch: outs[i][c] += sum over the npix[i] pixels of xs[i] (dense 16-bit NHWC, C[i] channels) for c < c_real[i]
 * (zero-padded tensors: only the real channels are written).  All tensors of one dtype.  The arrays are HOST arrays, read during the call. */
int spaa_channel_sum_nhwc16_multi(const void* const* xs, const int64_t* npix, const int32_t* C, const int32_t* c_real, float* const* outs, int dtype, int n,
                                  spaa_stream_t stream) {
    SPAA_CHECK_ARG(xs && npix && C && c_real && outs && n >= 1 && n <= kMaxSumJobs && (dtype == 1 || dtype == 2), "spaa_channel_sum_nhwc16_multi: bad arguments");
    SumJobs J;
    memset(&J, 0, sizeof(J));
    int64_t blocks = 1;
    for (int i = 0; i < n; ++i) {
        SPAA_CHECK_ARG(xs[i] && outs[i] && npix[i] > 0 && C[i] >= 8 && C[i] <= 256 && (C[i] & 7) == 0 && (256 % (C[i] >> 3)) == 0 && c_real[i] >= 1 && c_real[i] <= C[i],
                       "spaa_channel_sum_nhwc16_multi: bad tensor %d", i);
        J.x[i] = (const uint16_t*)xs[i]; J.out[i] = outs[i]; J.npix[i] = npix[i]; J.C[i] = C[i]; J.c_real[i] = c_real[i];
        const int nphase = 256 / (C[i] >> 3);
        const int64_t b = (npix[i] + nphase * 8 - 1) / (nphase * 8);
        if (b > blocks) blocks = b;
    }
    if (blocks > 2 * kNumSMs) blocks = 2 * kNumSMs;
    const dim3 grid((unsigned)blocks, (unsigned)n);
    if (dtype == 2) channel_sum_nhwc_multi_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(J);
    else channel_sum_nhwc_multi_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(J);
    SPAA_CHECK_LAUNCH("spaa_channel_sum_nhwc16_multi");
    return SPAA_OK;
}

}  // extern "C"

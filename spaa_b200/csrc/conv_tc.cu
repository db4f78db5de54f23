// tcgen05 / TMEM / TMA implicit-GEMM convolution for sm_100a (bf16 NHWC activations, fp32 accumulation in TMEM).
//
// One kernel family serves nn.Conv2d (stride 1 or 2), nn.ConvTranspose2d (stride 2) and the backward-data pass of each
// (the same "gather conv" of include/spaa_b200.h, spaa_conv_desc) for the layers whose channel counts are multiples
// of 32 -- conv2..conv5, conv*_s, skipConv2/3, transConv1/2 of /root/reference/src/python/models.py:18-46,223-252 --
// i.e. the work cuDNN does for those modules in the reference.
//
// Formulation.  Output pixels are tiled in 8 x 16 rectangles (M = 128 rows of the MMA).  For one filter tap the A
// operand of such a tile is a shifted 8 x 16 x BK box of the NHWC input: ONE 4-D TMA load (cp.async.bulk.tensor) lands
// it in shared memory already in the K-major 128B/64B-swizzled layout tcgen05.mma reads; convolution padding and ragged
// edge tiles are TMA out-of-bound zero fill; stride-2 convolutions use the TMA traversal stride; transposed /
// backward-of-strided convolutions are split into their 4 output-parity phases, each a stride-1 problem over a subset
// of the taps.  The B operand is the pre-packed bf16 weight slice [tap][Cout][Cin] (2-D TMA).  K loop = taps x Cin/BK.
//
// Kernel: persistent CTAs (one per SM), warp-specialised: warp 0 = TMA producer, warp 1 = MMA issuer (one elected
// lane issues tcgen05.mma, M=128, N=Cout, K=16) + TMEM owner, warps 2..5 = epilogue (tcgen05.ld of their 32-lane
// quadrant -> bias / residual add / ReLU / ReLU-mask of the backward pass -> bf16 -> global).  smem ring of NSTAGES
// {A,B} stages with full/empty mbarriers; TWO accumulators in TMEM (2 x Cout columns) so the epilogue of tile i
// overlaps the MMAs of tile i+1.
//
// That paragraph describes conv_tc_kernel, the first version (one TMA box per tap), kept as the fallback.  The product path is
// conv_halo_kernel further down: ONE halo box per tile whose taps are row-shifted descriptor views; weights resident in shared memory or
// streamed through their own ring; 1-4 accumulator buffers and 1-4 epilogue groups per CTA; two epilogues (the round-1 one for the wide layers
// and the split-precision mode, epilogue_nhwc16 with TMA tile stores for the narrow ones); halo_plan() picks the shape of every launch.
#include "common.cuh"
#include "tc_ptx.cuh"
#include "../../include/spaa_b200.h"
#include <mutex>
#include <type_traits>
#include <unordered_map>
#include <string>
#include <cstring>
#include <cstdlib>

using namespace spaa;

namespace {

constexpr int TH = 8, TW = 16, BM = TH * TW;      // output tile = 128 MMA rows
constexpr int kThreads = 192;                      // 6 warps
constexpr int kMaxTaps = 9, kMaxPhases = 4;
constexpr int kMaxKChunks = 24;                    // 256 channels / 64 per chunk x 6 part products

struct Tap { int8_t dy, dx, slot, pad; };
struct Phase {
    int32_t ntaps, py, px, Hph, Wph, tiles_y, tiles_x, tile_base;   // tile_base: first global tile index of this phase
    Tap taps[kMaxTaps];
};
struct TcParams {
    int32_t B, Cin, Hin, Win, Cout, Hout, Wout;
    int32_t in_step;            // TMA traversal stride of the input (conv stride), 1 for up-sampling phases
    int32_t out_step;           // output pixel step (2 for the parity phases of up == 2)
    int32_t nphases, total_tiles, kchunks;
    int32_t epi_flags, mask_mode;
    int64_t add_bs;             // 0: `add` is broadcast over the batch
    int64_t mask_bs;
    Phase ph[kMaxPhases];
    int32_t out_planar;         // 1: out / add are fp32 NCHW planes with Cout (real) channels; 0: 16-bit NHWC with Cout channels
    const float* bias;
    const void* add;            // 16-bit NHWC (same type as out) or fp32 planar
    const uint16_t* mask;       // 16-bit NHWC, fp16 or bf16: only the sign / zero test is used
    const uint16_t* mask2;
    void* out;
    void* out2;
};

using namespace spaa::tc;

// K-major operand tile in shared memory, rows of ROW_BYTES (= swizzle width), 8-row groups ROW_BYTES*8 apart.
// Field layout: cute/arch/mma_sm100_desc.hpp (SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46),
// version=1 [46,48), layout type [61,64): 2 = 128B swizzle, 4 = 64B, 6 = 32B.
template <int ROW_BYTES> SPAA_D uint64_t make_kmajor_desc(uint32_t smem_addr) {
    constexpr uint64_t layout = ROW_BYTES == 128 ? 2 : (ROW_BYTES == 64 ? 4 : 6);
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)(((ROW_BYTES * 8) >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= layout << 61;
    return d;
}
// kind::f16 instruction descriptor (InstrDescriptor): C fp32 [4,6)=1, A bf16 [7,10)=1, B bf16 [10,13)=1, both K-major,
// N>>3 at [17,23), M>>4 at [24,29)
SPAA_D uint32_t make_idesc(int M, int N, bool fp16) {
    const uint32_t fmt = fp16 ? 0u : 1u;          // F32F16Format: 0 = F16, 1 = BF16
    return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// 16-bit float helpers: T16 = __half or __nv_bfloat16
template <bool F16> SPAA_D float2 unpack2(uint32_t u) {
    if constexpr (F16) return __half22float2(*reinterpret_cast<const __half2*>(&u));
    else return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u));
}
template <bool F16> SPAA_D uint32_t pack2(float a, float b) {
    if constexpr (F16) { const __half2 h = __floats2half2_rn(a, b); return *reinterpret_cast<const uint32_t*>(&h); }
    else { const __nv_bfloat162 h = __floats2bfloat162_rn(a, b); return *reinterpret_cast<const uint32_t*>(&h); }
}
// x > 0 for an fp16 or bf16 bit pattern (both: sign bit 15, zero = all other bits clear; NaN counts as positive like `m > 0` never would,
// but activations are finite)
SPAA_D bool pos16(uint32_t bits) { return (bits & 0x8000u) == 0u && (bits & 0x7FFFu) != 0u; }

// 0xFFFF in each half of the result where that 16-bit float half of w is > 0, for fp16 AND bf16 bit patterns (masks are activations
// of either type, include/spaa_b200.h): both formats share the sign bit and the all-zero pattern, and a bf16 pattern read as fp16
// is a positive finite / subnormal fp16 unless its magnitude is >= 2^121 (never an activation), so ONE fp16 compare serves both
// (HSET2.BM; fp16 compares do not flush subnormals).
SPAA_D uint32_t posmask2(uint32_t w) { return __hgt2_mask(*reinterpret_cast<const __half2*>(&w), __float2half2_rn(0.f)); }

template <int BN, int BK> struct SmemLayout {
    static constexpr int kABytes = BM * BK * 2, kBBytes = BN * BK * 2, kStageBytes = kABytes + kBBytes;
    static constexpr int kStages = (200 * 1024 / kStageBytes) > 8 ? 8 : (200 * 1024 / kStageBytes);
    static constexpr int kBarOff = kStages * kStageBytes;                 // full[kStages], empty[kStages], tfull[2], tempty[2]
    static constexpr int kBiasOff = kBarOff + (2 * kStages + 4) * 8 + 16;
    static constexpr int kTotal = kBiasOff + BN * 4 + 1024;               // + slack for the 1024 B alignment of the ring
};

template <int BN, int BK, bool F16>
__global__ void __launch_bounds__(kThreads, 1) conv_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                                                              const __grid_constant__ TcParams P) {
    using L = SmemLayout<BN, BK>;
    constexpr int kStages = L::kStages;
    constexpr uint32_t kTmemCols = (2 * BN) < 32 ? 32 : 2 * BN;           // power of two >= 32 (BN in {32,64,128,256})
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* full = (uint64_t*)(smem + L::kBarOff);
    uint64_t* empty = full + kStages;
    uint64_t* tfull = empty + kStages;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = (uint32_t*)(tempty + 2);
    float* s_bias = (float*)(smem + L::kBiasOff);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&map_a);
        prefetch_tmap(&map_b);
        for (int i = 0; i < kStages; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(tfull + i, 1); mbar_init(tempty + i, 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
    for (int i = threadIdx.x; i < BN; i += kThreads) s_bias[i] = (P.bias && i < P.Cout) ? P.bias[i] : 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    auto decode = [&](int tile, int& ph, int& b, int& ty, int& tx) {
        ph = 0;
#pragma unroll
        for (int i = 1; i < kMaxPhases; ++i)
            if (i < P.nphases && tile >= P.ph[i].tile_base) ph = i;
        const Phase& F = P.ph[ph];
        int t = tile - F.tile_base;
        const int per_img = F.tiles_y * F.tiles_x;
        b = t / per_img;
        t -= b * per_img;
        ty = t / F.tiles_x;
        tx = t - ty * F.tiles_x;
    };

    if (warp == 0) {
        // ===================================== TMA producer ========================================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
                int ph, b, ty, tx;
                decode(tile, ph, b, ty, tx);
                const Phase& F = P.ph[ph];
                const int y0 = ty * TH * P.in_step, x0 = tx * TW * P.in_step;
                for (int t = 0; t < F.ntaps; ++t) {
                    const Tap tp = F.taps[t];
                    for (int kc = 0; kc < P.kchunks; ++kc) {
                        mbar_wait(empty + stage, phase ^ 1);
                        uint8_t* sa = smem + stage * L::kStageBytes;
                        uint8_t* sb = sa + L::kABytes;
                        mbar_expect_tx(full + stage, L::kStageBytes);
                        tma_load_4d(sa, &map_a, full + stage, kc * BK, x0 + tp.dx, y0 + tp.dy, b);
                        tma_load_2d(sb, &map_b, full + stage, kc * BK, tp.slot * BN);
                        if (++stage == kStages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================== MMA issuer ==========================================
        const uint32_t idesc = make_idesc(BM, BN, F16);
        int stage = 0;
        uint32_t phase = 0;
        int local = 0;
        for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x, ++local) {
            int ph, b, ty, tx;
            decode(tile, ph, b, ty, tx);
            const int nk = P.ph[ph].ntaps * P.kchunks;
            const int acc = local & 1;
            mbar_wait(tempty + acc, ((local >> 1) & 1) ^ 1);           // epilogue has drained this accumulator
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
            for (int kb = 0; kb < nk; ++kb) {
                mbar_wait(full + stage, phase);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t a_addr = smem_u32(smem + stage * L::kStageBytes);
                    const uint64_t adesc = make_kmajor_desc<BK * 2>(a_addr);
                    const uint64_t bdesc = make_kmajor_desc<BK * 2>(a_addr + L::kABytes);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k)                    // +32 B per K=16 step inside the swizzle atom
                        umma_bf16(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0);
                    umma_commit(empty + stage);                          // frees the smem stage when these MMAs retire
                    if (kb == nk - 1) umma_commit(tfull + acc);          // accumulator complete -> epilogue
                }
                __syncwarp();
                if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        // ===================================== epilogue ============================================
        const int q = warp & 3;                        // TMEM lane quadrant this warp may access
        const int row = q * 32 + lane;
        const int j = row / TW, i = row - j * TW;
        const int ef = P.epi_flags;
        int local = 0;
        for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x, ++local) {
            int ph, b, ty, tx;
            decode(tile, ph, b, ty, tx);
            const Phase& F = P.ph[ph];
            const int acc = local & 1;
            const int oy = (ty * TH + j) * P.out_step + F.py, ox = (tx * TW + i) * P.out_step + F.px;
            const bool valid = (ty * TH + j) < F.Hph && (tx * TW + i) < F.Wph && oy < P.Hout && ox < P.Wout;
            const int64_t pix = ((int64_t)oy * P.Wout + ox) * P.Cout;
            const int64_t o_off = (int64_t)b * P.Hout * P.Wout * P.Cout + pix;
            mbar_wait(tfull + acc, (local >> 1) & 1);
            tc_fence_after();
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                uint32_t r[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c0), r);
                if (!valid || c0 >= P.Cout) continue;
                float v[32];
#pragma unroll
                for (int k = 0; k < 32; ++k) v[k] = __uint_as_float(r[k]) + s_bias[c0 + k];
                if (P.out_planar) {
                    // fp32 NCHW planes, Cout (<= 32) real channels: conv6 forward, conv1 / conv1_s backward-data
                    const int64_t hw = (int64_t)P.Hout * P.Wout;
                    const int64_t p = (int64_t)oy * P.Wout + ox;
                    float* op = (float*)P.out + (int64_t)b * P.Cout * hw + p;
                    const float* ap = P.add ? (const float*)P.add + (int64_t)b * P.add_bs + p : nullptr;
#pragma unroll
                    for (int k = 0; k < 32; ++k) {
                        if (k < P.Cout) {
                            float y = v[k];
                            if (ap) y += __ldg(ap + k * hw);
                            if (ef & SPAA_EPI_RELU) y = fmaxf(y, 0.f);
                            if (ef & SPAA_EPI_CLAMP_MAX1) y = fminf(y, 1.f);
                            op[k * hw] = y;
                        }
                    }
                    continue;
                }
                if (P.add) {
                    const uint4* ap = reinterpret_cast<const uint4*>((const uint16_t*)P.add + (int64_t)b * P.add_bs + pix + c0);
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const uint4 u = __ldg(ap + g);
                        const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float2 f = unpack2<F16>(w4[e]);
                            v[g * 8 + e * 2] += f.x; v[g * 8 + e * 2 + 1] += f.y;
                        }
                    }
                }
                if (ef & SPAA_EPI_RELU) {
#pragma unroll
                    for (int k = 0; k < 32; ++k) v[k] = fmaxf(v[k], 0.f);
                }
                if (P.mask) {                                  // backward of ReLU: keep the gradient where the activation was > 0
                    const uint4* mp = reinterpret_cast<const uint4*>(P.mask + (int64_t)b * P.mask_bs + pix + c0);
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const uint4 u = __ldg(mp + g);
                        const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            if (!pos16(w4[e] & 0xFFFFu)) v[g * 8 + e * 2] = 0.f;
                            if (!pos16(w4[e] >> 16)) v[g * 8 + e * 2 + 1] = 0.f;
                        }
                    }
                }
                uint4* op = reinterpret_cast<uint4*>((uint16_t*)P.out + o_off + c0);
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    uint4 u;
                    u.x = pack2<F16>(v[g * 8 + 0], v[g * 8 + 1]); u.y = pack2<F16>(v[g * 8 + 2], v[g * 8 + 3]);
                    u.z = pack2<F16>(v[g * 8 + 4], v[g * 8 + 5]); u.w = pack2<F16>(v[g * 8 + 6], v[g * 8 + 7]);
                    op[g] = u;
                }
                if (P.out2) {
                    const uint4* mp = reinterpret_cast<const uint4*>(P.mask2 + (int64_t)b * P.mask_bs + pix + c0);
                    uint4* op2 = reinterpret_cast<uint4*>((uint16_t*)P.out2 + o_off + c0);
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const uint4 m = __ldg(mp + g);
                        const uint32_t w4[4] = {m.x, m.y, m.z, m.w};
                        uint32_t o4[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            o4[e] = pack2<F16>(pos16(w4[e] & 0xFFFFu) ? v[g * 8 + e * 2] : 0.f, pos16(w4[e] >> 16) ? v[g * 8 + e * 2 + 1] : 0.f);
                        op2[g] = make_uint4(o4[0], o4[1], o4[2], o4[3]);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty + acc);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

// ===============================================================================================================
// v2: halo-tile kernel.  Measured on B200 (profiles/r1_tc_v1_ncu.md): v1 above is bound by the rate at which TMA delivers
// smem rows (~3 clk per <=128 B row, ~43 B/clk/SM): it fetches one shifted 8x16 box per filter TAP, so every input pixel is
// pulled from L2 nine times (conv6: 1.9 GB of L2->SM traffic for 158 MB of HBM reads).  Here the A operand of an output tile
// is fetched ONCE per 64-channel chunk as a (TH+KH-1) x (TW+KW-1) halo box; the MMA of tap (r,s) reads the same smem through
// a descriptor whose start address is shifted by (r*halo_w + s) rows.  That needs every 8-row group of the MMA (one core
// matrix row group) to be 8 CONSECUTIVE x-pixels, hence the 16 x 8 output tile (row m = j*8 + i) and SBO = halo_w rows.
//   stride-2 convolutions: four input-parity planes (TMA element stride 2), a tap reads plane (r&1, s&1) at shift (r>>1, s>>1);
//   up-sampling (transposed / backward-of-strided) convolutions: the four output-parity phases share one input halo tile
//   and accumulate into four TMEM column ranges.
// Weights: all taps of small layers stay resident in shared memory for the life of the CTA; large layers stream one
// (tap, 64-channel) slice per stage through a second ring.
// ===============================================================================================================
constexpr int HTH = 16, HTW = 8;                   // output tile: 16 rows x 8 columns = 128 MMA rows

struct HTap { int16_t shift; int8_t slot; int8_t plane; };
struct HPhase { int32_t ntaps, py, px, pad_; HTap taps[kMaxTaps]; };
struct HaloParams {
    int32_t B, Cin, Hin, Win, Cout, Hout, Wout;
    int32_t up, stride;
    int32_t nphases, nplanes;
    int32_t halo_w, halo_h;            // pixels per plane box
    int32_t org_x, org_y;              // plane origin relative to the tile origin, in plane pixels
    int32_t tiles_x, tiles_y, total_tiles, kchunks;
    int32_t a_plane_bytes, a_stage_bytes, a_tx_bytes, b_slice_bytes;
    int32_t sa, sb, resident, nslots, use_base_off, ntap_total;
    int32_t ctas_per_sm;
    int32_t egroups, nbuf;             // epilogue warp groups (1 or 2); accumulator buffers in TMEM (1, 2 or 4): tile number k of a CTA uses buffer k % nbuf
    int32_t pair;                      // 1: the MMA issuer works on TWO tiles per weight pass (see conv_halo_kernel); a_pair_bytes = offset of the second tile's A planes in a stage
    int32_t a_pair_bytes;
    int32_t e_stage_bytes;             // output staging blocks of the 4 epilogue warps (16-bit NHWC output only)
    int32_t e_dbuf;                    // 1: two staging blocks per warp (epilogue_nhwc16)
    int32_t e_slots, e_nops;           // epilogue operand ring: e_slots slots per group (0: no ring) of e_slot_bytes = e_nops x 8 KB (2 KB for the fp32-planar residual)
    int32_t e_slot_bytes;
    uint32_t tap_tab[kMaxTaps * kMaxPhases];     // flattened (phase, tap) list, see the MMA issuer
    int32_t epi_flags, out_planar;
    int32_t split;                     // 1: bf16x3 split-precision operands (see the SPLIT template parameter of conv_halo_kernel)
    int32_t out_ps;                    // pixel stride (elements) of out / out2 / add / mask / mask2 for 16-bit NHWC output: Cout, or 3 * Cout when split
    int32_t a_chunk[kMaxKChunks];      // first input channel of K chunk kc (TMA coordinate): kc * BK, or the part table of the split mode
    int64_t add_bs, mask_bs;
    HPhase ph[kMaxPhases];
    const float* bias;
    const void* add;
    const uint16_t* mask;
    const uint16_t* mask2;
    void* out;
    void* out2;
};

template <int ROW_BYTES> SPAA_D uint64_t make_halo_desc(uint32_t smem_addr, uint32_t sbo_bytes, uint32_t use_base_off) {
    constexpr uint64_t layout = ROW_BYTES == 128 ? 2 : (ROW_BYTES == 64 ? 4 : 6);
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    if (use_base_off) d |= (uint64_t)((smem_addr >> 7) & 7) << 49;
    d |= layout << 61;
    return d;
}

// Tensor maps of the 16-bit NHWC outputs, one per output phase (the sub-lattice (py, px) of an up-sampling layer is its own strided view):
// m[ph] = out, m[4 + ph] = out2.  Box = {32 channels, 8 x, 4 y, 1 image} = the 2 KB staging block of one epilogue warp, 64-byte swizzle.
struct alignas(64) HaloOutMaps { CUtensorMap m[2 * kMaxPhases]; };
struct HaloNoMaps { int32_t unused; };             // kernels that keep the round-1 epilogue take this instead (1 KB less of launch parameters)
// which instantiations run the register-lean epilogue (see conv_halo_kernel)
template <int BN, int BK, bool SPLIT> constexpr bool halo_lean_v = !SPLIT && BN <= 64 && !(BN == 64 && BK == 64);
template <int BN, int BK, bool SPLIT> using HaloMapsT = std::conditional_t<halo_lean_v<BN, BK, SPLIT>, HaloOutMaps, HaloNoMaps>;

// ---------------------------------------------------------------------------------------------------------------
// Epilogue of the 16-bit NHWC layers (every mode but the split-precision one), round 2.
// The round-1 epilogue spent ~250-390 warp instructions per (32 rows x 32 channels) unit -- 64-bit index arithmetic recomputed per unit,
// 32 generic loads for the bias, a staged LDS + STG copy with per-row predicates -- on 8 epilogue warps per SM whose dependent chains ran at
// ~8 clk per instruction (ncu source page, profiles/r2_epilogue_sass.md): the narrow layers reached 0.35 of the HBM roofline with every pipe idle.
// This version: 32-bit element offsets with everything tile-invariant hoisted, bias through LDS.128, operands read from the ring 8 channels at
// a time right where they are used (half the registers, so more epilogue warps per SM), and the output block leaves through ONE TMA tile store
// per unit issued by lane 0 (edge clipping by the tensor map, no per-row predicates, no L1 round trip).  The accumulator buffer is released right
// after the last tcgen05.ld of the tile, not after its stores.
// ---------------------------------------------------------------------------------------------------------------
template <int BN, bool F16>
SPAA_D void epilogue_nhwc16(const HaloParams& P, const HaloOutMaps& OM, uint32_t tmem_base, uint32_t e_ring, uint32_t bias_u32, uint64_t* tfull,
                            uint64_t* tempty, int warp, int lane) {
    const int EG = P.egroups, NPH = P.nphases;
    const int eg = (warp - 2) >> 2, q = warp & 3;
    if (eg >= EG) return;
    const uint32_t acc_cols = (uint32_t)(NPH * BN);
    const int nacc_mask = P.nbuf - 1;                               // nbuf is 1, 2 or 4
    const int nacc_shift = P.nbuf == 4 ? 2 : (P.nbuf == 2 ? 1 : 0);
    const int per_img = P.tiles_x * P.tiles_y;
    const bool has_add = P.add != nullptr, has_mask = P.mask != nullptr, has_out2 = P.out2 != nullptr, has_bias = P.bias != nullptr;
    const bool has_m2 = P.mask2 != nullptr;                          // out2 = out * (mask2 > 0), or (SPAA_EPI_OUT2_BF16, no mask2) the bf16 rounding of the result
    const bool relu = (P.epi_flags & SPAA_EPI_RELU) != 0;
    constexpr int NCH = BN / 32;
    const int S = P.e_slots;                                        // 0: no operand ring
    const uint32_t slot_bytes = (uint32_t)P.e_slot_bytes;
    const int ci = lane >> 2, cc = lane & 3;                        // cooperative copy mapping: x pixel of the row group, 16-byte chunk
    const uint32_t e_warp = e_ring + (uint32_t)(eg * S) * slot_bytes + (uint32_t)q * 2048u;
    uint32_t own_off[4], coop_off[4];                               // byte offsets inside a 2 KB warp block: [row][16-byte chunk ^ ((row >> 1) & 3)]
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        own_off[g] = (uint32_t)lane * 64u + (uint32_t)((g ^ ((lane >> 1) & 3)) * 16);
        const int rk = 8 * g + ci;
        coop_off[g] = (uint32_t)rk * 64u + (uint32_t)((cc ^ ((rk >> 1) & 3)) * 16);
    }
    // output staging: TWO blocks per warp ([out | out2] each), so that the TMA store of one unit is still reading while the next unit is written
    const uint32_t stage_blk = has_out2 ? 4096u : 2048u, stage_flip = P.e_dbuf ? stage_blk : 0u;
    const uint32_t o_stage = e_ring + (uint32_t)(EG * S) * slot_bytes + (uint32_t)(eg * 4 + q) * (P.e_dbuf ? 2u : 1u) * stage_blk;
    uint32_t stage_sel = 0;
    const uint32_t mask_slot_off = has_add ? 8192u : 0u, mask2_slot_off = mask_slot_off + (has_mask ? 8192u : 0u);
    // geometry in 32-bit ELEMENT offsets (the host routes tensors of >= 2^31 elements elsewhere)
    const int up = P.up;
    const uint32_t ps = (uint32_t)P.out_ps;
    const uint32_t row_el = (uint32_t)P.Wout * ps;
    const uint32_t kstep = (uint32_t)up * row_el;
    const uint32_t add_bs = (uint32_t)P.add_bs, mask_bs = (uint32_t)P.mask_bs;
    const uint16_t* const add_p = (const uint16_t*)P.add;

    const int tstep = EG * (int)gridDim.x, tile0 = (int)blockIdx.x + eg * (int)gridDim.x;
    // tile -> (image, tile row, tile column) without per-tile divisions: a group's tiles are tstep apart
    const int g_db = tstep / per_img, g_rem = tstep - g_db * per_img;
    const int g_dty = g_rem / P.tiles_x, g_dtx = g_rem - g_dty * P.tiles_x;
    auto tile_step = [&](int& tb, int& tty, int& ttx) {
        ttx += g_dtx;
        if (ttx >= P.tiles_x) { ttx -= P.tiles_x; ++tty; }
        tty += g_dty;
        if (tty >= P.tiles_y) { tty -= P.tiles_y; ++tb; }
        tb += g_db;
    };
    const int b0 = tile0 / per_img, t0 = tile0 - b0 * per_img;
    const int ty0 = t0 / P.tiles_x, tx0 = t0 - ty0 * P.tiles_x;

    // operand prefetch: unit = (tile, phase, 32-channel chunk), S - 1 units ahead of the unit being written, one cp.async group per unit
    int pf_tile = tile0, pf_ph = 0, pf_c = 0, pf_slot = 0, pf_b = b0, pf_ty = ty0, pf_tx = tx0;
    auto issue = [&]() {
        if (pf_tile < P.total_tiles) {
            const int cox = (pf_tx * HTW + ci) * up + (pf_ph & 1);
            const int coy0 = (pf_ty * HTH + q * 4) * up + (pf_ph >> 1);
            if (cox < P.Wout && pf_c * 32 < P.Cout) {
                const uint32_t pix = (uint32_t)coy0 * row_el + (uint32_t)cox * ps + (uint32_t)(pf_c * 32 + cc * 8);
                const int nrow = P.Hout - coy0;                     // row k of the warp's four exists iff k * up < nrow
                uint32_t d = e_warp + (uint32_t)pf_slot * slot_bytes;
                if (has_add) {
                    const uint16_t* sp = add_p + ((uint32_t)pf_b * add_bs + pix);
#pragma unroll
                    for (int k = 0; k < 4; ++k) if (k * up < nrow) cp_async16(d + coop_off[k], sp + k * kstep);
                    d += 8192;
                }
                if (has_mask) {
                    const uint16_t* sp = P.mask + ((uint32_t)pf_b * mask_bs + pix);
#pragma unroll
                    for (int k = 0; k < 4; ++k) if (k * up < nrow) cp_async16(d + coop_off[k], sp + k * kstep);
                    d += 8192;
                }
                if (has_m2) {
                    const uint16_t* sp = P.mask2 + ((uint32_t)pf_b * mask_bs + pix);
#pragma unroll
                    for (int k = 0; k < 4; ++k) if (k * up < nrow) cp_async16(d + coop_off[k], sp + k * kstep);
                }
            }
            if (++pf_c == NCH) {
                pf_c = 0;
                if (++pf_ph == NPH) { pf_ph = 0; pf_tile += tstep; tile_step(pf_b, pf_ty, pf_tx); }
            }
            if (++pf_slot == S) pf_slot = 0;
        }
        cp_async_commit();
    };
    if (S) for (int d = 0; d < S - 1; ++d) issue();
    int cs = 0;                                                     // ring slot of the unit being consumed

    int local = eg;
    int b = b0, ty = ty0, tx = tx0;
    for (int tile = tile0; tile < P.total_tiles; tile += tstep, local += EG, tile_step(b, ty, tx)) {
        const int acc = local & nacc_mask;
        const uint32_t tpar = (uint32_t)((local >> nacc_shift) & 1);
        const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * acc_cols;
        const int x0 = tx * HTW, y0 = ty * HTH + q * 4;            // the warp's 8 x 4 pixel block in the phase's (sub-sampled) output view
        bool waited = false;
        for (int ph = 0; ph < NPH; ++ph) {
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                uint32_t sa16 = e_warp + (uint32_t)cs * slot_bytes;
                if (S) {
                    cp_async_wait(S - 2);                          // this unit's operands have landed (every lane waits for its own copies) ...
                    __syncwarp();                                   // ... and everybody's; all lanes are also done reading the previous unit's slot,
                    issue();                                        // which the prefetch S - 1 units ahead now overwrites
                    if (++cs == S) cs = 0;
                }
                if (!waited) { mbar_wait(tfull + acc, tpar); tc_fence_after(); waited = true; }
                uint32_t r[32];
                tmem_ld32(t_row + (uint32_t)(ph * BN + c0), r);
                if (ph == NPH - 1 && c0 == BN - 32) {              // last read of this accumulator buffer: hand it back to the MMA issuer
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tempty + acc);
                }
                if (c0 >= P.Cout) continue;
                if (lane == 0) {                                    // the TMA store that last read THIS staging block is done with it
                    if (stage_flip) bulk_wait_read1(); else bulk_wait_read0();
                }
                __syncwarp();
                const uint32_t stg = o_stage + stage_sel;
                uint4 ca, cm, cm2;                                  // operands of the 8-channel group being processed, fetched one group ahead
                ca = cm = cm2 = make_uint4(0, 0, 0, 0);
                if (has_add) ca = lds128(sa16 + own_off[0]);
                if (has_mask) cm = lds128(sa16 + mask_slot_off + own_off[0]);
                if (has_m2) cm2 = lds128(sa16 + mask2_slot_off + own_off[0]);
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    uint4 na, nm, nm2;
                    na = nm = nm2 = make_uint4(0, 0, 0, 0);
                    if (g < 3) {
                        if (has_add) na = lds128(sa16 + own_off[g + 1]);
                        if (has_mask) nm = lds128(sa16 + mask_slot_off + own_off[g + 1]);
                        if (has_m2) nm2 = lds128(sa16 + mask2_slot_off + own_off[g + 1]);
                    }
                    float v[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) v[k] = __uint_as_float(r[g * 8 + k]);
                    if (has_bias) {
                        const uint4 ba = lds128_const(bias_u32 + (uint32_t)(c0 + g * 8) * 4u), bb = lds128_const(bias_u32 + (uint32_t)(c0 + g * 8 + 4) * 4u);
                        v[0] += __uint_as_float(ba.x); v[1] += __uint_as_float(ba.y); v[2] += __uint_as_float(ba.z); v[3] += __uint_as_float(ba.w);
                        v[4] += __uint_as_float(bb.x); v[5] += __uint_as_float(bb.y); v[6] += __uint_as_float(bb.z); v[7] += __uint_as_float(bb.w);
                    }
                    if (has_add) {
                        const uint32_t w4[4] = {ca.x, ca.y, ca.z, ca.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float2 f = unpack2<F16>(w4[e]);
                            v[e * 2] += f.x; v[e * 2 + 1] += f.y;
                        }
                    }
                    if (relu) {
#pragma unroll
                        for (int k = 0; k < 8; ++k) v[k] = fmaxf(v[k], 0.f);
                    }
                    // pack to 16 bit, then apply the ReLU masks of the backward pass on the packed pairs (AND with a per-half "> 0" bit mask:
                    // the same bits as selecting 0.f before the conversion)
                    uint4 pk = make_uint4(pack2<F16>(v[0], v[1]), pack2<F16>(v[2], v[3]), pack2<F16>(v[4], v[5]), pack2<F16>(v[6], v[7]));
                    if (has_mask) { pk.x &= posmask2(cm.x); pk.y &= posmask2(cm.y); pk.z &= posmask2(cm.z); pk.w &= posmask2(cm.w); }
                    if (has_out2) {
                        if (has_m2)
                            sts128(stg + 2048u + own_off[g], make_uint4(pk.x & posmask2(cm2.x), pk.y & posmask2(cm2.y), pk.z & posmask2(cm2.z), pk.w & posmask2(cm2.w)));
                        else
                            sts128(stg + 2048u + own_off[g], make_uint4(pack2<false>(v[0], v[1]), pack2<false>(v[2], v[3]), pack2<false>(v[4], v[5]), pack2<false>(v[6], v[7])));
                    }
                    sts128(stg + own_off[g], pk);
                    ca = na; cm = nm; cm2 = nm2;
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    tma_store_4d(&OM.m[ph], stg, c0, x0, y0, b);
                    if (has_out2) tma_store_4d(&OM.m[kMaxPhases + ph], stg + 2048u, c0, x0, y0, b);
                    bulk_commit();
                }
                stage_sel ^= stage_flip;
            }
        }
    }
    cp_async_wait(0);
    if (lane == 0) bulk_wait0();
}

// SPLIT (bf16 only): "fp32-accurate" split-precision mode.  Every fp32 activation / gradient / weight value v is stored as THREE bf16 numbers
// h = bf16(v), m = bf16(v - h), l = bf16(v - h - m)  (3 x 8 = 24 significand bits, fp32's exponent range), an NHWC tensor of logical C channels
// as [h(C) | m(C) | l(C)] = 3C physical channels.  A product v*w is the sum of the six part products hh + hm + mh + mm + hl + lh (the dropped
// ml + lm + ll are < 2^-23 relative), i.e. the SAME implicit GEMM with a six times longer K: the producer walks the input's parts through the
// chunk table P.a_chunk, the weights are packed with 6C input channels in the matching order (smallest products first, so that the fp32
// accumulator in TMEM holds small values while the small terms arrive), and the MMA issuer is unchanged.  The epilogue sums the three parts of
// the residual operand in fp32, and writes its fp32 result v as three bf16 parts (masks act on every part; the sign of v is the sign of h).
template <int BN, int BK, bool F16, bool SPLIT, bool COPY2 = false>
__global__ void __launch_bounds__(BN >= 128 ? kThreads + 128 : ((SPLIT || (BN == 64 && BK == 64)) ? kThreads : 640), (BN <= 64 && (SPLIT || (BN == 64 && BK == 64))) ? 2 : 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const __grid_constant__ HaloParams P,
                 const __grid_constant__ HaloMapsT<BN, BK, SPLIT> OM) {
    static_assert(!(SPLIT && F16), "the split-precision mode stores bf16 parts");
    // narrow layers outside the split mode: register-lean epilogue with TMA stores (epilogue_nhwc16), up to four epilogue groups.  The wide layers keep
    // the round-1 epilogue: measured on the same box, the TMA-store epilogue cost them 4-14 % (conv4_s forward 74.0 -> 84.8 us: with one staging block
    // per warp every chunk waits for the previous store to leave shared memory, and a second block costs a stage of the weight ring).
    // (BN = 64 with 64-channel K chunks = the 128 -> 64-channel layers, whose weights are streamed: same-box, conv3 backward 33 -> 39 us with it)
    constexpr bool LEAN = halo_lean_v<BN, BK, SPLIT>;
    constexpr int NPART = SPLIT ? 3 : 1;
    constexpr int ROWB = BK * 2;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int SA = P.sa, SB = P.sb;
    uint8_t* a_ring = smem;
    uint8_t* b_base = smem + (size_t)SA * P.a_stage_bytes;
    const size_t b_bytes = P.resident ? (size_t)P.kchunks * P.nslots * P.b_slice_bytes : (size_t)SB * P.b_slice_bytes;
    uint8_t* e_ring = b_base + b_bytes;            // epilogue operand ring (see the epilogue)
    uint64_t* bars = (uint64_t*)(e_ring + (size_t)P.egroups * P.e_slots * P.e_slot_bytes + P.e_stage_bytes);
    uint64_t* a_full = bars;                       // [SA]
    uint64_t* a_empty = a_full + 8;                // [SA]
    uint64_t* b_full = a_empty + 8;                // [SB] (b_full[0] doubles as the "resident weights loaded" barrier)
    uint64_t* b_empty = b_full + 8;                // [SB]
    uint64_t* tfull = b_empty + 8;                 // [4]
    uint64_t* tempty = tfull + 4;                  // [4]
    uint32_t* tmem_slot = (uint32_t*)(tempty + 4);
    uint32_t* s_tap = tmem_slot + 4;               // [kMaxTaps * kMaxPhases]
    float* s_bias = (float*)(s_tap + kMaxTaps * kMaxPhases);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int NPH = P.nphases;
    for (int i = threadIdx.x; i < P.ntap_total; i += blockDim.x) s_tap[i] = P.tap_tab[i];
    const uint32_t acc_cols = (uint32_t)(NPH * BN);                  // TMEM columns of one accumulator buffer
    const int NACC = P.nbuf;                                         // accumulator buffers: the epilogue of tile i overlaps the MMAs of the tiles after it
    uint32_t tmem_cols = 32;
    while (tmem_cols < (uint32_t)P.nbuf * acc_cols) tmem_cols <<= 1;
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&map_a);
        prefetch_tmap(&map_b);
        for (int i = 0; i < 8; ++i) { mbar_init(a_full + i, 1); mbar_init(a_empty + i, 1); mbar_init(b_full + i, 1); mbar_init(b_empty + i, 1); }
        for (int i = 0; i < 4; ++i) { mbar_init(tfull + i, 1); mbar_init(tempty + i, 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
    // PDL: the next kernel of the stream may start its own prologue as soon as this grid's CTAs free their SMs; this grid's first access to global
    // memory (bias, then the TMA loads / operand copies / stores of the three roles) waits for the previous grid to complete
    pdl_launch_dependents();
    pdl_wait();
    for (int i = threadIdx.x; i < BN; i += blockDim.x) s_bias[i] = (P.bias && i < P.Cout) ? P.bias[i] : 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int per_img = P.tiles_x * P.tiles_y;

    if (warp == 0) {
        // ===================================== TMA producer ========================================
        if (lane == 0) {
            if (P.resident) {
                mbar_expect_tx(b_full, (uint32_t)b_bytes);
                for (int kc = 0; kc < P.kchunks; ++kc)
                    for (int sl = 0; sl < P.nslots; ++sl)
                        tma_load_2d(b_base + (size_t)(kc * P.nslots + sl) * P.b_slice_bytes, &map_b, b_full, kc * BK, sl * BN);
            }
            int sa = 0, sb = 0;
            uint32_t pa = 0, pb = 0;
            if (P.resident && !P.pair && P.kchunks == 1) {
                // narrow layers: ONE box set per tile and nothing else, so the per-tile cost of this loop is what bounds the layer once the epilogue
                // keeps up -- tile coordinates advance incrementally (the two integer divisions per tile were ~70 of its ~130 instructions)
                const int step = (int)gridDim.x;
                const int s_db = step / per_img, s_rem = step - s_db * per_img;
                const int s_dty = s_rem / P.tiles_x, s_dtx = s_rem - s_dty * P.tiles_x;
                int tb = (int)blockIdx.x / per_img;
                const int t0 = (int)blockIdx.x - tb * per_img;
                int tty = t0 / P.tiles_x, ttx = t0 - tty * P.tiles_x;
                const int ch0 = P.a_chunk[0], npl = P.nplanes;
                const uint32_t tx_bytes = (uint32_t)P.a_tx_bytes;
                for (int tile = blockIdx.x; tile < P.total_tiles; tile += step) {
                    const int xx = (ttx * HTW + P.org_x) * P.stride, yy = (tty * HTH + P.org_y) * P.stride;
                    mbar_wait(a_empty + sa, pa ^ 1);
                    uint8_t* dst = a_ring + (size_t)sa * P.a_stage_bytes;
                    mbar_expect_tx(a_full + sa, tx_bytes);
                    tma_load_4d(dst, &map_a, a_full + sa, ch0, xx, yy, tb);
                    if (npl == 4) {
                        tma_load_4d(dst + (size_t)P.a_plane_bytes, &map_a, a_full + sa, ch0, xx + 1, yy, tb);
                        tma_load_4d(dst + 2 * (size_t)P.a_plane_bytes, &map_a, a_full + sa, ch0, xx, yy + 1, tb);
                        tma_load_4d(dst + 3 * (size_t)P.a_plane_bytes, &map_a, a_full + sa, ch0, xx + 1, yy + 1, tb);
                    }
                    if (++sa == SA) { sa = 0; pa ^= 1; }
                    ttx += s_dtx;
                    if (ttx >= P.tiles_x) { ttx -= P.tiles_x; ++tty; }
                    tty += s_dty;
                    if (tty >= P.tiles_y) { tty -= P.tiles_y; ++tb; }
                    tb += s_db;
                }
            } else {
            // NT tiles per pipeline step: in pair mode one stage holds the halo boxes of TWO tiles (this CTA's tiles number 2k and 2k+1), and every
            // weight slice that arrives is used for both (see the MMA issuer)
            const int NT = P.pair ? 2 : 1;
            for (int tile = blockIdx.x; tile < P.total_tiles; tile += NT * (int)gridDim.x) {
                int bb[2], xx0[2], yy0[2];
                int nt = 0;
                for (int j = 0; j < NT; ++j) {
                    const int tj = tile + j * (int)gridDim.x;
                    if (tj >= P.total_tiles) break;
                    const int b = tj / per_img;
                    const int t = tj - b * per_img;
                    const int ty = t / P.tiles_x, tx = t - ty * P.tiles_x;
                    bb[j] = b; xx0[j] = (tx * HTW + P.org_x) * P.stride; yy0[j] = (ty * HTH + P.org_y) * P.stride;
                    ++nt;
                }
                for (int kc = 0; kc < P.kchunks; ++kc) {
                    mbar_wait(a_empty + sa, pa ^ 1);
                    uint8_t* dst = a_ring + (size_t)sa * P.a_stage_bytes;
                    mbar_expect_tx(a_full + sa, (uint32_t)(P.a_tx_bytes * nt));
                    const int ch0 = P.a_chunk[kc];
                    for (int j = 0; j < nt; ++j)
                        for (int pl = 0; pl < P.nplanes; ++pl)
                            tma_load_4d(dst + (size_t)j * P.a_pair_bytes + (size_t)pl * P.a_plane_bytes, &map_a, a_full + sa, ch0, xx0[j] + (pl & 1), yy0[j] + (pl >> 1), bb[j]);
                    if (++sa == SA) { sa = 0; pa ^= 1; }
                    if (!P.resident) {
                        for (int ph = 0; ph < NPH; ++ph)
                            for (int tp = 0; tp < P.ph[ph].ntaps; ++tp) {
                                mbar_wait(b_empty + sb, pb ^ 1);
                                mbar_expect_tx(b_full + sb, (uint32_t)P.b_slice_bytes);
                                tma_load_2d(b_base + (size_t)sb * P.b_slice_bytes, &map_b, b_full + sb, kc * BK, P.ph[ph].taps[tp].slot * BN);
                                if (++sb == SB) { sb = 0; pb ^= 1; }
                            }
                    }
                }
            }
            }
        }
    } else if (warp == 1) {
        // ===================================== MMA issuer ==========================================
        // ONE elected thread runs the whole loop.  The first version of this loop (every lane evaluating the waits, lane 0
        // building both 64-bit descriptors per MMA from the tap table in the constant bank) spent ~300 clk of issue overhead
        // per tcgen05.mma -- more than the 16..128 clk the MMA itself needs -- and was THE bottleneck of every layer
        // (profiles/r1_tc_issue_bound.md).  Now: descriptor high words are loop constants, low words are a base plus a
        // per-tap 16-byte-unit offset read from a small shared-memory table.
        if (elect_one()) {
            const uint32_t idesc = make_idesc(BM, BN, F16);
            constexpr uint32_t layout = ROWB == 128 ? 2u : (ROWB == 64 ? 4u : 6u);
            const uint32_t a_hi = (((uint32_t)(P.halo_w * ROWB) >> 4) & 0x3FFFu) | (1u << 14) | (layout << 29);
            const uint32_t b_hi = (((uint32_t)(8 * ROWB) >> 4) & 0x3FFFu) | (1u << 14) | (layout << 29);
            const uint32_t a_ring_lo = smem_u32(a_ring) >> 4, b_base_lo = smem_u32(b_base) >> 4;
            const uint32_t a_stage16 = (uint32_t)P.a_stage_bytes >> 4, b_slice16 = (uint32_t)P.b_slice_bytes >> 4;
            const int ntap = P.ntap_total;                   // <= kMaxTaps: each filter tap belongs to exactly one output phase
            uint32_t taps[kMaxTaps];
#pragma unroll
            for (int e = 0; e < kMaxTaps; ++e) taps[e] = e < ntap ? s_tap[e] : 0u;
            int sa = 0, sb = 0;
            uint32_t pa = 0, pb = 0;
            int local = 0;                                   // this CTA's tile counter: tile k uses accumulator buffer k % NACC, for the (k / NACC)-th time
            if (P.resident) { mbar_wait(b_full, 0); tc_fence_after(); }
            // Narrow layers (weights resident, one K chunk, no pair mode): a 128-pixel tile is 4-18 small MMAs, and the generic loop below spent ~23
            // instructions per tap on run-time tap counts, the resident / streamed branch and re-materialised constants (~270 per tile at ~6 clk each:
            // as long as the loads and the MMAs themselves).  Here the tap count is a compile-time constant and every per-tap quantity is a register.
            bool fast_done = false;
            if constexpr (BN <= 64 && !(BN == 64 && BK == 64)) {      // (the 128 -> 64-channel layers always stream their weights: no dead code in their kernels)
                if (P.resident && !P.pair && P.kchunks == 1) {
                    auto fast = [&](auto ntap_c) {
                        constexpr int NTAP = decltype(ntap_c)::value;
                        uint32_t ab[NTAP], df[NTAP];          // packed per-tap constants: A offset | weight slice address << 16;  TMEM column offset | first-of-phase << 16
#pragma unroll
                        for (int e = 0; e < NTAP; ++e) {
                            const uint32_t te = taps[e];
                            ab[e] = (te & 0xFFFFu) | ((b_base_lo + ((te >> 16) & 15u) * b_slice16) << 16);
                            df[e] = (((te >> 20) & 3u) * (uint32_t)BN) | (((te >> 22) & 1u) << 16);
                        }
                        const int nacc_mask = NACC - 1, nacc_shift = NACC == 4 ? 2 : (NACC == 2 ? 1 : 0);
                        for (int tile = blockIdx.x; tile < P.total_tiles; tile += (int)gridDim.x, ++local) {
                            const int acc = local & nacc_mask;
                            mbar_wait(tempty + acc, (uint32_t)(((local >> nacc_shift) & 1) ^ 1));
                            const uint32_t d_base = tmem_base + (uint32_t)acc * acc_cols;
                            mbar_wait(a_full + sa, pa);
                            tc_fence_after();
                            const uint32_t a_lo = a_ring_lo + (uint32_t)sa * a_stage16;
#pragma unroll
                            for (int e = 0; e < NTAP; ++e) {
                                const uint32_t at = a_lo + (ab[e] & 0xFFFFu), bt = ab[e] >> 16;
                                const uint32_t dt = d_base + (df[e] & 0xFFFFu), fresh = df[e] >> 16;
#pragma unroll
                                for (int k = 0; k < BK / 16; ++k)
                                    umma_bf16(dt, ((uint64_t)a_hi << 32) | (uint64_t)(at + 2 * k), ((uint64_t)b_hi << 32) | (uint64_t)(bt + 2 * k), idesc,
                                              (k == 0 && fresh) ? 0u : 1u);
                            }
                            umma_commit(a_empty + sa);
                            umma_commit(tfull + acc);
                            if (++sa == SA) { sa = 0; pa ^= 1; }
                        }
                    };
                    fast_done = true;
                    if (ntap == 9) fast(std::integral_constant<int, 9>{});
                    else if (ntap == 4) fast(std::integral_constant<int, 4>{});
                    else if (ntap == 1) fast(std::integral_constant<int, 1>{});
                    else fast_done = false;
                }
            }
            if (!fast_done) {
            // Pair mode (wide layers whose weights do not fit in shared memory: the K = 1152 / 2304 layers stream 590 KB of weights per 128-pixel
            // tile from L2, 85 % of the kernel's L2->SM bytes, and ran at the L2->SM bandwidth (~10 TB/s, profiles/r2_halo_ncu_full.md) with the tensor
            // pipe 63 % active): every weight slice that lands is multiplied with the A views of TWO tiles, into two accumulator buffers -- an
            // M = 256 step without a CTA pair -- which halves the weight traffic per output pixel.
            const int NT = P.pair ? 2 : 1;
            const uint32_t pair16 = (uint32_t)P.a_pair_bytes >> 4;
            for (int tile = blockIdx.x; tile < P.total_tiles; tile += NT * (int)gridDim.x, local += NT) {
                const int nt = (NT == 2 && tile + (int)gridDim.x < P.total_tiles) ? 2 : 1;
                uint32_t d_base[2];
                int acc[2];
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    if (j < nt) {
                        const int lj = local + j;
                        acc[j] = lj % NACC;
                        mbar_wait(tempty + acc[j], (uint32_t)(((lj / NACC) & 1) ^ 1));
                        d_base[j] = tmem_base + (uint32_t)acc[j] * acc_cols;
                    }
                }
                tc_fence_after();
                for (int kc = 0; kc < P.kchunks; ++kc) {
                    mbar_wait(a_full + sa, pa);
                    tc_fence_after();
                    const uint32_t a_lo = a_ring_lo + (uint32_t)sa * a_stage16;
                    const uint32_t b_kc_lo = b_base_lo + (uint32_t)(kc * P.nslots) * b_slice16;
#pragma unroll
                    for (int e = 0; e < kMaxTaps; ++e) {        // tap table lives in registers (compile-time indices after unrolling)
                        if (e < ntap) {
                            const uint32_t te = taps[e];            // [0,16) A offset in 16 B units, [16,20) weight slot, [20,22) phase, bit 22: first tap of its phase
                            uint32_t b_lo;
                            if (P.resident) b_lo = b_kc_lo + ((te >> 16) & 15u) * b_slice16;
                            else {
                                mbar_wait(b_full + sb, pb);
                                tc_fence_after();
                                b_lo = b_base_lo + (uint32_t)sb * b_slice16;
                            }
                            const uint32_t fresh = (kc == 0 && (te & (1u << 22))) ? 1u : 0u;
#pragma unroll
                            for (int j = 0; j < 2; ++j) {
                                if (j < nt) {
                                    const uint32_t d_tmem = d_base[j] + ((te >> 20) & 3u) * (uint32_t)BN;
                                    const uint32_t at = a_lo + (uint32_t)j * pair16 + (te & 0xFFFFu);
#pragma unroll
                                    for (int k = 0; k < BK / 16; ++k)
                                        umma_bf16(d_tmem, ((uint64_t)a_hi << 32) | (uint64_t)(at + 2 * k), ((uint64_t)b_hi << 32) | (uint64_t)(b_lo + 2 * k), idesc,
                                                  (k == 0 && fresh) ? 0u : 1u);
                                }
                            }
                            if (!P.resident) {
                                umma_commit(b_empty + sb);
                                if (++sb == SB) { sb = 0; pb ^= 1; }
                            }
                        }
                    }
                    umma_commit(a_empty + sa);
                    if (kc == P.kchunks - 1) {
                        umma_commit(tfull + acc[0]);
                        if (nt == 2) umma_commit(tfull + acc[1]);
                    }
                    if (++sa == SA) { sa = 0; pa ^= 1; }
                }
            }
            }   // !fast_done
        }
        __syncwarp();
    } else {
        // ===================================== epilogue ============================================
        // Operands that do not depend on the accumulator (residual, ReLU masks) are streamed through a per-warp cp.async ring
        // in shared memory, S-1 (tile, phase, 32-channel chunk) units AHEAD of the unit being written.  The first version fetched
        // them one chunk ahead into registers, so every tile paid one exposed global-memory latency in series (conv6 backward:
        // 1.4 us per 128-pixel tile = 30 % of HBM roofline), and every lane touched its own pixel row: 32 cache lines per
        // 128-bit load / store instruction (ncu: 32 sectors per request, L1 wavefronts at 50-58 % of peak on conv5 backward).
        // Now a warp moves its 32 rows x 64 B cooperatively -- instruction k covers rows 8k..8k+7 (8 consecutive x pixels), four
        // lanes per row -- for the operand copies and, through a 2 KB staging block, for the output stores.  A warp copies and
        // reads only ITS OWN 32 rows, so cp.async.wait_group + __syncwarp is all the synchronisation the ring needs.
        // Block layout: [row][16-byte chunk c ^ ((row >> 1) & 3)] (conflict-free for the row-wise and the cooperative accesses).
        // EG epilogue groups of 4 warps (one warp per TMEM lane quarter): with EG == 2 (wide layers) group g owns the tiles with
        // local index = g (mod 2), i.e. accumulator buffer g, and has its own operand ring and staging blocks -- a lone warp per
        // scheduler issues one dependent instruction every ~5 clk, and the ~500-instruction chain of a 32-channel chunk was what
        // bounded every layer between the HBM-bound and the MMA-bound ones (conv3: MMA pipe idle 80 % of the time).
        bool lean_done = false;
        if constexpr (LEAN) {
            if (!P.out_planar) {                                     // the 16-bit NHWC outputs of the narrow layers (fp16 / bf16 modes)
                epilogue_nhwc16<BN, F16>(P, OM, tmem_base, smem_u32(e_ring), smem_u32(s_bias), tfull, tempty, warp, lane);
                lean_done = true;
            }
        }
        if (!lean_done) {
        const int EG = P.egroups;
        const int eg = (warp - 2) >> 2;
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const int j = row >> 3, i = row & 7;
        const int ef = P.epi_flags;
        const bool has_add = P.add != nullptr, has_mask = P.mask != nullptr, has_out2 = P.out2 != nullptr;
        const bool planar = P.out_planar != 0, has_bias = P.bias != nullptr;
        const bool has_m2 = P.mask2 != nullptr;                  // (out2 without mask2: SPAA_EPI_OUT2_BF16, the bf16 rounding of the result)
        const int tstep = EG * (int)gridDim.x, tile0 = (int)blockIdx.x + eg * (int)gridDim.x;
        constexpr int NCH = BN / 32;
        const int nchu = planar ? 1 : NCH;                       // prefetch units per (tile, phase)
        const int S = P.e_slots;                                 // 0: no ring (no operand / wide planar residual)
        const uint32_t slot_bytes = (uint32_t)P.e_slot_bytes;
        const int ci = lane >> 2, cc = lane & 3;                 // cooperative mapping: x pixel within the row group, 16-byte chunk
        const uint32_t e_group = smem_u32(e_ring) + (uint32_t)eg * (uint32_t)S * slot_bytes;      // this group's ring
        const uint32_t e_warp = e_group + (uint32_t)q * 2048u;
        uint32_t own_off[4], coop_off[4];                        // byte offsets inside a 2 KB warp block
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            own_off[g] = (uint32_t)lane * 64u + (uint32_t)((g ^ ((lane >> 1) & 3)) * 16);
            const int rk = 8 * g + ci;
            coop_off[g] = (uint32_t)rk * 64u + (uint32_t)((cc ^ ((rk >> 1) & 3)) * 16);
        }
        const uint32_t o_stage = smem_u32(e_ring) + (uint32_t)(EG * S) * slot_bytes + (uint32_t)(eg * 4 + q) * (has_out2 ? 4096u : 2048u);     // [out | out2] staging of this warp
        const uint32_t e_row4 = e_group + (uint32_t)row * 4u;                // fp32 planar residual: [channel][row]
        struct Ops { uint4 a[4], m[4], m2[4]; };

        // tile -> (image, tile row, tile column) without per-tile divisions: a CTA's tiles are gridDim.x apart
        const int g_db = tstep / per_img, g_rem = tstep - g_db * per_img;
        const int g_dty = g_rem / P.tiles_x, g_dtx = g_rem - g_dty * P.tiles_x;
        auto tile_step = [&](int& tb, int& tty, int& ttx) {
            ttx += g_dtx;
            if (ttx >= P.tiles_x) { ttx -= P.tiles_x; ++tty; }
            tty += g_dty;
            if (tty >= P.tiles_y) { tty -= P.tiles_y; ++tb; }
            tb += g_db;
        };
        const int b0 = tile0 / per_img, t0 = tile0 - b0 * per_img;
        const int ty0 = t0 / P.tiles_x, tx0 = t0 - ty0 * P.tiles_x;
        int pf_tile = tile0, pf_ph = 0, pf_c = 0, pf_slot = 0, pf_b = b0, pf_ty = ty0, pf_tx = tx0;
        auto issue = [&]() {
            if (pf_tile < P.total_tiles) {
                const int oy = (pf_ty * HTH + j) * P.up + P.ph[pf_ph].py, ox = (pf_tx * HTW + i) * P.up + P.ph[pf_ph].px;
                if (oy < P.Hout && ox < P.Wout) {
                    if (planar) {
                        const int64_t hw = (int64_t)P.Hout * P.Wout;
                        const float* ap = (const float*)P.add + (int64_t)pf_b * P.add_bs + (int64_t)oy * P.Wout + ox;
                        const uint32_t d = e_row4 + (uint32_t)pf_slot * slot_bytes;
#pragma unroll
                        for (int k = 0; k < 4; ++k) if (k < P.Cout) cp_async4(d + k * 512, ap + k * hw);
                    }
                }
                if (!planar && pf_c * 32 < P.Cout) {
                    const int cox = (pf_tx * HTW + ci) * P.up + P.ph[pf_ph].px;
                    const int coy0 = (pf_ty * HTH + q * 4) * P.up + P.ph[pf_ph].py;
                    if (cox < P.Wout) {
                        const int64_t pix0 = ((int64_t)coy0 * P.Wout + cox) * P.out_ps + pf_c * 32 + cc * 8;
                        const int64_t kstep = (int64_t)P.up * P.Wout * P.out_ps;
                        uint32_t d = e_warp + (uint32_t)pf_slot * slot_bytes;
                        if (has_add) {
#pragma unroll
                            for (int part = 0; part < NPART; ++part) {          // split mode: the residual's h, m and l parts (P.Cout channels apart)
                                const uint16_t* sp = (const uint16_t*)P.add + (int64_t)pf_b * P.add_bs + pix0 + part * P.Cout;
#pragma unroll
                                for (int k = 0; k < 4; ++k) if (coy0 + k * P.up < P.Hout) cp_async16(d + coop_off[k], sp + k * kstep);
                                d += 8192;
                            }
                        }
                        if (has_mask) {
                            const uint16_t* sp = P.mask + (int64_t)pf_b * P.mask_bs + pix0;
#pragma unroll
                            for (int k = 0; k < 4; ++k) if (coy0 + k * P.up < P.Hout) cp_async16(d + coop_off[k], sp + k * kstep);
                            d += 8192;
                        }
                        if (has_m2) {
                            const uint16_t* sp = P.mask2 + (int64_t)pf_b * P.mask_bs + pix0;
#pragma unroll
                            for (int k = 0; k < 4; ++k) if (coy0 + k * P.up < P.Hout) cp_async16(d + coop_off[k], sp + k * kstep);
                        }
                    }
                }
                if (++pf_c == nchu) {
                    pf_c = 0;
                    if (++pf_ph == NPH) { pf_ph = 0; pf_tile += tstep; tile_step(pf_b, pf_ty, pf_tx); }
                }
                if (++pf_slot == S) pf_slot = 0;
            }
            cp_async_commit();
        };
        if (S) for (int d = 0; d < S - 1; ++d) issue();
        int cs = 0;                                              // ring slot of the unit being consumed

        int local = eg;
        int b = b0, ty = ty0, tx = tx0;
        for (int tile = tile0; tile < P.total_tiles && eg < EG; tile += tstep, local += EG, tile_step(b, ty, tx)) {
            const int acc = local % NACC;
            const uint32_t tpar = (uint32_t)((local / NACC) & 1);
            const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * acc_cols;
            bool waited = false;
            for (int ph = 0; ph < NPH; ++ph) {
                const int oy = (ty * HTH + j) * P.up + P.ph[ph].py, ox = (tx * HTW + i) * P.up + P.ph[ph].px;
                const bool valid = oy < P.Hout && ox < P.Wout;
                if (planar) {
                    // fp32 NCHW planes, Cout (<= 32) real channels: conv6 forward, conv1 / conv1_s backward-data
                    const int64_t hw = (int64_t)P.Hout * P.Wout;
                    const int64_t p = (int64_t)oy * P.Wout + ox;
                    float* op = (float*)P.out + (int64_t)b * P.Cout * hw + p;
                    const float* ap = has_add ? (const float*)P.add + (int64_t)b * P.add_bs + p : nullptr;
                    float av[4] = {0.f, 0.f, 0.f, 0.f};
                    if (S) {
                        cp_async_wait(S - 2);
                        if (valid) {
                            const uint32_t sa4 = e_row4 + (uint32_t)cs * slot_bytes;
#pragma unroll
                            for (int k = 0; k < 4; ++k) if (k < P.Cout) av[k] = lds32f(sa4 + k * 512);
                        }
                        issue();
                        if (++cs == S) cs = 0;
                    } else if (valid && ap && P.Cout <= 4) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) if (k < P.Cout) av[k] = __ldg(ap + k * hw);
                    }
                    if (!waited) { mbar_wait(tfull + acc, tpar); tc_fence_after(); waited = true; }
                    uint32_t r[32];
                    tmem_ld32(t_row + (uint32_t)(ph * BN), r);
                    if (!valid) continue;
                    if (P.Cout <= 4) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            if (k < P.Cout) {
                                float y = __uint_as_float(r[k]) + s_bias[k] + av[k];
                                if (ef & SPAA_EPI_RELU) y = fmaxf(y, 0.f);
                                if (ef & SPAA_EPI_CLAMP_MAX1) y = fminf(y, 1.f);
                                op[k * hw] = y;
                            }
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < 32; ++k) {
                            if (k < P.Cout) {
                                float y = __uint_as_float(r[k]) + s_bias[k];
                                if (ap) y += __ldg(ap + k * hw);
                                if (ef & SPAA_EPI_RELU) y = fmaxf(y, 0.f);
                                if (ef & SPAA_EPI_CLAMP_MAX1) y = fminf(y, 1.f);
                                op[k * hw] = y;
                            }
                        }
                    }
                    continue;
                }
                if constexpr (!LEAN) {
                // cooperative store mapping of this (tile, phase): instruction k writes rows 8k..8k+7 of the warp
                const int cox = (tx * HTW + ci) * P.up + P.ph[ph].px;
                const int coy0 = (ty * HTH + q * 4) * P.up + P.ph[ph].py;
                const int64_t co_off = ((int64_t)b * P.Hout * P.Wout + (int64_t)coy0 * P.Wout + cox) * P.out_ps + cc * 8;
                const int64_t kstep = (int64_t)P.up * P.Wout * P.out_ps;
                const bool cox_ok = cox < P.Wout;
                // NOT unrolled: with the chunk loop unrolled the BN = 256 kernel was 10 400 SASS instructions and its epilogue warps
                // (one per scheduler, nothing to hide a miss behind) spent 39 % of their samples in instruction-fetch stalls (ncu)
#pragma unroll 1
                for (int c0 = 0; c0 < BN; c0 += 32) {
                    Ops cur;
                    float asum[SPLIT ? 32 : 1];                      // split mode: h + m + l of the residual, summed in fp32
                    if (S) {
                        cp_async_wait(S - 2);
                        __syncwarp();
                        if (c0 < P.Cout) {
                            uint32_t sa16 = e_warp + (uint32_t)cs * slot_bytes;
                            if (has_add) {
                                if constexpr (SPLIT) {
#pragma unroll
                                    for (int k = 0; k < 32; ++k) asum[k] = 0.f;
#pragma unroll
                                    for (int part = 0; part < 3; ++part) {
#pragma unroll
                                        for (int g = 0; g < 4; ++g) {
                                            const uint4 t = lds128(sa16 + own_off[g]);
                                            const uint32_t w4[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
                                            for (int e = 0; e < 4; ++e) {
                                                const float2 f = unpack2<false>(w4[e]);
                                                asum[g * 8 + e * 2] += f.x; asum[g * 8 + e * 2 + 1] += f.y;
                                            }
                                        }
                                        sa16 += 8192;
                                    }
                                } else {
#pragma unroll
                                    for (int g = 0; g < 4; ++g) cur.a[g] = lds128(sa16 + own_off[g]);
                                    sa16 += 8192;
                                }
                            }
                            if (has_mask) {
#pragma unroll
                                for (int g = 0; g < 4; ++g) cur.m[g] = lds128(sa16 + own_off[g]);
                                sa16 += 8192;
                            }
                            if (has_m2) {
#pragma unroll
                                for (int g = 0; g < 4; ++g) cur.m2[g] = lds128(sa16 + own_off[g]);
                            }
                        }
                        issue();
                        if (++cs == S) cs = 0;
                    }
                    if (!waited) { mbar_wait(tfull + acc, tpar); tc_fence_after(); waited = true; }
                    uint32_t r[32];
                    tmem_ld32(t_row + (uint32_t)(ph * BN + c0), r);
                    if (c0 < P.Cout) {
                        float v[32];
                        if (has_bias) {
#pragma unroll
                            for (int k = 0; k < 32; ++k) v[k] = __uint_as_float(r[k]) + s_bias[c0 + k];
                        } else {
#pragma unroll
                            for (int k = 0; k < 32; ++k) v[k] = __uint_as_float(r[k]);
                        }
                        if (has_add) {
                            if constexpr (SPLIT) {
#pragma unroll
                                for (int k = 0; k < 32; ++k) v[k] += asum[k];
                            } else {
#pragma unroll
                                for (int g = 0; g < 4; ++g) {
                                    const uint32_t w4[4] = {cur.a[g].x, cur.a[g].y, cur.a[g].z, cur.a[g].w};
#pragma unroll
                                    for (int e = 0; e < 4; ++e) {
                                        const float2 f = unpack2<F16>(w4[e]);
                                        v[g * 8 + e * 2] += f.x; v[g * 8 + e * 2 + 1] += f.y;
                                    }
                                }
                            }
                        }
                        if (ef & SPAA_EPI_RELU) {
#pragma unroll
                            for (int k = 0; k < 32; ++k) v[k] = fmaxf(v[k], 0.f);
                        }
                        if constexpr (SPLIT) {
                            // three bf16 parts of the fp32 result, each staged and stored like the single 16-bit output below; the ReLU masks
                            // (sign of the mask operand's h part) zero every part
                            uint32_t mk[16], mk2[16];
                            if (has_mask) {
#pragma unroll
                                for (int g = 0; g < 4; ++g) {
                                    mk[g * 4 + 0] = posmask2(cur.m[g].x); mk[g * 4 + 1] = posmask2(cur.m[g].y);
                                    mk[g * 4 + 2] = posmask2(cur.m[g].z); mk[g * 4 + 3] = posmask2(cur.m[g].w);
                                }
                            }
                            if (has_out2) {
#pragma unroll
                                for (int g = 0; g < 4; ++g) {
                                    mk2[g * 4 + 0] = posmask2(cur.m2[g].x); mk2[g * 4 + 1] = posmask2(cur.m2[g].y);
                                    mk2[g * 4 + 2] = posmask2(cur.m2[g].z); mk2[g * 4 + 3] = posmask2(cur.m2[g].w);
                                }
                            }
#pragma unroll 1
                            for (int part = 0; part < 3; ++part) {
                                uint32_t pk[16];
#pragma unroll
                                for (int k = 0; k < 16; ++k) {
                                    pk[k] = pack2<false>(v[2 * k], v[2 * k + 1]);
                                    const float2 f = unpack2<false>(pk[k]);
                                    v[2 * k] -= f.x; v[2 * k + 1] -= f.y;              // exact: the next part rounds what this one left
                                    if (has_mask) pk[k] &= mk[k];
                                }
                                if (has_out2) {
#pragma unroll
                                    for (int g = 0; g < 4; ++g)
                                        sts128(o_stage + 2048u + own_off[g], make_uint4(pk[g * 4 + 0] & mk2[g * 4 + 0], pk[g * 4 + 1] & mk2[g * 4 + 1],
                                                                                        pk[g * 4 + 2] & mk2[g * 4 + 2], pk[g * 4 + 3] & mk2[g * 4 + 3]));
                                }
#pragma unroll
                                for (int g = 0; g < 4; ++g) sts128(o_stage + own_off[g], make_uint4(pk[g * 4 + 0], pk[g * 4 + 1], pk[g * 4 + 2], pk[g * 4 + 3]));
                                __syncwarp();
                                if (cox_ok) {
                                    uint16_t* op = (uint16_t*)P.out + co_off + c0 + part * P.Cout;
#pragma unroll
                                    for (int k = 0; k < 4; ++k)
                                        if (coy0 + k * P.up < P.Hout) *reinterpret_cast<uint4*>(op + k * kstep) = lds128(o_stage + coop_off[k]);
                                    if (has_out2) {
                                        uint16_t* op2 = (uint16_t*)P.out2 + co_off + c0 + part * P.Cout;
#pragma unroll
                                        for (int k = 0; k < 4; ++k)
                                            if (coy0 + k * P.up < P.Hout) *reinterpret_cast<uint4*>(op2 + k * kstep) = lds128(o_stage + 2048u + coop_off[k]);
                                    }
                                }
                                __syncwarp();
                            }
                            continue;
                        }
                        if constexpr (COPY2) {
                            // SPAA_EPI_OUT2_BF16 (its own instantiation: as a run-time branch it cost the wide layers 3-4 us each, taken or not -- this
                            // epilogue sits on its register cap): the bf16 rounding of the result goes to the second staging block FIRST, so that v[]
                            // dies when pk[] is formed
#pragma unroll
                            for (int g = 0; g < 4; ++g)
                                sts128(o_stage + 2048u + own_off[g],
                                       make_uint4(pack2<false>(v[g * 8 + 0], v[g * 8 + 1]), pack2<false>(v[g * 8 + 2], v[g * 8 + 3]),
                                                  pack2<false>(v[g * 8 + 4], v[g * 8 + 5]), pack2<false>(v[g * 8 + 6], v[g * 8 + 7])));
                        }
                        // pack to 16 bit, then apply the ReLU masks of the backward pass on the packed pairs (AND with a per-half
                        // "> 0" bit mask: same bits as selecting 0.f before the conversion, ~6x fewer instructions)
                        uint32_t pk[16];
#pragma unroll
                        for (int k = 0; k < 16; ++k) pk[k] = pack2<F16>(v[2 * k], v[2 * k + 1]);
                        if (has_mask) {
#pragma unroll
                            for (int g = 0; g < 4; ++g) {
                                pk[g * 4 + 0] &= posmask2(cur.m[g].x); pk[g * 4 + 1] &= posmask2(cur.m[g].y);
                                pk[g * 4 + 2] &= posmask2(cur.m[g].z); pk[g * 4 + 3] &= posmask2(cur.m[g].w);
                            }
                        }
                        if (has_out2 && has_m2) {
#pragma unroll
                            for (int g = 0; g < 4; ++g)
                                sts128(o_stage + 2048u + own_off[g],
                                       make_uint4(pk[g * 4 + 0] & posmask2(cur.m2[g].x), pk[g * 4 + 1] & posmask2(cur.m2[g].y),
                                                  pk[g * 4 + 2] & posmask2(cur.m2[g].z), pk[g * 4 + 3] & posmask2(cur.m2[g].w)));
                        }
#pragma unroll
                        for (int g = 0; g < 4; ++g) sts128(o_stage + own_off[g], make_uint4(pk[g * 4 + 0], pk[g * 4 + 1], pk[g * 4 + 2], pk[g * 4 + 3]));
                        __syncwarp();
                        if (cox_ok) {
                            uint16_t* op = (uint16_t*)P.out + co_off + c0;
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                if (coy0 + k * P.up < P.Hout) *reinterpret_cast<uint4*>(op + k * kstep) = lds128(o_stage + coop_off[k]);
                            if (has_out2) {
                                uint16_t* op2 = (uint16_t*)P.out2 + co_off + c0;
#pragma unroll
                                for (int k = 0; k < 4; ++k)
                                    if (coy0 + k * P.up < P.Hout) *reinterpret_cast<uint4*>(op2 + k * kstep) = lds128(o_stage + 2048u + coop_off[k]);
                            }
                        }
                        __syncwarp();                                    // the staging block is rewritten by the next chunk
                    }
                }
                }   // !LEAN
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty + acc);
        }
        cp_async_wait(0);
        }   // !lean_done
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

// ---------------------------------------------------------------------------------------------------------------
// weight packing: fp32 parameter (any layout, by strides) -> bf16 [slot = gather tap][BN rows = cout][Cin]
// ---------------------------------------------------------------------------------------------------------------
template <bool F16>
__global__ void pack_weights_kernel(const float* __restrict__ w, uint16_t* __restrict__ out, int KH, int KW, int Cin, int cin_real, int cin_off, int Cout,
                                    int BN, int flip, int64_t w_ts, int64_t w_cis, int64_t w_cos) {
    const int64_t total = (int64_t)KH * KW * BN * Cin;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(e % Cin);
        const int co = (int)((e / Cin) % BN);
        const int tap = (int)(e / ((int64_t)Cin * BN));
        const int r = tap / KW, s = tap - r * KW;
        const int wtap = flip ? (KH - 1 - r) * KW + (KW - 1 - s) : tap;
        const int ci = c - cin_off;
        const float v = (co < Cout && ci >= 0 && ci < cin_real) ? w[(int64_t)wtap * w_ts + (int64_t)ci * w_cis + (int64_t)co * w_cos] : 0.f;
        if constexpr (F16) { const __half h = __float2half_rn(v); out[e] = *reinterpret_cast<const uint16_t*>(&h); }
        else { const __nv_bfloat16 h = __float2bfloat16_rn(v); out[e] = *reinterpret_cast<const uint16_t*>(&h); }
    }
}

// Multi-tensor form: ONE launch re-packs every cached layer after an optimiser step (blockIdx.y = job).  A training step used to spend 34
// launches (one per layer and direction) on this.
struct PackJob {
    const float* w;
    uint16_t* out;
    int64_t w_ts, w_cis, w_cos, total;
    int32_t KH, KW, Cin, cin_real, cin_off, Cout, BN, flip, f16, pad_;
};
__global__ void pack_weights_multi_kernel(const PackJob* __restrict__ jobs) {
    const PackJob J = jobs[blockIdx.y];
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < J.total; e += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(e % J.Cin);
        const int co = (int)((e / J.Cin) % J.BN);
        const int tap = (int)(e / ((int64_t)J.Cin * J.BN));
        const int r = tap / J.KW, s2 = tap - r * J.KW;
        const int wtap = J.flip ? (J.KH - 1 - r) * J.KW + (J.KW - 1 - s2) : tap;
        const int ci = c - J.cin_off;
        const float v = (co < J.Cout && ci >= 0 && ci < J.cin_real) ? J.w[(int64_t)wtap * J.w_ts + (int64_t)ci * J.w_cis + (int64_t)co * J.w_cos] : 0.f;
        if (J.f16) { const __half h = __float2half_rn(v); J.out[e] = *reinterpret_cast<const uint16_t*>(&h); }
        else { const __nv_bfloat16 h = __float2bfloat16_rn(v); J.out[e] = *reinterpret_cast<const uint16_t*>(&h); }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
int bn_for(int cout) { return cout <= 32 ? 32 : (cout <= 64 ? 64 : (cout <= 128 ? 128 : 256)); }

bool tc_supported(const spaa_conv_desc* d, const char** why) {
    auto fail = [&](const char* m) { if (why) *why = m; return false; };
    if (d->in_dtype != 1 && d->in_dtype != 2) return fail("tensor-core path needs bf16 or fp16 input activations");
    const int np = d->split ? 3 : 1;              // physical channels per logical channel (bf16x3 split-precision operands)
    if (d->split && d->in_dtype != 1) return fail("split-precision operands are bf16 parts");
    const bool planar = d->out_dtype == 0;
    if (!planar && d->out_dtype != d->in_dtype) return fail("16-bit output must have the input's type");
    if (!(d->Cin == 16 || d->Cin == 32 || (d->Cin >= 64 && d->Cin % 64 == 0))) return fail("Cin must be 16, 32 or a multiple of 64");
    if (planar) {
        if (d->Cout < 1 || d->Cout > 32) return fail("fp32 planar output supports up to 32 channels");
        const int64_t hw = (int64_t)d->Hout * d->Wout;
        if (d->out_ps != 1 || d->out_cs != hw || (d->B > 1 && d->out_bs != hw * d->Cout)) return fail("fp32 output must be dense NCHW");
    } else {
        if (d->Cout % 32 != 0 || d->Cout > 256) return fail("Cout must be a multiple of 32 (<= 256)");
        if (d->out_cs != 1 || d->out_ps != np * d->Cout || (d->B > 1 && d->out_bs != (int64_t)d->Hout * d->Wout * d->Cout * np)) return fail("16-bit output must be dense NHWC");
    }
    if (d->Cin == 16 && bn_for(d->Cout) != 32) return fail("Cin == 16 is implemented for Cout <= 32");
    if (d->Cin == 32 && bn_for(d->Cout) > 64) return fail("Cin == 32 is implemented for Cout <= 64");
    if (!((d->up == 1 && (d->stride == 1 || d->stride == 2)) || (d->up == 2 && d->stride == 1))) return fail("unsupported stride / up combination");
    if (d->KH != d->KW || d->KH > 3 || d->pad_h != d->pad_w) return fail("square kernels up to 3x3 only");
    if (d->in_cs != 1 || d->in_ps != np * d->Cin || (d->B > 1 && d->in_bs != (int64_t)d->Hin * d->Win * d->Cin * np)) return fail("input must be dense NHWC");   // (a single image: any batch stride)
    if (d->split && 6 * d->Cin / (d->Cin >= 64 ? 64 : d->Cin) > kMaxKChunks) return fail("too many K chunks for the split-precision mode");
    return true;
}

template <int BN, int BK, bool F16>
int launch_tc(const CUtensorMap& ma, const CUtensorMap& mb, const TcParams& P, cudaStream_t st) {
    using L = SmemLayout<BN, BK>;
    static SmemOptIn opt;
    if (!opt.ensure(conv_tc_kernel<BN, BK, F16>, (size_t)L::kTotal)) {
        set_last_error("spaa_conv_tc_fwd: cannot reserve %d bytes of shared memory", L::kTotal);
        return SPAA_ERR_CUDA;
    }
    const int grid = P.total_tiles < kNumSMs ? P.total_tiles : kNumSMs;
    conv_tc_kernel<BN, BK, F16><<<grid, kThreads, L::kTotal, st>>>(ma, mb, P);
    return SPAA_OK;
}

// ---- v2 host side ---------------------------------------------------------------------------------------------
constexpr int kHaloBarBytes = (8 * 4 + 8) * 8 + 16 + kMaxTaps * kMaxPhases * 4;      // a_full/a_empty/b_full/b_empty [8] + tfull/tempty [4] + tmem slot

template <int BN, int BK, bool F16, bool SPLIT = false, bool COPY2 = false>
int launch_halo(const CUtensorMap& ma, const CUtensorMap& mb, const HaloParams& P, const HaloOutMaps& OM_full, size_t smem_bytes, cudaStream_t st) {
    HaloMapsT<BN, BK, SPLIT> OM;
    if constexpr (halo_lean_v<BN, BK, SPLIT>) OM = OM_full;
    else OM.unused = 0;
    static SmemOptIn opt;
    if (!opt.ensure(conv_halo_kernel<BN, BK, F16, SPLIT, COPY2>, smem_bytes, true)) {
        set_last_error("spaa_conv_tc_fwd: cannot reserve %zu bytes of shared memory", smem_bytes);
        return SPAA_ERR_CUDA;
    }
    const int slots = kNumSMs * P.ctas_per_sm;
    const int grid = P.total_tiles < slots ? P.total_tiles : slots;
    // Programmatic dependent launch (opt-in, $SPAA_PDL=1): this kernel's prologue overlaps the tail of the previous kernel in the stream
    // (tc_ptx.cuh: pdl_wait).  Measured on B200 inside the captured attack iteration: 343.4 it/s with it, 347.4 without (same box, same run) --
    // early-resident CTAs of the next layer take SM resources from the tail of the current one -- so it is OFF by default.
    static const int use_pdl = [] { const char* e = getenv("SPAA_PDL"); return e ? atoi(e) : 0; }();
    if (use_pdl) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned)grid);
        cfg.blockDim = dim3((unsigned)(64 + 128 * P.egroups));
        cfg.dynamicSmemBytes = smem_bytes;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        if (cudaLaunchKernelEx(&cfg, conv_halo_kernel<BN, BK, F16, SPLIT, COPY2>, ma, mb, P, OM) != cudaSuccess) {
            set_last_error("spaa_conv_tc_fwd: cudaLaunchKernelEx failed: %s", cudaGetErrorString(cudaGetLastError()));
            return SPAA_ERR_CUDA;
        }
        return SPAA_OK;
    }
    conv_halo_kernel<BN, BK, F16, SPLIT, COPY2><<<grid, 64 + 128 * P.egroups, smem_bytes, st>>>(ma, mb, P, OM);
    return SPAA_OK;
}

inline int floordiv2(int v) { return v >= 0 ? v / 2 : -((-v + 1) / 2); }

// The launch plan of the halo kernel for one layer: tap tables, tile grid, shared-memory / TMEM split, CTAs per SM, epilogue groups.  Depends on the
// descriptor and on WHICH epilogue operands are present only (host arithmetic, no CUDA call: spaa_conv_tc_plan exposes it to the CPU test-suite).
// Returns SPAA_ERR_UNSUPPORTED when the halo kernel does not cover the case (the caller then uses v1).
int halo_plan(const spaa_conv_desc* d, bool add, bool mask, bool mask2, bool out2, HaloParams& P, size_t& smem_bytes) {
    const int BN = bn_for(d->Cout);
    const int BK = d->Cin >= 64 ? 64 : d->Cin;
    const int nph = d->up * d->up;
    if (nph * BN > 512) return SPAA_ERR_UNSUPPORTED;
    memset(&P, 0, sizeof(P));
    P.B = d->B; P.Cin = d->Cin; P.Hin = d->Hin; P.Win = d->Win; P.Cout = d->Cout; P.Hout = d->Hout; P.Wout = d->Wout;
    const int np = d->split ? 3 : 1;
    const int cchunks = d->Cin / BK;               // K chunks of one part
    P.up = d->up; P.stride = d->stride; P.nphases = nph; P.nplanes = d->stride * d->stride; P.kchunks = (d->split ? 6 : 1) * cchunks;
    P.split = d->split ? 1 : 0;
    P.out_ps = np * d->Cout;
    if (P.kchunks > kMaxKChunks) return SPAA_ERR_UNSUPPORTED;
    {
        // part of the INPUT each of the six part products reads; the weights are packed with the matching part of the filter in the same order
        // (spaa_b200/ops.py: _split_weights): smallest products first -- l*h, h*l, m*m, m*h, h*m, h*h
        static const int a_part[6] = {2, 0, 1, 1, 0, 0};
        for (int kc = 0; kc < P.kchunks; ++kc)
            P.a_chunk[kc] = d->split ? a_part[kc / cchunks] * d->Cin + (kc % cchunks) * BK : kc * BK;
    }
    P.nslots = d->KH * d->KW;
    P.epi_flags = d->epi_flags; P.out_planar = d->out_dtype == 0 ? 1 : 0;
    P.add_bs = d->add_bs; P.mask_bs = d->mask_bs;
    // Measured on B200: the MMA unit applies the 128/64/32-byte swizzle to the ABSOLUTE shared-memory address bits (the same
    // function TMA used when it wrote the box), so a start address shifted by whole rows needs NO base-offset correction;
    // setting the descriptor's base_offset field to (addr >> 7) & 7 gives wrong results (tests/test_gpu_conv_tc.py).
    P.use_base_off = 0;
    // ---- tap tables: (plane, shift) of every tap; first pass finds the halo extent, second fills the tables ----
    int qminx = 1 << 20, qmaxx = -(1 << 20), qminy = 1 << 20, qmaxy = -(1 << 20);
    for (int pass = 0; pass < 2; ++pass) {
        if (pass == 1) {
            P.org_x = qminx; P.org_y = qminy;
            P.halo_w = HTW + (qmaxx - qminx); P.halo_h = HTH + (qmaxy - qminy);
        }
        for (int py = 0; py < d->up; ++py)
            for (int px = 0; px < d->up; ++px) {
                HPhase& F = P.ph[py * d->up + px];
                if (pass == 1) { F.py = py; F.px = px; F.ntaps = 0; }
                for (int r = 0; r < d->KH; ++r)
                    for (int s2 = 0; s2 < d->KW; ++s2) {
                        int qy, qx, plane = 0;
                        if (d->up > 1) {
                            const int vy = py + r - d->pad_h, vx = px + s2 - d->pad_w;
                            if ((vy & 1) || (vx & 1)) continue;
                            qy = floordiv2(vy); qx = floordiv2(vx);
                        } else if (d->stride == 2) {
                            const int ry = r - d->pad_h, rx = s2 - d->pad_w;
                            const int ppy = ry & 1, ppx = rx & 1;
                            qy = floordiv2(ry - ppy); qx = floordiv2(rx - ppx);
                            plane = ppy * 2 + ppx;
                        } else {
                            qy = r - d->pad_h; qx = s2 - d->pad_w;
                        }
                        if (pass == 0) {
                            qminx = qx < qminx ? qx : qminx; qmaxx = qx > qmaxx ? qx : qmaxx;
                            qminy = qy < qminy ? qy : qminy; qmaxy = qy > qmaxy ? qy : qmaxy;
                        } else {
                            HTap& t = F.taps[F.ntaps++];
                            t.shift = (int16_t)((qy - qminy) * P.halo_w + (qx - qminx));
                            t.slot = (int8_t)(r * d->KW + s2);
                            t.plane = (int8_t)plane;
                        }
                    }
            }
    }
    const int rowb16 = (BK * 2) >> 4;
    P.ntap_total = 0;
    for (int ph = 0; ph < nph; ++ph)
        for (int t = 0; t < P.ph[ph].ntaps; ++t) {
            const HTap& T = P.ph[ph].taps[t];
            const uint32_t plane16 = (uint32_t)T.plane * ((((uint32_t)(P.halo_h * P.halo_w * BK * 2) + 1023u) & ~1023u) >> 4);
            P.tap_tab[P.ntap_total++] = (plane16 + (uint32_t)T.shift * rowb16) | ((uint32_t)T.slot << 16) | ((uint32_t)ph << 20) | (t == 0 ? (1u << 22) : 0u);
        }
    const int Hgrid = (d->Hout + d->up - 1) / d->up, Wgrid = (d->Wout + d->up - 1) / d->up;
    P.tiles_y = (Hgrid + HTH - 1) / HTH; P.tiles_x = (Wgrid + HTW - 1) / HTW;
    P.total_tiles = d->B * P.tiles_y * P.tiles_x;
    if (P.total_tiles <= 0) { set_last_error("spaa_conv_tc_fwd: empty problem"); return SPAA_ERR_ARG; }
    const int rowb = BK * 2;
    P.a_tx_bytes = P.nplanes * P.halo_h * P.halo_w * rowb;
    P.a_plane_bytes = (P.halo_h * P.halo_w * rowb + 1023) & ~1023;
    P.b_slice_bytes = BN * rowb;
    // ---- shared-memory plan.  Layers whose accumulator is narrow (BN <= 64: the HBM-bound ones) run TWO CTAs per SM when two
    // TMEM allocations and two half-size rings fit: twice the epilogue warps and loads in flight per SM.
    const bool planar = d->out_dtype == 0;
    const bool narrow = BN <= 64 && !d->split && !(BN == 64 && BK == 64);      // kernels with the register-lean epilogue (conv_halo_kernel: LEAN)
    const bool lean = !planar && narrow;          // epilogue_nhwc16: 32-bit element offsets, TMA tile stores
    if (lean) {
        const int64_t img = (int64_t)d->Hout * d->Wout * d->Cout, lim = (int64_t)1 << 31;
        if ((int64_t)d->B * img >= lim || (add && (int64_t)(d->B - 1) * d->add_bs + img >= lim) || ((mask || mask2) && (int64_t)(d->B - 1) * d->mask_bs + img >= lim))
            return SPAA_ERR_UNSUPPORTED;
        if (d->up > 1 && (d->Hout < d->up || d->Wout < d->up)) return SPAA_ERR_UNSUPPORTED;
    }
    P.e_nops = planar ? ((add && d->Cout <= 4) ? 1 : 0) : ((add ? np : 0) + (mask ? 1 : 0) + (mask2 ? 1 : 0));
    static const int max_ctas = [] { const char* e = getenv("SPAA_TC_CTAS"); return e ? atoi(e) : 2; }();
    static const int max_eg = [] { const char* e = getenv("SPAA_TC_EG"); return e ? atoi(e) : 2; }();
    static const int e_kb = [] { const char* e = getenv("SPAA_TC_EKB"); return e ? atoi(e) : -1; }();
    static const int use_pair = [] { const char* e = getenv("SPAA_TC_PAIR"); return e ? atoi(e) : 1; }();
    static const int narrow_ctas = [] { const char* e = getenv("SPAA_TC_NCTAS"); return e ? atoi(e) : 0; }();      // experiments: force one narrow-layer plan
    static const int narrow_eg = [] { const char* e = getenv("SPAA_TC_NEG"); return e ? atoi(e) : 0; }();
    const int64_t res_bytes = (int64_t)P.kchunks * P.nslots * P.b_slice_bytes;
    // Pair mode (two tiles per weight pass, see the MMA issuer): for the wide single-phase layers whose weights are streamed, when the launch has
    // at least two tiles per CTA.  Planned first; if its two-tile A stages leave no room for a weight ring the single-tile plan follows.
    // BN = 128 only: four accumulator buffers keep the epilogue of one pair overlapped with the MMAs of the next.  With BN = 256 a pair fills all 512
    // TMEM columns, the overlap is lost and the layer gets SLOWER (measured: conv4_s forward 73.8 -> 100.5 us, conv4 forward unchanged), while
    // BN = 128 gains (conv5 forward 84.0 -> 72.5 us).  $SPAA_TC_PAIR=2 forces it for BN = 256 too.
    // $SPAA_TC_RESMAX (KB): weights up to this size stay resident in shared memory when ONE CTA per SM can hold them next to two input stages
    // (the 64 <-> 128-channel 3x3 layers: 147 KB); larger ones, or when that plan does not fit, are streamed per tile
    static const int res_max = [] { const char* e = getenv("SPAA_TC_RESMAX"); return e ? atoi(e) : 80; }();
    const bool pair_ok0 = use_pair && (BN == 128 || (BN == 256 && use_pair >= 2)) && nph == 1 && P.total_tiles >= 2 * kNumSMs;
    int ctas = 1;
    smem_bytes = 0;
    bool planned = false;
    if (narrow) {
        // ---- narrow layers (register-lean epilogue, up to 640 threads per CTA).  A plan = (CTAs per SM, epilogue groups, staging blocks per warp);
        // accumulator buffers = what TMEM holds for that many CTAs, at least one per group.  Measured per layer at the BASELINE shapes on one box
        // (tools/kbench.py under $SPAA_TC_NCTAS / $SPAA_TC_NEG, profiles/r2_narrow_plans.md): one CTA with four epilogue groups wins wherever its
        // rings leave >= 3 input stages (one producer / issuer pair with the whole shared memory, weights loaded once); the fp32-planar outputs
        // (short epilogue, no operand ring worth the name) prefer two CTAs with two groups; layers whose weights must be streamed keep the
        // round-1 shape (two CTAs, one group, everything else for the weight ring).
        struct Cand { int c, eg, dbuf; };
        static const Cand nhwc_c[] = {{1, 4, 1}, {1, 4, 0}, {1, 2, 1}, {2, 2, 1}, {1, 2, 0}, {2, 2, 0}, {2, 1, 1}, {2, 1, 0}, {1, 1, 0}};
        static const Cand planar_c[] = {{2, 2, 0}, {2, 1, 0}, {1, 2, 0}, {1, 1, 0}};
        static const Cand stream_c[] = {{2, 1, 1}, {2, 1, 0}, {1, 2, 0}, {1, 1, 0}};
        {
        const int lim_kb = 80;                                      // (larger resident weights were tried for the wide layers only, $SPAA_TC_RESMAX)
        const bool streamed = res_bytes > lim_kb * 1024;
        const Cand forced[] = {{narrow_ctas >= 2 ? 2 : 1, narrow_eg < 1 ? 1 : narrow_eg, 1}, {narrow_ctas >= 2 ? 2 : 1, narrow_eg < 1 ? 1 : narrow_eg, 0}};
        const Cand* cands = narrow_ctas ? forced : (streamed ? stream_c : (planar ? planar_c : nhwc_c));
        const int ncand = narrow_ctas ? 2 : (streamed ? 4 : (planar ? 4 : 9));
        P.pair = 0;
        P.a_pair_bytes = P.nplanes * P.a_plane_bytes;
        P.a_stage_bytes = P.a_pair_bytes;
        for (int pass = 0; pass < 2 && !planned; ++pass)            // pass 0: plans with a comfortable pipeline depth only; pass 1: anything that fits
            for (int ic = 0; ic < ncand && !planned; ++ic) {
                const Cand& C = cands[ic];
                int nb = (512 / C.c) / (nph * BN);
                if (nb < 1) continue;
                if (streamed && nb > 2) nb = 2;
                const int nbuf = nb >= 4 ? 4 : (nb >= 2 ? 2 : 1);
                if (C.eg > nbuf) continue;
                const int total = C.c == 2 ? 108 * 1024 : 222 * 1024;
                const int resident = res_bytes <= (C.c == 2 ? 40 : lim_kb) * 1024 ? 1 : 0;
                const int slot = planar ? 2048 : P.e_nops * 8192;       // (the fp32-planar residual of a tile is 4 channels x 128 rows x 4 bytes)
                int S = 0;
                if (P.e_nops) {
                    const int eb = (e_kb >= 0 ? e_kb : (C.c == 2 ? 32 : 48)) * 1024;
                    const int sl = eb / (slot * C.eg);
                    S = (resident || sl < 2) ? 2 : (sl > 8 ? 8 : sl);          // resident weights: minimal ring first, deepened below with what is left
                }
                const int stage = planar ? 0 : (1 + C.dbuf) * C.eg * 4 * (out2 ? 4096 : 2048);
                int64_t budget = (int64_t)total - (int64_t)C.eg * S * slot - stage;
                while (S > 2 && budget - 2 * P.a_stage_bytes < (resident ? res_bytes : 2 * (int64_t)P.b_slice_bytes)) {
                    --S;
                    budget += (int64_t)C.eg * slot;
                }
                int64_t sa, sb, bbytes;
                if (resident) {
                    bbytes = res_bytes; sb = 1;
                    sa = (budget - bbytes) / P.a_stage_bytes;
                    if (sa < (pass == 0 ? 3 : 2)) continue;
                    if (sa > 8) sa = 8;
                    // These layers run at (bytes in flight) / (memory latency): measured, conv6 forward 86 us with 69 KB of input stages per SM and 68 us
                    // with 92 KB.  Whatever shared memory is left goes to the pipeline that holds fewer TILES -- input stages or the operand ring
                    // (one slot = one 32-channel unit of one phase of one tile, S - 1 of them in flight per epilogue group).
                    int64_t left = budget - bbytes - sa * P.a_stage_bytes;
                    const int units = nph * (BN / 32);
                    for (;;) {
                        const bool can_a = sa < 8 && left >= P.a_stage_bytes;
                        const bool can_e = S > 0 && S < 8 && left >= (int64_t)C.eg * slot;
                        if (!can_a && !can_e) break;
                        const bool pick_e = can_e && (!can_a || (int64_t)C.eg * (S - 1) < sa * units);
                        if (pick_e) { ++S; left -= (int64_t)C.eg * slot; }
                        else { ++sa; left -= P.a_stage_bytes; }
                    }
                } else {
                    sa = 2;
                    sb = (budget - 2 * (int64_t)P.a_stage_bytes) / P.b_slice_bytes;
                    if (sb > 8) sb = 8;
                    if (sb < (pass == 0 ? 4 : 2)) continue;
                    bbytes = sb * P.b_slice_bytes;
                }
                P.nbuf = nbuf; P.egroups = C.eg; P.e_dbuf = (!planar && C.dbuf) ? 1 : 0; P.e_stage_bytes = stage; P.e_slots = S; P.e_slot_bytes = slot;
                P.resident = resident; P.sb = (int)sb; P.sa = (int)sa;
                ctas = C.c;
                smem_bytes = (size_t)P.sa * P.a_stage_bytes + (size_t)bbytes + (size_t)P.egroups * P.e_slots * slot + P.e_stage_bytes + kHaloBarBytes + BN * 4 + 1024;
                planned = true;
            }
        }
        if (!planned) return SPAA_ERR_UNSUPPORTED;
    }
    for (int res_try = 0; res_try < 2 && !planned; ++res_try) {
    const int res_lim = (res_try == 0 ? res_max : 80) * 1024;
    if (res_try == 1 && res_max <= 80) break;
    const bool pair_ok = pair_ok0 && res_bytes > res_lim;
    for (int try_pair = pair_ok ? 1 : 0; try_pair >= 0 && !planned; --try_pair) {
        P.pair = try_pair;
        P.a_pair_bytes = P.nplanes * P.a_plane_bytes;
        P.a_stage_bytes = (try_pair ? 2 : 1) * P.a_pair_bytes;
        // accumulator buffers: two when they fit in half of TMEM's 512 columns (so that two CTAs can share an SM) or in all of it for
        // the wide layers; a 4-phase BN = 64 layer (transConv1 forward) runs single-buffered in 256 columns with two CTAs per SM instead
        // of double-buffered alone on its SM; pair mode: four (BN = 128: two pairs in flight) or two (BN = 256)
        P.nbuf = try_pair ? (4 * BN <= 512 ? 4 : 2) : (2 * nph * BN <= (BN <= 64 ? 256 : 512) ? 2 : 1);
        uint32_t tmem_cols = 32;
        while (tmem_cols < (uint32_t)(P.nbuf * nph * BN)) tmem_cols <<= 1;
        // two epilogue groups for the wide layers (one CTA per SM), unless their rings and staging blocks would starve the weight ring
        // (measured: it pays where the MMA work per tile is short next to the epilogue's -- conv3 / skipConv3 / conv3_s / conv4 forward,
        // K per epilogue operand <= 640 -- and costs 3-8 % on the MMA-bound layers, whose issuing warp then shares its schedulers)
        const int k_per_pass = P.nslots * d->Cin * (d->split ? 6 : 1) / (1 + P.e_nops);
        P.egroups = (BN >= 128 && P.nbuf >= 2 && max_eg >= 2 && !(mask2 && BN == 256) && k_per_pass <= 640) ? 2 : 1;
        P.e_dbuf = 0;
        P.e_stage_bytes = planar ? 0 : (1 + P.e_dbuf) * P.egroups * 4 * (out2 ? 4096 : 2048);
        ctas = (BN <= 64 && 2 * tmem_cols <= 512 && max_ctas >= 2) ? 2 : 1;
        for (;; --ctas) {
            const int total = ctas == 2 ? 108 * 1024 : 222 * 1024;
            P.resident = res_bytes <= (ctas == 2 ? 40 * 1024 : res_lim) ? 1 : 0;
            P.e_slots = 0;
            if (P.e_nops) {
                const int eb = (e_kb >= 0 ? e_kb : (ctas == 2 ? 32 : (P.resident ? 48 : 32))) * 1024;
                const int sl = eb / (P.e_nops * 8192 * P.egroups);             // slots PER GROUP
                P.e_slots = sl < 2 ? 2 : (sl > 8 ? 8 : sl);
            }
            int budget = total - P.egroups * P.e_slots * P.e_nops * 8192 - P.e_stage_bytes;
            // big A stages (stride-2 layers load four parity planes): give the operand ring's depth back before giving up the kernel
            while (P.e_slots > 2 && budget - 2 * P.a_stage_bytes < (P.resident ? res_bytes : 2 * (int64_t)P.b_slice_bytes)) {
                --P.e_slots;
                budget += P.egroups * P.e_nops * 8192;
            }
            int64_t bbytes, sa;
            if (P.resident) {
                bbytes = res_bytes; P.sb = 1;
                sa = (budget - bbytes) / P.a_stage_bytes;
            } else {
                // streamed weights: one A stage feeds KH*KW*BK/16 MMAs but one weight slice only BK/16 -- the depth belongs to the B ring
                sa = 2;
                const int64_t sbn = (budget - 2 * (int64_t)P.a_stage_bytes) / P.b_slice_bytes;
                P.sb = (int)(sbn > 8 ? 8 : sbn);
                if (P.sb < (try_pair ? 3 : 2)) sa = 0;               // (pair mode is only worth it with a weight ring of depth >= 3)
                bbytes = (int64_t)P.sb * P.b_slice_bytes;
            }
            if (sa < 2) {
                if (P.e_dbuf) {                                      // single staging block per warp before giving up a CTA or an epilogue group
                    P.e_dbuf = 0;
                    P.e_stage_bytes /= 2;
                    ++ctas;
                    continue;
                }
                if (ctas == 1 && P.egroups == 2) {                   // a second epilogue group's ring and staging blocks do not fit: run with one
                    P.egroups = 1;
                    P.e_stage_bytes = planar ? 0 : 4 * (out2 ? 4096 : 2048);
                    ++ctas;                                          // (undo the loop's decrement: plan again with one CTA per SM)
                    continue;
                }
                if (ctas == 1) break;                                // this mode does not fit
                continue;
            }
            P.sa = (int)(sa > 6 ? 6 : sa);
            P.e_slot_bytes = P.e_nops * 8192;
            smem_bytes = (size_t)P.sa * P.a_stage_bytes + (size_t)bbytes + (size_t)P.egroups * P.e_slots * P.e_nops * 8192 + P.e_stage_bytes + kHaloBarBytes + BN * 4 + 1024;
            planned = true;
            break;
        }
    }
    }   // res_try
    if (!planned) return SPAA_ERR_UNSUPPORTED;
    P.ctas_per_sm = ctas;
    return SPAA_OK;
}

int conv_halo(const spaa_conv_desc* d, const void* in, const void* wpacked, const float* bias, const void* add, const void* mask, const void* mask2,
              void* out, void* out2, cudaStream_t st, EncodeTiledFn enc) {
    HaloParams P;
    size_t smem_bytes = 0;
    const int prc = halo_plan(d, add != nullptr, mask != nullptr, mask2 != nullptr, out2 != nullptr, P, smem_bytes);
    if (prc != SPAA_OK) return prc;
    const int BN = bn_for(d->Cout);
    const int BK = d->Cin >= 64 ? 64 : d->Cin;
    const bool f16 = d->in_dtype == 2;
    const int nph = d->up * d->up;
    const int np = d->split ? 3 : 1;
    const bool lean = d->out_dtype != 0 && !d->split && BN <= 64 && !(BN == 64 && BK == 64);
    P.bias = bias; P.add = add; P.mask = (const uint16_t*)mask; P.mask2 = (const uint16_t*)mask2; P.out = out; P.out2 = out2;

    CUtensorMap ma, mb;
    const CUtensorMapSwizzle swz = BK == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (BK == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    const CUtensorMapDataType dt = f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    {
        const cuuint64_t cphys = (cuuint64_t)np * d->Cin;          // physical channels of the NHWC input
        cuuint64_t dims[4] = {cphys, (cuuint64_t)d->Win, (cuuint64_t)d->Hin, (cuuint64_t)d->B};
        cuuint64_t strides[3] = {cphys * 2, (cuuint64_t)d->Win * cphys * 2, (cuuint64_t)d->Hin * d->Win * cphys * 2};
        cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)(P.halo_w * d->stride), (cuuint32_t)(P.halo_h * d->stride), 1};
        cuuint32_t es[4] = {1, (cuuint32_t)d->stride, (cuuint32_t)d->stride, 1};
        CUresult r = enc(&ma, dt, 4, const_cast<void*>(in), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_last_error("spaa_conv_tc_fwd: cuTensorMapEncodeTiled(halo input) failed with %d", (int)r); return SPAA_ERR_CUDA; }
    }
    {
        const cuuint64_t wk = (cuuint64_t)(d->split ? 6 : 1) * d->Cin;      // K extent of the packed weights
        cuuint64_t dims[2] = {wk, (cuuint64_t)d->KH * d->KW * BN};
        cuuint64_t strides[1] = {wk * 2};
        cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BN};
        cuuint32_t es[2] = {1, 1};
        CUresult r = enc(&mb, dt, 2, const_cast<void*>(wpacked), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_last_error("spaa_conv_tc_fwd: cuTensorMapEncodeTiled(weights) failed with %d", (int)r); return SPAA_ERR_CUDA; }
    }
    HaloOutMaps OM;
    memset(&OM, 0, sizeof(OM));
    if (lean) {
        // one strided view per output phase: pixel (y', x') of phase (py, px) is output pixel (up * y' + py, up * x' + px)
        const cuuint64_t C = (cuuint64_t)d->Cout;
        for (int o = 0; o < (out2 ? 2 : 1); ++o)
            for (int ph = 0; ph < nph; ++ph) {
                const int py = ph / d->up, px = ph % d->up;
                uint16_t* base = (uint16_t*)(o ? out2 : out) + ((int64_t)py * d->Wout + px) * d->Cout;
                cuuint64_t dims[4] = {C, (cuuint64_t)((d->Wout - px + d->up - 1) / d->up), (cuuint64_t)((d->Hout - py + d->up - 1) / d->up), (cuuint64_t)d->B};
                cuuint64_t strides[3] = {(cuuint64_t)d->up * C * 2, (cuuint64_t)d->up * d->Wout * C * 2, (cuuint64_t)d->Hout * d->Wout * C * 2};
                cuuint32_t box[4] = {32, (cuuint32_t)HTW, 4, 1};
                cuuint32_t es[4] = {1, 1, 1, 1};
                CUresult r = enc(&OM.m[o * kMaxPhases + ph], CU_TENSOR_MAP_DATA_TYPE_UINT16, 4, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                if (r != CUDA_SUCCESS) { set_last_error("spaa_conv_tc_fwd: cuTensorMapEncodeTiled(output, phase %d) failed with %d", ph, (int)r); return SPAA_ERR_CUDA; }
            }
    }
    int rc = SPAA_OK;
    // the bf16 copy of an fp16 output (SPAA_EPI_OUT2_BF16) is a run-time flag in the register-lean epilogue of the narrow layers and a separate
    // instantiation of the kernels that keep the round-1 epilogue
    const bool copy2_wide = out2 != nullptr && mask2 == nullptr && !lean;
#define SPAA_HALO_LAUNCH(BN_, BK_) rc = f16 ? ((copy2_wide && (BN_ >= 128 || (BN_ == 64 && BK_ == 64))) ? launch_halo<BN_, BK_, true, false, (BN_ >= 128 || (BN_ == 64 && BK_ == 64))>(ma, mb, P, OM, smem_bytes, st) \
                                                                                              : launch_halo<BN_, BK_, true>(ma, mb, P, OM, smem_bytes, st)) : \
    (d->split ? launch_halo<BN_, BK_, false, true>(ma, mb, P, OM, smem_bytes, st) : launch_halo<BN_, BK_, false>(ma, mb, P, OM, smem_bytes, st))
    if (BK == 64) {
        if (BN == 32) SPAA_HALO_LAUNCH(32, 64);
        else if (BN == 64) SPAA_HALO_LAUNCH(64, 64);
        else if (BN == 128) SPAA_HALO_LAUNCH(128, 64);
        else SPAA_HALO_LAUNCH(256, 64);
    } else if (BK == 32) {
        if (BN == 32) SPAA_HALO_LAUNCH(32, 32);
        else SPAA_HALO_LAUNCH(64, 32);
    } else {
        SPAA_HALO_LAUNCH(32, 16);
    }
#undef SPAA_HALO_LAUNCH
    return rc;
}

}  // namespace

namespace spaa { namespace tc {
EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    });
    return fn;
}
} }

extern "C" {

int spaa_conv_tc_supported(const spaa_conv_desc* d) { return (d && tc_supported(d, nullptr)) ? 1 : 0; }

int spaa_conv_tc_plan(const spaa_conv_desc* d, int has_add, int has_mask, int has_mask2, int32_t* plan) {
    SPAA_CHECK_ARG(d && plan, "spaa_conv_tc_plan: null argument");
    const char* why = "";
    SPAA_CHECK_ARG(tc_supported(d, &why), "spaa_conv_tc_plan: %s", why);
    HaloParams P;
    size_t smem = 0;
    const int rc = halo_plan(d, has_add != 0, has_mask != 0, has_mask2 != 0, has_mask2 != 0 || (d->epi_flags & SPAA_EPI_OUT2_BF16) != 0, P, smem);
    if (rc != SPAA_OK) return rc;
    plan[0] = P.ctas_per_sm; plan[1] = P.egroups; plan[2] = P.nbuf; plan[3] = P.sa; plan[4] = P.resident ? 0 : P.sb; plan[5] = P.e_slots;
    plan[6] = P.e_dbuf; plan[7] = P.pair; plan[8] = (int32_t)smem; plan[9] = P.total_tiles; plan[10] = 64 + 128 * P.egroups; plan[11] = P.resident;
    return SPAA_OK;
}

int64_t spaa_conv_tc_packed_elems(const spaa_conv_desc* d) {
    if (!d) return 0;
    return (int64_t)d->KH * d->KW * bn_for(d->Cout) * d->Cin * (d->split ? 6 : 1);
}

int spaa_conv_tc_pack_weights(const spaa_conv_desc* d, const float* w, int cin_real, int cin_offset, void* packed, spaa_stream_t stream) {
    SPAA_CHECK_ARG(d && w && packed && cin_real > 0 && cin_offset >= 0 && cin_offset + cin_real <= d->Cin, "spaa_conv_tc_pack_weights: bad arguments");
    const int BN = bn_for(d->Cout);
    const int64_t total = spaa_conv_tc_packed_elems(d);
    int64_t blocks = (total + 255) / 256;
    if (blocks > 4 * kNumSMs) blocks = 4 * kNumSMs;
    if (d->in_dtype == 2)
        pack_weights_kernel<true><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(w, (uint16_t*)packed, d->KH, d->KW, d->Cin, cin_real, cin_offset, d->Cout,
                                                                                     BN, d->flip, d->w_ts, d->w_cis, d->w_cos);
    else
        pack_weights_kernel<false><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(w, (uint16_t*)packed, d->KH, d->KW, d->Cin, cin_real, cin_offset, d->Cout,
                                                                                      BN, d->flip, d->w_ts, d->w_cis, d->w_cos);
    SPAA_CHECK_LAUNCH("spaa_conv_tc_pack_weights");
    return SPAA_OK;
}

int64_t spaa_conv_tc_pack_job_bytes(void) { return (int64_t)sizeof(PackJob); }

int spaa_conv_tc_pack_job(const spaa_conv_desc* d, const float* w, int cin_real, int cin_offset, void* packed, void* job_host) {
    SPAA_CHECK_ARG(d && w && packed && job_host && !d->split && cin_real > 0 && cin_offset >= 0 && cin_offset + cin_real <= d->Cin,
                   "spaa_conv_tc_pack_job: bad arguments");
    PackJob J;
    memset(&J, 0, sizeof(J));
    J.w = w; J.out = (uint16_t*)packed;
    J.w_ts = d->w_ts; J.w_cis = d->w_cis; J.w_cos = d->w_cos; J.total = spaa_conv_tc_packed_elems(d);
    J.KH = d->KH; J.KW = d->KW; J.Cin = d->Cin; J.cin_real = cin_real; J.cin_off = cin_offset; J.Cout = d->Cout; J.BN = bn_for(d->Cout);
    J.flip = d->flip; J.f16 = d->in_dtype == 2 ? 1 : 0;
    memcpy(job_host, &J, sizeof(J));
    return SPAA_OK;
}

int spaa_conv_tc_pack_weights_multi(const void* jobs_dev, int njobs, spaa_stream_t stream) {
    SPAA_CHECK_ARG(jobs_dev && njobs > 0 && njobs < 65536, "spaa_conv_tc_pack_weights_multi: bad arguments");
    pack_weights_multi_kernel<<<dim3(64, (unsigned)njobs), 256, 0, (cudaStream_t)stream>>>((const PackJob*)jobs_dev);
    SPAA_CHECK_LAUNCH("spaa_conv_tc_pack_weights_multi");
    return SPAA_OK;
}

int spaa_conv_tc_fwd(const spaa_conv_desc* d, const void* in, const void* wpacked, const float* bias, const void* add, const void* mask, const void* mask2,
                     void* out, void* out2, spaa_stream_t stream) {
    SPAA_CHECK_ARG(d && in && wpacked && out, "spaa_conv_tc_fwd: null argument");
    const char* why = "";
    SPAA_CHECK_ARG(tc_supported(d, &why), "spaa_conv_tc_fwd: %s", why);
    const bool copy2 = (d->epi_flags & SPAA_EPI_OUT2_BF16) != 0;
    SPAA_CHECK_ARG(copy2 ? (out2 != nullptr && mask2 == nullptr && d->in_dtype == 2 && d->out_dtype == 2 && !d->split)
                         : ((out2 == nullptr) == (mask2 == nullptr)),
                   "spaa_conv_tc_fwd: out2 and mask2 go together, except with SPAA_EPI_OUT2_BF16 (fp16 NHWC layers: out2 alone)");
    SPAA_CHECK_ARG(d->mask_mode == SPAA_MASK_NONE || mask, "spaa_conv_tc_fwd: mask_mode needs mask");
    const bool planar = d->out_dtype == 0;
    SPAA_CHECK_ARG(!(d->epi_flags & ~(SPAA_EPI_RELU | (planar ? SPAA_EPI_CLAMP_MAX1 : SPAA_EPI_OUT2_BF16))),
                   "spaa_conv_tc_fwd: epilogue flags: ReLU (and clamp for fp32 planar output, the bf16 copy for fp16 NHWC output) only");
    SPAA_CHECK_ARG(!mask || d->mask_mode == SPAA_MASK_POS, "spaa_conv_tc_fwd: only the ReLU mask (SPAA_MASK_POS) is implemented on the tensor-core path");
    if (planar) {
        SPAA_CHECK_ARG(!mask && !mask2, "spaa_conv_tc_fwd: masks are not implemented for fp32 planar output");
        SPAA_CHECK_ARG(!add || (d->add_ps == 1 && d->add_cs == (int64_t)d->Hout * d->Wout), "spaa_conv_tc_fwd: add must be fp32 NCHW planes");
    } else {
        const int np = d->split ? 3 : 1;
        SPAA_CHECK_ARG(!add || (d->add_cs == 1 && d->add_ps == np * d->Cout), "spaa_conv_tc_fwd: add must be dense NHWC");
        SPAA_CHECK_ARG(!(mask || mask2) || (d->mask_cs == 1 && d->mask_ps == np * d->Cout), "spaa_conv_tc_fwd: masks must be dense NHWC");
    }
    EncodeTiledFn enc = get_encode();
    if (!enc) { set_last_error("spaa_conv_tc_fwd: cuTensorMapEncodeTiled is unavailable in this driver"); return SPAA_ERR_CUDA; }
    static const int force_v1 = [] { const char* e = getenv("SPAA_TC_V1"); return e ? atoi(e) : 0; }();
    if (!force_v1) {
        const int rc2 = conv_halo(d, in, wpacked, bias, add, mask, mask2, out, out2, (cudaStream_t)stream, enc);
        if (rc2 == SPAA_OK) { SPAA_CHECK_LAUNCH("spaa_conv_tc_fwd (halo kernel)"); return SPAA_OK; }
        if (rc2 != SPAA_ERR_UNSUPPORTED) return rc2;
    }
    SPAA_CHECK_ARG(!d->split, "spaa_conv_tc_fwd: the split-precision mode is implemented by the halo kernel only (this shape needs the fallback kernel)");
    SPAA_CHECK_ARG(!copy2, "spaa_conv_tc_fwd: SPAA_EPI_OUT2_BF16 is implemented by the halo kernel only (this shape needs the fallback kernel)");
    const int BN = bn_for(d->Cout);
    const int BK = d->Cin >= 64 ? 64 : d->Cin;
    const bool f16 = d->in_dtype == 2;

    TcParams P;
    memset(&P, 0, sizeof(P));
    P.B = d->B; P.Cin = d->Cin; P.Hin = d->Hin; P.Win = d->Win; P.Cout = d->Cout; P.Hout = d->Hout; P.Wout = d->Wout;
    P.in_step = d->stride; P.out_step = d->up; P.kchunks = d->Cin / BK;
    P.epi_flags = d->epi_flags; P.mask_mode = mask ? d->mask_mode : SPAA_MASK_NONE;
    P.add_bs = d->add_bs; P.mask_bs = d->mask_bs;
    P.out_planar = planar ? 1 : 0;
    P.bias = bias; P.add = add; P.mask = (const uint16_t*)mask; P.mask2 = (const uint16_t*)mask2;
    P.out = out; P.out2 = out2;
    int tile_base = 0;
    P.nphases = d->up * d->up;
    for (int py = 0; py < d->up; ++py)
        for (int px = 0; px < d->up; ++px) {
            Phase& F = P.ph[py * d->up + px];
            F.py = py; F.px = px;
            F.Hph = (d->Hout - py + d->up - 1) / d->up;
            F.Wph = (d->Wout - px + d->up - 1) / d->up;
            F.tiles_y = (F.Hph + TH - 1) / TH; F.tiles_x = (F.Wph + TW - 1) / TW;
            F.tile_base = tile_base;
            tile_base += d->B * F.tiles_y * F.tiles_x;
            F.ntaps = 0;
            for (int r = 0; r < d->KH; ++r)
                for (int s = 0; s < d->KW; ++s) {
                    const int vy = py + r - d->pad_h, vx = px + s - d->pad_w;
                    if (d->up > 1 && ((vy % d->up) != 0 || (vx % d->up) != 0)) continue;
                    Tap& t = F.taps[F.ntaps++];
                    t.dy = (int8_t)(d->up > 1 ? vy / d->up : r - d->pad_h);
                    t.dx = (int8_t)(d->up > 1 ? vx / d->up : s - d->pad_w);
                    t.slot = (int8_t)(r * d->KW + s);
                }
        }
    P.total_tiles = tile_base;
    SPAA_CHECK_ARG(P.total_tiles > 0, "spaa_conv_tc_fwd: empty problem");

    CUtensorMap ma, mb;
    {
        cuuint64_t dims[4] = {(cuuint64_t)d->Cin, (cuuint64_t)d->Win, (cuuint64_t)d->Hin, (cuuint64_t)d->B};
        cuuint64_t strides[3] = {(cuuint64_t)d->Cin * 2, (cuuint64_t)d->Win * d->Cin * 2, (cuuint64_t)d->Hin * d->Win * d->Cin * 2};
        cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)(TW * d->stride), (cuuint32_t)(TH * d->stride), 1};
        cuuint32_t es[4] = {1, (cuuint32_t)d->stride, (cuuint32_t)d->stride, 1};
        CUresult r = enc(&ma, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(in), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         BK == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (BK == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_last_error("spaa_conv_tc_fwd: cuTensorMapEncodeTiled(input) failed with %d", (int)r); return SPAA_ERR_CUDA; }
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)d->Cin, (cuuint64_t)d->KH * d->KW * BN};
        cuuint64_t strides[1] = {(cuuint64_t)d->Cin * 2};
        cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BN};
        cuuint32_t es[2] = {1, 1};
        CUresult r = enc(&mb, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(wpacked), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         BK == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (BK == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_last_error("spaa_conv_tc_fwd: cuTensorMapEncodeTiled(weights) failed with %d", (int)r); return SPAA_ERR_CUDA; }
    }
    cudaStream_t st = (cudaStream_t)stream;
    int rc = SPAA_OK;
#define SPAA_TC_LAUNCH(BN_, BK_) rc = f16 ? launch_tc<BN_, BK_, true>(ma, mb, P, st) : launch_tc<BN_, BK_, false>(ma, mb, P, st)
    if (BK == 64) {
        if (BN == 32) SPAA_TC_LAUNCH(32, 64);
        else if (BN == 64) SPAA_TC_LAUNCH(64, 64);
        else if (BN == 128) SPAA_TC_LAUNCH(128, 64);
        else SPAA_TC_LAUNCH(256, 64);
    } else if (BK == 32) {
        if (BN == 32) SPAA_TC_LAUNCH(32, 32);
        else SPAA_TC_LAUNCH(64, 32);
    } else {
        SPAA_TC_LAUNCH(32, 16);
    }
#undef SPAA_TC_LAUNCH
    if (rc != SPAA_OK) return rc;
    SPAA_CHECK_LAUNCH("spaa_conv_tc_fwd");
    return SPAA_OK;
}

}  // extern "C"

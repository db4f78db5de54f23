// Attack-loop update kernels: per-sample gradient norms, masked normalised steps, clamp / quantise projection,
// best-so-far copies and the device-side decision logic (so the loops never synchronise with the host).
// Replaces /root/reference/src/python/projector_based_attack.py:275,290-328 and perc_al/__init__.py:193-245.
// All HBM-bound streaming kernels; vectorised (float4) where the row length allows.
#include "common.cuh"
#include "../../include/spaa_b200.h"

using namespace spaa;

namespace {

constexpr int kThreads = 256;
constexpr int kPerBlock = 8192;   // elements per block in row reductions (two 128-bit loads per operand and thread in flight)

SPAA_D bool in_range(float x, float lo, float hi) { return x >= lo && x <= hi; }   // torch.clamp backward: inclusive

__global__ void __launch_bounds__(kThreads) row_sqnorm_kernel(const float* __restrict__ g, const float* __restrict__ xc, float lo, float hi, int64_t n,
                                                              float* __restrict__ sq, float* __restrict__ partial, unsigned* __restrict__ counter) {
    __shared__ float red[32];
    __shared__ bool is_last;
    const int b = blockIdx.y;
    const float* gb = g + (int64_t)b * n;
    const float* xb = xc ? xc + (int64_t)b * n : nullptr;
    float s[1] = {0.f};
    if ((n & 3) == 0) {
        const float4* g4 = reinterpret_cast<const float4*>(gb);
        const float4* x4 = reinterpret_cast<const float4*>(xb);
#pragma unroll 2
        for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < (n >> 2); i += (int64_t)gridDim.x * blockDim.x) {
            float4 v = __ldg(g4 + i);
            if (xb) {
                const float4 x = __ldg(x4 + i);
                if (!in_range(x.x, lo, hi)) v.x = 0.f;
                if (!in_range(x.y, lo, hi)) v.y = 0.f;
                if (!in_range(x.z, lo, hi)) v.z = 0.f;
                if (!in_range(x.w, lo, hi)) v.w = 0.f;
            }
            s[0] += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
        }
    } else {
        for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
            float v = __ldg(gb + i);
            if (xb && !in_range(__ldg(xb + i), lo, hi)) v = 0.f;
            s[0] += v * v;
        }
    }
    row_reduce_finish<1>(s, b, partial, counter, sq, red, &is_last);
}

// x[b] += step * g[b] / sqrt(sq[b]);  step = step2[sel[b] != 0]; rows with step == 0 are left untouched.
__global__ void __launch_bounds__(kThreads) row_step_kernel(float* __restrict__ x, const float* __restrict__ g, const float* __restrict__ sq,
                                                            const uint8_t* __restrict__ sel, const float* __restrict__ step2, int use_clamp, float lo,
                                                            float hi, const float* __restrict__ base, int64_t base_bs, float* __restrict__ sum_out,
                                                            float* __restrict__ copy_dst, const uint8_t* __restrict__ copy_sel, int64_t n) {
    const int b = blockIdx.y;
    const int s = sel ? (sel[b] != 0) : 0;
    const float step = __ldg(step2 + s);
    const bool do_copy = copy_dst && copy_sel[b];
    if (step == 0.f && !sum_out && !do_copy) return;
    const float scale = step == 0.f ? 0.f : step / sqrtf(__ldg(sq + b));
    float* xb = x + (int64_t)b * n;
    const float* gb = g + (int64_t)b * n;
    if ((n & 3) == 0 && (base_bs & 3) == 0) {
        // 128-bit path (rows are 196 608 floats in the attack loops): same arithmetic per element
        float4* x4 = reinterpret_cast<float4*>(xb);
        const float4* g4 = reinterpret_cast<const float4*>(gb);
        const float4* b4 = base ? reinterpret_cast<const float4*>(base + (int64_t)b * base_bs) : nullptr;
        float4* s4 = sum_out ? reinterpret_cast<float4*>(sum_out + (int64_t)b * n) : nullptr;
        float4* c4 = do_copy ? reinterpret_cast<float4*>(copy_dst + (int64_t)b * n) : nullptr;
        for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < (n >> 2); i += (int64_t)gridDim.x * blockDim.x) {
            float4 xv = x4[i];
            if (step != 0.f) {
                float4 gv = __ldg(g4 + i);
                if (use_clamp) {
                    if (!in_range(xv.x, lo, hi)) gv.x = 0.f;
                    if (!in_range(xv.y, lo, hi)) gv.y = 0.f;
                    if (!in_range(xv.z, lo, hi)) gv.z = 0.f;
                    if (!in_range(xv.w, lo, hi)) gv.w = 0.f;
                }
                xv.x = xv.x + scale * gv.x; xv.y = xv.y + scale * gv.y; xv.z = xv.z + scale * gv.z; xv.w = xv.w + scale * gv.w;
                x4[i] = xv;
            }
            if (s4) { const float4 bv = __ldg(b4 + i); s4[i] = make_float4(bv.x + xv.x, bv.y + xv.y, bv.z + xv.z, bv.w + xv.w); }
            if (c4) c4[i] = xv;
        }
        return;
    }
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float xv = xb[i];
        if (step != 0.f) {
            float gv = __ldg(gb + i);
            if (use_clamp && !in_range(xv, lo, hi)) gv = 0.f;
            xv = xv + scale * gv;        // 0/0 -> NaN exactly as g / ||g|| does in the reference
            xb[i] = xv;
        }
        if (sum_out) sum_out[(int64_t)b * n + i] = __ldg(base + (int64_t)b * base_bs + i) + xv;
        if (do_copy) copy_dst[(int64_t)b * n + i] = xv;
    }
}

__global__ void __launch_bounds__(kThreads) masked_copy_kernel(float* __restrict__ dst, const float* __restrict__ src, const uint8_t* __restrict__ sel,
                                                               int64_t n) {
    const int b = blockIdx.y;
    if (!sel[b]) return;
    const float* s = src + (int64_t)b * n;
    float* d = dst + (int64_t)b * n;
    if ((n & 3) == 0) {
        for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < (n >> 2); i += (int64_t)gridDim.x * blockDim.x)
            reinterpret_cast<float4*>(d)[i] = __ldg(reinterpret_cast<const float4*>(s) + i);
    } else {
        for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) d[i] = __ldg(s + i);
    }
}

__global__ void __launch_bounds__(kThreads) select_cot_kernel(const float* __restrict__ g0, const float* __restrict__ g1, const uint8_t* __restrict__ sel,
                                                              const float* __restrict__ act, int mode, float* __restrict__ out, int64_t n) {
    const int b = blockIdx.y;
    const float* src = ((sel && sel[b]) ? g1 : g0) + (int64_t)b * n;
    const float* a = act ? act + (int64_t)b * n : nullptr;
    float* o = out + (int64_t)b * n;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float v = __ldg(src + i);
        if (a) {
            const float m = __ldg(a + i);
            if (mode == SPAA_MASK_POS) v = m > 0.f ? v : 0.f;
            else if (mode == SPAA_MASK_OPEN01) v = (m > 0.f && m < 1.f) ? v : 0.f;
            else if (mode == SPAA_MASK_LEAKY01) v = m > 0.f ? v : 0.1f * v;
        }
        o[i] = v;
    }
}

// same selection, written as the zero-padded 16-channel 16-bit NHWC tensor the tensor-core backward of conv6 reads
template <bool F16>
__global__ void __launch_bounds__(kThreads) select_cot_packed_kernel(const float* __restrict__ g0, const float* __restrict__ g1, const uint8_t* __restrict__ sel,
                                                                     const float* __restrict__ act, int mode, uint4* __restrict__ out16, int64_t HW) {
    const int b = blockIdx.y;
    const float* src = ((sel && sel[b]) ? g1 : g0) + (int64_t)b * 3 * HW;
    const float* a = act ? act + (int64_t)b * 3 * HW : nullptr;
#pragma unroll 4
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < HW; p += (int64_t)gridDim.x * blockDim.x) {
        float v[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            v[c] = __ldg(src + c * HW + p);
            if (a) {
                const float m = __ldg(a + c * HW + p);
                if (mode == SPAA_MASK_POS) v[c] = m > 0.f ? v[c] : 0.f;
                else if (mode == SPAA_MASK_OPEN01) v[c] = (m > 0.f && m < 1.f) ? v[c] : 0.f;
            }
        }
        uint4 lo = make_uint4(0u, 0u, 0u, 0u);
        if constexpr (F16) {
            const __half2 h0 = __floats2half2_rn(v[0], v[1]), h1 = __floats2half2_rn(v[2], 0.f);
            lo.x = *reinterpret_cast<const uint32_t*>(&h0); lo.y = *reinterpret_cast<const uint32_t*>(&h1);
        } else {
            const __nv_bfloat162 h0 = __floats2bfloat162_rn(v[0], v[1]), h1 = __floats2bfloat162_rn(v[2], 0.f);
            lo.x = *reinterpret_cast<const uint32_t*>(&h0); lo.y = *reinterpret_cast<const uint32_t*>(&h1);
        }
        uint4* o = out16 + ((int64_t)b * HW + p) * 2;
        o[0] = lo; o[1] = make_uint4(0u, 0u, 0u, 0u);
    }
}

// [x (Cx planes) | surf (Cs planes) | 0] fp32 NCHW -> zero-padded 16-channel 16-bit NHWC (the tensor-core conv1 / conv1_s operand in training)
template <bool F16>
__global__ void __launch_bounds__(kThreads) pack_nhwc16_kernel(const float* __restrict__ x, int Cx, const float* __restrict__ surf, int Cs, int64_t surf_bs,
                                                               uint4* __restrict__ out16, int64_t HW) {
    const int b = blockIdx.y;
    const float* xb = x + (int64_t)b * Cx * HW;
    const float* sb = surf ? surf + (int64_t)b * surf_bs : nullptr;
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < HW; p += (int64_t)gridDim.x * blockDim.x) {
        float v[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            float t = 0.f;
            if (c < Cx) t = __ldg(xb + (int64_t)c * HW + p);
            else if (sb && c < Cx + Cs) t = __ldg(sb + (int64_t)(c - Cx) * HW + p);
            v[c] = t;
        }
        uint32_t w[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if constexpr (F16) { const __half2 h = __floats2half2_rn(v[2 * k], v[2 * k + 1]); w[k] = *reinterpret_cast<const uint32_t*>(&h); }
            else { const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]); w[k] = *reinterpret_cast<const uint32_t*>(&h); }
        }
        uint4* o = out16 + ((int64_t)b * HW + p) * 2;
        o[0] = make_uint4(w[0], w[1], w[2], w[3]); o[1] = make_uint4(w[4], w[5], w[6], w[7]);
    }
}

// The same operand for the split-precision ("fp32-accurate") tensor-core mode: every value as three bf16 parts h = bf16(v), m = bf16(v - h),
// l = bf16(v - h - m), written as [h(16) | m(16) | l(16)] = 48 NHWC channels per pixel (include/spaa_b200.h, spaa_conv_desc.split)
__global__ void __launch_bounds__(kThreads) pack_nhwc16_split3_kernel(const float* __restrict__ x, int Cx, const float* __restrict__ surf, int Cs, int64_t surf_bs,
                                                                      uint4* __restrict__ out48, int64_t HW) {
    const int b = blockIdx.y;
    const float* xb = x + (int64_t)b * Cx * HW;
    const float* sb = surf ? surf + (int64_t)b * surf_bs : nullptr;
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < HW; p += (int64_t)gridDim.x * blockDim.x) {
        float v[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            float t = 0.f;
            if (c < Cx) t = __ldg(xb + (int64_t)c * HW + p);
            else if (sb && c < Cx + Cs) t = __ldg(sb + (int64_t)(c - Cx) * HW + p);
            v[c] = t;
        }
        uint4* o = out48 + ((int64_t)b * HW + p) * 6;
#pragma unroll
        for (int part = 0; part < 3; ++part) {
            uint32_t w[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);
                w[k] = *reinterpret_cast<const uint32_t*>(&h);
                const float2 f = __bfloat1622float2(h);
                v[2 * k] -= f.x; v[2 * k + 1] -= f.y;
            }
            o[2 * part] = make_uint4(w[0], w[1], w[2], w[3]); o[2 * part + 1] = make_uint4(w[4], w[5], w[6], w[7]);
        }
    }
}

// fp16 -> bf16 copy of an activation tensor (any dense layout: n elements, n % 8 == 0).  The 'fp16' training mode keeps its forward activations in fp16
// (the storage that meets the 2e-3 bar) while its gradients are bf16, and tcgen05 kind::f16 wants both backward-weight operands in one format: the X
// operand is re-rounded to bf16 right before the backward-weight launch (3 fewer mantissa bits in the weight gradient only).
__global__ void __launch_bounds__(kThreads) half_to_bf16_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int64_t n8) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
        const uint4 u = __ldg(src + i);
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
        uint32_t o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[k]));
            const __nv_bfloat162 h = __floats2bfloat162_rn(f.x, f.y);
            o[k] = *reinterpret_cast<const uint32_t*>(&h);
        }
        dst[i] = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

// PerC-AL projection (perc_al/__init__.py:211-215, :15-18)
__global__ void __launch_bounds__(kThreads) percal_project_kernel(const float* __restrict__ base, int64_t base_bs, float* __restrict__ delta,
                                                                  float* __restrict__ xq, float* __restrict__ xsum, float* __restrict__ l2sum, int64_t HW,
                                                                  float* __restrict__ partial, unsigned* __restrict__ counter) {
    __shared__ float red[32];
    __shared__ bool is_last;
    const int b = blockIdx.y;
    const float* bb = base + (int64_t)b * base_bs;
    float* db = delta + (int64_t)b * 3 * HW;
    float s[1] = {0.f};
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < HW; p += (int64_t)gridDim.x * blockDim.x) {
        float nn = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int64_t o = (int64_t)c * HW + p;
            const float x0 = __ldg(bb + o);
            const float cl = fminf(fmaxf(x0 + db[o], 0.f), 1.f);
            const float d = cl - x0;
            db[o] = d;
            const float xs = x0 + d;
            if (xsum) xsum[(int64_t)b * 3 * HW + o] = xs;
            xq[(int64_t)b * 3 * HW + o] = rintf(xs * 255.f) / 255.f;      // torch.round: half to even
            nn += d * d;
        }
        s[0] += sqrtf(nn);
    }
    row_reduce_finish<1>(s, b, partial, counter, l2sum, red, &is_last);
}

// sums[b] = sum_p ||x - ref||_2 ; optional gradient c * (x-ref)/||x-ref|| ADDED into grad for rows with sel (NULL = all),
// after zeroing grad entries where x is outside [0,1] when apply_clamp_mask (projector_based_attack.py:265,275).
__global__ void __launch_bounds__(kThreads) chan_l2_kernel(const float* __restrict__ x, const float* __restrict__ ref, int64_t ref_bs, int64_t HW, float c,
                                                           const uint8_t* __restrict__ sel, int apply_clamp, float* __restrict__ sums,
                                                           float* __restrict__ grad, float* __restrict__ partial, unsigned* __restrict__ counter) {
    __shared__ float red[32];
    __shared__ bool is_last;
    const int b = blockIdx.y;
    const float* xb = x + (int64_t)b * 3 * HW;
    const float* rb = ref + (int64_t)b * ref_bs;
    float* gb = grad ? grad + (int64_t)b * 3 * HW : nullptr;
    const bool add = grad && (!sel || sel[b]);
    float s[1] = {0.f};
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < HW; p += (int64_t)gridDim.x * blockDim.x) {
        float xv[3], d[3];
        float nn = 0.f;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            xv[ch] = __ldg(xb + ch * HW + p);
            d[ch] = xv[ch] - __ldg(rb + ch * HW + p);
            nn += d[ch] * d[ch];
        }
        const float nrm = sqrtf(nn);
        s[0] += nrm;
        if (gb) {
            const float inv = (add && nrm > 0.f) ? c / nrm : 0.f;
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                float gv = gb[ch * HW + p];
                if (apply_clamp && !in_range(xv[ch], 0.f, 1.f)) gv = 0.f;
                gb[ch * HW + p] = gv + inv * d[ch];
            }
        }
    }
    row_reduce_finish<1>(s, b, partial, counter, sums, red, &is_last);
}

// one warp per sample: argmax, softmax top-1 probability, logit margin
struct LogitInfo { int arg; float pmax; float real; float other; };
SPAA_D LogitInfo logit_info(const float* __restrict__ l, int ncls, int label) {
    const int lane = threadIdx.x & 31;
    float mx = -INFINITY; int arg = 0; float other = -INFINITY;
    for (int j = lane; j < ncls; j += 32) {
        const float v = __ldg(l + j);
        if (v > mx) { mx = v; arg = j; }
        if (j != label && v > other) other = v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float m2 = __shfl_xor_sync(0xffffffffu, mx, o);
        const int a2 = __shfl_xor_sync(0xffffffffu, arg, o);
        const float o2 = __shfl_xor_sync(0xffffffffu, other, o);
        if (m2 > mx || (m2 == mx && a2 < arg)) { mx = m2; arg = a2; }
        other = fmaxf(other, o2);
    }
    float se = 0.f;
    for (int j = lane; j < ncls; j += 32) se += expf(__ldg(l + j) - mx);
    se = warp_sum(se);
    LogitInfo r;
    r.arg = arg; r.pmax = 1.f / se; r.real = __ldg(l + label); r.other = other;
    return r;
}

__global__ void __launch_bounds__(kThreads) spaa_masks_kernel(const float* __restrict__ logits, int ncls, const int64_t* __restrict__ target, int targeted,
                                                              const float* __restrict__ stats, const float* __restrict__ prjl2sum, float inv_hw_cam,
                                                              float inv_hw_prj, float w_prjl2, float w_caml2, float w_camde, float d_thr, float p_thresh,
                                                              int B, uint8_t* __restrict__ use_col, uint8_t* __restrict__ succ, uint8_t* __restrict__ better,
                                                              float* __restrict__ col_loss, float* __restrict__ best_col) {
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= B) return;
    const int tgt = (int)target[b];
    const LogitInfo li = logit_info(logits + (int64_t)b * ncls, ncls, tgt);
    if ((threadIdx.x & 31) != 0) return;
    const float camde = stats[b * 4] * inv_hw_cam, caml2 = stats[b * 4 + 1] * inv_hw_cam;
    const float prjl2 = prjl2sum ? prjl2sum[b] * inv_hw_prj : 0.f;
    const float col = w_prjl2 * prjl2 + w_caml2 * caml2 + w_camde * camde;
    const bool high_conf = li.pmax > p_thresh;
    const bool high_pert = caml2 * 255.f > d_thr;
    bool s, u;
    if (targeted) { s = li.arg == tgt; u = s && high_conf && high_pert; }
    else { s = li.arg != tgt; u = s && high_pert; }
    const bool bt = u && (col < best_col[b]);
    if (bt) best_col[b] = col;
    use_col[b] = u; succ[b] = s; better[b] = bt; col_loss[b] = col;
}

__global__ void __launch_bounds__(kThreads) percal_masks_kernel(const float* __restrict__ logits, int ncls, const int64_t* __restrict__ labels, int mode,
                                                                float margin, const float* __restrict__ l2sum, float inv_hw, float d_thr, float p_thresh,
                                                                const float* __restrict__ stats, int B, uint8_t* __restrict__ isadv,
                                                                uint8_t* __restrict__ use_col, uint8_t* __restrict__ better, float* __restrict__ dis,
                                                                float* __restrict__ best_dis) {
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= B) return;
    const int lab = (int)labels[b];
    const LogitInfo li = logit_info(logits + (int64_t)b * ncls, ncls, lab);
    if ((threadIdx.x & 31) != 0) return;
    const bool high_pert = l2sum[b] * inv_hw * 255.f > d_thr;
    const bool high_conf = li.pmax > p_thresh;
    bool a, u;
    if (mode == 2) { a = (li.real - li.other) <= -margin; u = a && high_pert; }          // untargeted with confidence margin
    else if (mode == 1) { a = li.arg == lab; u = a && high_conf && high_pert; }          // targeted
    else { a = li.arg != lab; u = a && high_pert; }                                      // untargeted, plain argmax
    const float dv = sqrtf(stats[b * 4 + 2]);
    const bool bt = u && (dv < best_dis[b]);
    if (bt) best_dis[b] = dv;
    isadv[b] = a; use_col[b] = u; better[b] = bt; dis[b] = dv;
}

inline dim3 row_grid(int64_t n, int64_t B, int per_thread = 4) {
    int64_t g = (n / per_thread + kThreads - 1) / kThreads;
    const int64_t cap = ((int64_t)kNumSMs * 8 + B - 1) / B;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return dim3((unsigned)g, (unsigned)B);
}

}  // namespace

extern "C" {

int64_t spaa_rownorm_ws_bytes(int64_t B, int64_t n) { return row_reduce_ws_bytes(B, row_reduce_nblk(n, kPerBlock), 1); }

int spaa_row_sqnorm(const float* g, const float* x_for_clamp, float lo, float hi, int64_t B, int64_t n, float* sq, void* ws, spaa_stream_t stream) {
    SPAA_CHECK_ARG(g && sq && ws && B > 0 && B < 65536 && n > 0, "spaa_row_sqnorm: bad arguments");
    const int nblk = row_reduce_nblk(n, kPerBlock);
    float* partial = (float*)ws;
    unsigned* counter = (unsigned*)(partial + B * nblk);
    row_sqnorm_kernel<<<dim3(nblk, (unsigned)B), kThreads, 0, (cudaStream_t)stream>>>(g, x_for_clamp, lo, hi, n, sq, partial, counter);
    SPAA_CHECK_LAUNCH("spaa_row_sqnorm");
    return SPAA_OK;
}

int spaa_row_normalized_step(float* x, const float* g, const float* sq, const uint8_t* sel, const float* step2, int use_clamp_mask, float lo, float hi,
                             const float* base, int64_t base_bstride, float* sum_out, float* copy_dst, const uint8_t* copy_sel, int64_t B, int64_t n,
                             spaa_stream_t stream) {
    SPAA_CHECK_ARG(x && g && sq && step2 && B > 0 && B < 65536 && n > 0, "spaa_row_normalized_step: bad arguments");
    SPAA_CHECK_ARG((sum_out == nullptr) || base, "spaa_row_normalized_step: sum_out needs base");
    SPAA_CHECK_ARG((copy_dst == nullptr) == (copy_sel == nullptr), "spaa_row_normalized_step: copy_dst and copy_sel go together");
    row_step_kernel<<<row_grid(n, B, ((n & 3) == 0 && (base_bstride & 3) == 0) ? 4 : 1), kThreads, 0, (cudaStream_t)stream>>>(x, g, sq, sel, step2, use_clamp_mask, lo, hi, base, base_bstride, sum_out,
                                                                             copy_dst, copy_sel, n);
    SPAA_CHECK_LAUNCH("spaa_row_normalized_step");
    return SPAA_OK;
}

int spaa_masked_copy_rows(float* dst, const float* src, const uint8_t* sel, int64_t B, int64_t n, spaa_stream_t stream) {
    SPAA_CHECK_ARG(dst && src && sel && B > 0 && B < 65536 && n > 0, "spaa_masked_copy_rows: bad arguments");
    masked_copy_kernel<<<row_grid(n, B), kThreads, 0, (cudaStream_t)stream>>>(dst, src, sel, n);
    SPAA_CHECK_LAUNCH("spaa_masked_copy_rows");
    return SPAA_OK;
}

int spaa_select_cotangent(const float* g0, const float* g1, const uint8_t* sel, const float* act, int mask_mode, float* out, int64_t B, int64_t n,
                          spaa_stream_t stream) {
    SPAA_CHECK_ARG(g0 && out && B > 0 && B < 65536 && n > 0 && (!sel || g1), "spaa_select_cotangent: bad arguments");
    select_cot_kernel<<<row_grid(n, B, 1), kThreads, 0, (cudaStream_t)stream>>>(g0, g1, sel, act, mask_mode, out, n);
    SPAA_CHECK_LAUNCH("spaa_select_cotangent");
    return SPAA_OK;
}

int spaa_select_cotangent_packed(const float* g0, const float* g1, const uint8_t* sel, const float* act, int mask_mode, void* out16, int dtype, int64_t B,
                                 int64_t HW, spaa_stream_t stream) {
    SPAA_CHECK_ARG(g0 && out16 && B > 0 && B < 65536 && HW > 0 && (!sel || g1) && (dtype == 1 || dtype == 2), "spaa_select_cotangent_packed: bad arguments");
    if (dtype == 2) select_cot_packed_kernel<true><<<row_grid(HW, B, 1), kThreads, 0, (cudaStream_t)stream>>>(g0, g1, sel, act, mask_mode, (uint4*)out16, HW);
    else select_cot_packed_kernel<false><<<row_grid(HW, B, 1), kThreads, 0, (cudaStream_t)stream>>>(g0, g1, sel, act, mask_mode, (uint4*)out16, HW);
    SPAA_CHECK_LAUNCH("spaa_select_cotangent_packed");
    return SPAA_OK;
}

int spaa_half_to_bf16(const void* src, void* dst, int64_t n, spaa_stream_t stream) {
    SPAA_CHECK_ARG(src && dst && n > 0 && (n & 7) == 0, "spaa_half_to_bf16: n must be a positive multiple of 8");
    int64_t blocks = (n / 8 + kThreads - 1) / kThreads;
    if (blocks > 8 * kNumSMs) blocks = 8 * kNumSMs;
    half_to_bf16_kernel<<<(unsigned)blocks, kThreads, 0, (cudaStream_t)stream>>>((const uint4*)src, (uint4*)dst, n / 8);
    SPAA_CHECK_LAUNCH("spaa_half_to_bf16");
    return SPAA_OK;
}

int spaa_pack_nhwc16_split3(const float* x, int Cx, const float* surf, int Cs, int64_t surf_bstride, void* out48, int64_t B, int64_t HW, spaa_stream_t stream) {
    SPAA_CHECK_ARG(x && out48 && B > 0 && B < 65536 && HW > 0 && Cx >= 1 && Cs >= 0 && Cx + Cs <= 16 && (Cs == 0 || surf), "spaa_pack_nhwc16_split3: bad arguments");
    pack_nhwc16_split3_kernel<<<row_grid(HW, B, 1), kThreads, 0, (cudaStream_t)stream>>>(x, Cx, Cs ? surf : nullptr, Cs, surf_bstride, (uint4*)out48, HW);
    SPAA_CHECK_LAUNCH("spaa_pack_nhwc16_split3");
    return SPAA_OK;
}

int spaa_pack_nhwc16(const float* x, int Cx, const float* surf, int Cs, int64_t surf_bstride, void* out16, int dtype, int64_t B, int64_t HW,
                     spaa_stream_t stream) {
    SPAA_CHECK_ARG(x && out16 && Cx > 0 && Cs >= 0 && Cx + Cs <= 16 && (Cs == 0 || surf) && B > 0 && B < 65536 && HW > 0 && (dtype == 1 || dtype == 2),
                   "spaa_pack_nhwc16: bad arguments");
    if (dtype == 2) pack_nhwc16_kernel<true><<<row_grid(HW, B, 1), kThreads, 0, (cudaStream_t)stream>>>(x, Cx, Cs ? surf : nullptr, Cs, surf_bstride, (uint4*)out16, HW);
    else pack_nhwc16_kernel<false><<<row_grid(HW, B, 1), kThreads, 0, (cudaStream_t)stream>>>(x, Cx, Cs ? surf : nullptr, Cs, surf_bstride, (uint4*)out16, HW);
    SPAA_CHECK_LAUNCH("spaa_pack_nhwc16");
    return SPAA_OK;
}

int64_t spaa_percal_project_ws_bytes(int64_t B, int64_t HW) { return row_reduce_ws_bytes(B, row_reduce_nblk(HW, 1024), 1); }

int spaa_percal_project(const float* base, int64_t base_bstride, float* delta, float* xq, float* xsum, float* l2sum, int64_t B, int64_t HW, void* ws,
                        spaa_stream_t stream) {
    SPAA_CHECK_ARG(base && delta && xq && l2sum && ws && B > 0 && B < 65536 && HW > 0, "spaa_percal_project: bad arguments");
    const int nblk = row_reduce_nblk(HW, 1024);
    float* partial = (float*)ws;
    unsigned* counter = (unsigned*)(partial + B * nblk);
    percal_project_kernel<<<dim3(nblk, (unsigned)B), kThreads, 0, (cudaStream_t)stream>>>(base, base_bstride, delta, xq, xsum, l2sum, HW, partial, counter);
    SPAA_CHECK_LAUNCH("spaa_percal_project");
    return SPAA_OK;
}

int64_t spaa_chan_l2_ws_bytes(int64_t B, int64_t HW) { return row_reduce_ws_bytes(B, row_reduce_nblk(HW, 1024), 1); }

int spaa_chan_l2_fwd_bwd(const float* x, const float* ref, int64_t ref_bstride, int64_t B, int64_t HW, float c, const uint8_t* sel, int apply_clamp_mask,
                         float* sums, float* grad, void* ws, spaa_stream_t stream) {
    SPAA_CHECK_ARG(x && ref && sums && ws && B > 0 && B < 65536 && HW > 0, "spaa_chan_l2_fwd_bwd: bad arguments");
    const int nblk = row_reduce_nblk(HW, 1024);
    float* partial = (float*)ws;
    unsigned* counter = (unsigned*)(partial + B * nblk);
    chan_l2_kernel<<<dim3(nblk, (unsigned)B), kThreads, 0, (cudaStream_t)stream>>>(x, ref, ref_bstride, HW, c, sel, apply_clamp_mask, sums, grad, partial,
                                                                                  counter);
    SPAA_CHECK_LAUNCH("spaa_chan_l2_fwd_bwd");
    return SPAA_OK;
}

int spaa_attack_masks(const float* logits, int ncls, const int64_t* target, int targeted, const float* stats, const float* prjl2sum, int64_t HW_cam,
                      int64_t HW_prj, float w_prjl2, float w_caml2, float w_camde, float d_thr, float p_thresh, int64_t B, uint8_t* use_col, uint8_t* succ,
                      uint8_t* better, float* col_loss, float* best_col, spaa_stream_t stream) {
    SPAA_CHECK_ARG(logits && target && stats && use_col && succ && better && col_loss && best_col && B > 0 && ncls > 1 && HW_cam > 0,
                   "spaa_attack_masks: bad arguments");
    const int wpb = kThreads / 32;
    spaa_masks_kernel<<<(unsigned)((B + wpb - 1) / wpb), kThreads, 0, (cudaStream_t)stream>>>(
        logits, ncls, target, targeted, stats, prjl2sum, 1.f / (float)HW_cam, HW_prj > 0 ? 1.f / (float)HW_prj : 0.f, w_prjl2, w_caml2, w_camde, d_thr,
        p_thresh, (int)B, use_col, succ, better, col_loss, best_col);
    SPAA_CHECK_LAUNCH("spaa_attack_masks");
    return SPAA_OK;
}

int spaa_percal_masks(const float* logits, int ncls, const int64_t* labels, int mode, float margin, const float* l2sum, int64_t HW, float d_thr,
                      float p_thresh, const float* stats, int64_t B, uint8_t* isadv, uint8_t* use_col, uint8_t* better, float* dis, float* best_dis,
                      spaa_stream_t stream) {
    SPAA_CHECK_ARG(logits && labels && l2sum && stats && isadv && use_col && better && dis && best_dis && B > 0 && ncls > 1 && HW > 0 && mode >= 0 &&
                       mode <= 2,
                   "spaa_percal_masks: bad arguments");
    const int wpb = kThreads / 32;
    percal_masks_kernel<<<(unsigned)((B + wpb - 1) / wpb), kThreads, 0, (cudaStream_t)stream>>>(logits, ncls, labels, mode, margin, l2sum, 1.f / (float)HW,
                                                                                               d_thr, p_thresh, stats, (int)B, isadv, use_col, better, dis,
                                                                                               best_dis);
    SPAA_CHECK_LAUNCH("spaa_percal_masks");
    return SPAA_OK;
}

}  // extern "C"

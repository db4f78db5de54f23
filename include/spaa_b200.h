/* spaa_b200 -- C ABI of the B200-native SPAA hot path (libspaa_b200.so, sm_100a).
 *
 * The reference (BingyaoHuang/SPAA) is pure Python on stock PyTorch ops: it has no FFI of its own.  Each entry
 * point below replaces the PyTorch library calls the reference makes at the cited place
 * (paths relative to /root/reference/src/python/).  INTEGRATION.md shows the ctypes binding a maintainer adds.
 *
 * Conventions
 *  - every function returns 0 on success, <0 on error (never throws); spaa_last_error() gives a thread-local text
 *  - all pointers are DEVICE pointers unless named host_*; nothing is allocated or synchronised inside
 *  - `stream` is a cudaStream_t; calls are asynchronous and CUDA-graph capturable
 *  - images are fp32 NCHW planar ("plane" = H*W floats); conv activations are described by explicit strides
 */
#ifndef SPAA_B200_H
#define SPAA_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef void* spaa_stream_t; /* cudaStream_t */

const char* spaa_last_error(void);
int spaa_abi_version(void);

/* ----------------------------------------------------------------------------------------------------------
 * Colour: sRGB->Lab and the reference's CIEDE2000 variant
 * replaces perc_al/differential_color_functions.py:12-64 (rgb2xyz, xyz_lab, rgb2lab_diff) and :67-180
 * (hpf_diff, dhpf_diff, ahpf_diff, ciede2000_diff) and their autograd graphs.
 * Tensors are [B,3,HW] planar fp32.  *_bstride = elements between consecutive batch items (0 = broadcast).
 * -------------------------------------------------------------------------------------------------------- */
/* fast != 0: the hardware-approximation arithmetic of spaa_color_loss_fwd_bwd(fast = 1) -- a reference Lab image for that mode must be
 * computed with it, so that a pixel equal to its reference keeps dE = 0 exactly. */
int spaa_rgb2lab_fwd(const float* rgb, float* lab, int64_t B, int64_t HW, int fast, spaa_stream_t stream);
int spaa_rgb2lab_bwd(const float* rgb, const float* dlab, float* drgb, int64_t B, int64_t HW, spaa_stream_t stream);
int spaa_de2000_fwd(const float* lab1, int64_t lab1_bstride, const float* lab2, int64_t lab2_bstride, float* de,
                    int64_t B, int64_t HW, spaa_stream_t stream);
/* dlab1 / dlab2 may be NULL; cot is the [B,HW] cotangent of the dE map */
int spaa_de2000_bwd(const float* lab1, int64_t lab1_bstride, const float* lab2, int64_t lab2_bstride, const float* cot,
                    float* dlab1, float* dlab2, int64_t B, int64_t HW, spaa_stream_t stream);

/* Fused stealth-loss kernel of the attack loops: replaces projector_based_attack.py:279-283 (caml2, camdE) and
 * perc_al/__init__.py:197-199, plus the backward of those terms to the camera image.
 *   cam        [B,3,HW]  camera image (PCNet output / inputs+delta)
 *   ref_rgb    [*,3,HW]  un-attacked scene, ref_lab its precomputed Lab (spaa_rgb2lab_fwd)
 *   cam_is_lab2: 0 -> dE(lab(cam), ref_lab) (SPAA argument order), 1 -> dE(ref_lab, lab(cam)) (PerC-AL order)
 *   stats      [B,4] out: sum_p dE, sum_p ||cam-ref||_2, sum_p dE^2, 0        (sums over pixels, not means)
 *   grad       [B,3,HW] out or NULL: c_de * d(sum_p w_p dE_p)/dcam + c_l2 * d(sum_p ||.||_2)/dcam,
 *              with w_p = 1 (de_weighting 0) or w_p = dE_p (de_weighting 1: gradient of 0.5*sum dE^2)
 *   fast       0: IEEE division / square root and accurate powf, cbrtf, sincosf, expf (the arithmetic held to 1e-5 against the reference);
 *              1: hardware approximations (relative error ~1e-6) for the 16-bit tensor-core modes, whose camera image already carries
 *              ~3e-4 of rounding: the kernel is bound by instruction issue and executes ~40 % fewer instructions (measured +1.3 % it/s);
 *              ref_lab must then come from spaa_rgb2lab_fwd(fast = 1)
 *   ws         workspace of spaa_color_loss_ws_bytes(B,HW) bytes, zero-initialised once by the caller
 */
int64_t spaa_color_loss_ws_bytes(int64_t B, int64_t HW);
int spaa_color_loss_fwd_bwd(const float* cam, const float* ref_rgb, const float* ref_lab, int64_t ref_bstride,
                            int64_t B, int64_t HW, int cam_is_lab2, int de_weighting, float c_de, float c_l2, int fast,
                            float* stats, float* grad, void* ws, spaa_stream_t stream);

/* ----------------------------------------------------------------------------------------------------------
 * Warping: affine o thin-plate-spline sampling grid and bilinear grid_sample
 * replaces models.py:163-185 (WarpingNet.forward: F.affine_grid, pytorch_tps.tps_grid, 2x F.grid_sample),
 * pytorch_tps.py:29-106 (tps, tps_grid).  Grids are channel-planar [2,H,W] (x plane then y plane), in [-1,1].
 * -------------------------------------------------------------------------------------------------------- */
/* plain TPS grid (pytorch_tps.tps_grid); theta [(T+2),2] reduced form, ctrl [T,2] */
int spaa_tps_grid_fwd(const float* theta, const float* ctrl, int T, int H, int W, float* grid, spaa_stream_t stream);
/* coarse grid = grid_sample(affine_grid(affine; Hin x Win), tps_grid(theta; H x W)) */
int spaa_coarse_grid_fwd(const float* affine, const float* theta, const float* ctrl, int T, int Hin, int Win, int H,
                         int W, float* grid, spaa_stream_t stream);
/* gradients of the coarse grid wrt affine [2,3] and theta [(T+2),2]; ws: spaa_coarse_grid_ws_bytes(T,H,W), zeroed once.
 * affine may be NULL (then Hin/Win are ignored and this is the backward of spaa_tps_grid_fwd). */
int64_t spaa_coarse_grid_ws_bytes(int T, int H, int W);
int spaa_coarse_grid_bwd(const float* affine, const float* theta, const float* ctrl, int T, int Hin, int Win, int H,
                         int W, const float* dgrid, float* daffine, float* dtheta, void* ws, spaa_stream_t stream);
/* fine = clamp(coarse + refine, -1, 1) and its backward (models.py:176-178) */
int spaa_grid_finish_fwd(const float* coarse, const float* refine /*nullable*/, float* fine, int64_t n,
                         spaa_stream_t stream);
int spaa_grid_finish_bwd(const float* coarse, const float* refine /*nullable*/, const float* dfine, float* dsum,
                         int64_t n, spaa_stream_t stream);

/* out[b,c,y,x] = bilinear(img[b,c], grid[:,y,x]) (zeros padding, align_corners=True) [* mask[y,x]]
 *   clamp01: sample clamp(img,0,1) instead of img   (projector_based_attack.py:265)
 *   mask   : [H*W] or NULL                          (models.py:340)
 *   rough  : if non-NULL, out2[b,c] = out[b,c] * rough[b?,c] (models.py:342, "x*s"); out2/rough have their own
 *            batch strides so out2 can be the tail channels of a wider tensor
 *   grid_bstride: 0 when one grid serves the whole batch                                                    */
int spaa_grid_sample_fwd(const float* img, int64_t B, int C, int Hi, int Wi, const float* grid, int64_t grid_bstride,
                         int H, int W, int clamp01, const float* mask, float* out, const float* rough,
                         int64_t rough_bstride, float* out2, int64_t out2_bstride, spaa_stream_t stream);
/* The same 3-channel warp written as the zero-padded 16-channel 16-bit NHWC tensor [B,H,W,16] the tensor-core
 * convolutions read: channels 0-2 = bilinear(clamp(img)) * mask (conv1 input), 3-5 = rough (the surface image),
 * 6-8 = their product (conv1_s input is channels 3-8), 9-15 = 0.  dtype: 1 bf16, 2 fp16.  rough may be NULL. */
int spaa_grid_sample_fwd_packed(const float* img, int64_t B, int Hi, int Wi, const float* grid, int64_t grid_bstride, int H,
                                int W, int clamp01, const float* mask, const float* rough, int64_t rough_bstride, void* out16,
                                int dtype, spaa_stream_t stream);
/* Gather form of the same adjoint for a grid that stays FIXED over many calls (the attack loops: frozen model, one grid per attack).
 * spaa_warp_taps lists for every output pixel p and bilinear tap k (entry k*H*W + p) the input pixel it reads (ent_q; Hi*Wi = unused tap)
 * and its bilinear weight (ent_w).  The caller sorts the entries by ent_q once (stable), keeps p, the weight and mask[p] per entry (ent_p,
 * ent_w, ent_m) and the CSR offsets row_ptr[Hi*Wi + 1]; spaa_grid_sample_bwd_gather then computes
 *   dimg[b,c,q] = sum_{e in row q} ((dout[b,c,p_e] + dout2[b,c,p_e] * rough[b,c,p_e]) * m_e) * w_e          (no atomics, dimg fully overwritten)
 * and, when sq != NULL, sq[b] = sum_{c,q} dimg[b,c,q]^2 over the entries whose x_for_clamp[b,c,q] lies in [lo,hi] (all if NULL): the backward of
 * grid_sample (models.py:184) + clamp (projector_based_attack.py:265) and the norm of :307,315 in one pass.  ws: spaa_grid_sample_bwd_gather_ws_bytes, zeroed once. */
int spaa_warp_taps(const float* grid, const float* mask, int Hi, int Wi, int H, int W, int32_t* ent_q, float* ent_w,
                   spaa_stream_t stream);
int64_t spaa_grid_sample_bwd_gather_ws_bytes(int64_t B, int Hi, int Wi);
int spaa_grid_sample_bwd_gather(const float* dout, const float* dout2, int64_t dout2_bstride, const float* rough,
                                int64_t rough_bstride, const int32_t* row_ptr, const int32_t* ent_p, const float* ent_w,
                                const float* ent_m, int64_t B, int C, int Hi, int Wi, int H, int W, const float* x_for_clamp,
                                float lo, float hi, float* dimg, float* sq, void* ws, spaa_stream_t stream);
/* Tiled form of spaa_grid_sample_bwd_gather (bit-identical results): a block owns a 32 x 32 tile of input pixels and first copies the compact
 * rectangle of output pixels that contribute to it into shared memory.  boxes[tile] = (y0, x0, h, w) of that rectangle (int32 x 4 per tile, tiles
 * row-major over ceil(Hi/32) x ceil(Wi/32)); ent_l[e] = (p_e / W - y0) * w + (p_e % W - x0) for the tile of entry e's row; max_region = max h * w
 * (C * max_region floats of shared memory).  ws: spaa_grid_sample_bwd_gather_tiled_ws_bytes, zeroed once. */
int64_t spaa_grid_sample_bwd_gather_tiled_ws_bytes(int64_t B, int Hi, int Wi);
int spaa_grid_sample_bwd_gather_tiled(const float* dout, const float* dout2, int64_t dout2_bstride, const float* rough,
                                      int64_t rough_bstride, const int32_t* row_ptr, const int32_t* ent_l, const float* ent_w,
                                      const float* ent_m, const int32_t* boxes, int max_region, int64_t B, int C, int Hi, int Wi, int H,
                                      int W, const float* x_for_clamp, float lo, float hi, float* dimg, float* sq, void* ws,
                                      spaa_stream_t stream);
/* dimg (+)= scatter of (dout + dout2*rough) * mask ; dimg must be zero-filled by the caller (atomic accumulate);
 * clamp01: zero the gradient where img is outside [0,1] is NOT applied here (applied by the consumer). */
int spaa_grid_sample_bwd_input(const float* dout, const float* dout2, int64_t dout2_bstride, const float* rough,
                               int64_t rough_bstride, const float* mask, const float* grid, int64_t grid_bstride,
                               int64_t B, int C, int Hi, int Wi, int H, int W, float* dimg, spaa_stream_t stream);
/* dgrid[2,H,W] (grid_bstride==0: summed over the batch; else [B,2,H,W]) */
int spaa_grid_sample_bwd_grid(const float* dout, const float* dout2, int64_t dout2_bstride, const float* rough,
                              int64_t rough_bstride, const float* mask, const float* img, int clamp01, const float* grid,
                              int64_t grid_bstride, int64_t B, int C, int Hi, int Wi, int H, int W, float* dgrid,
                              spaa_stream_t stream);

/* ----------------------------------------------------------------------------------------------------------
 * Convolution / transposed convolution as one strided "gather conv":
 *   out[b,oy,ox,co] = epi( sum_{r,s,ci} in[b,(oy*stride+r-pad)/up,(ox*stride+s-pad)/up,ci] * w[r,s,ci,co] )
 * (terms whose division is inexact or whose coordinates are out of range are skipped).
 * up==1: nn.Conv2d (models.py:18-46,223-252 via F.conv2d); up>1: nn.ConvTranspose2d with the kernel flipped and
 * pad = k-1-padding; the backward-data passes of both are the same operation with swapped roles (see
 * spaa_b200/conv_plan.py).  Replaces cudnnConvolutionForward/BackwardData/BackwardFilter for those modules.
 * dtype: 0 = fp32 storage, 1 = bf16 storage (fp32 accumulate).  Addressing is by explicit element strides so NCHW
 * and NHWC tensors (and channel slices of wider tensors) can be read/written without repacking.
 * -------------------------------------------------------------------------------------------------------- */
enum {
    SPAA_EPI_RELU = 1,        /* v = max(v,0) */
    SPAA_EPI_LEAKY01 = 2,     /* v = v>0 ? v : 0.1 v                    (models.py:138) */
    SPAA_EPI_CLAMP_MAX1 = 4,  /* v = min(v,1) after the activation      (models.py:301) */
    SPAA_EPI_ADD_AFTER_ACT = 8, /* add `add` after the activation instead of before */
    SPAA_EPI_OUT2_BF16 = 16   /* spaa_conv_tc_fwd, fp16 NHWC output only: out2 (no mask2) receives the SAME fp32 result rounded to bf16 -- the
                                 operand format of the tensor-core backward-weight kernel in the fp16 training mode (gradients are bf16) */
};
enum {
    SPAA_MASK_NONE = 0,
    SPAA_MASK_POS = 1,    /* v *= (m > 0)              backward of ReLU expressed on its output */
    SPAA_MASK_LEAKY01 = 2,/* v *= (m > 0 ? 1 : 0.1)    backward of LeakyReLU(0.1) */
    SPAA_MASK_OPEN01 = 3  /* v *= (m > 0 && m < 1)     backward of clamp(relu(.),max=1) */
};
typedef struct spaa_conv_desc {
    int32_t in_dtype, out_dtype;         /* 0 fp32, 1 bf16, 2 fp16; add / out2 share out_dtype; masks: see below */
    int32_t B, Cin, Hin, Win, Cout, Hout, Wout;
    int32_t KH, KW, stride, up, pad_h, pad_w;
    int32_t flip;                        /* 1: tap (r,s) reads weight tap (KH-1-r, KW-1-s) */
    int64_t in_bs, in_ps, in_cs;         /* element strides: batch, pixel (iy*Win+ix), channel */
    int64_t w_ts, w_cis, w_cos;          /* fp32 weight strides: tap (r*KW+s), input channel, output channel */
    int64_t out_bs, out_ps, out_cs;      /* out and out2 */
    int64_t add_bs, add_ps, add_cs;      /* strides of `add` (add_bs = 0 broadcasts over the batch) */
    int64_t mask_bs, mask_ps, mask_cs;   /* mask and mask2 (mask_bs = 0 broadcasts) */
    int32_t epi_flags, mask_mode;
    int32_t split;                       /* tensor-core path only. 1: bf16x3 split-precision ("fp32-accurate") operands: every logical channel c of the
                                          * input / add / masks / 16-bit output is stored as three bf16 parts h, m, l (v = h + m + l to 24 bits) at
                                          * physical channels c, C + c, 2C + c of a dense NHWC tensor with 3C channels; Cin / Cout stay LOGICAL
                                          * counts, the *_ps strides are physical (3C); weights are packed with 6 * Cin input channels (the six part
                                          * products, see spaa_b200/ops.py:_split_weights).  fp32 planar outputs are unchanged. */
} spaa_conv_desc;
/* v = sum + bias; [v += add]; activation; [clamp]; [v += add if ADD_AFTER_ACT]; v *= mask(mask_mode); out = v;
 * out2 (nullable) = v * (mask2 > 0).
 * w: fp32, addressed by (w_ts, w_cis, w_cos) so nn.Conv2d [Cout,Cin,KH,KW] and nn.ConvTranspose2d
 * [Cin,Cout,KH,KW] parameters are read in place; bias: fp32 [Cout] or NULL; add / mask / mask2 / out2 nullable.
 * CUDA-core fp32 path (any shape).  */
int spaa_conv_fwd(const spaa_conv_desc* d, const void* in, const float* w, const float* bias, const void* add,
                  const void* mask, const void* mask2, void* out, void* out2, spaa_stream_t stream);
/* dw (fp32, += into a caller-zeroed buffer, addressed by d->w_* strides) = sum_pixels in(gathered) x dout ;
 * dbias[Cout] (fp32, +=, nullable) = sum_pixels dout.  `d` describes the FORWARD gather conv with up == 1 and
 * flip == 0; `dout` has the forward output's strides (d->out_*) and dtype d->out_dtype. */
int spaa_conv_bwd_weight(const spaa_conv_desc* d, const void* in, const void* dout, float* dw, float* dbias,
                         spaa_stream_t stream);
/* Tensor-core path of the same operation: tcgen05.mma (M=128 x N=Cout x K=16, fp32 accumulators in TMEM) fed by 4-D TMA
 * box loads of the NHWC input (one shifted box per filter tap; zero padding = TMA out-of-bound fill).  Requirements
 * (spaa_conv_tc_supported tells): bf16 or fp16 dense NHWC input with 16, 32 or 64k channels; output either the same
 * 16-bit type, dense NHWC, Cout a multiple of 32 (<= 256), or fp32 dense NCHW with Cout <= 32 (conv6 forward, conv1
 * backward); square kernels <= 3x3, (stride, up) in {(1,1), (2,1), (1,2)}.  Epilogue: bias, residual add, ReLU
 * (+ clamp for fp32 output), ReLU mask / mask2 / out2 as in spaa_conv_fwd; masks are 16-bit activations of EITHER
 * type (only their sign is used), so fp16 forward activations can mask bf16 gradients.
 * Weights are packed once per layer and direction into bf16 [tap][Cout][Cin] by spaa_conv_tc_pack_weights (which reads
 * the fp32 parameter through d->w_* / d->flip). */
int spaa_conv_tc_supported(const spaa_conv_desc* d);
/* The launch plan spaa_conv_tc_fwd would use for this layer (host arithmetic only, no CUDA call; SPAA_ERR_UNSUPPORTED when the halo kernel
 * does not cover it).  plan[12] = {CTAs per SM, epilogue warp groups, TMEM accumulator buffers, input stages, weight-ring slices (0 = weights
 * resident in shared memory), operand-ring slots per group, staging blocks per warp - 1, pair mode, dynamic shared memory bytes, tiles,
 * threads per CTA, weights resident}.  has_* say which epilogue operands the call will pass. */
int spaa_conv_tc_plan(const spaa_conv_desc* d, int has_add, int has_mask, int has_mask2, int32_t* plan);
int64_t spaa_conv_tc_packed_elems(const spaa_conv_desc* d);
/* cin_real / cin_offset: the fp32 parameter has cin_real input channels which sit at tensor channels
 * [cin_offset, cin_offset + cin_real) of a zero-padded d->Cin-channel NHWC activation (3-, 6-channel images padded to 16). */
int spaa_conv_tc_pack_weights(const spaa_conv_desc* d, const float* w, int cin_real, int cin_offset, void* packed,
                              spaa_stream_t stream);
/* Multi-tensor re-pack (training: after every optimiser step all layers' packed copies are refreshed by ONE launch instead of one launch per layer
 * and direction).  spaa_conv_tc_pack_job writes an opaque job record (spaa_conv_tc_pack_job_bytes() bytes, HOST memory) describing the call
 * spaa_conv_tc_pack_weights(d, w, cin_real, cin_offset, packed) would make; the caller uploads an array of such records to the device once and
 * launches spaa_conv_tc_pack_weights_multi(jobs_dev, njobs) whenever the fp32 parameters changed.  Replaces nothing in the reference (cuDNN reads the
 * fp32 parameters directly, models.py:223-252); it exists because the tensor-core kernels read 16-bit [tap][Cout][Cin] copies. */
int64_t spaa_conv_tc_pack_job_bytes(void);
int spaa_conv_tc_pack_job(const spaa_conv_desc* d, const float* w, int cin_real, int cin_offset, void* packed, void* job_host);
int spaa_conv_tc_pack_weights_multi(const void* jobs_dev, int njobs, spaa_stream_t stream);
int spaa_conv_tc_fwd(const spaa_conv_desc* d, const void* in, const void* wpacked, const float* bias, const void* add,
                     const void* mask, const void* mask2, void* out, void* out2, spaa_stream_t stream);
/* Tensor-core backward-weight (training): tcgen05.mma with BOTH operands MN-major (the contraction index is the pixel),
 * fp32 accumulation in TMEM over all pixel tiles of a CTA, one atomic flush per CTA.  `d` describes the forward gather
 * conv as for spaa_conv_bwd_weight; x (gathered operand) and dy (pointwise operand) are dense 16-bit NHWC with 16, 32, 64,
 * 128 or 256 channels, both bf16 or both fp16 (tcgen05 kind::f16 traps on mixed operand formats); cx_real / cx_off / cy_real describe zero-padded operands (real X channels
 * at [cx_off, cx_off + cx_real), real DY channels [0, cy_real)); dw is addressed by d->w_ts / w_cis (X channel) / w_cos
 * (DY channel).  d->split: x and dy point at ONE bf16 part each (Cin / Cout logical channels, pixel strides 3 * Cin / 3 * Cout) of split-precision
 * operands; the caller launches the six part products (they all accumulate into dw).  Replaces cudnnConvolutionBackwardFilter for models.py:18-46,223-252 under train_network.py:304-320. */
int spaa_conv_wgrad_tc_supported(const spaa_conv_desc* d);
int spaa_conv_wgrad_tc(const spaa_conv_desc* d, const void* x, const void* dy, float* dw, int cx_real, int cx_off, int cy_real,
                       spaa_stream_t stream);
/* The same backward-weight GEMM accumulated into a SCRATCH gradient of layout [tap][real X channel][d->Cout] (fp32, +=, 16-byte aligned,
 * spaa_conv_wgrad_tc_scratch_elems floats, zero before the first use): the DY channel is contiguous there, so the kernel's final flush is four
 * 16-byte reductions per thread instead of 16 scalar atomics.  spaa_wgrad_scatter_multi then adds the scratch gradients of up to 24 layers into
 * the parameter gradients (dw[tap * w_ts + cx * w_xs + cy * w_ys] += scratch[(tap * cx_real + cx) * Cy + cy] for cy < cy_real) in ONE launch and
 * zeroes the scratch for the next pass.  Its array arguments are HOST arrays, read during the call. */
int64_t spaa_conv_wgrad_tc_scratch_elems(const spaa_conv_desc* d, int cx_real);
int spaa_conv_wgrad_tc_scratch(const spaa_conv_desc* d, const void* x, const void* dy, float* scratch, int cx_real, int cx_off, int cy_real, spaa_stream_t stream);
int spaa_wgrad_scatter_multi(const float* const* scratch, float* const* dw, const int32_t* ntap, const int32_t* cx_real, const int32_t* Cy, const int32_t* cy_real,
                             const int64_t* w_ts, const int64_t* w_xs, const int64_t* w_ys, int n, spaa_stream_t stream);
/* out[c] += sum over pixels of a dense 16-bit NHWC tensor [npix][C] (C in 8..256, power of two); dtype 1 bf16, 2 fp16 */
int spaa_channel_sum_nhwc16(const void* x, int dtype, int64_t npix, int C, float* out, spaa_stream_t stream);
/* n <= 24 of these sums in ONE launch (the bias gradients of a whole backward pass): outs[i][c] += sum over the npix[i] pixels of xs[i]
 * (dense 16-bit NHWC with C[i] channels) for c < c_real[i] (zero-padded tensors: only the real channels are written).  One dtype for all.
 * xs / npix / C / c_real / outs are HOST arrays, read during the call. */
int spaa_channel_sum_nhwc16_multi(const void* const* xs, const int64_t* npix, const int32_t* C, const int32_t* c_real, float* const* outs, int dtype, int n,
                                  spaa_stream_t stream);
/* out[c] += sum_{b,p} x[b,p,c]  (bias gradient; x addressed by element strides, dtype 0 fp32 / 1 bf16) */
int spaa_channel_sum(const void* x, int dtype, int64_t B, int C, int64_t HW, int64_t bs, int64_t ps, int64_t cs, float* out,
                     spaa_stream_t stream);

/* ----------------------------------------------------------------------------------------------------------
 * Training loss: replaces train_network.py:367-392 (compute_loss: F.l1_loss, F.mse_loss, 1-SSIM) and
 * pytorch_ssim/__init__.py:24-61 (_ssim: 2 replicate pads + 5 depthwise 11x11 Gaussian convs) fwd+bwd.
 *   pred,target [N,H,W] planes (N = B*C); sums[4] out = sum|d|, sum d^2, sum ssim_map, 0
 *   grad (nullable) = w_l1*sign(d)/numel + w_l2*2d/numel - w_ssim * d(mean ssim)/dpred
 *   cot_map (nullable, [N,H,W]): per-pixel cotangent of the SSIM map used INSTEAD of the uniform -w_ssim/numel
 *                                (the mask / weights / size_average=False branches of pytorch_ssim :54-67)
 *   ssim_map (nullable, [N,H,W]): the SSIM map itself
 *   ws: spaa_ssim_l1_ws_bytes(N,H,W) bytes, zeroed once by the caller
 * -------------------------------------------------------------------------------------------------------- */
int64_t spaa_ssim_l1_ws_bytes(int64_t N, int H, int W);
int spaa_ssim_l1_fwd_bwd(const float* pred, const float* target, int64_t N, int H, int W, float w_l1, float w_l2,
                         float w_ssim, const float* cot_map, float* sums, float* ssim_map, float* grad, void* ws,
                         spaa_stream_t stream);

/* ----------------------------------------------------------------------------------------------------------
 * Attack-loop updates: replaces projector_based_attack.py:275,290-328 and perc_al/__init__.py:193-245
 * (per-sample gradient norms, masked normalised steps, clamp/quantise, best-so-far bookkeeping, decision masks).
 * Rows are samples: tensors are [B,n] fp32 with n = C*H*W.  Row selectors are uint8 [B] on the device.
 * -------------------------------------------------------------------------------------------------------- */
/* sq[b] = sum_n g[b,n]^2, counting g as zero where x_for_clamp (nullable) is outside [lo,hi] (torch.clamp backward).
 * ws: spaa_rownorm_ws_bytes(B,n) bytes, zeroed once by the caller. */
int64_t spaa_rownorm_ws_bytes(int64_t B, int64_t n);
int spaa_row_sqnorm(const float* g, const float* x_for_clamp /*nullable*/, float lo, float hi, int64_t B, int64_t n,
                    float* sq, void* ws, spaa_stream_t stream);
/* x[b] += step * g[b] / sqrt(sq[b]) with step = step2[sel[b] != 0] (step2: DEVICE float[2]; sel NULL -> step2[0]);
 * rows whose step is exactly 0 are not touched.  No epsilon: a zero gradient gives NaN as in the reference (:307,315).
 * use_clamp_mask: gradient entries where x (before the step) is outside [lo,hi] count as zero.
 * sum_out (nullable): sum_out[b] = base[b] + x[b] (after the step) for every row   (perc_al/__init__.py:197 input)
 * copy_dst/copy_sel (nullable): copy_dst[b] = x[b] (after the step) where copy_sel[b] != 0   (:323) */
int spaa_row_normalized_step(float* x, const float* g, const float* sq, const uint8_t* sel, const float* step2,
                             int use_clamp_mask, float lo, float hi, const float* base, int64_t base_bstride,
                             float* sum_out, float* copy_dst, const uint8_t* copy_sel, int64_t B, int64_t n,
                             spaa_stream_t stream);
/* dst[b] = src[b] where sel[b] != 0 */
int spaa_masked_copy_rows(float* dst, const float* src, const uint8_t* sel, int64_t B, int64_t n, spaa_stream_t stream);
/* out[b] = mask(act[b]; mask_mode) * (sel[b] ? g1[b] : g0[b])  -- picks each sample's cotangent (adversarial or
 * stealth) and applies the backward of the network's output activation in the same pass. act/sel nullable. */
int spaa_select_cotangent(const float* g0, const float* g1, const uint8_t* sel, const float* act, int mask_mode,
                          float* out, int64_t B, int64_t n, spaa_stream_t stream);
/* The same selection for 3-channel images [B,3,HW], written as zero-padded 16-channel 16-bit NHWC [B,HW,16]
 * (the operand of the tensor-core backward of the network's last convolution).  dtype: 1 bf16, 2 fp16. */
int spaa_select_cotangent_packed(const float* g0, const float* g1, const uint8_t* sel, const float* act, int mask_mode,
                                 void* out16, int dtype, int64_t B, int64_t HW, spaa_stream_t stream);
/* [x | surf | 0]: fp32 NCHW images x [B,Cx,HW] and surf [B or 1,Cs,HW] (nullable, Cx+Cs <= 16) written as ONE zero-padded
 * 16-channel 16-bit NHWC tensor [B,HW,16] -- the operand of the tensor-core conv1 / conv1_s (forward and backward-weight)
 * when the images come from the nn.Module API instead of the fused warp kernel.  dtype: 1 bf16, 2 fp16. */
int spaa_pack_nhwc16(const float* x, int Cx, const float* surf, int Cs, int64_t surf_bstride, void* out16, int dtype, int64_t B,
                     int64_t HW, spaa_stream_t stream);
/* dst (bf16) = src (fp16), n elements (a multiple of 8), same layout: the X operand of the tensor-core backward-weight kernel in the 'fp16' training mode
 * (fp16 forward activations, bf16 gradients; tcgen05 kind::f16 needs both operands in one format).  No reference counterpart (cuDNN wgrad reads fp32). */
int spaa_half_to_bf16(const void* src, void* dst, int64_t n, spaa_stream_t stream);
/* The same operand for the split-precision ("fp32-accurate", spaa_conv_desc.split) tensor-core mode: [B,HW,48] bf16 = three parts h, m, l of every
 * one of the 16 (zero-padded) channels, [h(16) | m(16) | l(16)] per pixel, v = h + m + l to 24 significand bits.  Stands where the reference's fp32
 * F.conv2d reads its fp32 input (models.py:223,231,239). */
int spaa_pack_nhwc16_split3(const float* x, int Cx, const float* surf, int Cs, int64_t surf_bstride, void* out48, int64_t B, int64_t HW,
                            spaa_stream_t stream);
/* PerC-AL projection: delta = clamp(base+delta,0,1)-base ; xsum (nullable) = base+delta ;
 * xq = round((base+delta)*255)/255 ; l2sum[b] = sum_p ||delta[b,:,p]||_2   (perc_al/__init__.py:211-215, :15-18).
 * Tensors [B,3,HW]; base_bstride may be 0. */
int64_t spaa_percal_project_ws_bytes(int64_t B, int64_t HW);
int spaa_percal_project(const float* base, int64_t base_bstride, float* delta, float* xq, float* xsum, float* l2sum,
                        int64_t B, int64_t HW, void* ws, spaa_stream_t stream);
/* sums[b] = sum_p ||x[b,:,p] - ref[b,:,p]||_2 (projector_based_attack.py:275).  grad (nullable, in/out):
 * entries where x is outside [0,1] are zeroed first when apply_clamp_mask, then c*(x-ref)/||x-ref|| is added on
 * rows with sel[b] != 0 (sel NULL = all rows). */
int64_t spaa_chan_l2_ws_bytes(int64_t B, int64_t HW);
int spaa_chan_l2_fwd_bwd(const float* x, const float* ref, int64_t ref_bstride, int64_t B, int64_t HW, float c,
                         const uint8_t* sel, int apply_clamp_mask, float* sums, float* grad, void* ws,
                         spaa_stream_t stream);
/* SPAA decision logic on the device (projector_based_attack.py:290-299, 318-320):
 *   logits [B,ncls]; target [B]; stats from spaa_color_loss_fwd_bwd; prjl2sum [B] (nullable)
 *   outputs (uint8 [B]): use_col (=mask_best_adv), succ (=mask_succ_adv), better (=mask_best);
 *   col_loss [B] = w_prjl2*prjl2 + w_caml2*caml2 + w_camde*camde ; best_col [B] updated in place */
int spaa_attack_masks(const float* logits, int ncls, const int64_t* target, int targeted, const float* stats,
                      const float* prjl2sum, int64_t HW_cam, int64_t HW_prj, float w_prjl2, float w_caml2, float w_camde,
                      float d_thr, float p_thresh, int64_t B, uint8_t* use_col, uint8_t* succ, uint8_t* better,
                      float* col_loss, float* best_col, spaa_stream_t stream);
/* PerC-AL decision logic (perc_al/__init__.py:215-242).  mode 0: untargeted argmax != label; 1: targeted
 * (argmax == label, p_top1 > p_thresh); 2: untargeted with logit margin (real - best other <= -margin).
 * dis[b] = sqrt(stats[b,2]) (the L2 norm of the dE map); best_dis updated in place. */
int spaa_percal_masks(const float* logits, int ncls, const int64_t* labels, int mode, float margin, const float* l2sum,
                      int64_t HW, float d_thr, float p_thresh, const float* stats, int64_t B, uint8_t* isadv,
                      uint8_t* use_col, uint8_t* better, float* dis, float* best_dis, spaa_stream_t stream);

/* ----------------------------------------------------------------------------------------------------------
 * Classifier pre-processing (the step on either side of the external classifier): replaces classifier.py:55-59
 * (centre crop, F.interpolate(mode='area'), ImageNet normalise) and its autograd graph.
 *   img [B,3,H,W] fp32; crop rectangle (top, left, crop_h, crop_w) per img_proc.py:126-132; out [B,3,out_h,out_w]
 *   (nhwc = 0) or [B,out_h,out_w,3] (nhwc = 1, what a channels_last cuDNN network reads) or, nhwc = 2, the 2x2 space-to-depth fold
 *   [B,out_h/2+3,out_w/2+3,16]: cell (I,J), channel (dy*2+dx)*3+c = pixel (2(I-2)+dy, 2(J-2)+dx) channel c, channels 12..15 and the border cells
 *   (2 before, 1 after each axis) zero -- the input of the 4x4 stride-1 form of a 7x7 stride-2 pad-3 stem convolution (out_h, out_w even);
 *   host_mean3 / host_std3: HOST float[3].
 *   bwd: dimg [B,3,H,W] = adjoint applied to dout (zero outside the crop, every element written). Supported resize factors: shrink <= 3x, enlarge <= 2x.
 * -------------------------------------------------------------------------------------------------------- */
int spaa_clf_preprocess_fwd(const float* img, int64_t B, int H, int W, int top, int left, int crop_h, int crop_w, int out_h,
                            int out_w, const float* host_mean3, const float* host_std3, int nhwc, float* out, spaa_stream_t stream);
int spaa_clf_preprocess_bwd(const float* dout, int64_t B, int H, int W, int top, int left, int crop_h, int crop_w, int out_h,
                            int out_w, const float* host_std3, int nhwc, float* dimg, spaa_stream_t stream);

/* ----------------------------------------------------------------------------------------------------------
 * Fused ReLU + max-pooling on fp32 NHWC activations, forward and adjoint: the step right after the first convolution of the
 * external classifier (torchvision ResNet stem relu -> maxpool(3,2,1), VGG ReLU -> MaxPool2d(2,2), Inception MaxPool2d(3,2);
 * networks built at classifier.py:22-33).  Replaces ATen's relu / max_pool2d(channels_last) forward and their autograd
 * (threshold_backward, max_pool2d_with_indices_backward) in the attack engines' PRIVATE copy of the frozen classifier.
 *   x [N,H,W,C] fp32 (C % 4 == 0); y [N,Ho,Wo,C]; idx [N,Ho,Wo,C] uint8: tap number (r*k + q, row-major, first maximum wins as in
 *   ATen) of the selected element, 255 = none (relu != 0 and the window maximum was <= 0); Ho = (H + 2*pad - k)/stride + 1.
 *   bias [C] (nullable): per-channel constant added before the ReLU (the preceding convolution's bias: max(x) + b == max(x + b)).
 *   bwd: dx [N,H,W,C] = adjoint applied to dy (every element written; no atomics).
 * spaa_bias_act_nhwc: y = act(x + bias[c] + res) over n = N*H*W*C elements (bias, res nullable; relu != 0: clamp at 0): replaces the
 *   `output.add_(bias)` that follows every biased cuDNN convolution in PyTorch, `out += identity` and `relu` of torchvision's
 *   BasicBlock / BasicConv2d / VGG `features` (its adjoint is ATen's threshold_backward on the saved output).
 * -------------------------------------------------------------------------------------------------------- */
int spaa_relu_maxpool_nhwc_fwd(const float* x, const float* bias, int64_t N, int H, int W, int C, int k, int stride, int pad,
                               int Ho, int Wo, int relu, float* y, uint8_t* idx, spaa_stream_t stream);
int spaa_bias_act_nhwc(const float* x, const float* bias, const float* res, int64_t n, int C, int relu, float* y,
                       spaa_stream_t stream);
int spaa_relu_maxpool_nhwc_bwd(const float* dy, const uint8_t* idx, int64_t N, int H, int W, int C, int k, int stride, int pad,
                               int Ho, int Wo, float* dx, spaa_stream_t stream);

/* ----------------------------------------------------------------------------------------------------------
 * Optimiser: replaces optim.Adam.step over the parameter groups of train_network.py:253-255,145 with one
 * launch over a flat fp32 buffer split into segments with their own lr / weight decay.
 *   seg_end[nseg] (int64, device): exclusive end offsets; seg_lr / seg_wd [nseg] (float, device)
 * -------------------------------------------------------------------------------------------------------- */
int spaa_adam_step(float* param, const float* grad, float* m, float* v, int64_t n, const int64_t* seg_end,
                   const float* seg_lr, const float* seg_wd, int nseg, float beta1, float beta2, float eps, int step,
                   float grad_scale, spaa_stream_t stream);
/* Same update with the step-dependent scalars read from DEVICE memory, so that the launch can be recorded once in a CUDA
 * graph and replayed every step: bias_corr2 = {1 - beta1^t, sqrt(1 - beta2^t)} (float[2], device); seg_lr as above (the
 * caller refreshes both buffers before each replay).  Replaces the same optim.Adam.step / scheduler.step calls
 * (train_network.py:318-320,355-357). */
int spaa_adam_step_dev(float* param, const float* grad, float* m, float* v, int64_t n, const int64_t* seg_end,
                       const float* seg_lr, const float* seg_wd, int nseg, float beta1, float beta2, float eps,
                       const float* bias_corr2, float grad_scale, spaa_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SPAA_B200_H */

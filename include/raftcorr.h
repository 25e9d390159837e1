/*
 * raftcorr.h -- C ABI of the B200-native RAFT correlation hot path (libraftcorr_b200.so).
 *
 * Drop-in boundary for andrei-iosif/optical-flow-dnn's correlation subsystem.  Every entry point
 * names the reference interface it replaces (paths relative to the reference root):
 *
 *   raft/core/corr.py:18-43      CorrBlock.__init__  (volume + pyramid)      -> raftcorr_build_fwd
 *   raft/core/corr.py:45-91      CorrBlock.__call__  (lookup)                -> raftcorr_lookup_fwd
 *   raft/core/extractor.py:231,239 + raft.py:114-119  conv2 head + CorrBlock  -> raftcorr_fnet_head_build_fwd
 *   autograd of corr.py:81       grid_sampler_2d_backward x L x iters        -> raftcorr_lookup_bwd
 *   autograd of corr.py:31-43    avg_pool2d_backward, mul, 2 x matmul bwd    -> raftcorr_build_bwd
 *   the two rows above for ALL refinement iterations of a backward pass       -> raftcorr_lookup_bwd_batched
 *                                (raft.py:141-144 x corr.py:81, then :41-43)     + raftcorr_build_bwd_folded
 *   raft/alt_cuda_corr/correlation.cpp:23-33  alt_cuda_corr.forward          -> raftcorr_alt_forward
 *   raft/alt_cuda_corr/correlation.cpp:36-49  alt_cuda_corr.backward         -> raftcorr_alt_backward
 *   raft/core/corr.py:124-140    AlternateCorrBlock.__init__ (feature pyramid)-> raftcorr_alt_pyramid
 *   raft/core/corr.py:142-167    AlternateCorrBlock.__call__ (all levels)    -> raftcorr_alt_lookup_fwd / _bwd
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless stated otherwise;
 *   - every compute call takes the CUDA device ordinal and the stream to launch on and is
 *     asynchronous with respect to the host; no call allocates or frees device memory
 *     (work space is queried with *_workspace_bytes and supplied by the caller);
 *   - return value 0 = success, anything else = error; raftcorr_last_error() returns a
 *     thread-local, human-readable description of the last failure on the calling thread;
 *   - the library keeps no mutable global state and is re-entrant: it may be called from one
 *     host thread per device concurrently (nn.DataParallel's threading model, raft/train.py:71);
 *   - all tensors are fp32.  Feature maps are NCHW [B,D,H,W] as produced by the reference's
 *     encoders (raft/core/raft.py:110-115) unless the name says nhwc;
 *   - coords are [B,2,H,W], channel 0 = x, channel 1 = y, in level-0 pixels (corr.py:57).
 */
#ifndef RAFTCORR_H_
#define RAFTCORR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RAFTCORR_ABI_VERSION 1
#define RAFTCORR_MAX_LEVELS 8

/* Opaque here: the caller passes its cudaStream_t (or 0 for the legacy default stream). */
typedef void* raftcorr_stream_t;

/* Where each pyramid level lives inside ONE flat device buffer.
 * Level l is a dense array [B][N][h[l]][pitch[l]] of fp32, N = H*W source pixels, starting
 * offset[l] floats from the buffer base.  (h[l], w[l]) = cascaded floor halving of (H, W)
 * (F.avg_pool2d(2, stride=2), corr.py:42); pitch[l] >= w[l] is the row pitch in floats
 * (multiple of 8 so every row starts on a 32-byte sector and satisfies TMA's 16-byte stride
 * rule).  The reference's corr_pyramid[l] ([B*N,1,h,w], corr.py:35-43) is the strided view
 * base+offset[l] with strides (h*pitch, -, pitch, 1). */
typedef struct raftcorr_pyramid_layout {
    int32_t B, H, W, L;
    int32_t h[RAFTCORR_MAX_LEVELS];
    int32_t w[RAFTCORR_MAX_LEVELS];
    int32_t pitch[RAFTCORR_MAX_LEVELS];
    int64_t offset[RAFTCORR_MAX_LEVELS];
    int64_t total_floats;
    int32_t interleaved; /* 0: rows one after the other, element (y,x) at y*pitch + x.
                            1: ROW PAIRS interleaved, element (y,x) at (y/2)*(2*pitch) + 2*x + (y%2), ceil(h/2) pair rows
                               per image (an odd last row is paired with zeros).  A (K+1)x(K+1) lookup window is then
                               ~(K+1)/2 runs of 2(K+1) floats instead of K+1 runs of K+1: fewer, longer DRAM bursts.
                            2: ROW QUADS interleaved, element (y,x) at (y/4)*(4*pitch) + 4*x + (y%4), ceil(h/4) groups per
                               image (the last group is completed with rows of zeros).  One column of a quad is 16
                               bytes, so a window starts on TMA's 16-byte grid at ANY x: a (K+1)x(K+1) window is 3 or 4
                               runs of exactly 4(K+1) floats (160 B at r=4), ~10.6 DRAM fills of 64 B instead of ~12.4
                               (pairs) or ~17 (rows).  Default of the tensor-core builds.
                               The tensor-core build writes and the lookup reads any of them; gradient pyramids and the
                               fp32 build use 0. */
    int32_t reserved_;
} raftcorr_pyramid_layout;

/* Build precision modes for raftcorr_build_fwd. */
enum {
    RAFTCORR_BUILD_FP32_SIMT = 0, /* exact fp32 FFMA accumulation (CUDA cores) */
    RAFTCORR_BUILD_TF32_TC = 1,   /* tcgen05.mma kind::tf32, fp32 accumulate in TMEM, TMA fed,
                                     pooling fused into the epilogue */
    RAFTCORR_BUILD_F16_TC = 2,    /* the same kernel on BLOCK-SCALED fp16 operands (kind::f16): fp16 has TF32's 11
                                     significant bits; one power-of-two scale per 64-pixel operand block (64 consecutive
                                     source pixels / one row pair of an 8x32 target tile) keeps everything down to 2^-29
                                     of the block maximum at full precision and is undone by two scalars in the epilogue.
                                     fp32 accumulate; half the operand bytes, MMA instructions and accumulator
                                     updates of the TF32 build (which the 1 kW power cap turns into time).
                                     raftcorr_build_fwd only, D <= 256; the backward and the on-the-fly entry points
                                     treat it as RAFTCORR_BUILD_TF32_TC. */
    RAFTCORR_BUILD_F16X3_TC = 3   /* fp32-class accuracy on the tensor cores: every (block-scaled) feature is split
                                     into hi = fp16(x) and lo = fp16(x - hi) and the K axis carries [hi|lo|hi] x
                                     [hi|hi|lo], so one fp32 accumulation forms hi*hi + lo*hi + hi*lo -- the product
                                     to ~2^-21 instead of 2^-11 (measured ~1e-6 of max|V|, the fp32 FFMA build's
                                     class) at 3x the MMA work and operand bytes of RAFTCORR_BUILD_F16_TC, still
                                     ~10x faster than RAFTCORR_BUILD_FP32_SIMT.  Forward only, D <= 256; the
                                     backward runs as RAFTCORR_BUILD_TF32_TC. */
};

int raftcorr_abi_version(void);
const char* raftcorr_last_error(void);

/* Tuning knob (process-wide CUDA device limit, not per call): cudaLimitMaxL2FetchGranularity for
 * `device`, in bytes (32, 64 or 128).  The lookup is a sparse gather of 40-byte runs; the default
 * 64-byte DRAM->L2 fill granularity doubles its DRAM traffic.  Returns the previous value through
 * *previous when non-NULL. */
int raftcorr_set_l2_fetch_granularity(int bytes, int device, int* previous);

/* Host-only helper: fill *layout for a [B,D,H,W] pair of feature maps and L levels.
 * Fails if L < 1, L > RAFTCORR_MAX_LEVELS or a level would collapse below 2x2 (the reference
 * divides by (W_l - 1), utils.py:74-75, and yields NaN there).  `interleaved`: see the struct. */
int raftcorr_pyramid_layout_init(raftcorr_pyramid_layout* layout, int B, int H, int W, int L, int interleaved);

/* CorrBlock.__init__ (corr.py:18-43): pyramid <- all-pairs <f1,f2>/sqrt(D) + L-1 poolings. */
size_t raftcorr_build_fwd_workspace_bytes(int B, int D, int H, int W, int mode);
int raftcorr_build_fwd(const float* fmap1, const float* fmap2, int D, float* pyramid,
                       const raftcorr_pyramid_layout* layout, int mode, void* workspace,
                       size_t workspace_bytes, int device, raftcorr_stream_t stream);

/* SURVEY.md 8(f) rank 2 -- the feature encoder's output head fused into the build's prologue:
 *   fmap = conv2(x) (1x1, raft/core/extractor.py:172,231 BasicEncoder 128 -> 256; :287,341 SmallEncoder 96 -> 128),
 *   fmap1, fmap2 = split(fmap) (:239), .float() (raft/core/raft.py:114-115), CorrBlock(fmap1, fmap2) (raft.py:116-119)
 * as ONE call: x [2B][Cin][H][W] is the head's input (output of layer3; images 0..B-1 are frame 1, B..2B-1 frame 2,
 * the order of extractor.py:215-217), weight [D][Cin] (conv2.weight [D, Cin, 1, 1]), bias [D] or NULL.  A tcgen05
 * TF32 product (fp32 accumulate; x and W rounded to nearest, the precision cuDNN's default gives this convolution)
 * writes the build's pixel-major, TF32-rounded operand buffers directly -- the NCHW feature maps never exist --
 * and the tensor-core build follows on the same stream.  TF32 build only; `layout` describes B pairs.
 * Needs Cin % 4 == 0, (H*W) % 4 == 0, D <= 256 and ceil(Cin/32)*roundup(D,32) <= 1024 (the weight stays resident in
 * shared memory): raftcorr_fnet_head_supported() returns 1 when the shapes qualify, else call conv2 yourself and
 * use raftcorr_build_fwd.  Work space: raftcorr_fnet_head_build_workspace_bytes(); after the call its first
 * raftcorr_build_fwd_workspace_bytes() bytes hold the two operand buffers [B][H*W][roundup(D,32)]. */
int raftcorr_fnet_head_supported(int Cin, int D, int H, int W);
size_t raftcorr_fnet_head_build_workspace_bytes(int B, int Cin, int D, int H, int W);
int raftcorr_fnet_head_build_fwd(const float* x, const float* weight, const float* bias, int Cin, int D, float* pyramid,
                                 const raftcorr_pyramid_layout* layout, void* workspace, size_t workspace_bytes,
                                 int device, raftcorr_stream_t stream);

/* The same head in front of the on-the-fly path: conv2 + split + .float() + AlternateCorrBlock.__init__
 * (extractor.py:231,239, raft.py:114-117, corr.py:124-140) -- frame-1 maps land in f1_nhwc, frame-2 maps in level 0 of
 * the NHWC feature pyramid (both TF32-rounded, as raftcorr_alt_pyramid writes them in mode RAFTCORR_BUILD_TF32_TC),
 * the pooled levels follow.  Needs D % 32 == 0 and raftcorr_fnet_head_supported(); work space (the rounded weight):
 * raftcorr_fnet_head_alt_workspace_bytes(). */
size_t raftcorr_fnet_head_alt_workspace_bytes(int Cin, int D);
int raftcorr_fnet_head_alt_pyramid(const float* x, const float* weight, const float* bias, int B, int Cin, int D, int H,
                                   int W, int L, float* f1_nhwc, float* f2_pyramid_nhwc, void* workspace,
                                   size_t workspace_bytes, int device, raftcorr_stream_t stream);

/* CorrBlock.__call__ (corr.py:45-91): out[B, L*(2r+1)^2, H, W]. */
int raftcorr_lookup_fwd(const float* pyramid, const raftcorr_pyramid_layout* layout, const float* coords,
                        int radius, float* out, int device, raftcorr_stream_t stream);

/* The same lookup with the per-pyramid setup (TMA tensor maps, launch geometry) done once: RAFT calls
 * the lookup 12-32 times per volume (raft.py:141-144).  The plan is a caller-owned HOST blob, valid as
 * long as `pyramid` and `layout` are; initialise it after raftcorr_build_fwd, reuse it for every call. */
#define RAFTCORR_LOOKUP_PLAN_BYTES 4096
typedef struct raftcorr_lookup_plan {
    unsigned char opaque[RAFTCORR_LOOKUP_PLAN_BYTES];
} raftcorr_lookup_plan;
int raftcorr_lookup_plan_init(raftcorr_lookup_plan* plan, const float* pyramid, const raftcorr_pyramid_layout* layout,
                              int radius, int device);
int raftcorr_lookup_fwd_planned(const raftcorr_lookup_plan* plan, const float* coords, float* out, int device,
                                raftcorr_stream_t stream);

/* SURVEY.md 8(f) rank 1 -- the lookup fused with its only consumer, the motion encoder's first 1x1 convolution:
 *   out[B, Cout, H, W] = relu( bias + W * lookup(coords) )
 * replaces `corr = corr_fn(coords1)` (raft/core/raft.py:144) followed by `F.relu(self.convc1(corr))`
 * (raft/core/update.py:135 SmallMotionEncoder, :170 BasicMotionEncoder; convc1 = nn.Conv2d(L*K*K, Cout, 1), :116,:158).
 * The [B, L*K*K, H, W] lookup result never reaches HBM: its K*K samples per level become rows of a TF32 operand tile
 * in shared memory and the product runs on tcgen05 (fp32 accumulate), level by level, behind the gather.
 *   weight [Cout][L*K*K] fp32 (conv weight [Cout, L*K*K, 1, 1], the reference's channel order); pack it once per
 *   weight update with raftcorr_convc1_pack_weight into a caller-owned device buffer of raftcorr_convc1_weight_bytes.
 *   bias [Cout] or NULL; relu != 0 applies max(., 0).  Cout: multiple of 16 up to 128, or of 32 up to 256
 *   (RAFT: 96 and 256).  The plan must describe at most 4 levels. */
size_t raftcorr_convc1_weight_bytes(int levels, int radius, int cout);
int raftcorr_convc1_pack_weight(const float* weight, int levels, int radius, int cout, void* packed, int device,
                                raftcorr_stream_t stream);
int raftcorr_lookup_convc1_fwd(const raftcorr_lookup_plan* plan, const float* coords, const void* packed_weight,
                               const float* bias, int cout, int relu, float* out, int device,
                               raftcorr_stream_t stream);

/* SURVEY.md 8(f) rank 4 -- RAFT.upsample_flow (raft/core/raft.py:70-81), called after every refinement iteration
 * (raft.py:161-176; twice per iteration by the uncertainty model, with the same mask):
 *   out[b, c, 8h+i, 8w+j] = sum_{k<9} softmax_k(mask[b, 64k + 8i + j, h, w]) * 8 * flow[b, c, h + k/3 - 1, w + k%3 - 1]
 * flow [B,C,H,W], mask [B,576,H,W], out [B,C,8H,8W], all fp32 contiguous; C in 1..4 (flow and flow variance can
 * share one pass over the mask: concatenate them along C).  One read of the mask, one write of the result.
 * _bwd: grad_out [B,C,8H,8W] -> dmask [B,576,H,W] (overwritten) and dflow [B,C,H,W] (overwritten). */
int raftcorr_upsample_flow_fwd(const float* flow, const float* mask, int B, int C, int H, int W, float* out, int device,
                               raftcorr_stream_t stream);
int raftcorr_upsample_flow_bwd(const float* flow, const float* mask, const float* grad_out, int B, int C, int H, int W,
                               float* dmask, float* dflow, int device, raftcorr_stream_t stream);

/* Adjoint of the lookup: dpyramid (same layout, ACCUMULATED into -- zero it once, then call once
 * per refinement iteration) += scatter of grad_out[B, L*(2r+1)^2, H, W].
 * dcoords [B,2,H,W] is optional (NULL to skip); when given, pyramid must be the forward pyramid. */
int raftcorr_lookup_bwd(const float* grad_out, const float* coords, const float* pyramid, float* dpyramid,
                        const raftcorr_pyramid_layout* layout, int radius, float* dcoords, int device,
                        raftcorr_stream_t stream);

/* Adjoint of the build.  dpyramid is CONSUMED (pooled-level gradients are folded into level 0
 * in place; in TF32 mode level 0 is also rounded to TF32); dfmap1/dfmap2 [B,D,H,W] are overwritten.
 * mode as in raftcorr_build_fwd: TF32 runs the two products on tcgen05 (D <= 256, else CUDA cores). */
size_t raftcorr_build_bwd_workspace_bytes(int B, int D, int H, int W, int mode);
int raftcorr_build_bwd(float* dpyramid, const raftcorr_pyramid_layout* layout, const float* fmap1,
                       const float* fmap2, int D, float* dfmap1, float* dfmap2, int mode, void* workspace,
                       size_t workspace_bytes, int device, raftcorr_stream_t stream);

/* The adjoints of ALL lookups of one backward pass in one launch, with the adjoint of the pooling cascade folded in
 * (autograd of corr.py:45-91 for every refinement iteration of raft.py:141-144 + avg_pool2d_backward of corr.py:41-43):
 *   level 0 of dpyramid  <-  [accumulate ? level 0 : 0] + sum_i fold(scatter(grad_outs[i], coords[i]))
 * where fold adds every pooled level's gradient into level 0 (dV0[y][x] += dV_l[y >> l][x >> l] / 4^l).  A source
 * pixel's gradient image (all levels) is accumulated in shared memory and written once: no atomics, no zero-filled
 * gradient pyramid, no fold passes.  grad_outs / coords: HOST arrays of `count` device pointers
 * (count <= RAFTCORR_LOOKUP_BWD_MAX_BATCH; call again with accumulate = 1 for more).  Level 0 is written in the
 * row-major gradient layout (raftcorr_pyramid_layout_init(.., interleaved = 0) of the same B, H, W, L) and is OVERWRITTEN
 * unless accumulate != 0; the other levels of dpyramid are not touched.  round_tf32 != 0 rounds the stored values to
 * TF32 (set it on the last call when raftcorr_build_bwd_folded runs in a tensor-core mode).  Feed the result to
 * raftcorr_build_bwd_folded.  raftcorr_lookup_bwd_batched_supported() returns 1 when a pixel's gradient pyramid fits
 * in shared memory (L <= 4; e.g. up to ~100x200 feature maps), else use raftcorr_lookup_bwd per iteration. */
#define RAFTCORR_LOOKUP_BWD_MAX_BATCH 16
int raftcorr_lookup_bwd_batched_supported(const raftcorr_pyramid_layout* layout, int radius);
int raftcorr_lookup_bwd_batched(const float* const* grad_outs, const float* const* coords, int count, float* dpyramid,
                                const raftcorr_pyramid_layout* layout, int radius, int accumulate, int round_tf32,
                                int device, raftcorr_stream_t stream);
/* raftcorr_build_bwd for a dpyramid whose level 0 already holds the FOLDED gradient (raftcorr_lookup_bwd_batched):
 * only the two products run; the pooled levels of dpyramid are not read. */
/* 1 when raftcorr_build_bwd(_folded) runs its products on the tensor cores for this (D, mode), i.e. when the last
 * raftcorr_lookup_bwd_batched call of a pass should set round_tf32. */
int raftcorr_build_bwd_wants_tf32(int D, int mode);
int raftcorr_build_bwd_folded(float* dpyramid, const raftcorr_pyramid_layout* layout, const float* fmap1,
                              const float* fmap2, int D, float* dfmap1, float* dfmap2, int mode, void* workspace,
                              size_t workspace_bytes, int device, raftcorr_stream_t stream);

/* alt_cuda_corr.forward (correlation.cpp:23-33, correlation_kernel.cu:260-286), same tensors:
 * fmap1 [B,H1,W1,C], fmap2 [B,H2,W2,C] (NHWC), coords [B,N,H1,W1,2] (x,y interleaved),
 * corr [B,N,(2r+1)^2,H1,W1] is OVERWRITTEN (the reference allocates zeros and accumulates). */
int raftcorr_alt_forward(const float* fmap1_nhwc, const float* fmap2_nhwc, const float* coords, int B, int N,
                         int H1, int W1, int H2, int W2, int C, int radius, float* corr, int device,
                         raftcorr_stream_t stream);

/* alt_cuda_corr.backward (correlation.cpp:36-49, correlation_kernel.cu:288-324).
 * fmap1_grad [B,H1,W1,C], fmap2_grad [B,H2,W2,C] and coords_grad [B,N,H1,W1,2] must be
 * zero-initialised by the caller (the reference's launcher does torch::zeros, :305-307) and are
 * accumulated into; coords_grad stays zero as in the reference (:307) and may be NULL. */
int raftcorr_alt_backward(const float* fmap1_nhwc, const float* fmap2_nhwc, const float* coords,
                          const float* corr_grad, int B, int N, int H1, int W1, int H2, int W2, int C,
                          int radius, float* fmap1_grad, float* fmap2_grad, float* coords_grad, int device,
                          raftcorr_stream_t stream);

/* AlternateCorrBlock.__init__ (corr.py:124-140), done once instead of per call (corr.py:158-159):
 * f1_nhwc [B,H,W,D] <- transpose(fmap1); f2_pyramid_nhwc <- levels l<L of avg-pooled fmap2, NHWC,
 * level l at float offset raftcorr_alt_level_offset(B,D,H,W,l), dense [B][H_l][W_l][D]. */
int64_t raftcorr_alt_level_offset(int B, int D, int H, int W, int level);
/* mode RAFTCORR_BUILD_TF32_TC additionally rounds every stored feature to TF32 (round to nearest), which is what
 * the tensor-core lookup below multiplies; RAFTCORR_BUILD_FP32_SIMT stores exact fp32. */
int raftcorr_alt_pyramid(const float* fmap1, const float* fmap2, int B, int D, int H, int W, int L, int mode,
                         float* f1_nhwc, float* f2_pyramid_nhwc, int device, raftcorr_stream_t stream);

/* AlternateCorrBlock.__call__ (corr.py:142-167) fused over all L levels, incl. the 1/sqrt(D).
 * mode RAFTCORR_BUILD_FP32_SIMT: exact fp32 dot products on the CUDA cores, one warp per source pixel (the
 *   arithmetic of correlation_kernel.cu:18-119); no work space.
 * mode RAFTCORR_BUILD_TF32_TC: blocks of 8x16 source pixels against the bounding box of their windows as TF32
 *   tcgen05 GEMM tiles, windows gathered from the accumulators; needs D % 32 == 0, a pyramid built with the same
 *   mode, and raftcorr_alt_lookup_workspace_bytes() bytes of 16-byte aligned device work space. */
size_t raftcorr_alt_lookup_workspace_bytes(int B, int H, int W, int L, int mode);
int raftcorr_alt_lookup_fwd(const float* f1_nhwc, const float* f2_pyramid_nhwc, const float* coords, int B,
                            int D, int H, int W, int L, int radius, int mode, float* out, void* workspace,
                            size_t workspace_bytes, int device, raftcorr_stream_t stream);

/* The tensor-core lookup on fp16 operands (TF32's 11 significant bits; half the operand bytes the kernel is bound by).
 * raftcorr_alt_f16_prepare converts the NHWC tensors of raftcorr_alt_pyramid once per AlternateCorrBlock into a
 * caller-owned buffer of raftcorr_alt_f16_bytes(): every (tensor, image) gets one power-of-two scale (its largest
 * magnitude lands in [2^14, 2^15), so 2^29 of dynamic range inside an image stay at full precision; the lattice tiles
 * of this kernel sit at arbitrary offsets, so there are no per-tile scales as in RAFTCORR_BUILD_F16_TC).
 * raftcorr_alt_lookup_fwd_f16 == raftcorr_alt_lookup_fwd(mode RAFTCORR_BUILD_TF32_TC) on that buffer: same work space,
 * same result to rounding.  D % 64 == 0.  The backward keeps reading the fp32 tensors. */
size_t raftcorr_alt_f16_bytes(int B, int D, int H, int W, int L);
int raftcorr_alt_f16_prepare(const float* f1_nhwc, const float* f2_pyramid_nhwc, int B, int D, int H, int W, int L,
                             void* f16_operands, int device, raftcorr_stream_t stream);
int raftcorr_alt_lookup_fwd_f16(const void* f16_operands, const float* coords, int B, int D, int H, int W, int L,
                                int radius, float* out, void* workspace, size_t workspace_bytes, int device,
                                raftcorr_stream_t stream);

/* Gradient of raftcorr_alt_lookup_fwd w.r.t. the NHWC feature tensors (accumulated into
 * df1_nhwc [B,H,W,D] and df2_pyramid_nhwc, same layout as the forward pyramid; zero them first). */
int raftcorr_alt_lookup_bwd(const float* f1_nhwc, const float* f2_pyramid_nhwc, const float* coords,
                            const float* grad_out, int B, int D, int H, int W, int L, int radius,
                            float* df1_nhwc, float* df2_pyramid_nhwc, int device, raftcorr_stream_t stream);

/* Fold the pooled-level gradients of df2_pyramid_nhwc down to level 0 (adjoint of corr.py:138-139,
 * in place) and transpose both gradients back to NCHW [B,D,H,W] (overwritten). */
int raftcorr_alt_grad_finish(const float* df1_nhwc, float* df2_pyramid_nhwc, int B, int D, int H, int W, int L,
                             float* dfmap1, float* dfmap2, int device, raftcorr_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* RAFTCORR_H_ */

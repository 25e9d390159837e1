// Shared host/device helpers for libraftcorr_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "raftcorr.h"

namespace raftcorr {

// ---- error plumbing (thread-local, no global mutable state shared between callers) ----------
char* error_buffer();  // defined in api.cu
int fail(const char* fmt, ...);

#define RC_CUDA(expr)                                                                          \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess)                                                                 \
            return ::raftcorr::fail("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                                    __LINE__);                                                 \
    } while (0)

#define RC_REQUIRE(cond, ...)                         \
    do {                                              \
        if (!(cond)) return ::raftcorr::fail(__VA_ARGS__); \
    } while (0)

// RAII device guard: callers may sit on any current device (DataParallel threads).
struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int device) {
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
        if (prev != device && cudaSetDevice(device) != cudaSuccess) ok = false;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Per-level view of the pyramid handed to kernels by value.
struct LevelView {
    float* base;          // level start
    int h, w, pitch;      // rows, valid cols, row pitch (floats)
    int il;               // raftcorr_pyramid_layout::interleaved: 0 plain rows, 1 row pairs, 2 row quads interleaved
    int64_t img_stride;   // floats between consecutive source pixels: round_up(h, rows per group) * pitch
    // offset of element (y, x) inside one image
    __host__ __device__ __forceinline__ int64_t at(int y, int x) const {
        return il == 2 ? (int64_t)(y >> 2) * (4 * pitch) + 4 * x + (y & 3)
             : il ? (int64_t)(y >> 1) * (2 * pitch) + 2 * x + (y & 1) : (int64_t)y * pitch + x;
    }
    __host__ __device__ __forceinline__ int group_rows() const { return il == 2 ? 4 : il ? 2 : 1; }
};

// Per-level interleave code of a layout (one code for all levels today; kept as the single place that decides).
// Layout 2 = row quads on every level: the build's 8x32 target tile holds a whole quad of level 1 but only 2 rows of
// level 2 and 1 row of level 3, so its epilogue carries those rows across the 4 bands of a work unit and stores whole
// quads (partial-sector stores cost as much as full ones on the SM's store path).
__host__ __device__ __forceinline__ int level_interleave(int interleaved, int level) {
    (void)level;
    return interleaved;
}
// rows interleaved per group for a per-level code (0 / 1 / 2 -> 1 / 2 / 4)
__host__ __device__ __forceinline__ int layout_group_rows(int il) { return il == 2 ? 4 : il ? 2 : 1; }

struct PyramidView {
    LevelView lvl[RAFTCORR_MAX_LEVELS];
    int B, N, L;  // N = H*W source pixels
};

inline PyramidView make_view(float* base, const raftcorr_pyramid_layout& lay) {
    PyramidView v{};
    v.B = lay.B;
    v.N = lay.H * lay.W;
    v.L = lay.L;
    for (int l = 0; l < lay.L; ++l) {
        v.lvl[l].base = base + lay.offset[l];
        v.lvl[l].h = lay.h[l];
        v.lvl[l].w = lay.w[l];
        v.lvl[l].pitch = lay.pitch[l];
        v.lvl[l].il = level_interleave(lay.interleaved, l);
        const int rg = layout_group_rows(v.lvl[l].il);
        v.lvl[l].img_stride = (int64_t)((lay.h[l] + rg - 1) / rg * rg) * lay.pitch[l];
    }
    return v;
}

// ---- launchers implemented in the individual .cu files ------------------------------------------
int launch_nchw_to_nhwc(const float* in, float* out, int B, int D, int HW, int Dout, bool round_tf32, cudaStream_t s);
int launch_nchw_to_nhwc_pair(const float* in1, float* out1, const float* in2, float* out2, int B, int D, int HW,
                             int Dout, bool round, cudaStream_t s);
int tc_feature_pitch(int D);  // K padding of the tensor-core build (multiple of 32 floats)
int launch_nhwc_to_nchw(const float* in, float* out, int B, int D, int HW, cudaStream_t s);
int launch_pool_nhwc(const float* in, float* out, int B, int Hi, int Wi, int D, bool round_tf32, cudaStream_t s);
int launch_pool_adjoint_nhwc(const float* gout, float* gin, int B, int Hi, int Wi, int D, cudaStream_t s);
int launch_pool_level(const LevelView& in, const LevelView& out, int64_t imgs, cudaStream_t s);
int launch_pool_adjoint_level(const LevelView& coarse, const LevelView& fine, int64_t imgs, bool round_tf32_out,
                              cudaStream_t s);
int launch_round_level(const LevelView& lv, int64_t imgs, cudaStream_t s);
bool build_bwd_tc_supported(int D);
size_t build_bwd_tc_workspace_bytes(int B, int D, int H, int W);
int launch_build_bwd_tc(const PyramidView& dpv, const float* f1, const float* f2, int D, float* df1, float* df2,
                        void* workspace, int device, cudaStream_t s);

int launch_build_simt(const float* f1, const float* f2, int D, const PyramidView& pv, cudaStream_t s);
int launch_build_tc(const float* f1_nhwc, const float* f2_nhwc, int D, float* pyramid,
                    const raftcorr_pyramid_layout& lay, int device, cudaStream_t s);
int tc_feature_pitch_f16(int D);  // K padding of the fp16 build (multiple of 64 halves)
int launch_build_tc_ex(const void* f1_ops, const void* f2_ops, int D, float* pyramid, const raftcorr_pyramid_layout& lay,
                       int device, cudaStream_t s, bool f16, const float* inv_a, const float* inv_b, bool split = false);
// NCHW fp32 maps -> block-scaled fp16 operand buffers [B][H*W][Dp] + inverse block scales (prep.cu)
size_t f16_scale_floats(int B, int H, int W);   // floats in inv_a followed by inv_b
int launch_f16_operands(const float* f1, const float* f2, int B, int D, int H, int W, void* f1h, void* f2h, float* inv_a,
                        float* inv_b, cudaStream_t s, bool split = false);
int launch_f16_operands_from_nhwc(const float* f1n, const float* f2n, int B, int D, int Dp32, int H, int W, void* f1h,
                                  void* f2h, float* inv_a, float* inv_b, cudaStream_t s);
int launch_build_bwd_gemms(const PyramidView& dpv, const float* f1, const float* f2, int D, float* df1,
                           float* df2, cudaStream_t s);

int launch_lookup_fwd(const PyramidView& pv, const float* coords, int radius, float* out, cudaStream_t s);
int launch_lookup_bwd(const float* grad_out, const float* coords, const PyramidView* pv_fwd,
                      const PyramidView& dpv, int radius, float* dcoords, cudaStream_t s);
// all lookups of a backward pass at once, pooled levels folded on chip (lookup_bwd_batched.cu); writes level 0 of dpv
bool lookup_bwd_batched_supported(const PyramidView& dpv, int radius);
int launch_lookup_bwd_batched(const float* const* gouts, const float* const* coords, int count, const PyramidView& dpv,
                              int radius, bool accumulate, bool round_out, int device, cudaStream_t s);

// lookup fused with the motion encoder's first 1x1 convolution (lookup_conv.cu)
size_t conv_weight_bytes(int levels, int radius, int cout);
int launch_pack_conv_weight(const float* weight, int levels, int radius, int cout, float* packed, cudaStream_t s);
int launch_lookup_conv(const void* plan_blob, const float* coords, const float* packed_w, const float* bias, int cout,
                       int relu, float* out, int device, cudaStream_t s);

// encoder output head -> build operands (fnet_head.cu)
bool fnet_head_supported(int Cin, int D, int HW);
size_t fnet_head_weight_bytes(int Cin, int D);
int launch_fnet_head(const float* x, const float* weight, const float* bias, int B, int Cin, int D, int HW, float* f1n,
                     float* f2n, float* w_rounded, int device, cudaStream_t s);

// convex 8x flow upsampling (upsample.cu)
int launch_upsample_fwd(const float* flow, const float* mask, int B, int C, int H, int W, float* out, cudaStream_t s);
int launch_upsample_bwd(const float* flow, const float* mask, const float* gout, int B, int C, int H, int W, float* dmask,
                        float* dflow, cudaStream_t s);

struct AltLevels {
    const float* f2[RAFTCORR_MAX_LEVELS];  // NHWC [B][h][w][C]
    float* df2[RAFTCORR_MAX_LEVELS];       // gradients (bwd only)
    int h[RAFTCORR_MAX_LEVELS], w[RAFTCORR_MAX_LEVELS];
    int L;
};
// coords addressing: value(b, c, p) = coords[b * batch_stride + c * comp_stride + p * pix_stride]
struct CoordsView {
    const float* ptr;
    int64_t batch_stride, comp_stride, pix_stride;
};
int launch_alt_fwd(const float* f1_nhwc, const AltLevels& lv, const CoordsView& cv, int B, int NC, int H1, int W1,
                   int C, int radius, float scale, float* out, cudaStream_t s);
size_t alt_tc_workspace_bytes(int B, int H, int W, int L);
int launch_alt_fwd_tc(const float* f1_nhwc, const AltLevels& lv, const CoordsView& cv, int B, int H, int W, int C,
                      int radius, float scale, float* out, void* workspace, size_t workspace_bytes, int device,
                      cudaStream_t s);
int launch_alt_fwd_tc_ex(const void* f1_nhwc, const AltLevels& lv, const CoordsView& cv, int B, int H, int W, int C,
                         int radius, float scale, float* out, void* workspace, size_t workspace_bytes, int device,
                         cudaStream_t s, bool f16, const float* inv);
// fp32 NHWC tensors -> fp16 copies scaled by one power of two per image and tensor (prep.cu).  `count` tensors of
// per_img[t] elements per image; src/dst element offsets off[t] (same numbers for both buffers);
// inv [count][B] floats, maxbits [count][B] scratch.
int launch_alt_f16_operands(const float* src_f1, const float* src_f2, void* dst_f1, void* dst_f2, const int64_t* off,
                            const int64_t* per_img, int count, int B, float* inv, unsigned* maxbits, cudaStream_t s);
int launch_alt_bwd(const float* f1_nhwc, const AltLevels& lv, const CoordsView& cv, const float* grad_out, int B,
                   int NC, int H1, int W1, int C, int radius, float scale, float* df1_nhwc, cudaStream_t s);

}  // namespace raftcorr

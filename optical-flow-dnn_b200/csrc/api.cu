// extern "C" surface of libraftcorr_b200.so -- see include/raftcorr.h for the contract and the
// reference interfaces each entry point replaces.
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "lookup_common.cuh"

namespace raftcorr {

char* error_buffer() {
    static thread_local char buf[512] = {0};
    return buf;
}

int fail(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(error_buffer(), 512, fmt, ap);
    va_end(ap);
    return 1;
}

namespace {

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int check_layout(const raftcorr_pyramid_layout* lay) {
    RC_REQUIRE(lay != nullptr, "layout is NULL");
    RC_REQUIRE(lay->L >= 1 && lay->L <= RAFTCORR_MAX_LEVELS, "layout: bad level count %d", lay->L);
    RC_REQUIRE(lay->B >= 0 && lay->H >= 1 && lay->W >= 1, "layout: bad sizes B=%d H=%d W=%d", lay->B, lay->H, lay->W);
    for (int l = 0; l < lay->L; ++l)
        RC_REQUIRE(lay->pitch[l] >= lay->w[l] && lay->h[l] >= 1 && lay->w[l] >= 1, "layout: bad level %d", l);
    return 0;
}

int alt_levels(const float* f2pyr, float* df2pyr, int B, int D, int H, int W, int L, AltLevels* out) {
    RC_REQUIRE(L >= 1 && L <= RAFTCORR_MAX_LEVELS, "alt: bad level count %d", L);
    out->L = L;
    int h = H, w = W;
    int64_t off = 0;
    for (int l = 0; l < L; ++l) {
        RC_REQUIRE(h >= 1 && w >= 1, "alt: level %d is empty (H=%d, W=%d)", l, H, W);
        out->f2[l] = f2pyr ? f2pyr + off : nullptr;
        out->df2[l] = df2pyr ? df2pyr + off : nullptr;
        out->h[l] = h;
        out->w[l] = w;
        off += (int64_t)align_up((size_t)B * h * w * D, 64);
        h /= 2;
        w /= 2;
    }
    return 0;
}

}  // namespace
}  // namespace raftcorr

using namespace raftcorr;

extern "C" {

int raftcorr_abi_version(void) { return RAFTCORR_ABI_VERSION; }

const char* raftcorr_last_error(void) { return error_buffer(); }

int raftcorr_set_l2_fetch_granularity(int bytes, int device, int* previous) {
    RC_REQUIRE(bytes == 32 || bytes == 64 || bytes == 128, "L2 fetch granularity must be 32, 64 or 128 (got %d)", bytes);
    DeviceGuard guard(device);
    RC_REQUIRE(guard.ok, "cannot select device %d", device);
    size_t prev = 0;
    RC_CUDA(cudaDeviceGetLimit(&prev, cudaLimitMaxL2FetchGranularity));
    if (previous) *previous = (int)prev;
    RC_CUDA(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)bytes));
    return 0;
}

int raftcorr_pyramid_layout_init(raftcorr_pyramid_layout* lay, int B, int H, int W, int L, int interleaved) {
    RC_REQUIRE(lay != nullptr, "layout is NULL");
    RC_REQUIRE(L >= 1 && L <= RAFTCORR_MAX_LEVELS, "num_levels must be in [1, %d] (got %d)", RAFTCORR_MAX_LEVELS, L);
    RC_REQUIRE(B >= 0 && H >= 1 && W >= 1, "bad sizes B=%d H=%d W=%d", B, H, W);
    std::memset(lay, 0, sizeof(*lay));
    lay->B = B; lay->H = H; lay->W = W; lay->L = L;
    RC_REQUIRE(interleaved >= 0 && interleaved <= 2, "pyramid layout: interleaved must be 0 (rows), 1 (row pairs) or 2 (row quads), got %d", interleaved);
    lay->interleaved = interleaved;
    int h = H, w = W;
    int64_t off = 0;
    const int64_t N = (int64_t)H * W;
    for (int l = 0; l < L; ++l) {
        // the reference normalises by (W_l - 1) and (H_l - 1) (utils.py:74-75): a 1-pixel level is NaN there
        RC_REQUIRE(h >= 2 && w >= 2, "pyramid level %d would be %dx%d; the reference divides by (size-1) there "
                   "(raft/core/utils/utils.py:74-75). Use fewer levels or larger feature maps", l, h, w);
        lay->h[l] = h;
        lay->w[l] = w;
        lay->pitch[l] = (w + 7) / 8 * 8;
        lay->offset[l] = off;
        const int rg = layout_group_rows(level_interleave(interleaved, l));
        const int64_t rows = (h + rg - 1) / rg * rg;  // the last group is completed with rows of zeros
        off += (int64_t)align_up((size_t)(B * N * rows * lay->pitch[l]), 64);
        h /= 2;
        w /= 2;
    }
    lay->total_floats = off;
    return 0;
}

size_t raftcorr_build_fwd_workspace_bytes(int B, int D, int H, int W, int mode) {
    if (mode == RAFTCORR_BUILD_TF32_TC)  // two NHWC (K-major), TF32-rounded, K-padded copies of the feature maps
        return 2 * align_up((size_t)B * tc_feature_pitch(D) * H * W * sizeof(float), 1024);
    if (mode == RAFTCORR_BUILD_F16_TC || mode == RAFTCORR_BUILD_F16X3_TC) {  // two block-scaled fp16 copies + inverse scales
        const size_t terms = mode == RAFTCORR_BUILD_F16X3_TC ? 3 : 1;
        return 2 * align_up((size_t)B * terms * tc_feature_pitch_f16(D) * H * W * 2, 1024) +
               align_up(f16_scale_floats(B, H, W) * sizeof(float), 1024);
    }
    return 0;
}

int raftcorr_build_fwd(const float* fmap1, const float* fmap2, int D, float* pyramid,
                       const raftcorr_pyramid_layout* lay, int mode, void* workspace, size_t workspace_bytes,
                       int device, raftcorr_stream_t stream) {
    if (int rc = check_layout(lay)) return rc;
    if (lay->B == 0) return 0;  // empty batch: nothing to do (and no storage behind the pointers)
    RC_REQUIRE(fmap1 && fmap2 && pyramid, "build_fwd: NULL tensor");
    RC_REQUIRE(D >= 1, "build_fwd: bad feature dim %d", D);
    DeviceGuard guard(device);
    RC_REQUIRE(guard.ok, "build_fwd: cannot select device %d", device);
    cudaStream_t s = (cudaStream_t)stream;
    PyramidView pv = make_view(pyramid, *lay);
    if (mode == RAFTCORR_BUILD_FP32_SIMT) {
        RC_REQUIRE(!lay->interleaved, "build_fwd: the fp32 build writes the row-major pyramid layout only");
        return launch_build_simt(fmap1, fmap2, D, pv, s);
    }
    if (mode == RAFTCORR_BUILD_TF32_TC) {
        const size_t need = raftcorr_build_fwd_workspace_bytes(lay->B, D, lay->H, lay->W, mode);
        RC_REQUIRE(workspace && workspace_bytes >= need, "build_fwd: workspace too small (%zu < %zu)", workspace_bytes, need);
        float* f1n = (float*)workspace;
        float* f2n = (float*)((char*)workspace + need / 2);
        const int HW = lay->H * lay->W;
        const int Dp = tc_feature_pitch(D);
        if (int rc = launch_nchw_to_nhwc_pair(fmap1, f1n, fmap2, f2n, lay->B, D, HW, Dp, true, s)) return rc;
        return launch_build_tc(f1n, f2n, D, pyramid, *lay, device, s);
    }
    if (mode == RAFTCORR_BUILD_F16_TC || mode == RAFTCORR_BUILD_F16X3_TC) {
        const bool split = mode == RAFTCORR_BUILD_F16X3_TC;
        const size_t need = raftcorr_build_fwd_workspace_bytes(lay->B, D, lay->H, lay->W, mode);
        RC_REQUIRE(workspace && workspace_bytes >= need, "build_fwd: workspace too small (%zu < %zu)", workspace_bytes, need);
        const size_t ops = align_up((size_t)lay->B * (split ? 3 : 1) * tc_feature_pitch_f16(D) * lay->H * lay->W * 2, 1024);
        void* f1h = workspace;
        void* f2h = (char*)workspace + ops;
        float* inv_a = (float*)((char*)workspace + 2 * ops);
        float* inv_b = inv_a + (size_t)lay->B * 2 * ((lay->H * lay->W + 127) / 128);
        if (int rc = launch_f16_operands(fmap1, fmap2, lay->B, D, lay->H, lay->W, f1h, f2h, inv_a, inv_b, s, split)) return rc;
        return launch_build_tc_ex(f1h, f2h, D, pyramid, *lay, device, s, true, inv_a, inv_b, split);
    }
    return fail("build_fwd: unknown mode %d", mode);
}

int raftcorr_fnet_head_supported(int Cin, int D, int H, int W) {
    return (H >= 1 && W >= 1 && fnet_head_supported(Cin, D, H * W)) ? 1 : 0;
}

size_t raftcorr_fnet_head_build_workspace_bytes(int B, int Cin, int D, int H, int W) {
    if (B < 0 || Cin < 1 || D < 1 || H < 1 || W < 1) return 0;
    // [TF32 operands the head writes][rounded weight][fp16 block-scaled operands + scales for the build (D <= 256)]
    return raftcorr_build_fwd_workspace_bytes(B, D, H, W, RAFTCORR_BUILD_TF32_TC) + fnet_head_weight_bytes(Cin, D) +
           (D <= 256 ? raftcorr_build_fwd_workspace_bytes(B, D, H, W, RAFTCORR_BUILD_F16_TC) : 0);
}

int raftcorr_fnet_head_build_fwd(const float* x, const float* weight, const float* bias, int Cin, int D, float* pyramid,
                                 const raftcorr_pyramid_layout* lay, void* workspace, size_t workspace_bytes, int device,
                                 raftcorr_stream_t stream) {
    if (int rc = check_layout(lay)) return rc;
    if (lay->B == 0) return 0;
    RC_REQUIRE(x && weight && pyramid, "fnet_head_build_fwd: NULL tensor");
    RC_REQUIRE(Cin >= 1 && D >= 1, "fnet_head_build_fwd: bad sizes Cin=%d D=%d", Cin, D);
    const size_t ops = raftcorr_build_fwd_workspace_bytes(lay->B, D, lay->H, lay->W, RAFTCORR_BUILD_TF32_TC);
    const size_t need = raftcorr_fnet_head_build_workspace_bytes(lay->B, Cin, D, lay->H, lay->W);
    RC_REQUIRE(workspace && workspace_bytes >= need, "fnet_head_build_fwd: workspace too small (%zu < %zu)", workspace_bytes,
               need);
    DeviceGuard guard(device);
    RC_REQUIRE(guard.ok, "fnet_head_build_fwd: cannot select device %d", device);
    cudaStream_t s = (cudaStream_t)stream;
    float* f1n = (float*)workspace;
    float* f2n = (float*)((char*)workspace + ops / 2);
    float* wr = (float*)((char*)workspace + ops);
    if (int rc = launch_fnet_head(x, weight, bias, lay->B, Cin, D, lay->H * lay->W, f1n, f2n, wr, device, s)) return rc;
    // Measured (scripts/fnet_head_time.py): the extra conversion pays from about 64x64 feature maps on (55x128: 504 vs
    // 556 us, 47x156 B=16: 1118 vs 1268 us) and costs ~5 us at 46x62.  RAFTCORR_HEAD_F16=0 / 1 force the choice.
    const char* fe = std::getenv("RAFTCORR_HEAD_F16");
    const bool want_f16 = fe ? fe[0] != '0' : lay->H * lay->W >= 4096;
    if (D > 256 || !want_f16) return launch_build_tc(f1n, f2n, D, pyramid, *lay, device, s);
    // the build runs on block-scaled fp16 operands (RAFTCORR_BUILD_F16_TC): convert the head's operands (still in L2)
    char* h16 = (char*)workspace + ops + fnet_head_weight_bytes(Cin, D);
    const size_t hops = align_up((size_t)lay->B * tc_feature_pitch_f16(D) * lay->H * lay->W * 2, 1024);
    float* inv_a = (float*)(h16 + 2 * hops);
    float* inv_b = inv_a + (size_t)lay->B * 2 * ((lay->H * lay->W + 127) / 128);
    if (int rc = launch_f16_operands_from_nhwc(f1n, f2n, lay->B, D, tc_feature_pitch(D), lay->H, lay->W, h16, h16 + hops,
                                               inv_a, inv_b, s))
        return rc;
    return launch_build_tc_ex(h16, h16 + hops, D, pyramid, *lay, device, s, true, inv_a, inv_b);
}

size_t raftcorr_fnet_head_alt_workspace_bytes(int Cin, int D) {
    if (Cin < 1 || D < 1) return 0;
    return fnet_head_weight_bytes(Cin, D);
}

int raftcorr_fnet_head_alt_pyramid(const float* x, const float* weight, const float* bias, int B, int Cin, int D, int H,
                                   int W, int L, float* f1_nhwc, float* f2_pyr, void* workspace, size_t workspace_bytes,
                                   int device, raftcorr_stream_t stream) {
    RC_REQUIRE(B == 0 || (x && weight && f1_nhwc && f2_pyr), "fnet_head_alt_pyramid: NULL tensor");
    RC_REQUIRE(D % 32 == 0, "fnet_head_alt_pyramid: the TF32 on-the-fly path needs D %% 32 == 0 (got %d)", D);
    AltLevels lv{};
    if (int rc = alt_levels(f2_pyr, nullptr, B, D, H, W, L, &lv)) return rc;
    if (B == 0) return 0;
    const size_t need = fnet_head_weight_bytes(Cin, D);
    RC_REQUIRE(workspace && workspace_bytes >= need, "fnet_head_alt_pyramid: workspace too small (%zu < %zu)", workspace_bytes,
               need);
    DeviceGuard guard(device);
    RC_REQUIRE(guard.ok, "fnet_head_alt_pyramid: cannot select device %d", device);
    cudaStream_t s = (cudaStream_t)stream;
    // frame-1 maps -> f1_nhwc, frame-2 maps -> level 0 of the feature pyramid, both NHWC and TF32-rounded (D == Dp)
    if (int rc = launch_fnet_head(x, weight, bias, B, Cin, D, H * W, f1_nhwc, const_cast<float*>(lv.f2[0]), (float*)workspace,
                                  device, s))
        return rc;
    for (int l = 1; l < L; ++l)  // corr.py:138-139
        if (int rc = launch_pool_nhwc(lv.f2[l - 1], const_cast<float*>(lv.f2[l]), B, lv.h[l - 1], lv.w[l - 1], D, true, s))
            return rc;
    return 0;
}

int raftcorr_lookup_fwd(const float* pyramid, const raftcorr_pyramid_layout* lay, const float* coords, int radius,
                        float* out, int device, raftcorr_stream_t stream) {
    if (int rc = check_layout(lay)) return rc;
    if (lay->B == 0) return 0;
    RC_REQUIRE(pyramid && coords && out, "lookup_fwd: NULL tensor");
    DeviceGuard guard(device);
    RC_REQUIRE(guard.ok, "lookup_fwd: cannot select device %d", device);
    PyramidView pv = make_view(const_cast<float*>(pyramid), *lay);
    return launch_lookup_fwd(pv, coords, radius, out, (cudaStream_t)stream);
}

int raftcorr_lookup_plan_init(raftcorr_lookup_plan* plan, const float* pyramid, const raftcorr_pyramid_layout* lay,
                              int radius, int device) {
    RC_REQUIRE(plan != nullptr, "lookup_plan_init: plan is NULL");
    if (int rc = check_layout(lay)) return rc;
    std::memset(plan, 0, sizeof(*plan));
    if (lay->B == 0) return 0;
    RC_REQUIRE(pyramid, "lookup_plan_init: NULL pyramid");
    DeviceGuard guard(device);
    RC_REQUIRE(guard.ok, "lookup_plan_init: cannot select device %d", device);
    PyramidView pv = make_view(const_cast<float*>(pyramid), *lay);
    return lookup_plan_init(plan->opaque, pv, radius);
}

int raftcorr_lookup_fwd_planned(const raftcorr_lookup_plan* plan, const float* coords, float* out, int device,
                                raftcorr_stream_t stream) {
    RC_REQUIRE(plan != nullptr, "lookup_fwd_planned: plan is NULL");
    bool empty = true;
    for (int i = 0; i < 16 && empty; ++i) empty = plan->opaque[i] == 0;
    if (empty) return 0;  // plan of an empty batch
    RC_REQUIRE(coords && out, "lookup_fwd_planned: NULL tensor");
    DeviceGuard guard(device);
    RC_REQUIRE(guard.ok, "lookup_fwd_planned: cannot select device %d", device);
    return lookup_plan_run(plan->opaque, coords, out, (cudaStream_t)stream);
}

size_t raftcorr_convc1_weight_bytes(int levels, int radius, int cout) {
    if (levels < 1 || levels > 4 || radius < 1 || radius > 4 || cout < 1) return 0;
    return conv_weight_bytes(levels, radius, cout);
}

int raftcorr_convc1_pack_weight(const float* weight, int levels, int radius, int cout, void* packed, int device,
                                raftcorr_stream_t stream) {
    RC_REQUIRE(weight && packed, "convc1_pack_weight: NULL tensor");
    RC_REQUIRE(levels >= 1 && levels <= 4 && radius >= 1 && radius <= 4 && cout >= 1, "convc1_pack_weight: bad sizes");
    DeviceGuard guard(device);
    RC_REQUIRE(guard.ok, "convc1_pack_weight: cannot select device %d", device);
    return launch_pack_conv_weight(weight, levels, radius, cout, (float*)packed, (cudaStream_t)stream);
}

int raftcorr_lookup_convc1_fwd(const raftcorr_lookup_plan* plan, const float* coords, const void* packed_weight,
                               const float* bias, int cout, int relu, float* out, int device,
                               raftcorr_stream_t stream) {
    RC_REQUIRE(plan != nullptr, "lookup_convc1_fwd: plan is NULL");
    bool empty = true;
    for (int i = 0; i < 16 && empty; ++i) empty = plan->opaque[i] == 0;
    if (empty) return 0;  // plan of an empty batch
    RC_REQUIRE(coords && packed_weight && out, "lookup_convc1_fwd: NULL tensor");
    DeviceGuard guard(device);
    RC_REQUIRE(guard.ok, "lookup_convc1_fwd: cannot select device %d", device);
    return launch_lookup_conv(plan->opaque, coords, (const float*)packed_weight, bias, cout, relu, out, device,
                              (cudaStream_t)stream);
}

int raftcorr_upsample_flow_fwd(const float* flow, const float* mask, int B, int C, int H, int W, float* out, int device,
                               raftcorr_stream_t stream) {
    RC_REQUIRE(B >= 0 && H >= 1 && W >= 1, "upsample_flow: bad sizes B=%d H=%d W=%d", B, H, W);
    if (B == 0) return 0;
    RC_REQUIRE(flow && mask && out, "upsample_flow: NULL tensor");
    DeviceGuard guard(device);
    RC_REQUIRE(guard.ok, "upsample_flow: cannot select device %d", device);
    return launch_upsample_fwd(flow, mask, B, C, H, W, out, (cudaStream_t)stream);
}

int raftcorr_upsample_flow_bwd(const float* flow, const float* mask, const float* grad_out, int B, int C, int H, int W,
                               float* dmask, float* dflow, int device, raftcorr_stream_t stream) {
    RC_REQUIRE(B >= 0 && H >= 1 && W >= 1, "upsample_flow_bwd: bad sizes B=%d H=%d W=%d", B, H, W);
    if (B == 0) return 0;
    RC_REQUIRE(flow && mask && grad_out && dmask && dflow, "upsample_flow_bwd: NULL tensor");
    DeviceGuard guard(device);
    RC_REQUIRE(guard.ok, "upsample_flow_bwd: cannot select device %d", device);
    cudaStream_t s = (cudaStream_t)stream;
    RC_CUDA(cudaMemsetAsync(dflow, 0, sizeof(float) * (size_t)B * C * H * W, s));
    return launch_upsample_bwd(flow, mask, grad_out, B, C, H, W, dmask, dflow, s);
}

int raftcorr_lookup_bwd(const float* grad_out, const float* coords, const float* pyramid, float* dpyramid,
                        const raftcorr_pyramid_layout* lay, int radius, float* dcoords, int device,
                        raftcorr_stream_t stream) {
    if (int rc = check_layout(lay)) return rc;
    if (lay->B == 0) return 0;
    RC_REQUIRE(grad_out && coords && dpyramid, "lookup_bwd: NULL tensor");
    RC_REQUIRE(!dcoords || pyramid, "lookup_bwd: dcoords requested without the forward pyramid");
    DeviceGuard guard(device);
    RC_REQUIRE(guard.ok, "lookup_bwd: cannot select device %d", device);
    cudaStream_t s = (cudaStream_t)stream;
    // gradient pyramids are always row-major (they feed the backward GEMMs); `lay` describes the forward pyramid
    raftcorr_pyramid_layout glay;
    if (int rc = raftcorr_pyramid_layout_init(&glay, lay->B, lay->H, lay->W, lay->L, 0)) return rc;
    PyramidView dpv = make_view(dpyramid, glay);
    PyramidView fpv{};
    if (dcoords) {
        fpv = make_view(const_cast<float*>(pyramid), *lay);
        RC_CUDA(cudaMemsetAsync(dcoords, 0, sizeof(float) * 2 * (size_t)lay->B * lay->H * lay->W, s));
    }
    return launch_lookup_bwd(grad_out, coords, dcoords ? &fpv : nullptr, dpv, radius, dcoords, s);
}

int raftcorr_lookup_bwd_batched_supported(const raftcorr_pyramid_layout* lay, int radius) {
    if (check_layout(lay)) return 0;
    raftcorr_pyramid_layout glay;
    if (raftcorr_pyramid_layout_init(&glay, lay->B, lay->H, lay->W, lay->L, 0)) return 0;
    PyramidView dpv = make_view(nullptr, glay);
    return lookup_bwd_batched_supported(dpv, radius) ? 1 : 0;
}

int raftcorr_lookup_bwd_batched(const float* const* grad_outs, const float* const* coords, int count, float* dpyramid,
                                const raftcorr_pyramid_layout* lay, int radius, int accumulate, int round_tf32,
                                int device, raftcorr_stream_t stream) {
    if (int rc = check_layout(lay)) return rc;
    if (lay->B == 0) return 0;
    RC_REQUIRE(grad_outs && coords && dpyramid, "lookup_bwd_batched: NULL argument");
    DeviceGuard guard(device);
    RC_REQUIRE(guard.ok, "lookup_bwd_batched: cannot select device %d", device);
    raftcorr_pyramid_layout glay;  // gradient pyramids are always row-major
    if (int rc = raftcorr_pyramid_layout_init(&glay, lay->B, lay->H, lay->W, lay->L, 0)) return rc;
    PyramidView dpv = make_view(dpyramid, glay);
    return launch_lookup_bwd_batched(grad_outs, coords, count, dpv, radius, accumulate != 0, round_tf32 != 0, device,
                                     (cudaStream_t)stream);
}

int raftcorr_build_bwd_wants_tf32(int D, int mode) {
    return (mode != RAFTCORR_BUILD_FP32_SIMT && build_bwd_tc_supported(D)) ? 1 : 0;
}

size_t raftcorr_build_bwd_workspace_bytes(int B, int D, int H, int W, int mode) {
    if (mode != RAFTCORR_BUILD_FP32_SIMT && build_bwd_tc_supported(D))
        return build_bwd_tc_workspace_bytes(B, D, H, W);
    return 0;
}

static int build_bwd_impl(float* dpyramid, const raftcorr_pyramid_layout* lay, const float* fmap1, const float* fmap2,
                          int D, float* dfmap1, float* dfmap2, int mode, void* workspace, size_t workspace_bytes,
                          int device, raftcorr_stream_t stream, bool folded) {
    if (int rc = check_layout(lay)) return rc;
    if (lay->B == 0) return 0;
    RC_REQUIRE(dpyramid && fmap1 && fmap2 && dfmap1 && dfmap2, "build_bwd: NULL tensor");
    RC_REQUIRE(mode >= RAFTCORR_BUILD_FP32_SIMT && mode <= RAFTCORR_BUILD_F16X3_TC, "build_bwd: unknown mode %d", mode);
    DeviceGuard guard(device);
    RC_REQUIRE(guard.ok, "build_bwd: cannot select device %d", device);
    cudaStream_t s = (cudaStream_t)stream;
    raftcorr_pyramid_layout glay;  // gradient pyramids are always row-major, whatever the forward layout is
    if (int rc = raftcorr_pyramid_layout_init(&glay, lay->B, lay->H, lay->W, lay->L, 0)) return rc;
    PyramidView dpv = make_view(dpyramid, glay);
    const int64_t imgs = (int64_t)lay->B * lay->H * lay->W;
    // the fp16 forward mode shares the TF32 backward (gradient GEMMs read the fp32 maps); D > 256: CUDA-core GEMMs
    const bool tc = mode != RAFTCORR_BUILD_FP32_SIMT && build_bwd_tc_supported(D);
    if (!folded) {
        for (int l = lay->L - 1; l >= 1; --l)  // adjoint of corr.py:41-43, coarse to fine
            if (int rc = launch_pool_adjoint_level(dpv.lvl[l], dpv.lvl[l - 1], imgs, tc && l == 1, s)) return rc;
    }
    if (!tc) return launch_build_bwd_gemms(dpv, fmap1, fmap2, D, dfmap1, dfmap2, s);
    if (lay->L == 1 && !folded)
        if (int rc = launch_round_level(dpv.lvl[0], imgs, s)) return rc;
    const size_t need = build_bwd_tc_workspace_bytes(lay->B, D, lay->H, lay->W);
    RC_REQUIRE(workspace && workspace_bytes >= need, "build_bwd: workspace too small (%zu < %zu)", workspace_bytes, need);
    return launch_build_bwd_tc(dpv, fmap1, fmap2, D, dfmap1, dfmap2, workspace, device, s);
}

int raftcorr_build_bwd(float* dpyramid, const raftcorr_pyramid_layout* lay, const float* fmap1, const float* fmap2,
                       int D, float* dfmap1, float* dfmap2, int mode, void* workspace, size_t workspace_bytes,
                       int device, raftcorr_stream_t stream) {
    return build_bwd_impl(dpyramid, lay, fmap1, fmap2, D, dfmap1, dfmap2, mode, workspace, workspace_bytes, device,
                          stream, false);
}

int raftcorr_build_bwd_folded(float* dpyramid, const raftcorr_pyramid_layout* lay, const float* fmap1,
                              const float* fmap2, int D, float* dfmap1, float* dfmap2, int mode, void* workspace,
                              size_t workspace_bytes, int device, raftcorr_stream_t stream) {
    return build_bwd_impl(dpyramid, lay, fmap1, fmap2, D, dfmap1, dfmap2, mode, workspace, workspace_bytes, device,
                          stream, true);
}

int raftcorr_alt_forward(const float* fmap1, const float* fmap2, const float* coords, int B, int N, int H1, int W1,
                         int H2, int W2, int C, int radius, float* corr, int device, raftcorr_stream_t stream) {
    RC_REQUIRE(B == 0 || (fmap1 && fmap2 && coords && corr), "alt_forward: NULL tensor");
    RC_REQUIRE(B >= 0 && N >= 1 && H1 >= 1 && W1 >= 1 && H2 >= 1 && W2 >= 1, "alt_forward: bad sizes");
    if (B == 0) return 0;
    DeviceGuard guard(device);
    RC_REQUIRE(guard.ok, "alt_forward: cannot select device %d", device);
    AltLevels lv{};
    lv.L = 1; lv.f2[0] = fmap2; lv.h[0] = H2; lv.w[0] = W2;
    const int64_t N1 = (int64_t)H1 * W1;
    CoordsView cv{coords, 2 * N1, 1, 2};  // [B,N,H1,W1,2]
    return launch_alt_fwd(fmap1, lv, cv, B, N, H1, W1, C, radius, 1.0f, corr, (cudaStream_t)stream);
}

int raftcorr_alt_backward(const float* fmap1, const float* fmap2, const float* coords, const float* corr_grad, int B,
                          int N, int H1, int W1, int H2, int W2, int C, int radius, float* fmap1_grad,
                          float* fmap2_grad, float* coords_grad, int device, raftcorr_stream_t stream) {
    (void)coords_grad;  // identically zero in the reference (correlation_kernel.cu:307); caller zero-fills
    RC_REQUIRE(B == 0 || (fmap1 && fmap2 && coords && corr_grad && fmap1_grad && fmap2_grad), "alt_backward: NULL tensor");
    RC_REQUIRE(B >= 0 && N >= 1 && H1 >= 1 && W1 >= 1 && H2 >= 1 && W2 >= 1, "alt_backward: bad sizes");
    if (B == 0) return 0;
    DeviceGuard guard(device);
    RC_REQUIRE(guard.ok, "alt_backward: cannot select device %d", device);
    AltLevels lv{};
    lv.L = 1; lv.f2[0] = fmap2; lv.df2[0] = fmap2_grad; lv.h[0] = H2; lv.w[0] = W2;
    const int64_t N1 = (int64_t)H1 * W1;
    CoordsView cv{coords, 2 * N1, 1, 2};
    return launch_alt_bwd(fmap1, lv, cv, corr_grad, B, N, H1, W1, C, radius, 1.0f, fmap1_grad, (cudaStream_t)stream);
}

int64_t raftcorr_alt_level_offset(int B, int D, int H, int W, int level) {
    int h = H, w = W;
    int64_t off = 0;
    for (int l = 0; l < level; ++l) {
        off += (int64_t)align_up((size_t)B * h * w * D, 64);
        h /= 2;
        w /= 2;
    }
    return off;
}

int raftcorr_alt_pyramid(const float* fmap1, const float* fmap2, int B, int D, int H, int W, int L, int mode,
                         float* f1_nhwc, float* f2_pyr, int device, raftcorr_stream_t stream) {
    RC_REQUIRE(mode == RAFTCORR_BUILD_FP32_SIMT || mode == RAFTCORR_BUILD_TF32_TC, "alt_pyramid: unknown mode %d", mode);
    const bool round = mode == RAFTCORR_BUILD_TF32_TC;  // the tensor-core lookup wants round-to-nearest TF32 operands
    RC_REQUIRE(B == 0 || (fmap1 && fmap2 && f1_nhwc && f2_pyr), "alt_pyramid: NULL tensor");
    AltLevels lv{};
    if (int rc = alt_levels(f2_pyr, nullptr, B, D, H, W, L, &lv)) return rc;
    if (B == 0) return 0;
    DeviceGuard guard(device);
    RC_REQUIRE(guard.ok, "alt_pyramid: cannot select device %d", device);
    cudaStream_t s = (cudaStream_t)stream;
    if (int rc = launch_nchw_to_nhwc_pair(fmap1, f1_nhwc, fmap2, const_cast<float*>(lv.f2[0]), B, D, H * W, D, round, s))
        return rc;
    for (int l = 1; l < L; ++l)  // corr.py:138-139 (fmap2 side; the pooled fmap1 copies are never read, :158)
        if (int rc = launch_pool_nhwc(lv.f2[l - 1], const_cast<float*>(lv.f2[l]), B, lv.h[l - 1], lv.w[l - 1], D, round, s))
            return rc;
    return 0;
}

size_t raftcorr_alt_lookup_workspace_bytes(int B, int H, int W, int L, int mode) {
    return mode == RAFTCORR_BUILD_TF32_TC ? alt_tc_workspace_bytes(B, H, W, L) : 0;
}

int raftcorr_alt_lookup_fwd(const float* f1_nhwc, const float* f2_pyr, const float* coords, int B, int D, int H,
                            int W, int L, int radius, int mode, float* out, void* workspace, size_t workspace_bytes,
                            int device, raftcorr_stream_t stream) {
    RC_REQUIRE(B == 0 || (f1_nhwc && f2_pyr && coords && out), "alt_lookup_fwd: NULL tensor");
    RC_REQUIRE(mode == RAFTCORR_BUILD_FP32_SIMT || mode == RAFTCORR_BUILD_TF32_TC, "alt_lookup_fwd: unknown mode %d", mode);
    AltLevels lv{};
    if (int rc = alt_levels(f2_pyr, nullptr, B, D, H, W, L, &lv)) return rc;
    if (B == 0) return 0;
    DeviceGuard guard(device);
    RC_REQUIRE(guard.ok, "alt_lookup_fwd: cannot select device %d", device);
    const int64_t N1 = (int64_t)H * W;
    CoordsView cv{coords, 2 * N1, N1, 1};  // [B,2,H,W]
    const float scale = 1.0f / sqrtf((float)D);  // corr.py:167
    if (mode == RAFTCORR_BUILD_TF32_TC)
        return launch_alt_fwd_tc(f1_nhwc, lv, cv, B, H, W, D, radius, scale, out, workspace, workspace_bytes, device,
                                 (cudaStream_t)stream);
    return launch_alt_fwd(f1_nhwc, lv, cv, B, 1, H, W, D, radius, scale, out, (cudaStream_t)stream);
}

// ---- fp16 operands for the tensor-core on-the-fly lookup ---------------------------------------------------------
// layout of the caller's buffer: [fmap1 halves][pyramid halves (same element offsets as the fp32 pyramid)][inv][maxbits]
namespace {
struct AltF16Layout { size_t f1, f2, inv, bits, total; };
AltF16Layout alt_f16_layout(int B, int D, int H, int W, int L) {
    AltF16Layout a{};
    a.f1 = 0;
    a.f2 = align_up((size_t)B * H * W * D * 2, 1024);
    a.inv = a.f2 + align_up((size_t)raftcorr_alt_level_offset(B, D, H, W, L) * 2, 1024);
    a.bits = a.inv + align_up((size_t)(1 + L) * B * sizeof(float), 256);
    a.total = a.bits + align_up((size_t)(1 + L) * B * sizeof(unsigned), 256);
    return a;
}
}  // namespace

size_t raftcorr_alt_f16_bytes(int B, int D, int H, int W, int L) {
    if (B < 0 || D < 1 || H < 1 || W < 1 || L < 1 || L > 4) return 0;
    return alt_f16_layout(B, D, H, W, L).total;
}

int raftcorr_alt_f16_prepare(const float* f1_nhwc, const float* f2_pyr, int B, int D, int H, int W, int L, void* f16_ops,
                             int device, raftcorr_stream_t stream) {
    RC_REQUIRE(B == 0 || (f1_nhwc && f2_pyr && f16_ops), "alt_f16_prepare: NULL tensor");
    RC_REQUIRE(D % 64 == 0, "alt_f16_prepare: needs D %% 64 == 0 (got %d)", D);
    AltLevels lv{};
    if (int rc = alt_levels(f2_pyr, nullptr, B, D, H, W, L, &lv)) return rc;
    if (B == 0) return 0;
    DeviceGuard guard(device);
    RC_REQUIRE(guard.ok, "alt_f16_prepare: cannot select device %d", device);
    const AltF16Layout a = alt_f16_layout(B, D, H, W, L);
    int64_t off[5] = {0}, per[5] = {0};
    per[0] = (int64_t)H * W * D;
    for (int l = 0; l < L; ++l) {
        off[1 + l] = lv.f2[l] - f2_pyr;
        per[1 + l] = (int64_t)lv.h[l] * lv.w[l] * D;
    }
    char* base = (char*)f16_ops;
    return launch_alt_f16_operands(f1_nhwc, f2_pyr, base + a.f1, base + a.f2, off, per, 1 + L, B, (float*)(base + a.inv),
                                   (unsigned*)(base + a.bits), (cudaStream_t)stream);
}

int raftcorr_alt_lookup_fwd_f16(const void* f16_ops, const float* coords, int B, int D, int H, int W, int L, int radius,
                                float* out, void* workspace, size_t workspace_bytes, int device, raftcorr_stream_t stream) {
    RC_REQUIRE(B == 0 || (f16_ops && coords && out), "alt_lookup_fwd_f16: NULL tensor");
    RC_REQUIRE(D % 64 == 0, "alt_lookup_fwd_f16: needs D %% 64 == 0 (got %d)", D);
    const AltF16Layout a = alt_f16_layout(B, D, H, W, L);
    const char* base = (const char*)f16_ops;
    AltLevels lv{};
    if (int rc = alt_levels((const float*)(base + a.f2), nullptr, B, D, H, W, L, &lv)) return rc;
    if (B == 0) return 0;
    // alt_levels computed FLOAT offsets; the halves live at the same ELEMENT offsets
    for (int l = 0; l < L; ++l)
        lv.f2[l] = (const float*)(base + a.f2 + 2 * (size_t)raftcorr_alt_level_offset(B, D, H, W, l));
    DeviceGuard guard(device);
    RC_REQUIRE(guard.ok, "alt_lookup_fwd_f16: cannot select device %d", device);
    const int64_t N1 = (int64_t)H * W;
    CoordsView cv{coords, 2 * N1, N1, 1};  // [B,2,H,W]
    return launch_alt_fwd_tc_ex(base + a.f1, lv, cv, B, H, W, D, radius, 1.0f / sqrtf((float)D), out, workspace, workspace_bytes,
                                device, (cudaStream_t)stream, true, (const float*)(base + a.inv));
}

int raftcorr_alt_lookup_bwd(const float* f1_nhwc, const float* f2_pyr, const float* coords, const float* grad_out,
                            int B, int D, int H, int W, int L, int radius, float* df1_nhwc, float* df2_pyr,
                            int device, raftcorr_stream_t stream) {
    RC_REQUIRE(B == 0 || (f1_nhwc && f2_pyr && coords && grad_out && df1_nhwc && df2_pyr), "alt_lookup_bwd: NULL tensor");
    AltLevels lv{};
    if (int rc = alt_levels(f2_pyr, df2_pyr, B, D, H, W, L, &lv)) return rc;
    if (B == 0) return 0;
    DeviceGuard guard(device);
    RC_REQUIRE(guard.ok, "alt_lookup_bwd: cannot select device %d", device);
    const int64_t N1 = (int64_t)H * W;
    CoordsView cv{coords, 2 * N1, N1, 1};
    const float scale = 1.0f / sqrtf((float)D);
    return launch_alt_bwd(f1_nhwc, lv, cv, grad_out, B, 1, H, W, D, radius, scale, df1_nhwc, (cudaStream_t)stream);
}

int raftcorr_alt_grad_finish(const float* df1_nhwc, float* df2_pyr, int B, int D, int H, int W, int L, float* dfmap1,
                             float* dfmap2, int device, raftcorr_stream_t stream) {
    RC_REQUIRE(B == 0 || (df1_nhwc && df2_pyr && dfmap1 && dfmap2), "alt_grad_finish: NULL tensor");
    AltLevels lv{};
    if (int rc = alt_levels(df2_pyr, df2_pyr, B, D, H, W, L, &lv)) return rc;
    if (B == 0) return 0;
    DeviceGuard guard(device);
    RC_REQUIRE(guard.ok, "alt_grad_finish: cannot select device %d", device);
    cudaStream_t s = (cudaStream_t)stream;
    for (int l = L - 1; l >= 1; --l)
        if (int rc = launch_pool_adjoint_nhwc(lv.df2[l], lv.df2[l - 1], B, lv.h[l - 1], lv.w[l - 1], D, s)) return rc;
    if (int rc = launch_nhwc_to_nchw(df1_nhwc, dfmap1, B, D, H * W, s)) return rc;
    return launch_nhwc_to_nchw(lv.df2[0], dfmap2, B, D, H * W, s);
}

}  // extern "C"

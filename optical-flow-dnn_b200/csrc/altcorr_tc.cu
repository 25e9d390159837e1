// On-the-fly correlation on the tensor cores (AlternateCorrBlock.__call__, raft/core/corr.py:142-167;
// the arithmetic of raft/alt_cuda_corr/correlation_kernel.cu:18-119 for all levels at once).
//
// The reference (and altcorr.cu) evaluates, per source pixel and level, the (K+1)^2 lattice of dot products
// it needs -- 102 k FMAs per pixel at D=256, r=4, L=4, on the CUDA cores.  Neighbouring source pixels look at
// overlapping windows, so here a CTA takes a block of 8x16 source pixels, finds the bounding box of all
// their windows on a level (alt_bbox_kernel), and computes the dot products of the 128 pixels with EVERY
// lattice point of that box as a 128 x 256 x D TF32 GEMM per 8x32 lattice tile on tcgen05: fmap1 block and
// fmap2 tile arrive by TMA as 128B-swizzled K-major operands straight from the NHWC feature tensors (4-D
// boxes; rows/columns outside the level are zero-filled by the TMA unit = grid_sample's zero padding),
// accumulators live in TMEM (two 256-column stages).  For smooth flow the box is (16+K+1) x (8+K+1) points,
// i.e. 2-3x more dot products than strictly needed, at ~30x the arithmetic rate.  Incoherent coordinates
// only cost time (the box is clipped to the level, so the worst case is the all-pairs product), never
// correctness.
//
// Epilogue (warps 4-7, TMEM lane == source pixel == thread): each accumulator row of 32 lattice points is
// copied to a thread-private shared-memory row, the thread picks the K+1 points of ITS window with its own
// sub-pixel weights and accumulates the K*K outputs of the level in a thread-private shared-memory column;
// contributions are linear, so windows that straddle lattice tiles simply add up.  No thread ever reads
// another thread's staging data: the epilogue needs no barrier besides the TMEM full/empty pair.
#include <climits>

#include "tc_util.cuh"

namespace raftcorr {

namespace {

constexpr int kSrcH = 8, kSrcW = 16;   // source block: 128 pixels = TMEM lanes, m = sy * 16 + sx
constexpr int kTgtH = 8, kTgtW = 32;   // lattice tile: 256 points = UMMA N, n = ty * 32 + tx
constexpr int kTileM = kSrcH * kSrcW;
constexpr int kTileN = kTgtH * kTgtW;
constexpr int kBlockK = 32, kUmmaK = 8;
constexpr int kStages = 3;
constexpr int kABytes = kTileM * kBlockK * 4;  // 16 KB
constexpr int kBBytes = kTileN * kBlockK * 4;  // 32 KB
constexpr int kStageBytes = kABytes + kBBytes;
constexpr int kRowPitch = kTgtW + 1;           // staged accumulator row of one pixel (+1 float: conflict-free)
constexpr int kThreads = 256;                  // warp 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4-7 epilogue
constexpr int kEpilogueThreads = 128;
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kInstrDesc = make_idesc_tf32(kTileM, kTileN, false, false);

__device__ __forceinline__ float clamp_coord(float v) { return fminf(fmaxf(v, -1048576.f), 1048576.f); }

struct AltTcMaps {
    CUtensorMap a;
    CUtensorMap b[4];
};

struct AltTcParams {
    const float* coords;
    int64_t cb, cc, cp;  // coords[b * cb + comp * cc + pixel * cp]
    int4* bbox;          // [tiles][L]: x, y = first lattice point (clipped to the level), z, w = lattice tiles in x, y
    float* out;          // [B][L*K*K][H][W]
    int B, H, W, L, k_blocks;
    int tiles_x, tiles_y, total_tiles;
    int lvl_h[4], lvl_w[4];
    float inv_scale[4];
    float scale;
};

constexpr int smem_bytes(int R) {
    return kStages * kStageBytes + kTileM * kRowPitch * 4 + (2 * R + 1) * (2 * R + 1) * kTileM * 4 + 1024 + 256;
}

// ---- pass 1: bounding box of the block's windows, per level ---------------------------------------------
template <int R>
__global__ void __launch_bounds__(kTileM) alt_bbox_kernel(const AltTcParams P) {
    constexpr int K = 2 * R + 1;
    __shared__ int red[4][4];
    const int t = blockIdx.x;
    const int tx = t % P.tiles_x, ty = (t / P.tiles_x) % P.tiles_y, b = t / (P.tiles_x * P.tiles_y);
    const int m = threadIdx.x, lane = m & 31, warp = m >> 5;
    const int y = ty * kSrcH + (m >> 4), x = tx * kSrcW + (m & 15);
    const bool valid = y < P.H && x < P.W;
    const int64_t pix = (int64_t)y * P.W + x;
    const float cx = clamp_coord(valid ? __ldg(P.coords + b * P.cb + pix * P.cp) : 0.f);
    const float cy = clamp_coord(valid ? __ldg(P.coords + b * P.cb + P.cc + pix * P.cp) : 0.f);
    for (int l = 0; l < P.L; ++l) {
        const int fx = __float2int_rd(cx * P.inv_scale[l]), fy = __float2int_rd(cy * P.inv_scale[l]);
        int x0 = valid ? fx - R : INT_MAX, x1 = valid ? fx - R + K : INT_MIN;  // lattice columns x0 .. x0 + K
        int y0 = valid ? fy - R : INT_MAX, y1 = valid ? fy - R + K : INT_MIN;
        x0 = __reduce_min_sync(0xffffffffu, x0);
        x1 = __reduce_max_sync(0xffffffffu, x1);
        y0 = __reduce_min_sync(0xffffffffu, y0);
        y1 = __reduce_max_sync(0xffffffffu, y1);
        if (lane == 0) { red[warp][0] = x0; red[warp][1] = x1; red[warp][2] = y0; red[warp][3] = y1; }
        __syncthreads();
        if (m == 0) {
            for (int w = 1; w < 4; ++w) {
                x0 = min(x0, red[w][0]); x1 = max(x1, red[w][1]);
                y0 = min(y0, red[w][2]); y1 = max(y1, red[w][3]);
            }
            // points outside the level are zero: clip (this also bounds the work for wild coordinates)
            x0 = max(x0, 0); x1 = min(x1, P.lvl_w[l] - 1);
            y0 = max(y0, 0); y1 = min(y1, P.lvl_h[l] - 1);
            int4 bb = make_int4(0, 0, 0, 0);
            if (x1 >= x0 && y1 >= y0) bb = make_int4(x0, y0, (x1 - x0) / kTgtW + 1, (y1 - y0) / kTgtH + 1);
            P.bbox[(int64_t)t * P.L + l] = bb;
        }
        __syncthreads();
    }
}

// ---- pass 2: GEMM tiles + gather -------------------------------------------------------------------------
template <int R>
__global__ void __launch_bounds__(kThreads, 1)
alt_fwd_tc_kernel(const __grid_constant__ AltTcMaps maps, const AltTcParams P) {
    constexpr int K = 2 * R + 1, KK = K * K;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // 128B-swizzle atoms want 1024-byte alignment
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t stage_base = smem_base;
    float* rows = reinterpret_cast<float*>(smem_gen + kStages * kStageBytes);             // [128][kRowPitch]
    float* acc = rows + kTileM * kRowPitch;                                               // [KK][128]
    const uint32_t bar_base = smem_base + kStages * kStageBytes + (kTileM * kRowPitch + KK * kTileM) * 4;
    const uint32_t full_bar = bar_base, empty_bar = bar_base + 8 * kStages;
    const uint32_t tfull_bar = bar_base + 16 * kStages, tempty_bar = tfull_bar + 16, tmem_slot = tempty_bar + 16;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&maps.a);
        for (int l = 0; l < P.L; ++l) prefetch_tensormap(&maps.b[l]);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(full_bar + 8 * s, 1);
            mbar_init(empty_bar + 8 * s, 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull_bar + 8 * a, 1);
            mbar_init(tempty_bar + 8 * a, kEpilogueThreads);
        }
        fence_mbar_init();
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));
    const int tiles_per_b = P.tiles_x * P.tiles_y;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int t = blockIdx.x; t < P.total_tiles; t += gridDim.x) {
                const int tx = t % P.tiles_x, ty = (t / P.tiles_x) % P.tiles_y, b = t / tiles_per_b;
                for (int l = 0; l < P.L; ++l) {
                    const int4 bb = P.bbox[(int64_t)t * P.L + l];
                    for (int c = 0; c < bb.w; ++c)
                        for (int a = 0; a < bb.z; ++a)
                            for (int kb = 0; kb < P.k_blocks; ++kb) {
                                mbar_wait(empty_bar + 8 * stage, phase ^ 1);
                                const uint32_t sa = stage_base + stage * kStageBytes;
                                mbar_expect_tx(full_bar + 8 * stage, kStageBytes);
                                tma_load_4d(sa, &maps.a, full_bar + 8 * stage, kb * kBlockK, tx * kSrcW, ty * kSrcH, b);
                                tma_load_4d(sa + kABytes, &maps.b[l], full_bar + 8 * stage, kb * kBlockK, bb.x + a * kTgtW,
                                            bb.y + c * kTgtH, b);
                                if (++stage == kStages) { stage = 0; phase ^= 1; }
                            }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            int stage = 0, it = 0;
            uint32_t phase = 0;
            for (int t = blockIdx.x; t < P.total_tiles; t += gridDim.x) {
                for (int l = 0; l < P.L; ++l) {
                    const int4 bb = P.bbox[(int64_t)t * P.L + l];
                    const int ntiles = bb.z * bb.w;
                    for (int n = 0; n < ntiles; ++n, ++it) {
                        const int as = it & 1;
                        mbar_wait(tempty_bar + 8 * as, ((it >> 1) & 1) ^ 1);  // epilogue has drained this accumulator
                        tc_fence_after();
                        const uint32_t d_tmem = tmem_base + as * kTileN;
                        for (int kb = 0; kb < P.k_blocks; ++kb) {
                            mbar_wait(full_bar + 8 * stage, phase);
                            tc_fence_after();
                            const uint32_t sa = stage_base + stage * kStageBytes;
                            const uint64_t adesc = make_smem_desc(sa), bdesc = make_smem_desc(sa + kABytes);
#pragma unroll
                            for (int k = 0; k < kBlockK / kUmmaK; ++k)
                                tc_mma_tf32(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), kInstrDesc,
                                            (kb | k) != 0 ? 1u : 0u);
                            tc_commit(empty_bar + 8 * stage);
                            if (++stage == kStages) { stage = 0; phase ^= 1; }
                        }
                        tc_commit(tfull_bar + 8 * as);
                    }
                }
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue: gather the windows out of the accumulator tiles =====================
        const int ew = warp & 3;  // TMEM lane quarter this warp may access
        const int m = ew * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(ew * 32) << 16;
        float* my_row = rows + m * kRowPitch;
        float* my_acc = acc + m;
        int it = 0;
        for (int t = blockIdx.x; t < P.total_tiles; t += gridDim.x) {
            const int tx = t % P.tiles_x, ty = (t / P.tiles_x) % P.tiles_y, b = t / tiles_per_b;
            const int y = ty * kSrcH + (m >> 4), x = tx * kSrcW + (m & 15);
            const bool valid = y < P.H && x < P.W;
            const int64_t pix = (int64_t)y * P.W + x;
            const float cx = clamp_coord(valid ? __ldg(P.coords + b * P.cb + pix * P.cp) : 0.f);
            const float cy = clamp_coord(valid ? __ldg(P.coords + b * P.cb + P.cc + pix * P.cp) : 0.f);
            for (int l = 0; l < P.L; ++l) {
                const int4 bb = P.bbox[(int64_t)t * P.L + l];
                const float xs = cx * P.inv_scale[l], ys = cy * P.inv_scale[l];
                const float xf = floorf(xs), yf = floorf(ys);
                const float ax = xs - xf, ay = ys - yf;
                const int x0 = (int)xf - R, y0 = (int)yf - R;  // first lattice column / row of this pixel's window
#pragma unroll 9
                for (int ch = 0; ch < KK; ++ch) my_acc[ch * kTileM] = 0.f;
                for (int c = 0; c < bb.w; ++c)
                    for (int a = 0; a < bb.z; ++a, ++it) {
                        const int as = it & 1;
                        mbar_wait(tfull_bar + 8 * as, (it >> 1) & 1);
                        tc_fence_after();
                        const uint32_t t_acc = tmem_base + lane_addr + as * kTileN;
                        const int c0 = x0 - (bb.x + a * kTgtW);  // window start column inside this lattice tile
                        const int jr = (bb.y + c * kTgtH) - y0;  // window row index of the tile's first row
                        const bool cols_hit = valid && c0 + K >= 0 && c0 < kTgtW;
#pragma unroll 1
                        for (int r = 0; r < kTgtH; ++r) {
                            const int j = jr + r;  // lattice row j of the window: top tap of output row j, bottom tap of j-1
                            const bool need = cols_hit && j >= 0 && j <= K;
                            if (__any_sync(0xffffffffu, need)) {
                                float v[32];
                                tmem_ld32(t_acc + r * kTgtW, v);
                                tmem_ld_wait();
                                if (need) {
#pragma unroll
                                    for (int q = 0; q < 32; ++q) my_row[q] = v[q];
                                    // horizontal lerps of the K window columns on this lattice row (points outside
                                    // this tile count as zero here; the neighbouring tile adds them)
                                    float prev = (c0 >= 0) ? my_row[c0] : 0.f;
                                    float hl[K];
#pragma unroll
                                    for (int i = 0; i < K; ++i) {
                                        const int cc = c0 + i + 1;
                                        const float next = (cc >= 0 && cc < kTgtW) ? my_row[cc] : 0.f;
                                        hl[i] = prev + ax * (next - prev);
                                        prev = next;
                                    }
                                    if (j < K) {
                                        const float wt = 1.f - ay;
#pragma unroll
                                        for (int i = 0; i < K; ++i) my_acc[(i * K + j) * kTileM] += wt * hl[i];
                                    }
                                    if (j >= 1) {
#pragma unroll
                                        for (int i = 0; i < K; ++i) my_acc[(i * K + j - 1) * kTileM] += ay * hl[i];
                                    }
                                }
                            }
                        }
                        tc_fence_before();
                        mbar_arrive(tempty_bar + 8 * as);  // every tcgen05.ld of this thread on the stage has completed
                    }
                if (valid) {
                    float* dst = P.out + (((int64_t)b * P.L + l) * KK * P.H + y) * P.W + x;
                    const int64_t plane = (int64_t)P.H * P.W;
#pragma unroll 9
                    for (int ch = 0; ch < KK; ++ch) dst[ch * plane] = my_acc[ch * kTileM] * P.scale;
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

template <int R>
int run_alt_fwd_tc(const AltTcMaps& maps, const AltTcParams& P, int sms, cudaStream_t s) {
    alt_bbox_kernel<R><<<P.total_tiles, kTileM, 0, s>>>(P);
    RC_CUDA(cudaGetLastError());
    constexpr int smem = smem_bytes(R);
    static_assert(smem <= 232448, "on-the-fly tensor-core kernel does not fit shared memory");
    RC_CUDA(cudaFuncSetAttribute(alt_fwd_tc_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int grid = P.total_tiles < sms ? P.total_tiles : sms;
    alt_fwd_tc_kernel<R><<<grid, kThreads, smem, s>>>(maps, P);
    RC_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace

size_t alt_tc_workspace_bytes(int B, int H, int W, int L) {
    return (size_t)B * ceil_div(H, kSrcH) * ceil_div(W, kSrcW) * (size_t)L * sizeof(int4);
}

// f1_nhwc / lv.f2[]: TF32-rounded NHWC features with C % 32 == 0; out: [B][L*K*K][H][W]
int launch_alt_fwd_tc(const float* f1_nhwc, const AltLevels& lv, const CoordsView& cv, int B, int H, int W, int C,
                      int radius, float scale, float* out, void* workspace, size_t workspace_bytes, int device,
                      cudaStream_t s) {
    RC_REQUIRE(lv.L >= 1 && lv.L <= 4, "alt tf32: 1..4 levels supported (got %d)", lv.L);
    RC_REQUIRE(radius >= 1 && radius <= 4, "alt tf32: radius %d not supported (1..4)", radius);
    RC_REQUIRE(C % kBlockK == 0 && C >= kBlockK, "alt tf32: feature dim must be a multiple of %d (got %d)", kBlockK, C);
    RC_REQUIRE(workspace && workspace_bytes >= alt_tc_workspace_bytes(B, H, W, lv.L), "alt tf32: workspace too small");
    RC_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "alt tf32: workspace must be 16-byte aligned");
    EncodeTiledFn fn;
    if (int rc = get_encode_fn(&fn)) return rc;
    AltTcMaps maps;
    {
        cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
        cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
        cuuint32_t box[4] = {kBlockK, kSrcW, kSrcH, 1};
        if (int rc = encode_f32_map(fn, &maps.a, (void*)f1_nhwc, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B,
                                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, "fmap1 (NHWC)"))
            return rc;
    }
    for (int l = 0; l < 4; ++l) {
        const int ll = l < lv.L ? l : 0;
        cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)lv.w[ll], (cuuint64_t)lv.h[ll], (cuuint64_t)B};
        cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)lv.w[ll] * C * 4, (cuuint64_t)lv.h[ll] * lv.w[ll] * C * 4};
        cuuint32_t box[4] = {kBlockK, kTgtW, kTgtH, 1};
        if (int rc = encode_f32_map(fn, &maps.b[l], (void*)lv.f2[ll], 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B,
                                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, "fmap2 level (NHWC)"))
            return rc;
    }
    AltTcParams P{};
    P.coords = cv.ptr; P.cb = cv.batch_stride; P.cc = cv.comp_stride; P.cp = cv.pix_stride;
    P.bbox = reinterpret_cast<int4*>(workspace);
    P.out = out;
    P.B = B; P.H = H; P.W = W; P.L = lv.L; P.k_blocks = C / kBlockK;
    P.tiles_x = ceil_div(W, kSrcW); P.tiles_y = ceil_div(H, kSrcH);
    const int64_t total = (int64_t)B * P.tiles_x * P.tiles_y;
    RC_REQUIRE(total < (1LL << 31), "alt tf32: too many source blocks");
    P.total_tiles = (int)total;
    for (int l = 0; l < lv.L; ++l) {
        P.lvl_h[l] = lv.h[l];
        P.lvl_w[l] = lv.w[l];
        P.inv_scale[l] = 1.0f / (float)(1 << l);
    }
    P.scale = scale;
    int sms = 0;
    RC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    switch (radius) {
        case 1: return run_alt_fwd_tc<1>(maps, P, sms, s);
        case 2: return run_alt_fwd_tc<2>(maps, P, sms, s);
        case 3: return run_alt_fwd_tc<3>(maps, P, sms, s);
        default: return run_alt_fwd_tc<4>(maps, P, sms, s);
    }
}

}  // namespace raftcorr

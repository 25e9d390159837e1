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
#include <cstdlib>

#include "tc_util.cuh"

namespace raftcorr {

namespace {

constexpr int kSrcH = 8, kSrcW = 16;   // source block: 128 pixels = TMEM lanes, m = sy * 16 + sx
constexpr int kTgtH = 8, kTgtW = 32;   // lattice tile: 256 points = UMMA N, n = ty * 32 + tx  (shape 0: 8 rows x 32 cols)
constexpr int kSqH = 16, kSqW = 16;    // shape 1: 16 x 16, n = ty * 16 + tx -- covers the small boxes of the coarse levels in one tile
constexpr int kHalfH = 4, kHalfW = 32;  // resident-A kernel: 4 x 32 lattice tiles = 128 accumulator columns (rows stay 32 wide: the gather cost is per row)
constexpr int kHalfN = kHalfH * kHalfW;
constexpr int kTileM = kSrcH * kSrcW;
constexpr int kTileN = kTgtH * kTgtW;
constexpr int kBlockK = 32, kUmmaK = 8;
constexpr int kStages = 3;
constexpr int kABytes = kTileM * kBlockK * 4;  // 16 KB
constexpr int kBBytes = kTileN * kBlockK * 4;  // 32 KB
constexpr int kStageBytes = kABytes + kBBytes;
constexpr int kRowPitch = kTgtW + 4;           // staged accumulator row of one pixel (16-byte aligned, bank-skewed)
// A warp's 32 staging rows + 16 floats of skew: the rows of lanes 16..31 (the second source row of the TMEM quarter, whose
// windows start at the same columns as those of lanes 0..15) are shifted by half the banks, so the per-lane reads
// lr[c0 + i] (c0 ~ lane & 15) of the two half-warps no longer collide (ncu: 26 % of the wavefronts were 2-way conflicts).
constexpr int kWarpRows = 32 * kRowPitch + 16;
constexpr int kRowsFloats = 4 * kWarpRows;     // one set of staging rows (4 TMEM lane quarters)
constexpr int kThreads = 384;                  // warp 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4-11 epilogue
constexpr int kEpilogueThreads = 256;
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kInstrDesc = make_idesc_tf32(kTileM, kTileN, false, false);

__device__ __forceinline__ float clamp_coord(float v) { return fminf(fmaxf(v, -1048576.f), 1048576.f); }

struct AltTcMaps {
    CUtensorMap a;
    CUtensorMap b[4];    // 8 x 32 lattice tiles (resident-A kernel: 8 x 16)
    CUtensorMap bsq[4];  // 16 x 16 lattice tiles
};

struct AltTcParams {
    const float* coords;
    int64_t cb, cc, cp;  // coords[b * cb + comp * cc + pixel * cp]
    int4* bbox;          // [tiles][L]: x, y = first lattice point (clipped to the level); z = tiles in x | shape << 16; w = tiles in y
    float* out;          // [B][L*K*K][H][W]
    int B, H, W, L, k_blocks;
    int tiles_x, tiles_y, total_tiles;
    int lvl_h[4], lvl_w[4];
    float inv_scale[4];
    float scale;
    int allow_square;  // pick 16 x 16 lattice tiles where they cover a box with fewer tiles than 8 x 32
    int bb_w, bb_h;    // primary lattice tile shape the bounding boxes are cut into
    const float* f1;   // [B][H][W][C] (resident-A kernel reads it directly)
    int C;
    // fp16 operands with one power-of-two scale per image and tensor (launch_alt_f16_operands): inv[b] for fmap1,
    // inv[(1 + l) * B + b] for level l of the fmap2 pyramid; k_elems = operand elements per 128-byte k-block
    int f16, k_elems;
    const float* inv;
};

constexpr uint32_t kInstrDescF16 = (1u << 4) | ((uint32_t)(kTileN >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
__device__ __forceinline__ void tc_mma_f16_alt(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(kInstrDescF16), "r"(accumulate)
        : "memory");
}

constexpr int smem_bytes(int R) {
    return kStages * kStageBytes + 2 * kRowsFloats * 4 + (2 * R + 1) * (2 * R + 1) * kTileM * 4 + 1024 + 256;
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
            if (x1 >= x0 && y1 >= y0) {
                const int nx0 = (x1 - x0) / P.bb_w + 1, ny0 = (y1 - y0) / P.bb_h + 1;  // 8 x 32 (8 x 16) tiling
                const int nx1 = (x1 - x0) / kSqW + 1, ny1 = (y1 - y0) / kSqH + 1;    // 16 x 16 tiling
                bb = (P.allow_square && nx1 * ny1 < nx0 * ny0) ? make_int4(x0, y0, nx1 | (1 << 16), ny1)
                                                               : make_int4(x0, y0, nx0, ny0);
            }
            P.bbox[(int64_t)t * P.L + l] = bb;
        }
        __syncthreads();
    }
}

// Epilogue of one warp: TMEM lane quarter `ew`, window columns [I0, I0 + IN) of its 32 source pixels.
template <int R, int I0, int IN, bool V2>
__device__ __forceinline__ void alt_tc_epilogue(const AltTcParams& P, int ew, int lane, uint32_t tmem_base,
                                                uint32_t tfull_bar, uint32_t tempty_bar, float* rows, float* acc) {
    constexpr int K = 2 * R + 1;
    const int m = ew * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(ew * 32) << 16;
    float* my_row = rows + ew * kWarpRows + lane * kRowPitch + (lane >= 16 ? 16 : 0);
    const uint32_t my_row_s = smem_u32(my_row);
    float* my_acc = acc + (I0 * K) * kTileM + m;  // this warp's output channels: [I0 * K, (I0 + IN) * K)
    const int tiles_per_b = P.tiles_x * P.tiles_y;
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
            const float ax = xs - xf, ay = ys - yf, wt = 1.f - ay;
            const int x0 = (int)xf - R + I0, y0 = (int)yf - R;  // first lattice column (of this warp's share) / row
#pragma unroll
            for (int ch = 0; ch < IN * K; ++ch) my_acc[ch * kTileM] = 0.f;
            const bool sq = bb.z >> 16;
            const int ntx = bb.z & 0xffff;
            // lattice tile: th rows of tw points (V2: always 8 x 16 = 128 accumulator columns)
            const int tw = V2 ? kHalfW : (sq ? kSqW : kTgtW), th = V2 ? kHalfH : (sq ? kSqH : kTgtH);
            const int rows_per_ld = 32 / tw;  // lattice rows per 32-column TMEM load
            constexpr int kAccCols = V2 ? kHalfN : kTileN;
            for (int c = 0; c < bb.w; ++c)
                for (int a = 0; a < ntx; ++a, ++it) {
                    const int as = it & 1;
                    mbar_wait(tfull_bar + 8 * as, (it >> 1) & 1);
                    tc_fence_after();
                    const uint32_t t_acc = tmem_base + lane_addr + as * kAccCols;
                    const int c0 = x0 - (bb.x + a * tw);  // first needed column inside this lattice tile
                    const int jr = (bb.y + c * th) - y0;  // window row index of the tile's first row
                    const bool cols_hit = valid && c0 + IN >= 0 && c0 < tw;
                    // Lattice row j of the window is the top tap of output row j and the bottom tap of row j-1.  The
                    // rows of a tile come in order, so output row j-1 is finished from two consecutive rows with ONE
                    // shared-memory update; the first / last row a tile holds leave partial sums (tiles add up).
                    float hp[IN];
                    bool have_prev = false;
                    int j_prev = 0;
#pragma unroll 1
                    for (int q32 = 0; q32 < kAccCols / 32; ++q32) {
                        const int j_first = jr + q32 * rows_per_ld;
                        const bool need = cols_hit && j_first + rows_per_ld - 1 >= 0 && j_first <= K;
                        if (__any_sync(0xffffffffu, need)) {
                            float v[32];
                            tmem_ld32(t_acc + q32 * 32, v);
                            tmem_ld_wait();
                            if (need) {
#pragma unroll
                                for (int q = 0; q < 8; ++q)
                                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(my_row_s + 16 * q), "f"(v[4 * q]),
                                                 "f"(v[4 * q + 1]), "f"(v[4 * q + 2]), "f"(v[4 * q + 3])
                                                 : "memory");
                                for (int rr = 0; rr < rows_per_ld; ++rr) {
                                    const int j = j_first + rr;
                                    if (j < 0 || j > K) continue;
                                    const volatile float* lr = my_row + rr * tw;
                                    // horizontal lerps of this warp's window columns (points outside this tile count as
                                    // zero here; the neighbouring tile adds them)
                                    float prev = (c0 >= 0) ? lr[c0] : 0.f;
                                    float hl[IN];
#pragma unroll
                                    for (int i = 0; i < IN; ++i) {
                                        const int cc = c0 + i + 1;
                                        const float next = (cc >= 0 && cc < tw) ? lr[cc] : 0.f;
                                        hl[i] = prev + ax * (next - prev);
                                        prev = next;
                                    }
                                    if (have_prev) {
#pragma unroll
                                        for (int i = 0; i < IN; ++i) my_acc[(i * K + j - 1) * kTileM] += wt * hp[i] + ay * hl[i];
                                    } else if (j >= 1) {
#pragma unroll
                                        for (int i = 0; i < IN; ++i) my_acc[(i * K + j - 1) * kTileM] += ay * hl[i];
                                    }
#pragma unroll
                                    for (int i = 0; i < IN; ++i) hp[i] = hl[i];
                                    have_prev = true;
                                    j_prev = j;
                                }
                            }
                        }
                    }
                    tc_fence_before();
                    mbar_arrive(tempty_bar + 8 * as);  // every tcgen05.ld of this thread on the stage has completed
                    if (have_prev && j_prev < K) {
#pragma unroll
                        for (int i = 0; i < IN; ++i) my_acc[(i * K + j_prev) * kTileM] += wt * hp[i];
                    }
                }
            if (valid) {
                float* dst = P.out + ((((int64_t)b * P.L + l) * K * K + I0 * K) * P.H + y) * P.W + x;
                const int64_t plane = (int64_t)P.H * P.W;
                const float osc = P.f16 ? P.scale * __ldg(P.inv + b) * __ldg(P.inv + (1 + l) * P.B + b) : P.scale;
#pragma unroll
                for (int ch = 0; ch < IN * K; ++ch) dst[ch * plane] = my_acc[ch * kTileM] * osc;
            }
        }
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
    float* rows = reinterpret_cast<float*>(smem_gen + kStages * kStageBytes);             // [2][4 warps][32 rows of kRowPitch + skew]
    float* acc = rows + 2 * kRowsFloats;                                           // [KK][128]
    const uint32_t bar_base = smem_base + kStages * kStageBytes + (2 * kRowsFloats + KK * kTileM) * 4;
    const uint32_t full_bar = bar_base, empty_bar = bar_base + 8 * kStages;
    const uint32_t tfull_bar = bar_base + 16 * kStages, tempty_bar = tfull_bar + 16, tmem_slot = tempty_bar + 16;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&maps.a);
        for (int l = 0; l < P.L; ++l) {
            prefetch_tensormap(&maps.b[l]);
            prefetch_tensormap(&maps.bsq[l]);
        }
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
                    const bool sq = bb.z >> 16;
                    const int ntx = bb.z & 0xffff, tw = sq ? kSqW : kTgtW, th = sq ? kSqH : kTgtH;
                    const CUtensorMap* mb = sq ? &maps.bsq[l] : &maps.b[l];
                    for (int c = 0; c < bb.w; ++c)
                        for (int a = 0; a < ntx; ++a)
                            for (int kb = 0; kb < P.k_blocks; ++kb) {
                                mbar_wait(empty_bar + 8 * stage, phase ^ 1);
                                const uint32_t sa = stage_base + stage * kStageBytes;
                                mbar_expect_tx(full_bar + 8 * stage, kStageBytes);
                                tma_load_4d(sa, &maps.a, full_bar + 8 * stage, kb * P.k_elems, tx * kSrcW, ty * kSrcH, b);
                                tma_load_4d(sa + kABytes, mb, full_bar + 8 * stage, kb * P.k_elems, bb.x + a * tw, bb.y + c * th, b);
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
                    const int ntiles = (bb.z & 0xffff) * bb.w;
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
                                if (P.f16)  // K = 16 halves: the same 32 bytes per instruction
                                    tc_mma_f16_alt(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), (kb | k) != 0 ? 1u : 0u);
                                else
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
        // two warps per TMEM lane quarter: each evaluates half of the window columns of its 32 pixels
        constexpr int KH = (K + 1) / 2;
        if (warp < 8)
            alt_tc_epilogue<R, 0, KH, false>(P, warp & 3, lane, tmem_base, tfull_bar, tempty_bar, rows, acc);
        else
            alt_tc_epilogue<R, KH, K - KH, false>(P, warp & 3, lane, tmem_base, tfull_bar, tempty_bar, rows + kRowsFloats, acc);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

// ---- pass 2, resident-A variant (D <= 256; A/B switch RAFTCORR_ALT_TC=resident, NOT the default) -----------
// The kernel above streams the 128 x D fmap1 block again for every lattice tile (1/3 of its operand bytes, and
// it runs at the chip's L2->SM throughput).  Here the block is written ONCE per source block into tensor
// memory -- lane = source pixel, one 32-bit column per feature, columns 256..511 -- by four loader warps
// (tcgen05.st of rows they read straight from the NHWC tensor), and the MMAs take A from TMEM
// (tcgen05.mma [d], [a], b-desc).  That leaves 256 columns for the accumulators: two stages of 128, i.e. 8 x 16
// lattice tiles, which also fit the small boxes of the coarse levels more tightly.  Only fmap2 tiles stream
// through shared memory (16 KB per k-block, 6 stages).  Bit-compatible results; measured ~30 % slower than the
// streaming kernel (the 128-column tiles double the MMA and hand-over count), kept as the working example of
// tcgen05.mma with the A operand in TMEM.
constexpr int kV2Stages = 6;
constexpr int kV2BBytes = kHalfN * kBlockK * 4;  // 16 KB
constexpr int kV2Threads = 512;                   // + warps 12-15: fmap1 block -> TMEM
constexpr uint32_t kV2ACol = 256;                 // first TMEM column of the resident fmap1 block
constexpr uint32_t kInstrDescHalf = make_idesc_tf32(kTileM, kHalfN, false, false);
constexpr int smem_bytes_v2(int R) {
    return kV2Stages * kV2BBytes + 2 * kRowsFloats * 4 + (2 * R + 1) * (2 * R + 1) * kTileM * 4 + 1024 + 256;
}

template <int R>
__global__ void __launch_bounds__(kV2Threads, 1)
alt_fwd_tc_resident_kernel(const __grid_constant__ AltTcMaps maps, const AltTcParams P) {
    constexpr int K = 2 * R + 1, KK = K * K;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t stage_base = smem_base;
    float* rows = reinterpret_cast<float*>(smem_gen + kV2Stages * kV2BBytes);  // [2][4 warps][32 rows of kRowPitch + skew]
    float* acc = rows + 2 * kRowsFloats;                                 // [KK][128]
    const uint32_t bar_base = smem_base + kV2Stages * kV2BBytes + (2 * kRowsFloats + KK * kTileM) * 4;
    const uint32_t full_bar = bar_base, empty_bar = bar_base + 8 * kV2Stages;
    const uint32_t tfull_bar = bar_base + 16 * kV2Stages, tempty_bar = tfull_bar + 16;
    const uint32_t afull_bar = tempty_bar + 16, aempty_bar = afull_bar + 8, tmem_slot = aempty_bar + 8;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0)
        for (int l = 0; l < P.L; ++l) prefetch_tensormap(&maps.b[l]);
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kV2Stages; ++s) {
            mbar_init(full_bar + 8 * s, 1);
            mbar_init(empty_bar + 8 * s, 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull_bar + 8 * a, 1);
            mbar_init(tempty_bar + 8 * a, kEpilogueThreads);
        }
        mbar_init(afull_bar, kTileM);  // 128 loader threads
        mbar_init(aempty_bar, 1);      // one tcgen05.commit
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
        // ===================== TMA producer (fmap2 tiles only) =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int t = blockIdx.x; t < P.total_tiles; t += gridDim.x) {
                const int b = t / tiles_per_b;
                for (int l = 0; l < P.L; ++l) {
                    const int4 bb = P.bbox[(int64_t)t * P.L + l];
                    for (int c = 0; c < bb.w; ++c)
                        for (int a = 0; a < bb.z; ++a)
                            for (int kb = 0; kb < P.k_blocks; ++kb) {
                                mbar_wait(empty_bar + 8 * stage, phase ^ 1);
                                mbar_expect_tx(full_bar + 8 * stage, kV2BBytes);
                                tma_load_4d(stage_base + stage * kV2BBytes, &maps.b[l], full_bar + 8 * stage, kb * kBlockK,
                                            bb.x + a * kHalfW, bb.y + c * kHalfH, b);
                                if (++stage == kV2Stages) { stage = 0; phase ^= 1; }
                            }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer: A from tensor memory =====================
        if (lane == 0) {
            int stage = 0, it = 0, blk = 0;
            uint32_t phase = 0;
            for (int t = blockIdx.x; t < P.total_tiles; t += gridDim.x, ++blk) {
                mbar_wait(afull_bar, blk & 1);  // this block's fmap1 rows are in TMEM
                tc_fence_after();
                for (int l = 0; l < P.L; ++l) {
                    const int4 bb = P.bbox[(int64_t)t * P.L + l];
                    const int ntiles = bb.z * bb.w;
                    for (int n = 0; n < ntiles; ++n, ++it) {
                        const int as = it & 1;
                        mbar_wait(tempty_bar + 8 * as, ((it >> 1) & 1) ^ 1);
                        tc_fence_after();
                        const uint32_t d_tmem = tmem_base + as * kHalfN;
                        for (int kb = 0; kb < P.k_blocks; ++kb) {
                            mbar_wait(full_bar + 8 * stage, phase);
                            tc_fence_after();
                            const uint64_t bdesc = make_smem_desc(stage_base + stage * kV2BBytes);
#pragma unroll
                            for (int k = 0; k < kBlockK / kUmmaK; ++k)
                                tc_mma_tf32_ts(d_tmem, tmem_base + kV2ACol + kb * kBlockK + k * kUmmaK, bdesc + (uint64_t)(2 * k),
                                               kInstrDescHalf, (kb | k) != 0 ? 1u : 0u);
                            tc_commit(empty_bar + 8 * stage);
                            if (++stage == kV2Stages) { stage = 0; phase ^= 1; }
                        }
                        tc_commit(tfull_bar + 8 * as);
                    }
                }
                tc_commit(aempty_bar);  // all MMAs that read this block's A have retired -> the loaders may overwrite it
            }
        }
    } else if (warp >= 12) {
        // ===================== fmap1 block -> tensor memory =====================
        const int ew = warp & 3;
        const int m = ew * 32 + lane;
        const uint32_t a_addr = tmem_base + ((uint32_t)(ew * 32) << 16) + kV2ACol;
        int blk = 0;
        for (int t = blockIdx.x; t < P.total_tiles; t += gridDim.x, ++blk) {
            const int tx = t % P.tiles_x, ty = (t / P.tiles_x) % P.tiles_y, b = t / tiles_per_b;
            const int y = ty * kSrcH + (m >> 4), x = tx * kSrcW + (m & 15);
            const bool valid = y < P.H && x < P.W;
            const float4* src = reinterpret_cast<const float4*>(P.f1 + (((int64_t)b * P.H + y) * P.W + x) * P.C);
            mbar_wait(aempty_bar, (blk & 1) ^ 1);  // previous block's MMAs are done with the columns
            tc_fence_after();
            for (int kb = 0; kb < P.k_blocks; ++kb) {
                float v[32];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float4 f = valid ? __ldg(src + kb * 8 + q) : make_float4(0.f, 0.f, 0.f, 0.f);
                    v[4 * q] = f.x; v[4 * q + 1] = f.y; v[4 * q + 2] = f.z; v[4 * q + 3] = f.w;
                }
                tmem_st32(a_addr + kb * kBlockK, v);
            }
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(afull_bar);
        }
    } else if (warp >= 4) {
        constexpr int KH = (K + 1) / 2;
        if (warp < 8)
            alt_tc_epilogue<R, 0, KH, true>(P, warp & 3, lane, tmem_base, tfull_bar, tempty_bar, rows, acc);
        else
            alt_tc_epilogue<R, KH, K - KH, true>(P, warp & 3, lane, tmem_base, tfull_bar, tempty_bar, rows + kRowsFloats, acc);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

template <int R>
int run_alt_fwd_tc(const AltTcMaps& maps, const AltTcParams& P, int sms, bool resident, cudaStream_t s) {
    alt_bbox_kernel<R><<<P.total_tiles, kTileM, 0, s>>>(P);
    RC_CUDA(cudaGetLastError());
    if (resident) {
        constexpr int smem2 = smem_bytes_v2(R);
        RC_CUDA(cudaFuncSetAttribute(alt_fwd_tc_resident_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2));
        const int grid2 = P.total_tiles < sms ? P.total_tiles : sms;
        alt_fwd_tc_resident_kernel<R><<<grid2, kV2Threads, smem2, s>>>(maps, P);
        RC_CUDA(cudaGetLastError());
        return 0;
    }
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
    return launch_alt_fwd_tc_ex(f1_nhwc, lv, cv, B, H, W, C, radius, scale, out, workspace, workspace_bytes, device, s, false,
                                nullptr);
}

// f16: f1_nhwc / lv.f2[] point at fp16 NHWC copies (C % 64 == 0) scaled per image; inv: see AltTcParams
int launch_alt_fwd_tc_ex(const void* f1_nhwc, const AltLevels& lv, const CoordsView& cv, int B, int H, int W, int C,
                         int radius, float scale, float* out, void* workspace, size_t workspace_bytes, int device,
                         cudaStream_t s, bool f16, const float* inv) {
    const int esz = f16 ? 2 : 4, k_elems = 128 / esz;
    const CUtensorMapDataType dt = f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    RC_REQUIRE(C % k_elems == 0, "alt tensor-core lookup: feature dim must be a multiple of %d (got %d)", k_elems, C);
    RC_REQUIRE(lv.L >= 1 && lv.L <= 4, "alt tf32: 1..4 levels supported (got %d)", lv.L);
    RC_REQUIRE(radius >= 1 && radius <= 4, "alt tf32: radius %d not supported (1..4)", radius);
    RC_REQUIRE(workspace && workspace_bytes >= alt_tc_workspace_bytes(B, H, W, lv.L), "alt tf32: workspace too small");
    RC_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "alt tf32: workspace must be 16-byte aligned");
    EncodeTiledFn fn;
    if (int rc = get_encode_fn(&fn)) return rc;
    AltTcMaps maps;
    {
        cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
        cuuint64_t strides[3] = {(cuuint64_t)C * esz, (cuuint64_t)W * C * esz, (cuuint64_t)H * W * C * esz};
        cuuint32_t box[4] = {(cuuint32_t)k_elems, kSrcW, kSrcH, 1};
        if (int rc = encode_map(fn, &maps.a, dt, (void*)f1_nhwc, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, "fmap1 (NHWC)"))
            return rc;
    }
    // RAFTCORR_ALT_TC=resident selects the kernel that keeps the fmap1 block in tensor memory (A/B switch, D <= 256;
    // measured slower: 3.2 us per 128-point tile vs 4.4 us per 256-point tile, see profiles/README.md)
    const char* ve = std::getenv("RAFTCORR_ALT_TC");
    const bool resident = !f16 && C <= 256 && ve && ve[0] == 'r';
    for (int l = 0; l < 4; ++l) {
        const int ll = l < lv.L ? l : 0;
        cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)lv.w[ll], (cuuint64_t)lv.h[ll], (cuuint64_t)B};
        cuuint64_t strides[3] = {(cuuint64_t)C * esz, (cuuint64_t)lv.w[ll] * C * esz, (cuuint64_t)lv.h[ll] * lv.w[ll] * C * esz};
        cuuint32_t box[4] = {(cuuint32_t)k_elems, (cuuint32_t)(resident ? kHalfW : kTgtW), (cuuint32_t)(resident ? kHalfH : kTgtH), 1};
        if (int rc = encode_map(fn, &maps.b[l], dt, (void*)lv.f2[ll], 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, "fmap2 level (NHWC)"))
            return rc;
        cuuint32_t box_sq[4] = {(cuuint32_t)k_elems, kSqW, kSqH, 1};
        if (int rc = encode_map(fn, &maps.bsq[l], dt, (void*)lv.f2[ll], 4, dims, strides, box_sq, CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, "fmap2 level (NHWC), square tiles"))
            return rc;
    }
    AltTcParams P{};
    P.coords = cv.ptr; P.cb = cv.batch_stride; P.cc = cv.comp_stride; P.cp = cv.pix_stride;
    P.bbox = reinterpret_cast<int4*>(workspace);
    P.out = out;
    P.B = B; P.H = H; P.W = W; P.L = lv.L; P.k_blocks = C / k_elems;
    P.f16 = f16 ? 1 : 0; P.k_elems = k_elems; P.inv = inv;
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
    { const char* e = std::getenv("RAFTCORR_ALT_SQUARE"); P.allow_square = !resident && !(e && e[0] == '0'); }
    P.bb_w = resident ? kHalfW : kTgtW;
    P.bb_h = resident ? kHalfH : kTgtH;
    P.f1 = (const float*)f1_nhwc;
    P.C = C;
    int sms = 0;
    RC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    switch (radius) {
        case 1: return run_alt_fwd_tc<1>(maps, P, sms, resident, s);
        case 2: return run_alt_fwd_tc<2>(maps, P, sms, resident, s);
        case 3: return run_alt_fwd_tc<3>(maps, P, sms, resident, s);
        default: return run_alt_fwd_tc<4>(maps, P, sms, resident, s);
    }
}

}  // namespace raftcorr

// On-the-fly correlation on the tensor cores (AlternateCorrBlock.__call__, raft/core/corr.py:142-167;
// the arithmetic of raft/alt_cuda_corr/correlation_kernel.cu:18-119 for all levels at once).
//
// The reference (and altcorr.cu) evaluates, per source pixel and level, the (K+1)^2 lattice of dot products
// it needs -- 102 k FMAs per pixel at D=256, r=4, L=4, on the CUDA cores.  Neighbouring source pixels look at
// overlapping windows, so here a CTA takes a block of 8x16 source pixels, finds the bounding box of all
// their windows on a level (alt_bbox_kernel), and computes the dot products of the 128 pixels with EVERY
// lattice point of that box as 128 x N x D GEMMs (fp16 or TF32 operands, N <= 256 lattice points per tile, shapes
// below) on tcgen05: fmap1 block and
// fmap2 tile arrive by TMA as 128B-swizzled K-major operands straight from the NHWC feature tensors (4-D
// boxes; rows/columns outside the level are zero-filled by the TMA unit = grid_sample's zero padding),
// accumulators live in TMEM (two 256-column stages).  For smooth flow the box is (16+K+1) x (8+K+1) points,
// i.e. 2-3x more dot products than strictly needed, at ~30x the arithmetic rate.  Incoherent coordinates
// only cost time (the box is clipped to the level, so the worst case is the all-pairs product), never
// correctness.
//
// Epilogue (round 2: window extraction; TMEM lane == source pixel).  Round 1 had 8 epilogue warps that copied each
// accumulator row into a per-pixel shared-memory row, re-read the K+1 points of the window, and added the row's share
// to K*K shared-memory accumulators (read-modify-write): ~5.5 k warp instructions in a dependent chain per lattice
// tile, 7.3 k cycles per tile against 2 k cycles of MMA (ncu: tensor pipe 33 %).  Now the work is split in two:
//   * STAGER warps (two per TMEM lane quarter, warps 4-7 and 16-19) read the accumulator chunks (32 columns) that some
//     window of their 32 pixels touches (tcgen05.ld.x32) and each lane copies ONLY the points of ITS OWN window -- a
//     32-bit lane mask, one predicated st.shared per column with a static register and a static address offset -- into
//     the pixel's (K+1) x (K+1) window image in shared memory.  No read-back, no accumulation: the tiles of a level
//     fill disjoint parts of the window image.
//   * COMPUTE warps (warps 8-15: one thread per pixel and half of the output rows) wait until the level is staged, read
//     the window image with static offsets, form the bilinear taps in registers and write the K*K outputs of the level
//     straight to global memory.  Points outside the level were never staged: border pixels take a predicated path.
// Stagers and compute warps meet on one mbarrier pair per (source block, level); the accumulator stage goes back to
// the MMA warp as soon as the stagers have copied the window points out of it.
#include <climits>
#include <cstdlib>

#include "tc_util.cuh"

namespace raftcorr {

namespace {

constexpr int kSrcH = 8, kSrcW = 16;   // source block: 128 pixels = TMEM lanes, m = sy * 16 + sx
constexpr int kTgtH = 8, kTgtW = 32;   // largest lattice tile: 256 points = UMMA N (8 rows x 32 columns, n = ty * 32 + tx)
// Lattice tile shapes: 32, 16 or 8 points wide (so that a 32-column accumulator chunk is 1, 2 or 4 whole tile rows) and
// th = 4 .. 256 / width rows high in steps of 4 (the fmap2 tile arrives as th / 4 TMA boxes of [width][4 rows]): UMMA
// N = width * th, a multiple of 32.  alt_bbox_kernel picks, per (source block, level), the width and height that cover
// the box with the fewest operand bytes (fmap1 block + lattice points); bbox.z = tiles in x | width code << 16 | th << 20.
// Default: widths 32 and 16 (8-point-wide tiles cut the redundancy further but not the time: profiles/README.md).
constexpr int kBoxRows = 4;
__host__ __device__ constexpr int shape_w(int z) { return 32 >> ((z >> 16) & 3); }
__host__ __device__ constexpr int shape_h(int z) { return (z >> 20) & 63; }
constexpr int kTileM = kSrcH * kSrcW;
constexpr int kTileN = kTgtH * kTgtW;
constexpr int kBlockK = 32, kUmmaK = 8;
// operand ring: 3 stages (measured at 1080p, r = 3 / r = 4: 4 stages 170 / 197 us, 3 stages 170 / 198, 2 stages 197 / 224 --
// the ring is not what bounds the kernel)
__host__ __device__ constexpr int ring_stages(int) { return 3; }
constexpr int kABytes = kTileM * kBlockK * 4;  // 16 KB
constexpr int kBBytes = kTileN * kBlockK * 4;  // 32 KB
constexpr int kStageBytes = kABytes + kBBytes;
constexpr int kStagersPerQuarter = 2;         // stager warps per TMEM lane quarter (they split the chunks of a tile)
constexpr int kStagerWarps = 4 * kStagersPerQuarter, kComputeWarps = 8;
constexpr int kThreads = 32 * 20;              // warp 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4-7 + 16-19 stagers, 8-15 compute
// Window image of one pixel: K+1 rows of K+2 floats, pixel pitch rounded up to 3 (mod 4).  Lane = pixel in both roles; the
// 32 lanes of a TMEM quarter are 2 source rows (sy) of 16 pixels (sx).  The stagers' store of tile column c goes to window
// column c - c0 with c0 ~ c00 + sx and to window row wj ~ wj0 - sy, i.e. to bank (pitch - 1) * sx + (16 * pitch - row pitch)
// * sy + const: pitch - 1 = 2 * odd keeps the 16 sx apart, the odd row pitch moves the second source row to the odd banks
// (with K+1 floats per row and pitch (K+1)^2 + 3 ncu counted 35 % of the shared-memory wavefronts as conflicts: lane
// (sx, 1) on lane (sx + 1, 0)).  The compute warps read one pixel per lane at an odd pitch: conflict-free.
__host__ __device__ constexpr int win_row(int R) { return 2 * R + 3; }
__host__ __device__ constexpr int win_pitch(int R) { return (((2 * R + 2) * win_row(R)) & ~3) + 3; }
constexpr uint32_t kTmemCols = 512;

__device__ __forceinline__ float clamp_coord(float v) { return fminf(fmaxf(v, -1048576.f), 1048576.f); }

struct AltTcMaps {
    CUtensorMap a;
    CUtensorMap b[3][4];  // [width code: 32 / 16 / 8 points][level]: boxes of [width][4 rows] lattice points x one k-block
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
    int allow_square;  // number of lattice tile widths to choose from: 1 = 32 points only, 2 = + 16, 3 = + 8
    int allow_short;   // tile heights fitted to the box (N < 256) instead of always 256 points per tile
    // fp16 operands with one power-of-two scale per image and tensor (launch_alt_f16_operands): inv[b] for fmap1,
    // inv[(1 + l) * B + b] for level l of the fmap2 pyramid; k_elems = operand elements per 128-byte k-block
    int f16, k_elems;
    const float* inv;
    int stages;  // depth of the operand ring
};

__host__ __device__ constexpr uint32_t idesc_f16(int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24); }
__device__ __forceinline__ void tc_mma_f16_alt(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

constexpr int smem_bytes(int R, int stages) {
    return stages * kStageBytes + win_pitch(R) * kTileM * 4 + 1024 + 256;
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
                const int bw = x1 - x0 + 1, bh = y1 - y0 + 1;
                int best = INT_MAX;
                for (int wc = 0; wc < P.allow_square; ++wc) {
                    const int tw = 32 >> wc, th_max = kTileN / tw;
                    const int nx = (bw + tw - 1) / tw, ny = (bh + th_max - 1) / th_max;
                    // rows spread evenly over the ny tiles, rounded up to whole TMA boxes
                    const int th = P.allow_short ? min(th_max, (((bh + ny - 1) / ny) + kBoxRows - 1) / kBoxRows * kBoxRows) : th_max;
                    // operand bytes per tile and k-block: 128 fmap1 rows + tw * th fmap2 rows of 128 B; + a constant for
                    // the per-tile hand-overs, so that ties go to fewer, larger tiles
                    const int cost = nx * ny * (kTileM + 32 + tw * th);
                    if (cost < best) {
                        best = cost;
                        bb = make_int4(x0, y0, nx | (wc << 16) | (th << 20), ny);
                    }
                }
            }
            P.bbox[(int64_t)t * P.L + l] = bb;
        }
        __syncthreads();
    }
}

// window hand-over between stagers and compute warps, one (source block, level) at a time
struct AltWinBars {
    uint32_t full, empty;  // shared-memory addresses: full = kStagerWarps arrivals, empty = kComputeWarps arrivals
};

// One lattice tile: copy the window points of this lane's pixel out of the accumulator chunks q32 = sub, sub + S, ...
// c0 = first window column relative to the tile, jr = window row of the tile's first row.
template <int R, int TW>
__device__ __forceinline__ void alt_stage_tile(uint32_t t_acc, int sub, int nchunks, uint32_t win_s, int c0, int jr, bool valid) {
    constexpr int K = 2 * R + 1, KP = K + 1, WR = win_row(R);
    constexpr int RPL = 32 / TW;  // rows of the tile per 32-column chunk
    uint32_t cm = 0;  // bit c: tile column c belongs to this lane's window
    if (valid && c0 + K >= 0 && c0 < TW) {
        const uint32_t w = (1u << KP) - 1u;
        cm = c0 >= 0 ? (w << c0) : (w >> (-c0));
        if (TW < 32) cm &= (1u << (TW & 31)) - 1u;
    }
#pragma unroll 1
    for (int q32 = sub; q32 < nchunks; q32 += kStagersPerQuarter) {  // a tile of N < 256 points ends early (later columns are stale)
        const int wj = jr + q32 * RPL;  // window row of the chunk's first tile row
        uint32_t mask = 0u;
#pragma unroll
        for (int r = 0; r < RPL; ++r) mask |= (unsigned)(wj + r) <= (unsigned)K ? (cm << ((r * TW) & 31)) : 0u;
        if (!__any_sync(0xffffffffu, mask != 0u)) continue;
        float v[32];
        tmem_ld32(t_acc + q32 * 32, v);
        tmem_ld_wait();
        // element cc of the chunk = tile row q32 * RPL + cc / TW, column cc % TW -> window (wj + cc / TW, cc % TW - c0)
        const uint32_t base = win_s + (uint32_t)((wj * WR - c0) * 4);
#pragma unroll
        for (int cc = 0; cc < 32; ++cc) {
            const int off = (cc / TW) * WR + (cc % TW);
            if (mask & (1u << cc))
                asm volatile("st.shared.f32 [%0], %1;" ::"r"(base + 4u * (uint32_t)off), "f"(v[cc]) : "memory");
        }
    }
}

template <int R>
__device__ __forceinline__ void alt_tc_stager(const AltTcParams& P, int ew, int sub, int lane, uint32_t tmem_base,
                                              uint32_t tfull_bar, uint32_t tempty_bar, uint32_t win_base, AltWinBars wb) {
    const int m = ew * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(ew * 32) << 16;
    const uint32_t win_s = win_base + (uint32_t)(m * win_pitch(R) * 4);
    const int tiles_per_b = P.tiles_x * P.tiles_y;
    int it = 0;
    uint32_t n = 0;  // (source block, level) steps so far
    for (int t = blockIdx.x; t < P.total_tiles; t += gridDim.x) {
        const int tx = t % P.tiles_x, ty = (t / P.tiles_x) % P.tiles_y, b = t / tiles_per_b;
        const int y = ty * kSrcH + (m >> 4), x = tx * kSrcW + (m & 15);
        const bool valid = y < P.H && x < P.W;
        const int64_t pix = (int64_t)y * P.W + x;
        const float cx = clamp_coord(valid ? __ldg(P.coords + b * P.cb + pix * P.cp) : 0.f);
        const float cy = clamp_coord(valid ? __ldg(P.coords + b * P.cb + P.cc + pix * P.cp) : 0.f);
        for (int l = 0; l < P.L; ++l, ++n) {
            const int4 bb = P.bbox[(int64_t)t * P.L + l];
            const int x0 = __float2int_rd(cx * P.inv_scale[l]) - R, y0 = __float2int_rd(cy * P.inv_scale[l]) - R;
            const int ntx = bb.z & 0xffff, tw = shape_w(bb.z), th = shape_h(bb.z), nchunks = th * tw / 32;
            mbar_wait(wb.empty, (n & 1u) ^ 1u);  // the compute warps have read the previous level's window images
            for (int c = 0; c < bb.w; ++c)
                for (int a = 0; a < ntx; ++a, ++it) {
                    const int as = it & 1;
                    mbar_wait(tfull_bar + 8 * as, (it >> 1) & 1);
                    tc_fence_after();
                    const uint32_t t_acc = tmem_base + lane_addr + as * kTileN;
                    const int c0 = x0 - (bb.x + a * tw), jr = (bb.y + c * th) - y0;
                    if (tw == 32)
                        alt_stage_tile<R, 32>(t_acc, sub, nchunks, win_s, c0, jr, valid);
                    else if (tw == 16)
                        alt_stage_tile<R, 16>(t_acc, sub, nchunks, win_s, c0, jr, valid);
                    else
                        alt_stage_tile<R, 8>(t_acc, sub, nchunks, win_s, c0, jr, valid);
                    tc_fence_before();
                    mbar_arrive(tempty_bar + 8 * as);  // every tcgen05.ld of this thread on the stage has completed
                }
            __syncwarp();
            if (lane == 0) mbar_arrive(wb.full);
        }
    }
}

// Output rows [J0, J0 + JN) of one pixel's level: lattice rows J0 .. J0 + JN of its window image.
template <int R, int J0, int JN, bool CHECK>
__device__ __forceinline__ void alt_window_rows(const float* wp, float ax, float ay, float osc, int x0, int y0, int lw, int lh,
                                                float* dst, int64_t plane) {
    constexpr int K = 2 * R + 1, KP = K + 1, WR = win_row(R);
    const float wt = 1.f - ay;
    float hp[K];
#pragma unroll
    for (int jj = 0; jj <= JN; ++jj) {
        const int j = J0 + jj;
        float v[KP];
#pragma unroll
        for (int c = 0; c < KP; ++c) v[c] = wp[j * WR + c];
        if (CHECK) {  // points outside the level are zero (grid_sample zero padding) and were never staged
            const bool row_ok = (unsigned)(y0 + j) < (unsigned)lh;
#pragma unroll
            for (int c = 0; c < KP; ++c) v[c] = (row_ok && (unsigned)(x0 + c) < (unsigned)lw) ? v[c] : 0.f;
        }
        float hl[K];
#pragma unroll
        for (int i = 0; i < K; ++i) hl[i] = v[i] + ax * (v[i + 1] - v[i]);
        if (jj > 0) {
#pragma unroll
            for (int i = 0; i < K; ++i) dst[(int64_t)(i * K + j - 1) * plane] = (wt * hp[i] + ay * hl[i]) * osc;
        }
#pragma unroll
        for (int i = 0; i < K; ++i) hp[i] = hl[i];
    }
}

template <int R, int J0, int JN>
__device__ __forceinline__ void alt_tc_compute(const AltTcParams& P, int ew, int lane, const float* win, AltWinBars wb) {
    constexpr int K = 2 * R + 1;
    const int m = ew * 32 + lane;
    const float* wp = win + m * win_pitch(R);
    const int tiles_per_b = P.tiles_x * P.tiles_y;
    const int64_t plane = (int64_t)P.H * P.W;
    uint32_t n = 0;
    for (int t = blockIdx.x; t < P.total_tiles; t += gridDim.x) {
        const int tx = t % P.tiles_x, ty = (t / P.tiles_x) % P.tiles_y, b = t / tiles_per_b;
        const int y = ty * kSrcH + (m >> 4), x = tx * kSrcW + (m & 15);
        const bool valid = y < P.H && x < P.W;
        const int64_t pix = (int64_t)y * P.W + x;
        const float cx = clamp_coord(valid ? __ldg(P.coords + b * P.cb + pix * P.cp) : 0.f);
        const float cy = clamp_coord(valid ? __ldg(P.coords + b * P.cb + P.cc + pix * P.cp) : 0.f);
        for (int l = 0; l < P.L; ++l, ++n) {
            const float xs = cx * P.inv_scale[l], ys = cy * P.inv_scale[l];
            const float xf = floorf(xs), yf = floorf(ys);
            const float ax = xs - xf, ay = ys - yf;
            const int x0 = (int)xf - R, y0 = (int)yf - R;
            const int lw = P.lvl_w[l], lh = P.lvl_h[l];
            const float osc = P.f16 ? P.scale * __ldg(P.inv + b) * __ldg(P.inv + (1 + l) * P.B + b) : P.scale;
            mbar_wait(wb.full, n & 1u);  // every stager warp has finished the level
            if (valid) {
                float* dst = P.out + (((int64_t)b * P.L + l) * K * K * P.H + y) * P.W + x;
                const bool inside = x0 >= 0 && x0 + K < lw && y0 >= 0 && y0 + K < lh;
                if (inside)
                    alt_window_rows<R, J0, JN, false>(wp, ax, ay, osc, x0, y0, lw, lh, dst, plane);
                else
                    alt_window_rows<R, J0, JN, true>(wp, ax, ay, osc, x0, y0, lw, lh, dst, plane);
            }
            __syncwarp();  // every lane has read its window image
            if (lane == 0) mbar_arrive(wb.empty);
        }
    }
}

// ---- pass 2: GEMM tiles + gather -------------------------------------------------------------------------
template <int R>
__global__ void __launch_bounds__(kThreads, 1)
alt_fwd_tc_kernel(const __grid_constant__ AltTcMaps maps, const AltTcParams P) {
    constexpr int K = 2 * R + 1;
    const int kStages = P.stages;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // 128B-swizzle atoms want 1024-byte alignment
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t stage_base = smem_base;
    float* win = reinterpret_cast<float*>(smem_gen + kStages * kStageBytes);  // [128 pixels][win_pitch]: window images
    const uint32_t win_base = smem_base + kStages * kStageBytes;
    const uint32_t bar_base = win_base + ((win_pitch(R) * kTileM * 4 + 15) & ~15);
    const uint32_t full_bar = bar_base, empty_bar = bar_base + 8 * kStages;
    const uint32_t tfull_bar = bar_base + 16 * kStages, tempty_bar = tfull_bar + 16, tmem_slot = tempty_bar + 16;
    const uint32_t wfull_bar = tmem_slot + 16, wempty_bar = wfull_bar + 8;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&maps.a);
        for (int l = 0; l < P.L; ++l) {
            for (int wc = 0; wc < 3; ++wc) prefetch_tensormap(&maps.b[wc][l]);
        }
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(full_bar + 8 * s, 1);
            mbar_init(empty_bar + 8 * s, 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull_bar + 8 * a, 1);
            mbar_init(tempty_bar + 8 * a, 32 * kStagerWarps);
        }
        mbar_init(wfull_bar, kStagerWarps);    // lane 0 of every stager warp, once per (source block, level)
        mbar_init(wempty_bar, kComputeWarps);  // lane 0 of every compute warp
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
                    const int ntx = bb.z & 0xffff, tw = shape_w(bb.z), th = shape_h(bb.z);
                    const CUtensorMap* mb = &maps.b[(bb.z >> 16) & 3][l];
                    const uint32_t stage_tx = kABytes + (uint32_t)(tw * th) * 128u;  // fmap1 block + the tile's lattice points x 128 B
                    const uint32_t box_bytes = (uint32_t)(tw * kBoxRows) * 128u;
                    for (int c = 0; c < bb.w; ++c)
                        for (int a = 0; a < ntx; ++a)
                            for (int kb = 0; kb < P.k_blocks; ++kb) {
                                mbar_wait(empty_bar + 8 * stage, phase ^ 1);
                                const uint32_t sa = stage_base + stage * kStageBytes;
                                mbar_expect_tx(full_bar + 8 * stage, stage_tx);
                                tma_load_4d(sa, &maps.a, full_bar + 8 * stage, kb * P.k_elems, tx * kSrcW, ty * kSrcH, b);
                                for (int r4 = 0; r4 * kBoxRows < th; ++r4)  // the tile's rows, one box of 4 at a time
                                    tma_load_4d(sa + kABytes + r4 * box_bytes, mb, full_bar + 8 * stage, kb * P.k_elems, bb.x + a * tw,
                                                bb.y + c * th + r4 * kBoxRows, b);
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
                    const int tile_n = shape_w(bb.z) * shape_h(bb.z);  // UMMA N = the tile's lattice points (32 .. 256)
                    const uint32_t idesc = P.f16 ? idesc_f16(tile_n) : make_idesc_tf32(kTileM, tile_n, false, false);
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
                                    tc_mma_f16_alt(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
                                else
                                    tc_mma_tf32(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
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
        const int ew = warp & 3;  // TMEM lane quarter
        const AltWinBars wb{wfull_bar, wempty_bar};
        constexpr int J0N = (K + 1) / 2;  // output rows of the first compute half
        if (warp < 8)
            alt_tc_stager<R>(P, ew, 0, lane, tmem_base, tfull_bar, tempty_bar, win_base, wb);
        else if (warp >= 16)
            alt_tc_stager<R>(P, ew, 1, lane, tmem_base, tfull_bar, tempty_bar, win_base, wb);
        else if (warp < 12)
            alt_tc_compute<R, 0, J0N>(P, ew, lane, win, wb);
        else
            alt_tc_compute<R, J0N, K - J0N>(P, ew, lane, win, wb);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

template <int R>
int run_alt_fwd_tc(const AltTcMaps& maps, AltTcParams P, int sms, cudaStream_t s) {
    alt_bbox_kernel<R><<<P.total_tiles, kTileM, 0, s>>>(P);
    RC_CUDA(cudaGetLastError());
    static_assert(smem_bytes(R, ring_stages(R)) <= 232448, "on-the-fly tensor-core kernel does not fit shared memory");
    P.stages = ring_stages(R);
    if (const char* e = std::getenv("RAFTCORR_ALT_STAGES")) {  // A/B: a shallower ring
        const int v = std::atoi(e);
        if (v >= 2 && v < P.stages) P.stages = v;
    }
    const int smem = smem_bytes(R, P.stages);
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
    for (int l = 0; l < 4; ++l) {
        const int ll = l < lv.L ? l : 0;
        cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)lv.w[ll], (cuuint64_t)lv.h[ll], (cuuint64_t)B};
        cuuint64_t strides[3] = {(cuuint64_t)C * esz, (cuuint64_t)lv.w[ll] * C * esz, (cuuint64_t)lv.h[ll] * lv.w[ll] * C * esz};
        for (int wc = 0; wc < 3; ++wc) {
            cuuint32_t box[4] = {(cuuint32_t)k_elems, (cuuint32_t)(32 >> wc), kBoxRows, 1};
            if (int rc = encode_map(fn, &maps.b[wc][l], dt, (void*)lv.f2[ll], 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B,
                                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, "fmap2 level (NHWC)"))
                return rc;
        }
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
    { const char* e = std::getenv("RAFTCORR_ALT_SQUARE"); P.allow_square = (e && e[0] == '0') ? 1 : (e && e[0] == '3') ? 3 : 2; }
    { const char* e = std::getenv("RAFTCORR_ALT_SHORT"); P.allow_short = !(e && e[0] == '0'); }
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

// Correlation lookup fused with the motion encoder's first 1x1 convolution (+ bias, ReLU)  -- SURVEY.md 8(f) rank 1.
//
//   out[b, co, p] = relu( bias[co] + sum_ch W[co, ch] * lookup(coords)[b, ch, p] )
//   (raft/core/corr.py:45-91 followed by raft/core/update.py:116,135 (small) / :158,170 (basic): the lookup result has
//   no other reader, so its [B, L*K^2, H, W] tensor never has to exist in HBM.)
//
// One persistent CTA per SM, tile = 128 source pixels, pipelined by pyramid LEVEL (the level's K^2 channels are one
// K-slice of the product):
//   warps 0-15  gather : the TMA lookup of lookup_tma.cu (one box per (pixel, level), zero padding by the TMA unit,
//                        lane = (x offset, group of vertically adjacent outputs)); a warp owns 8 pixels of the tile and
//                        a ring of two 4-box batches (issued by four lanes at once) that runs across levels AND tiles.  The K^2 bilinear samples of a
//                        pixel are rounded to TF32 and written as row `pixel` of the level's A operand tile
//                        ([128 px][<= 96 ch], K-major, 128B swizzle) instead of an output tensor; two such tiles, so the
//                        gather of level l+1 overlaps the MMAs of level l.
//   warp 16     MMA    : tcgen05.mma kind::tf32, M = 128 pixels, N = Cout (two halves of <= 128 when Cout > 128), K = 8;
//                        fp32 accumulators in TMEM, two accumulator stages (epilogue of tile i overlaps tile i+1).
//   warp 17     weights: TMA loads of the packed weight k-blocks ([<= 128 co][32 ch], 16 KB stages, 3-deep ring; every
//                        CTA streams the whole matrix once per tile, all L2 hits).
//   warps 18-21 epilogue: TMEM lane = pixel: + bias, ReLU, 128-byte coalesced stores per output channel.
#include <cstdlib>
#include <cstring>

#include "lookup_tma_common.cuh"
#include "tc_util.cuh"

namespace raftcorr {

namespace {

constexpr int kGatherWarps = 16;
constexpr int kConvTileM = 128;                          // source pixels per tile = TMEM lanes
constexpr int kPixPerWarp = kConvTileM / kGatherWarps;   // 8
constexpr int kBatch = 4;                                // patch boxes issued together (by kBatch lanes)
constexpr int kSlots = 2 * kBatch;                       // ring: the batch being evaluated + the batch in flight
constexpr int kWStages = 3;
constexpr int kWStageBytes = 128 * 128;                  // [<= 128 co][32 ch] tf32
constexpr int kKBlockBytes = kConvTileM * 128;           // one 32-channel k-block of the A tile: 16 KB
constexpr int kConvThreads = (kGatherWarps + 2 + 4) * 32;  // 704
constexpr int kMmaWarp = kGatherWarps, kWeightWarp = kGatherWarps + 1, kEpiWarp0 = kGatherWarps + 2;
constexpr uint32_t kConvTmemCols = 512;

template <int R>
struct ConvGeo {
    static constexpr int K = 2 * R + 1, KK = K * K;
    static constexpr int KB = (KK + 31) / 32;      // 128-byte k-blocks per level
    static constexpr int KSTEPS = (KK + 7) / 8;    // K = 8 MMA steps per level (channels beyond K^2 are zero on both sides)
    static constexpr int A_BYTES = KB * kKBlockBytes;
};

struct ConvParams {
    const float* coords;
    const float* bias;  // may be null
    float* out;
    int B, N, tiles_per_img, total_tiles;
    int cout, nh, halves;  // N of one MMA = nh = cout / halves
    int relu;
    float inv_scale[4];
};

__device__ __forceinline__ float to_tf32(float x) {  // round to nearest (ties away), finite inputs: what cvt.rna does
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}
__device__ __forceinline__ float to_tf32_exact(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}

// Waits of the hot loops: no clock reads inside the loop (the shared mbar_wait spends six instructions per poll on
// its time-out), a poll counter instead; the suspend-time hint lets one poll sleep in hardware for up to that long.
__device__ __forceinline__ void mbar_wait_spin(uint32_t bar, uint32_t parity) {
    uint32_t ok, polls = 0;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity), "r"(2000u)
            : "memory");
        if (!ok && ++polls > (1u << 26)) __trap();  // protocol bug -> trapped launch instead of a hang
    } while (!ok);
}

// Long waits (accumulator / operand tile hand-overs): sleep between polls so that waiting warps do not take issue
// slots from the gather warps.
__device__ __forceinline__ void mbar_wait_sleep(uint32_t bar, uint32_t parity, uint32_t ns) {
    uint32_t polls = 0;
    while (!mbar_try_wait(bar, parity)) {
        __nanosleep(ns);
        if (++polls > (1u << 24)) __trap();
    }
}

// patch slot bytes per pyramid layout code (raftcorr_pyramid_layout::interleaved): layout 2 mixes quad and pair boxes
template <int R, int LAY>
struct ConvSlot {
    static constexpr int value = LAY == 2 ? (QuadGeo<R>::SLOT > TmaGeo<R, true>::SLOT ? QuadGeo<R>::SLOT : TmaGeo<R, true>::SLOT)
                                          : TmaGeo<R, LAY == 1>::SLOT;
};

// LAY = raftcorr_pyramid_layout::interleaved: 0 plain rows, 1 row pairs, 2 quads for levels 0-1 + pairs deeper
template <int R, int LG, int LAY>
__global__ void __launch_bounds__(kConvThreads, 1)
lookup_conv_kernel(const __grid_constant__ LookupMaps maps, const __grid_constant__ CUtensorMap map_w, const ConvParams P) {
    constexpr bool IL = LAY == 1;
    using G = TmaGeo<R, LAY != 0>;   // LAY 2: the pair geometry of its deeper levels
    using Q = QuadGeo<R>;
    using C = ConvGeo<R>;
    constexpr int SLOT = ConvSlot<R, LAY>::value;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
    // [2 A tiles] [kWStages weight stages] [gather warps x kSlots patch boxes] [barriers]
    const uint32_t a_base = base;
    const uint32_t w_base = a_base + 2 * C::A_BYTES;
    const uint32_t patch_base = w_base + kWStages * kWStageBytes;
    const uint32_t bar_base = patch_base + kGatherWarps * kSlots * SLOT;
    const uint32_t patch_bar = bar_base;                                  // [kGatherWarps][kSlots]
    const uint32_t a_full = patch_bar + 8 * kGatherWarps * kSlots;        // [2]
    const uint32_t a_empty = a_full + 16;                                 // [2]
    const uint32_t w_full = a_empty + 16;                                 // [kWStages]
    const uint32_t w_empty = w_full + 8 * kWStages;                       // [kWStages]
    const uint32_t tfull = w_empty + 8 * kWStages;                        // [2]
    const uint32_t tempty = tfull + 16;                                   // [2]
    const uint32_t tmem_slot = tempty + 16;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int first = blockIdx.x, stride = gridDim.x;
    const int my_tiles = first < P.total_tiles ? (P.total_tiles - first + stride - 1) / stride : 0;

    if (warp < kGatherWarps) {
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < kSlots; ++i) mbar_init(patch_bar + 8 * (warp * kSlots + i), 1);
        }
    } else if (warp == kMmaWarp) {
        if (lane == 0) {
            for (int i = 0; i < 2; ++i) {
                mbar_init(a_full + 8 * i, kGatherWarps);
                mbar_init(a_empty + 8 * i, 1);
                mbar_init(tfull + 8 * i, 1);
                mbar_init(tempty + 8 * i, 128);
            }
            for (int i = 0; i < kWStages; ++i) {
                mbar_init(w_full + 8 * i, 1);
                mbar_init(w_empty + 8 * i, 1);
            }
        }
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(kConvTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else if (warp == kWeightWarp && lane == 0) {
        prefetch_tensormap(&map_w);
#pragma unroll
        for (int l = 0; l < LG; ++l) {
            prefetch_tensormap(&maps.m[l]);
            if (LAY) prefetch_tensormap(&maps.ms[l]);
        }
    }
    // the channels a level does not have (K^2 .. 32 * KB) must read as zero, not as stale shared memory
    {
        float4* a4 = reinterpret_cast<float4*>(gen);
        for (int i = threadIdx.x; i < 2 * C::A_BYTES / 16; i += kConvThreads) a4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    fence_mbar_init();
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - base));
    // Programmatic dependent launch: everything above (96 KB of zero fill, barrier init, TMEM allocation) overlaps the
    // tail of the previous kernel in the stream; its results (the coordinates) are visible from here on.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    if (warp < kGatherWarps) {
        // ===================== gather: patches -> bilinear samples -> A tile rows =====================
        const uint32_t my_patch = patch_base + warp * (kSlots * SLOT);
        const float* my_patch_gen = reinterpret_cast<const float*>(gen + (my_patch - base));
        const uint32_t my_bar = patch_bar + 8 * warp * kSlots;
        const int jg = lane / G::K, li = lane - jg * G::K;
        const int j0 = jg * G::RPL;
        const bool lane_on = jg < G::GROUPS;
        const int lane_off = j0 * G::BC + li;
        const int lane_ch = li * G::K + j0;  // channel inside the level: x offset major (corr.py:66-78)
        int il_row[2][G::RPL + 1];
        if (IL) {
#pragma unroll
            for (int sy = 0; sy < 2; ++sy)
#pragma unroll
                for (int rr = 0; rr <= G::RPL; ++rr) {
                    const int Y = sy + j0 + rr;
                    il_row[sy][rr] = (Y >> 1) * G::BC + (Y & 1);
                }
        }
        // channel part of the A tile address: k-block, 16-byte chunk inside the 128-byte row, word inside the chunk
        uint32_t ch_off[G::RPL], ch_chunk[G::RPL];
#pragma unroll
        for (int k = 0; k < G::RPL; ++k) {
            const int ch = lane_ch + k;
            ch_off[k] = (uint32_t)(ch >> 5) * kKBlockBytes + (uint32_t)(ch & 3) * 4;
            ch_chunk[k] = (uint32_t)(ch & 31) >> 2;
        }

        // layout 2 evaluation (lookup_fwd_quad_kernel's scheme): half-warp = pixel of a pair, lane = window column;
        // A tile address parts of the K outputs of this lane's column (channel col * K + j, x offset major)
        const int half = lane >> 4, col = lane & 15;
        uint32_t q_off[LAY == 2 ? G::K : 1], q_chunk[LAY == 2 ? G::K : 1];
        if (LAY == 2) {
#pragma unroll
            for (int jj = 0; jj < G::K; ++jj) {
                const int ch = (col < G::K ? col : 0) * G::K + jj;
                q_off[jj] = (uint32_t)(ch >> 5) * kKBlockBytes + (uint32_t)(ch & 3) * 4;
                q_chunk[jj] = (uint32_t)(ch & 31) >> 2;
            }
        }

        // coordinates: lanes 0..7 hold the warp's pixels of the tile being consumed, lanes 16..23 those of the next tile
        auto tile_origin = [&](int tile_iter, int& b, int& p_base) {  // batch item and first pixel of this warp's rows
            const int t = first + tile_iter * stride;
            b = t / P.tiles_per_img;
            p_base = (t - b * P.tiles_per_img) * kConvTileM + warp * kPixPerWarp;
        };
        auto load_coords = [&](int tile_iter, float& x, float& y) {
            x = y = 0.f;
            if (tile_iter < my_tiles && (lane & 15) < kPixPerWarp) {
                int b, p_base;
                tile_origin(tile_iter, b, p_base);
                const int p = p_base + (lane & 15);
                if (p < P.N) {
                    x = clamp_coord(__ldg(P.coords + ((int64_t)b * 2 + 0) * P.N + p));
                    y = clamp_coord(__ldg(P.coords + ((int64_t)b * 2 + 1) * P.N + p));
                }
            }
        };
        float cx, cy;
        load_coords(lane < 16 ? 0 : 1, cx, cy);

        // Patch boxes are issued in BATCHES of kBatch consecutive pixels of one level, by kBatch lanes at once (the
        // address arithmetic runs once for the batch; only the TMA instruction itself is serialised per lane), one
        // batch ahead of the batch being evaluated: the ring holds two batches.
        const char* map_base = reinterpret_cast<const char*>(&maps);  // m[l] at 128 * l, ms[l] at 512 + 128 * l
        constexpr int kBatchesPerTile = LG * (kPixPerWarp / kBatch);
        const int total_batches = my_tiles * kBatchesPerTile;
        int i_n = 0, i_ti = 0, i_l = 0, i_jb = 0, i_b = 0, i_pbase = 0;  // issue cursor
        if (my_tiles > 0) tile_origin(0, i_b, i_pbase);
        auto issue_batch = [&](int c_ti) {
            if (i_n >= total_batches) return;
            const int j = i_jb + (lane & (kBatch - 1));
            const int src = j + (i_ti != c_ti ? 16 : 0);
            const float x = __shfl_sync(0xffffffffu, cx, src), y = __shfl_sync(0xffffffffu, cy, src);
            const int p = i_pbase + j;
            if (lane < kBatch && p < P.N) {
                const float inv = __int_as_float((127 - i_l) << 23);  // 2^-level, exact (the plan has one level group)
                const int x0 = __float2int_rd(x * inv) - R;
                const int y0 = __float2int_rd(y * inv) - R;
                const int slot = (i_n & 1) * kBatch + lane;
                const uint32_t bar = my_bar + 8 * slot;
                const uint32_t dst = my_patch + slot * SLOT;
                fence_proxy_async();  // the slot was read through the generic proxy two batches ago
                if (LAY == 2) {
                    // row quads: exact box of K+1 columns at any x, GMAX or GMAX-1 quads
                    const int nq = Q::quads(y0 & 3);
                    mbar_expect_tx(bar, nq * Q::ROW_BYTES);
                    tma_load_3d(dst, reinterpret_cast<const CUtensorMap*>(map_base + (nq == Q::GMAX ? 0 : 512) + 128 * i_l), bar,
                                4 * x0, y0 >> 2, i_b * P.N + p);
                } else if (LAY) {
                    const bool tall = y0 & 1;  // a window starting on an odd row spans one more pair row
                    mbar_expect_tx(bar, tall ? G::BOX_BYTES : G::BOX_BYTES - G::BC * 4);
                    tma_load_3d(dst, reinterpret_cast<const CUtensorMap*>(map_base + (tall ? 0 : 512) + 128 * i_l), bar,
                                2 * (x0 & ~1), y0 >> 1, i_b * P.N + p);
                } else {
                    mbar_expect_tx(bar, G::BOX_BYTES);
                    tma_load_3d(dst, reinterpret_cast<const CUtensorMap*>(map_base + 128 * i_l), bar, x0 & ~3, y0,
                                i_b * P.N + p);
                }
            }
            ++i_n;
            i_jb += kBatch;
            if (i_jb == kPixPerWarp) {
                i_jb = 0;
                if (++i_l == LG) {
                    i_l = 0;
                    if (++i_ti < my_tiles) tile_origin(i_ti, i_b, i_pbase);
                }
            }
        };
        issue_batch(0);

        int c_n = 0;              // consume cursor (batch index)
        uint32_t slot_phase = 0;  // one parity bit per ring slot (pixels past the end of an image skip their slot's turn)
        const uint32_t a_row0 = (uint32_t)warp * 1024;  // rows warp * 8 + j of an A tile: one 8-row swizzle group per warp
        for (int ti = 0; ti < my_tiles; ++ti) {
            int b, p_base;
            tile_origin(ti, b, p_base);
#pragma unroll
            for (int l = 0; l < LG; ++l) {
                const int g = ti * LG + l;
                const uint32_t abuf = a_base + (uint32_t)(g & 1) * C::A_BYTES + a_row0;
                const float inv = 1.0f / (float)(1 << l);
                if (l == LG - 1) {  // the issue cursor is about to run into the next tile: fetch its coordinates
                    float nx, ny;
                    load_coords(ti + 1, nx, ny);
                    if (lane >= 16) { cx = nx; cy = ny; }
                }
                mbar_wait_sleep(a_empty + 8 * (g & 1), ((g >> 1) & 1) ^ 1, 500);  // the MMAs that read this A tile have retired
#pragma unroll 1
                for (int jb = 0; jb < kPixPerWarp; jb += kBatch, ++c_n) {
                    issue_batch(ti);
                    if constexpr (LAY == 2) {
#pragma unroll 1
                        for (int k = 0; k < kBatch; k += 2) {
                            if (p_base + jb + k >= P.N) break;                    // warp-uniform; pixels are consecutive
                            const bool two = p_base + jb + k + 1 < P.N;          // the pair's second pixel exists
                            const int slot0 = (c_n & 1) * kBatch + k;
                            mbar_wait_spin(my_bar + 8 * slot0, (slot_phase >> slot0) & 1u);
                            slot_phase ^= 1u << slot0;
                            if (two) {
                                mbar_wait_spin(my_bar + 8 * (slot0 + 1), (slot_phase >> (slot0 + 1)) & 1u);
                                slot_phase ^= 1u << (slot0 + 1);
                            }
                            const int j = jb + k + half;                         // this half-warp's pixel (A tile row warp * 8 + j)
                            const float x = __shfl_sync(0xffffffffu, cx, j) * inv, y = __shfl_sync(0xffffffffu, cy, j) * inv;
                            if (col < G::K && (half == 0 || two)) {
                                const float xf = floorf(x), yf = floorf(y);
                                const float ax = x - xf, ay = y - yf;
                                const float* pb = my_patch_gen + (slot0 + half) * (SLOT / 4);
                                float hh[4 * Q::GMAX > 2 * G::BR ? 4 * Q::GMAX : 2 * G::BR];
                                if (true) {
                                    // row quads: the lane's column and the next as 16-byte chunks (rows 4g .. 4g+3)
                                    const float4* cp = reinterpret_cast<const float4*>(pb) + col;
#pragma unroll
                                    for (int gq = 0; gq < Q::GMAX; ++gq) {
                                        const float4 a = cp[gq * Q::K1], c = cp[gq * Q::K1 + 1];
                                        hh[4 * gq + 0] = a.x + ax * (c.x - a.x);
                                        hh[4 * gq + 1] = a.y + ax * (c.y - a.y);
                                        hh[4 * gq + 2] = a.z + ax * (c.z - a.z);
                                        hh[4 * gq + 3] = a.w + ax * (c.w - a.w);
                                    }
                                    // the window starts at row sy of its first quad: rotate it to row 0 (two select stages)
                                    const int sy = ((int)yf - R) & 3;
                                    if (sy & 1) {
#pragma unroll
                                        for (int i = 0; i < Q::K1 + 2; ++i) hh[i] = hh[i + 1];
                                    }
                                    if (sy & 2) {
#pragma unroll
                                        for (int i = 0; i < Q::K1; ++i) hh[i] = hh[i + 2];
                                    }
                                } else {
                                    // row pairs: 8-byte chunks (rows 2g, 2g+1); the box starts at the even column <= x0
                                    const float2* cp = reinterpret_cast<const float2*>(pb) + (((int)xf - R) & 1) + col;
#pragma unroll
                                    for (int gp = 0; gp < G::BR; ++gp) {
                                        const float2 a = cp[gp * (G::BC / 2)], c = cp[gp * (G::BC / 2) + 1];
                                        hh[2 * gp + 0] = a.x + ax * (c.x - a.x);
                                        hh[2 * gp + 1] = a.y + ax * (c.y - a.y);
                                    }
                                    if (((int)yf - R) & 1) {
#pragma unroll
                                        for (int i = 0; i < Q::K1; ++i) hh[i] = hh[i + 1];
                                    }
                                }
                                const uint32_t row = abuf + (uint32_t)j * 128;  // A tile row warp * 8 + j, swizzled by j
#pragma unroll
                                for (int jj = 0; jj < G::K; ++jj) {
                                    const float v = to_tf32(hh[jj] + ay * (hh[jj + 1] - hh[jj]));
                                    asm volatile("st.shared.f32 [%0], %1;" ::"r"(row + q_off[jj] + ((q_chunk[jj] ^ (uint32_t)j) << 4)), "f"(v)
                                                 : "memory");
                                }
                            }
                        }
                    } else {
#pragma unroll 1
                    for (int k = 0; k < kBatch; ++k) {
                        const int j = jb + k;
                        if (p_base + j < P.N) {  // warp-uniform
                            const int slot = (c_n & 1) * kBatch + k;
                            mbar_wait_spin(my_bar + 8 * slot, (slot_phase >> slot) & 1u);
                            slot_phase ^= 1u << slot;
                            const float* pbuf = my_patch_gen + slot * (G::SLOT / 4);
                            const float x = __shfl_sync(0xffffffffu, cx, j) * inv, y = __shfl_sync(0xffffffffu, cy, j) * inv;
                            const float xf = floorf(x), yf = floorf(y);
                            const float ax = x - xf, ay = y - yf;
                            const float* Pp = IL ? pbuf + 2 * ((((int)xf - R) & 1) + li) : pbuf + (((int)xf - R) & 3) + lane_off;
                            const int sy = IL ? (((int)yf - R) & 1) : 0;
                            if (lane_on) {
                                float h[G::RPL + 1];
#pragma unroll
                                for (int rr = 0; rr <= G::RPL; ++rr) {
                                    if (j0 + rr <= G::K) {
                                        const int o = IL ? (sy ? il_row[1][rr] : il_row[0][rr]) : rr * G::BC;
                                        const float a = Pp[o], c = Pp[o + (IL ? 2 : 1)];
                                        h[rr] = a + ax * (c - a);
                                    } else {
                                        h[rr] = 0.f;
                                    }
                                }
                                const uint32_t row = abuf + (uint32_t)j * 128;  // A tile row warp * 8 + j, swizzled by j
#pragma unroll
                                for (int kk = 0; kk < G::RPL; ++kk)
                                    if (j0 + kk < G::K) {
                                        const float v = to_tf32(h[kk] + ay * (h[kk + 1] - h[kk]));
                                        asm volatile("st.shared.f32 [%0], %1;" ::"r"(row + ch_off[kk] + ((ch_chunk[kk] ^ (uint32_t)j) << 4)),
                                                     "f"(v)
                                                     : "memory");
                                    }
                            }
                        }
                    }
                    }
                    __syncwarp();  // all lanes are done with this half of the ring before it is refilled
                }
                // this warp's 8 rows of the level are complete: publish them to the tensor core (async proxy)
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(a_full + 8 * (g & 1));
            }
            // the next tile's coordinates move to the lower half-warp
            cx = __shfl_sync(0xffffffffu, cx, (lane & 15) + 16);
            cy = __shfl_sync(0xffffffffu, cy, (lane & 15) + 16);
        }
    } else if (warp == kMmaWarp) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            const uint32_t idesc = make_idesc_tf32(kConvTileM, P.nh, false, false);
            int ws = 0;
            uint32_t wph = 0;
            for (int ti = 0; ti < my_tiles; ++ti) {
                const int acc = ti & 1;
                mbar_wait_sleep(tempty + 8 * acc, ((ti >> 1) & 1) ^ 1, 2000);
                tc_fence_after();
                for (int l = 0; l < LG; ++l) {
                    const int g = ti * LG + l;
                    mbar_wait_sleep(a_full + 8 * (g & 1), (g >> 1) & 1, 1000);
                    tc_fence_after();
                    const uint32_t abuf = a_base + (uint32_t)(g & 1) * C::A_BYTES;
#pragma unroll 1
                    for (int kb = 0; kb < C::KB; ++kb) {
                        const int nk = C::KSTEPS - 4 * kb < 4 ? C::KSTEPS - 4 * kb : 4;
                        const uint64_t adesc = make_smem_desc(abuf + kb * kKBlockBytes);
                        for (int h = 0; h < P.halves; ++h) {
                            mbar_wait_spin(w_full + 8 * ws, wph);
                            tc_fence_after();
                            const uint64_t bdesc = make_smem_desc(w_base + ws * kWStageBytes);
                            const uint32_t d_tmem = tmem_base + acc * 256 + h * P.nh;
                            for (int k = 0; k < nk; ++k)
                                tc_mma_tf32(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc,
                                            (l | kb | k) != 0 ? 1u : 0u);
                            tc_commit(w_empty + 8 * ws);
                            if (++ws == kWStages) { ws = 0; wph ^= 1; }
                        }
                    }
                    tc_commit(a_empty + 8 * (g & 1));  // the gather may overwrite this A tile
                }
                tc_commit(tfull + 8 * acc);
            }
        }
    } else if (warp == kWeightWarp) {
        // ===================== weight loader =====================
        if (lane == 0) {
            int ws = 0;
            uint32_t wph = 0;
            for (int ti = 0; ti < my_tiles; ++ti)
                for (int l = 0; l < LG; ++l)
                    for (int kb = 0; kb < C::KB; ++kb)
                        for (int h = 0; h < P.halves; ++h) {
                            mbar_wait_sleep(w_empty + 8 * ws, wph ^ 1, 1000);
                            mbar_expect_tx(w_full + 8 * ws, (uint32_t)P.nh * 128);
                            tma_load_3d(w_base + ws * kWStageBytes, &map_w, w_full + 8 * ws, 0, h * P.nh, l * C::KB + kb);
                            if (++ws == kWStages) { ws = 0; wph ^= 1; }
                        }
        }
    } else {
        // ===================== epilogue: + bias, ReLU, channel-major stores =====================
        const int ew = warp & 3;  // TMEM lane quarter this warp may access
        const int m = ew * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(ew * 32) << 16;
        for (int ti = 0; ti < my_tiles; ++ti) {
            const int t = first + ti * stride;
            const int b = t / P.tiles_per_img;
            const int p = (t - b * P.tiles_per_img) * kConvTileM + m;
            const bool ok = p < P.N;
            const int acc = ti & 1;
            mbar_wait_sleep(tfull + 8 * acc, (ti >> 1) & 1, 8000);
            tc_fence_after();
            float* dst = P.out + (int64_t)b * P.cout * P.N + p;
            const float* bias = P.bias;
            for (int c0 = 0; c0 < P.cout; c0 += 16) {
                float v[16];
                tmem_ld16(tmem_base + lane_addr + acc * 256 + c0, v);
                tmem_ld_wait();
                if (c0 + 16 >= P.cout) {  // the accumulator has been read completely
                    tc_fence_before();
                    mbar_arrive(tempty + 8 * acc);
                }
                if (bias) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] += __ldg(bias + i);
                    bias += 16;
                }
                if (P.relu) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
                }
                if (ok) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) dst[(int64_t)i * P.N] = v[i];
                }
                dst += (int64_t)16 * P.N;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kConvTmemCols) : "memory");
    }
}

// weight [Cout][L * K^2] fp32 -> packed [L * KB][Cout][32] TF32 (round to nearest), channels beyond K^2 zero
__global__ void pack_conv_weight_kernel(const float* __restrict__ w, float* __restrict__ packed, int cout, int L, int KK, int KB) {
    const int64_t total = (int64_t)L * KB * cout * 32;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i & 31);
        const int co = (int)((i >> 5) % cout);
        const int blk = (int)((i >> 5) / cout);
        const int l = blk / KB, ch = (blk % KB) * 32 + c;
        packed[i] = ch < KK ? to_tf32_exact(w[(int64_t)co * L * KK + l * KK + ch]) : 0.f;
    }
}

template <int R, int LG, int LAY>
int run_lookup_conv(const LookupMaps& maps, const CUtensorMap& map_w, const ConvParams& P, int device, cudaStream_t s) {
    using C = ConvGeo<R>;
    constexpr size_t smem = 1024 + 2 * (size_t)C::A_BYTES + kWStages * kWStageBytes +
                            (size_t)kGatherWarps * kSlots * ConvSlot<R, LAY>::value +
                            8 * (kGatherWarps * kSlots + 2 * kWStages + 8) + 16;
    static_assert(smem <= 232448, "lookup+conv pipeline does not fit shared memory");
    auto kern = lookup_conv_kernel<R, LG, LAY>;
    static bool attr_done[64] = {};
    if (device < 0 || device >= 64 || !attr_done[device]) {
        RC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (device >= 0 && device < 64) attr_done[device] = true;
    }
    int sms = 0;
    RC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    const int grid = P.total_tiles < sms ? P.total_tiles : sms;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kConvThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    RC_CUDA(cudaLaunchKernelEx(&cfg, kern, maps, map_w, P));
    return 0;
}

template <int R, int LAY>
int dispatch_lookup_conv(int lg, const LookupMaps& maps, const CUtensorMap& map_w, const ConvParams& P, int device,
                         cudaStream_t s) {
    switch (lg) {
        case 1: return run_lookup_conv<R, 1, LAY>(maps, map_w, P, device, s);
        case 2: return run_lookup_conv<R, 2, LAY>(maps, map_w, P, device, s);
        case 3: return run_lookup_conv<R, 3, LAY>(maps, map_w, P, device, s);
        default: return run_lookup_conv<R, 4, LAY>(maps, map_w, P, device, s);
    }
}

inline int conv_kblocks(int radius) { return ((2 * radius + 1) * (2 * radius + 1) + 31) / 32; }

}  // namespace

size_t conv_weight_bytes(int levels, int radius, int cout) {
    return (size_t)levels * conv_kblocks(radius) * cout * 32 * sizeof(float);
}

int launch_pack_conv_weight(const float* weight, int levels, int radius, int cout, float* packed, cudaStream_t s) {
    const int KK = (2 * radius + 1) * (2 * radius + 1);
    const int64_t total = (int64_t)levels * conv_kblocks(radius) * cout * 32;
    const int grid = (int)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184);
    pack_conv_weight_kernel<<<grid, 256, 0, s>>>(weight, packed, cout, levels, KK, conv_kblocks(radius));
    RC_CUDA(cudaGetLastError());
    return 0;
}

int launch_lookup_conv(const void* plan_blob, const float* coords, const float* packed_w, const float* bias, int cout,
                       int relu, float* out, int device, cudaStream_t s) {
    LookupPlanData plan;
    std::memcpy(&plan, plan_blob, sizeof(plan));
    RC_REQUIRE(plan.magic == kPlanMagic, "lookup_convc1: lookup plan is not initialised");
    RC_REQUIRE(plan.groups == 1, "lookup_convc1: at most 4 pyramid levels");
    RC_REQUIRE(cout >= 16 && cout <= 256 && cout % 16 == 0 && (cout <= 128 || cout % 32 == 0),
               "lookup_convc1: Cout must be a multiple of 16 up to 128, or a multiple of 32 up to 256 (got %d)", cout);
    const GroupView& g = plan.g[0];
    const int lg = plan.lg[0];
    EncodeTiledFn fn;
    if (int rc = get_encode_fn(&fn)) return rc;
    ConvParams P{};
    P.coords = coords; P.bias = bias; P.out = out;
    P.B = g.B; P.N = g.N;
    P.tiles_per_img = ceil_div(g.N, kConvTileM);
    const int64_t total = (int64_t)g.B * P.tiles_per_img;
    RC_REQUIRE(total < (1LL << 31), "lookup_convc1: too many tiles");
    P.total_tiles = (int)total;
    P.cout = cout;
    P.halves = cout > 128 ? 2 : 1;
    P.nh = cout / P.halves;
    P.relu = relu;
    for (int l = 0; l < 4; ++l) P.inv_scale[l] = g.inv_scale[l];
    CUtensorMap map_w;
    {
        const int KB = conv_kblocks(plan.radius);
        cuuint64_t dims[3] = {32, (cuuint64_t)cout, (cuuint64_t)lg * KB};
        cuuint64_t strides[2] = {128, (cuuint64_t)cout * 128};
        cuuint32_t box[3] = {32, (cuuint32_t)P.nh, 1};
        if (int rc = encode_f32_map(fn, &map_w, (void*)packed_w, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B,
                                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, "packed convc1 weight"))
            return rc;
    }
    const int lay = g.lvl[0].il;  // level 0's code identifies the layout: 0 rows, 1 pairs, 2 quads (+ pairs deeper)
    const LookupMaps& maps = plan.maps[0];
#define RAFTCORR_LC_CASE(R_)                                                                    \
    case R_:                                                                                    \
        return lay == 2 ? dispatch_lookup_conv<R_, 2>(lg, maps, map_w, P, device, s)            \
             : lay      ? dispatch_lookup_conv<R_, 1>(lg, maps, map_w, P, device, s)            \
                        : dispatch_lookup_conv<R_, 0>(lg, maps, map_w, P, device, s);
    switch (plan.radius) {
        RAFTCORR_LC_CASE(1)
        RAFTCORR_LC_CASE(2)
        RAFTCORR_LC_CASE(3)
        RAFTCORR_LC_CASE(4)
        default: return fail("lookup_convc1: radius %d not supported (1..4)", plan.radius);
    }
#undef RAFTCORR_LC_CASE
}

}  // namespace raftcorr

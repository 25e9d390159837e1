// Correlation lookup forward, TMA variant (CorrBlock.__call__, raft/core/corr.py:45-91).
//
// A (pixel, level) patch is exactly what a tiled TMA load describes: a box at a signed element
// coordinate of the 3-D tensor [B*N][H_l][W_l], with out-of-range elements ZERO-FILLED by the hardware
// -- grid_sample's padding_mode='zeros' for free, and no per-element address / bounds arithmetic on
// the SM: one elected lane issues ONE cp.async.bulk.tensor per level.  The only hardware rule is that
// the innermost start coordinate is 16-byte aligned (measured: scripts/probe/tma_probe.cu traps with
// 'illegal instruction' otherwise), so the box starts at x0 & ~3 and is K+1+3 columns wide (rounded to
// 4); the 0..3 column shift is applied when the outputs read their taps.
// Each warp owns two patch buffers and two mbarriers and keeps the next pixel's boxes in flight while
// it evaluates the current pixel; results are transposed through a [channels][32 px] shared tile and
// stored channel-major with full 128-byte lines, as in lookup.cu.
#include <cstdlib>
#include <cstring>

#include "lookup_tma_common.cuh"

namespace raftcorr {

namespace {

constexpr int kDefaultRingDepth = 3;
constexpr int kBarBytes = 32;  // per-warp mbarrier block: up to 3 barriers, padded

// kBufs = patch ring depth per warp.  3 (two pixels in flight while one is evaluated) fits two CTAs per SM
// at r=4; 2 fits three CTAs (24 warps) with the same number of patches in flight per SM.
template <int R, int LG, int kBufs, bool IL>
__global__ void __launch_bounds__(kWarps * 32, kBufs == 2 ? 3 : 2)
lookup_fwd_tma_kernel(const __grid_constant__ LookupMaps maps, const GroupView g, const float* __restrict__ coords,
                      float* __restrict__ out) {
    using G = TmaGeo<R, IL>;
    constexpr int CH = LG * G::KK;
    constexpr int PIX_PER_WARP = kPix / kWarps;
    constexpr int WARP_PATCH = kBufs * LG * G::SLOT;  // bytes: this warp's ring of patch buffers
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 127u) & ~127u;
    uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
    // [kWarps][kBufs][LG][SLOT] patches | [kWarps][kBufs] mbarriers | out_s [CH][kTilePitch]
    const uint32_t bar_base = base + kWarps * WARP_PATCH;
    float* out_s = reinterpret_cast<float*>(gen + kWarps * WARP_PATCH + kWarps * kBarBytes);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.y, p0 = blockIdx.x * kPix;
    const int N = g.N;
    const uint32_t my_patch = base + warp * WARP_PATCH;
    const float* my_patch_gen = reinterpret_cast<const float*>(gen + warp * WARP_PATCH);
    const uint32_t my_bar = bar_base + warp * kBarBytes;

    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < kBufs; ++i) mbar_init(my_bar + 8 * i, 1);
        fence_mbar_init();
        if (warp == 0) {
#pragma unroll
            for (int l = 0; l < LG; ++l) {
                prefetch_tensormap(&maps.m[l]);
                if (IL) prefetch_tensormap(&maps.ms[l]);
            }
        }
    }
    __syncwarp();

    // lanes run along x inside a patch row (fewest bank conflicts); lanes >= K * GROUPS idle during evaluation
    const int jg = lane / G::K, li = lane - jg * G::K;
    const int j0 = jg * G::RPL;
    const bool lane_on = jg < G::GROUPS;
    const int lane_off = j0 * G::BC + li;       // first tap of this lane inside a patch (row-major patches)
    const int lane_ch = li * G::K + j0;         // output channel: i (x offset) major, corr.py:66-78
    // interleaved patches: tap (row Y, col X) sits at (Y >> 1) * BC + 2 * X + (Y & 1); the lane's rows are
    // Y = (y0 & 1) + j0 + rr, so two small tables (window starting on an even / odd row) cover it
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

    // Programmatic dependent launch: everything above overlaps the tail of the previous kernel in the stream (the
    // launch carries cudaLaunchAttributeProgrammaticStreamSerialization); from here on its results are visible.
    // Our own dependents may be scheduled as soon as every CTA of this grid has got this far.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    float cxs[PIX_PER_WARP], cys[PIX_PER_WARP];
#pragma unroll
    for (int n = 0; n < PIX_PER_WARP; ++n) {
        const int p = p0 + warp + kWarps * n;
        const bool v = p < N;
        cxs[n] = clamp_coord(v ? __ldg(coords + ((int64_t)b * 2 + 0) * N + p) : 0.f);
        cys[n] = clamp_coord(v ? __ldg(coords + ((int64_t)b * 2 + 1) * N + p) : 0.f);
    }

    auto issue = [&](int n, int buf) {  // lane 0 only
        const int img = b * N + p0 + warp + kWarps * n;
        if (IL) {
            // a window starting on an even row spans K1/2 pair rows, on an odd row K1/2 + 1: two boxes per level
            int y0[LG], short_boxes = 0;
#pragma unroll
            for (int l = 0; l < LG; ++l) {
                y0[l] = __float2int_rd(cys[n] * g.inv_scale[l]) - R;
                short_boxes += !(y0[l] & 1);
            }
            mbar_expect_tx(my_bar + 8 * buf, LG * G::BOX_BYTES - short_boxes * (G::BC * 4));
#pragma unroll
            for (int l = 0; l < LG; ++l) {
                const int x0 = __float2int_rd(cxs[n] * g.inv_scale[l]) - R;
                // inner coordinate counts interleaved floats: 2 * column, 16-byte aligned at even columns
                tma_load_3d(my_patch + (buf * LG + l) * G::SLOT, (y0[l] & 1) ? &maps.m[l] : &maps.ms[l], my_bar + 8 * buf,
                            2 * (x0 & ~1), y0[l] >> 1, img);
            }
            return;
        }
        mbar_expect_tx(my_bar + 8 * buf, LG * G::BOX_BYTES);
#pragma unroll
        for (int l = 0; l < LG; ++l) {
            const int x0 = __float2int_rd(cxs[n] * g.inv_scale[l]) - R;
            const int y0 = __float2int_rd(cys[n] * g.inv_scale[l]) - R;
            tma_load_3d(my_patch + (buf * LG + l) * G::SLOT, &maps.m[l], my_bar + 8 * buf, x0 & ~3, y0, img);
        }
    };

    // prologue: kBufs-1 pixels in flight before the first one is consumed
    if (lane == 0) {
#pragma unroll
        for (int n = 0; n < kBufs - 1 && n < PIX_PER_WARP; ++n)
            if (p0 + warp + kWarps * n < N) issue(n, n);
    }

#pragma unroll
    for (int n = 0; n < PIX_PER_WARP; ++n) {
        const int q = warp + kWarps * n;
        const int p = p0 + q;
        if (p >= N) break;
        const int buf = n % kBufs;
        constexpr int kAhead = kBufs - 1;
        if (lane == 0 && n + kAhead < PIX_PER_WARP && p + kAhead * kWarps < N) {
            fence_proxy_async();  // that buffer was read through the generic proxy one iteration ago
            issue(n + kAhead, (n + kAhead) % kBufs);
        }
        mbar_wait(my_bar + 8 * buf, (n / kBufs) & 1);
        const float* pbuf = my_patch_gen + buf * (LG * G::SLOT / 4);
#pragma unroll
        for (int l = 0; l < LG; ++l) {
            const float x = cxs[n] * g.inv_scale[l], y = cys[n] * g.inv_scale[l];
            const float xf = floorf(x);
            const float ax = x - xf, ay = y - floorf(y);
            const float* P = IL ? pbuf + l * (G::SLOT / 4) + 2 * ((((int)xf - R) & 1) + li)
                                : pbuf + l * (G::SLOT / 4) + (((int)xf - R) & 3) + lane_off;
            const int sy = IL ? (((int)floorf(y) - R) & 1) : 0;
            if (lane_on) {
                float h[G::RPL + 1];
#pragma unroll
                for (int rr = 0; rr <= G::RPL; ++rr) {
                    if (j0 + rr <= G::K) {  // the last group may own fewer than RPL rows
                        const int o = IL ? (sy ? il_row[1][rr] : il_row[0][rr]) : rr * G::BC;
                        const float a = P[o], c = P[o + (IL ? 2 : 1)];
                        h[rr] = a + ax * (c - a);
                    } else {
                        h[rr] = 0.f;
                    }
                }
                float* o = out_s + (l * G::KK + lane_ch) * kTilePitch + q;
#pragma unroll
                for (int k = 0; k < G::RPL; ++k)
                    if (j0 + k < G::K) o[k * kTilePitch] = h[k] + ay * (h[k + 1] - h[k]);
            }
        }
        __syncwarp();  // all lanes are done with `buf` before lane 0 refills it next iteration
    }
    __syncthreads();
    const int p = p0 + lane;
    if (p < N) {
        float* dst = out + ((int64_t)b * g.ctot + g.ch0) * N + p;
        for (int ch = warp; ch < CH; ch += kWarps) dst[(int64_t)ch * N] = out_s[ch * kTilePitch + lane];
    }
}


// ---------------------------------------------------------------------------------------------------------------
// Layout 2 (raftcorr_pyramid_layout::interleaved == 2: row quads on every level): the default of the tensor-core builds.
//
// ncu on the row-pair kernel above (profiles/r02_ncu_full_r2a_*.json) showed it bound by its own instruction stream, not
// by DRAM (53 % DRAM cycles active, 121 warp instructions per (pixel, level) window, 16 warps per SM), and
// scripts/gather_roofline.py measured the same for the first quad kernel: the gather alone 27 us, evaluation + output
// alone 31 us.  Here
//  * the box is exact (K+1 columns at any x, GMAX or GMAX-1 quads): no alignment shift on either axis at evaluation time;
//  * a warp evaluates TWO pixels x four levels per pass: 4 lanes per (pixel, level) window, each lane owns 2-3 window
//    COLUMNS, reads them as 16-byte quads (LDS.128, compile-time offsets), forms the horizontal lerps of the quad rows
//    and the vertical lerps of the K+3 quad-relative rows that can be an output; the window's start row inside its
//    quad (sy) only moves the DESTINATION channel (one address subtraction) and predicates the stores -- no dynamic
//    register indexing, all 32 lanes busy, ~115 warp instructions per pixel instead of ~480;
//  * the 8 boxes of a pixel pair are issued by 8 lanes at once (each arrives on the pair's mbarrier with its own byte
//    count), so the coordinate / tensor-map arithmetic runs once per pair;
//  * PERSISTENT: the CTA walks tiles of 16 consecutive source pixels (one pixel pair per warp and tile); the ring of
//    patch boxes runs ACROSS tiles, so only the very first pair of a CTA pays the coordinates -> box -> data latency.
constexpr int kQPix = 16;    // source pixels per tile: a 64-byte output segment per channel
constexpr int kQPitch = 17;  // out tile row pitch (floats)

// Wait of the hot loop: no clock reads per poll (mbar_wait spends six instructions per poll on its time-out and waiting
// warps take issue slots from evaluating ones); the suspend-time hint lets a poll sleep in hardware; a protocol bug
// still becomes a trapped launch instead of a hang.
__device__ __forceinline__ void mbar_wait_polls(uint32_t bar, uint32_t parity) {
    uint32_t ok, polls = 0;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity), "r"(4000u)
            : "memory");
        if (!ok && ++polls > (1u << 24)) __trap();
    } while (!ok);
}

// EXP (measurement only, scripts/gather_roofline.py): 0 = the kernel; 1 = the gather alone (same boxes, same waits, no
// evaluation, no output: what the memory system charges for this box list); 2 = evaluation + output alone (no boxes).
template <int R, int LG, int EXP = 0>
__global__ void __launch_bounds__(kWarps * 32, 2)
lookup_fwd_quad_kernel(const __grid_constant__ LookupMaps maps, const GroupView g, const float* __restrict__ coords,
                       float* __restrict__ out, const int tiles_per_img, const int total_tiles) {
    using Q = QuadGeo<R>;
    constexpr int CH = LG * Q::KK;
    constexpr int PPW = kQPix / kWarps;               // pixels per warp and tile: one pair
    constexpr int kPairSlots = 2;                     // ring: the pair being evaluated + the pair in flight
    constexpr int PAIR_BYTES = PPW * 4 * Q::SLOT;     // 2 pixels x 4 level slots
    constexpr int WARP_PATCH = kPairSlots * PAIR_BYTES;
    // window columns per lane: K columns over the 4 lanes of a window
    constexpr int CBASE = Q::K / 4, CREM = Q::K % 4, CMAX = CBASE + (CREM ? 1 : 0);
    static_assert(PPW == 2 && LG <= 4, "a pass is 2 pixels x 4 levels x 4 lanes");
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 127u) & ~127u;
    uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
    // [kWarps][2 pairs][2 px][4 levels][SLOT] patches | [kWarps][2] mbarriers | out_s [CH][kQPitch]
    const uint32_t bar_base = base + kWarps * WARP_PATCH;
    float* out_s = reinterpret_cast<float*>(gen + kWarps * WARP_PATCH + kWarps * kBarBytes);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int N = g.N;
    const uint32_t my_patch = base + warp * WARP_PATCH;
    const float* my_patch_gen = reinterpret_cast<const float*>(gen + warp * WARP_PATCH);
    const uint32_t my_bar = bar_base + warp * kBarBytes;
    const int first = blockIdx.x, stride = gridDim.x;
    const int my_tiles = first < total_tiles ? (total_tiles - first + stride - 1) / stride : 0;

    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < kPairSlots; ++i) mbar_init(my_bar + 8 * i, PPW * 4);  // one arrival per (pixel, level slot)
        fence_mbar_init();
        if (warp == 0) {
#pragma unroll
            for (int l = 0; l < LG; ++l) {
                prefetch_tensormap(&maps.m[l]);
                prefetch_tensormap(&maps.ms[l]);
            }
        }
    }
    __syncwarp();

    // roles of a lane.  Issue: lane i < 8 -> pixel i >> 2 of the pair, level i & 3.  Evaluation: window = lane >> 2
    // (same numbering), sub = lane & 3 -> window columns [c0, c0 + cn).
    const int w_px = (lane >> 4) & 1, w_lv = (lane >> 2) & 3, sub = lane & 3;
    const int c0 = sub * CBASE + (sub < CREM ? sub : CREM), cn = CBASE + (sub < CREM ? 1 : 0);
    const int i_px = (lane >> 2) & 1, i_lv = lane & 3;

    asm volatile("griddepcontrol.wait;" ::: "memory");  // PDL: see lookup_fwd_tma_kernel
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    auto tile_origin = [&](int it, int& b, int& p0) {  // integer division: only where a cursor is set up
        const int t = first + it * stride;
        b = t / tiles_per_img;
        p0 = (t - b * tiles_per_img) * kQPix;
    };
    const int img_span = tiles_per_img * kQPix, step_span = stride * kQPix;
    auto tile_advance = [&](int& b, int& p0) {  // (b, p0) of the cursor's next tile without dividing
        p0 += step_span;
        while (p0 >= img_span) { p0 -= img_span; ++b; }
    };
    // lanes 0..1: coordinates of the warp's pixels (tile pixel q = warp + kWarps * n) of the tile being consumed; lanes
    // 2..3: the same for the next tile; (nx, ny): the tile after that, on its way from memory (merged in at the next
    // tile switch, a whole tile after the load was issued)
    auto load_coords_at = [&](int it, int b, int p0, float& x, float& y) {
        x = y = 0.f;
        if (it < my_tiles && lane < 2 * PPW) {
            const int p = p0 + warp + kWarps * (lane & (PPW - 1));
            if (p < N) {
                x = clamp_coord(__ldg(coords + ((int64_t)b * 2 + 0) * N + p));
                y = clamp_coord(__ldg(coords + ((int64_t)b * 2 + 1) * N + p));
            }
        }
    };
    float cx, cy, nx, ny;
    int b = 0, p0 = 0;           // tile being consumed
    int i_b = 0, i_p0 = 0;       // next tile (the one whose boxes are requested while this one is evaluated)
    int l_b = 0, l_p0 = 0;       // coordinates cursor: tile it + 2
    if (my_tiles > 0) tile_origin(0, b, p0);
    i_b = b; i_p0 = p0;
    tile_advance(i_b, i_p0);
    load_coords_at(lane < PPW ? 0 : 1, lane < PPW ? b : i_b, lane < PPW ? p0 : i_p0, cx, cy);
    l_b = i_b; l_p0 = i_p0;
    tile_advance(l_b, l_p0);
    load_coords_at(2, l_b, l_p0, nx, ny);
    const float my_inv = g.inv_scale[i_lv < LG ? i_lv : 0];          // issue role
    const char* map_base = reinterpret_cast<const char*>(&maps);     // m[l] at 128 * l, ms[l] at 512 + 128 * l

    // request the 8 boxes of a pixel pair: tile (tb, tp0), coordinates in lanes src0 / src0 + 1, ring slot `slot`
    auto issue_pair = [&](int tb, int tp0, int src0, int slot) {
        const float x = __shfl_sync(0xffffffffu, cx, src0 + i_px), y = __shfl_sync(0xffffffffu, cy, src0 + i_px);
        if (lane < PPW * 4) {
            const int p = tp0 + warp + kWarps * i_px;
            const uint32_t bar = my_bar + 8 * slot;
            if (EXP == 2 || i_lv >= LG || p >= N) {
                mbar_arrive(bar);  // nothing to fetch for this slot: keep the arrival count
            } else {
                const int x0 = __float2int_rd(x * my_inv) - R;
                const int y0 = __float2int_rd(y * my_inv) - R;
                const int nq = Q::quads(y0 & 3);  // the window spans GMAX or GMAX-1 row quads
                fence_proxy_async();  // the slot was read through the generic proxy when it was last evaluated
                mbar_expect_tx(bar, nq * Q::ROW_BYTES);
                tma_load_3d(my_patch + slot * PAIR_BYTES + (i_px * 4 + i_lv) * Q::SLOT,
                            reinterpret_cast<const CUtensorMap*>(map_base + (nq == Q::GMAX ? 0 : 512) + 128 * i_lv), bar, 4 * x0,
                            y0 >> 2, tb * N + p);
            }
        }
    };
    if (my_tiles > 0) issue_pair(b, p0, 0, 0);

    for (int it = 0; it < my_tiles; ++it) {
        const int slot = it & 1;
        if (it + 1 < my_tiles) issue_pair(i_b, i_p0, PPW, slot ^ 1);
        const int q = warp + kWarps * w_px;        // this lane's pixel inside the tile
        const float xq = __shfl_sync(0xffffffffu, cx, w_px), yq = __shfl_sync(0xffffffffu, cy, w_px);
        mbar_wait_polls(my_bar + 8 * slot, (it >> 1) & 1);
        if (EXP != 1 && w_lv < LG && p0 + q < N) {
            const float inv = g.inv_scale[w_lv];
            const float x = xq * inv, y = yq * inv;
            const float xf = floorf(x), yf = floorf(y);
            const float ax = x - xf, ay = y - yf;
            const int sy = ((int)yf - R) & 3;
            // column chunk = 16 bytes (rows 4g .. 4g+3); the lane's first column
            const float4* cp = reinterpret_cast<const float4*>(my_patch_gen + (slot * PAIR_BYTES + (w_px * 4 + w_lv) * Q::SLOT) / 4) + c0;
            // output row j = Y - sy of window column c is channel c * K + j (x offset major, corr.py:66-78)
            float* o = out_s + ((w_lv * Q::KK + c0 * Q::K - sy) * kQPitch + q);
            bool ok[Q::NV];
#pragma unroll
            for (int Y = 0; Y < Q::NV; ++Y) ok[Y] = Y < 3 ? (sy <= Y) : (Y >= Q::K ? (sy > Y - Q::K) : true);
            float4 a[Q::GMAX];
#pragma unroll
            for (int gq = 0; gq < Q::GMAX; ++gq) a[gq] = cp[gq * Q::K1];
#pragma unroll
            for (int ci = 0; ci < CMAX; ++ci) {
                if (ci < cn) {
                    float4 c[Q::GMAX];
#pragma unroll
                    for (int gq = 0; gq < Q::GMAX; ++gq) c[gq] = cp[gq * Q::K1 + ci + 1];
                    float hh[4 * Q::GMAX];
#pragma unroll
                    for (int gq = 0; gq < Q::GMAX; ++gq) {
                        hh[4 * gq + 0] = a[gq].x + ax * (c[gq].x - a[gq].x);
                        hh[4 * gq + 1] = a[gq].y + ax * (c[gq].y - a[gq].y);
                        hh[4 * gq + 2] = a[gq].z + ax * (c[gq].z - a[gq].z);
                        hh[4 * gq + 3] = a[gq].w + ax * (c[gq].w - a[gq].w);
                    }
#pragma unroll
                    for (int Y = 0; Y < Q::NV; ++Y) {
                        const float v = hh[Y] + ay * (hh[Y + 1] - hh[Y]);
                        if (ok[Y]) o[(ci * Q::K + Y) * kQPitch] = v;
                    }
#pragma unroll
                    for (int gq = 0; gq < Q::GMAX; ++gq) a[gq] = c[gq];
                }
            }
        }
        __syncthreads();  // (also: every lane of every warp is done with its patch slot)
        if (EXP == 1) {
            // gather alone: nothing was evaluated, nothing leaves
        } else if ((N & 3) == 0 && p0 + kQPix <= N) {
            // 4 lanes x 4 pixels per channel row: 64-byte segments leave as 128-bit stores
            float* dst = out + ((int64_t)b * g.ctot + g.ch0) * N + p0;
            for (int idx = threadIdx.x; idx < CH * 4; idx += kWarps * 32) {
                const int ch = idx >> 2, qd = idx & 3;
                const float* sp = out_s + ch * kQPitch + 4 * qd;
                *reinterpret_cast<float4*>(dst + (int64_t)ch * N + 4 * qd) = make_float4(sp[0], sp[1], sp[2], sp[3]);
            }
        } else {
            const int px = threadIdx.x & (kQPix - 1), p = p0 + px;
            if (p < N) {
                float* dst = out + ((int64_t)b * g.ctot + g.ch0) * N + p;
                for (int ch = threadIdx.x / kQPix; ch < CH; ch += kWarps * 32 / kQPix) dst[(int64_t)ch * N] = out_s[ch * kQPitch + px];
            }
        }
        // tile switch: next tile's coordinates move to lanes 0..1, the pending ones behind them
        {
            const float sx = __shfl_sync(0xffffffffu, cx, (lane & (PPW - 1)) + PPW);
            const float sy2 = __shfl_sync(0xffffffffu, cy, (lane & (PPW - 1)) + PPW);
            cx = lane < PPW ? sx : nx;
            cy = lane < PPW ? sy2 : ny;
            tile_advance(l_b, l_p0);
            load_coords_at(it + 3, l_b, l_p0, nx, ny);
            b = i_b; p0 = i_p0;
            tile_advance(i_b, i_p0);
        }
        __syncthreads();  // out_s is free for the next tile
    }
}

inline bool lookup_pdl() {  // RAFTCORR_LOOKUP_PDL=0 launches without programmatic stream serialization (A/B switch)
    static const bool on = [] {
        const char* e = std::getenv("RAFTCORR_LOOKUP_PDL");
        return !(e && e[0] == '0');
    }();
    return on;
}

template <int R, int LG, int kBufs, bool IL>
int run_fwd_tma_bufs(const LookupMaps& maps, const GroupView& g, const float* coords, float* out, cudaStream_t s) {
    using G = TmaGeo<R, IL>;
    const size_t smem = 128 + (size_t)kWarps * kBufs * LG * G::SLOT + kWarps * kBarBytes + sizeof(float) * LG * G::KK * kTilePitch;
    auto kern = lookup_fwd_tma_kernel<R, LG, kBufs, IL>;
    // the opt-in is per function AND per device (nn.DataParallel drives several devices from one process)
    static bool attr_done[64] = {};
    int dev = 0;
    RC_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_done[dev]) {
        RC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (dev >= 0 && dev < 64) attr_done[dev] = true;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(ceil_div(g.N, kPix), g.B);
    cfg.blockDim = dim3(kWarps * 32);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = lookup_pdl() ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    RC_CUDA(cudaLaunchKernelEx(&cfg, kern, maps, g, coords, out));
    return 0;
}


template <int R, int LG>
int run_fwd_quad(const LookupMaps& maps, const GroupView& g, const float* coords, float* out, cudaStream_t s) {
    using Q = QuadGeo<R>;
    const size_t smem = 128 + (size_t)kWarps * 2 * 2 * 4 * Q::SLOT + kWarps * kBarBytes + sizeof(float) * LG * Q::KK * kQPitch;
    auto kern = lookup_fwd_quad_kernel<R, LG>;
    static bool attr_done[64] = {};
    int dev = 0;
    RC_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_done[dev]) {
        RC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (dev >= 0 && dev < 64) attr_done[dev] = true;
    }
    if (R == 4 && LG == 4) {
        // RAFTCORR_LOOKUP_EXP=1|2: timing-only ablations of the bench shape (the results are NOT a lookup)
        const char* e = std::getenv("RAFTCORR_LOOKUP_EXP");
        if (e && (e[0] == '1' || e[0] == '2')) {
            kern = e[0] == '1' ? lookup_fwd_quad_kernel<R, LG, (R == 4 && LG == 4) ? 1 : 0>
                               : lookup_fwd_quad_kernel<R, LG, (R == 4 && LG == 4) ? 2 : 0>;
            RC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        }
    }
    static int sm_count[64] = {};
    int sms = (dev >= 0 && dev < 64) ? sm_count[dev] : 0;
    if (!sms) {
        RC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        if (dev >= 0 && dev < 64) sm_count[dev] = sms;
    }
    const int tiles_per_img = ceil_div(g.N, kQPix);
    const int64_t total = (int64_t)tiles_per_img * g.B;
    RC_REQUIRE(total < (1LL << 31), "lookup: too many tiles");
    const int ctas = 2 * sms;  // two resident CTAs per SM (shared memory), each walking total / ctas tiles
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(total < ctas ? total : ctas));
    cfg.blockDim = dim3(kWarps * 32);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = lookup_pdl() ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    RC_CUDA(cudaLaunchKernelEx(&cfg, kern, maps, g, coords, out, tiles_per_img, (int)total));
    return 0;
}

inline int lookup_ring_depth() {  // RAFTCORR_LOOKUP_BUFS=2|3 (A/B switch)
    static const int depth = [] {
        const char* e = std::getenv("RAFTCORR_LOOKUP_BUFS");
        return (e && e[0] == '2') ? 2 : (e && e[0] == '3') ? 3 : kDefaultRingDepth;
    }();
    return depth;
}

template <int R, int LG>
int run_fwd_tma(const LookupMaps& maps, const GroupView& g, const float* coords, float* out, cudaStream_t s) {
    if (g.lvl[0].il == 2) return run_fwd_quad<R, LG>(maps, g, coords, out, s);
    if (g.lvl[0].il) return run_fwd_tma_bufs<R, LG, 3, true>(maps, g, coords, out, s);
    return lookup_ring_depth() == 2 ? run_fwd_tma_bufs<R, LG, 2, false>(maps, g, coords, out, s)
                                    : run_fwd_tma_bufs<R, LG, 3, false>(maps, g, coords, out, s);
}

template <int R>
int dispatch_fwd_tma(int lg, const LookupMaps& maps, const GroupView& g, const float* coords, float* out,
                     cudaStream_t s) {
    switch (lg) {
        case 1: return run_fwd_tma<R, 1>(maps, g, coords, out, s);
        case 2: return run_fwd_tma<R, 2>(maps, g, coords, out, s);
        case 3: return run_fwd_tma<R, 3>(maps, g, coords, out, s);
        default: return run_fwd_tma<R, 4>(maps, g, coords, out, s);
    }
}

}  // namespace

// ---- plans: tensor maps + launch geometry encoded once per pyramid, reused by every refinement iteration
// (LookupPlanData: lookup_tma_common.cuh)
int lookup_plan_init(void* plan_blob, const PyramidView& pv, int radius) {
    RC_REQUIRE(radius >= 1 && radius <= 4, "lookup: radius %d not supported (1..4)", radius);
    EncodeTiledFn fn;
    if (int rc = get_encode_fn(&fn)) return rc;
    LookupPlanData plan;
    std::memset(&plan, 0, sizeof(plan));
    plan.magic = kPlanMagic;
    plan.radius = radius;
    const int K1 = 2 * radius + 2;
    for (int l0 = 0; l0 < pv.L; l0 += 4) {
        const int gi = plan.groups++;
        const int lg = pv.L - l0 < 4 ? pv.L - l0 : 4;
        plan.lg[gi] = lg;
        plan.g[gi] = make_group(pv, l0, lg, radius);
        for (int l = 0; l < lg; ++l) {
            const LevelView& lv = pv.lvl[l0 + l];
            // box geometry of this level's layout == TmaGeo (rows / row pairs), QuadGeo (row quads)
            const bool quad = lv.il == 2, il = lv.il == 1;
            const int BC = quad ? 4 * K1 : il ? 2 * ((K1 + 1 + 1) / 2 * 2) : (K1 + 3 + 3) / 4 * 4;
            const int BR = quad ? (K1 + 2) / 4 + 1 : il ? K1 / 2 + 1 : K1;
            const int rg = lv.group_rows();
            // logical width = W_l (not the pitch): columns in the row padding read as zeros, like everything outside
            // row groups: [rg * W_l floats][ceil(H_l / rg) groups]; the rows completing the last group hold zeros
            cuuint64_t dims[3] = {(cuuint64_t)(rg * lv.w), (cuuint64_t)((lv.h + rg - 1) / rg), (cuuint64_t)pv.B * pv.N};
            cuuint64_t strides[2] = {(cuuint64_t)lv.pitch * 4 * rg, (cuuint64_t)lv.img_stride * 4};
            cuuint32_t box[3] = {(cuuint32_t)BC, (cuuint32_t)BR, 1};
            if (int rc = encode_f32_map(fn, &plan.maps[gi].m[l], (void*)lv.base, 3, dims, strides, box,
                                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, "pyramid level"))
                return rc;
            if (il || quad) box[1] = (cuuint32_t)(BR - 1);
            if (int rc = encode_f32_map(fn, &plan.maps[gi].ms[l], (void*)lv.base, 3, dims, strides, box,
                                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, "pyramid level"))
                return rc;
        }
        for (int l = lg; l < 4; ++l) {
            plan.maps[gi].m[l] = plan.maps[gi].m[0];
            plan.maps[gi].ms[l] = plan.maps[gi].ms[0];
        }
    }
    std::memcpy(plan_blob, &plan, sizeof(plan));
    return 0;
}

int lookup_plan_run(const void* plan_blob, const float* coords, float* out, cudaStream_t s) {
    LookupPlanData plan;  // local, properly aligned copy (the caller's blob is plain bytes)
    std::memcpy(&plan, plan_blob, sizeof(plan));
    RC_REQUIRE(plan.magic == kPlanMagic, "lookup plan is not initialised");
    const bool tma = use_tma_lookup();
    for (int gi = 0; gi < plan.groups; ++gi) {
        int rc;
        if (!tma) {
            rc = launch_lookup_fwd_group_cpasync(plan.radius, plan.lg[gi], plan.g[gi], coords, out, s);
            if (rc) return rc;
            continue;
        }
        switch (plan.radius) {
            case 1: rc = dispatch_fwd_tma<1>(plan.lg[gi], plan.maps[gi], plan.g[gi], coords, out, s); break;
            case 2: rc = dispatch_fwd_tma<2>(plan.lg[gi], plan.maps[gi], plan.g[gi], coords, out, s); break;
            case 3: rc = dispatch_fwd_tma<3>(plan.lg[gi], plan.maps[gi], plan.g[gi], coords, out, s); break;
            default: rc = dispatch_fwd_tma<4>(plan.lg[gi], plan.maps[gi], plan.g[gi], coords, out, s); break;
        }
        if (rc) return rc;
    }
    return 0;
}

int launch_lookup_fwd_tma(const PyramidView& pv, const float* coords, int radius, float* out, cudaStream_t s) {
    alignas(64) unsigned char blob[RAFTCORR_LOOKUP_PLAN_BYTES];
    if (int rc = lookup_plan_init(blob, pv, radius)) return rc;
    return lookup_plan_run(blob, coords, out, s);
}

}  // namespace raftcorr

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
// ncu on the row-pair kernel above (profiles/r02_ncu_lookup_pair_*.json) showed it bound by its own instruction stream,
// not by DRAM (53 % DRAM cycles active, 121 warp instructions per (pixel, level) window, 16 warps per SM): per window
// ~25 useful instructions and ~80 of addressing (per-tap row tables, alignment shifts, lane-0 box arithmetic).  Here
//  * the box is exact (K+1 columns at any x, 3 or 4 quads): no alignment shift on either axis at evaluation time;
//  * lane = window COLUMN, half-warp = pyramid level: a lane reads its column and the next as 16-byte quads
//    (LDS.128, compile-time offsets), forms the horizontal lerps of all 4*GMAX quad rows and the vertical lerps of the
//    K+3 quad-relative rows that can be an output; the window's start row inside its quad (sy) only moves the
//    DESTINATION channel (one address subtraction) and predicates the stores -- no dynamic register indexing;
//  * the LG boxes of a pixel are issued by LG lanes at once (each arrives on the pixel's mbarrier with its own byte
//    count), so the coordinate / tensor-map arithmetic runs once per pixel instead of once per level.
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
// TPX = source pixels per tile (32: one 128-byte line per channel; 16 with a 2-deep ring fits three CTAs per SM), MINB =
// resident CTAs per SM the register allocation is sized for
template <int R, int LG, int kBufs, int EXP = 0, int TPX = kPix, int MINB = 2>
__global__ void __launch_bounds__(kWarps * 32, MINB)
lookup_fwd_quad_kernel(const __grid_constant__ LookupMaps maps, const GroupView g, const float* __restrict__ coords,
                       float* __restrict__ out, const int tiles_per_img, const int total_tiles) {
    using Q = QuadGeo<R>;
    constexpr int SLOT = Q::SLOT;
    constexpr int CH = LG * Q::KK;
    constexpr int PIX_PER_WARP = TPX / kWarps;
    constexpr int QP = TPX + 1;  // out tile row pitch (floats): conflict-free column writes
    constexpr int WARP_PATCH = kBufs * LG * SLOT;  // bytes: this warp's ring of patch buffers
    static_assert(kBufs - 1 <= PIX_PER_WARP, "the issue cursor may run at most one tile ahead");
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 127u) & ~127u;
    uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
    // [kWarps][kBufs][LG][SLOT] patches | [kWarps][kBufs] mbarriers | out_s [CH][QP]
    const uint32_t bar_base = base + kWarps * WARP_PATCH;
    float* out_s = reinterpret_cast<float*>(gen + kWarps * WARP_PATCH + kWarps * kBarBytes);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int N = g.N;
    const uint32_t my_patch = base + warp * WARP_PATCH;
    const float* my_patch_gen = reinterpret_cast<const float*>(gen + warp * WARP_PATCH);
    const uint32_t my_bar = bar_base + warp * kBarBytes;
    // PERSISTENT: the CTA walks tiles of kPix consecutive source pixels (first, first + stride, ...); a warp's ring of
    // patch boxes runs ACROSS tiles, so only the very first pixels of a CTA pay the coordinates -> box -> data latency
    // chain (ncu on the one-tile-per-CTA version: half of a CTA's life was that chain, DRAM 53 % busy)
    const int first = blockIdx.x, stride = gridDim.x;
    const int my_tiles = first < total_tiles ? (total_tiles - first + stride - 1) / stride : 0;

    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < kBufs; ++i) mbar_init(my_bar + 8 * i, LG);  // one arrival (+ its bytes) per level box
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

    const int half = lane >> 4, col = lane & 15;  // evaluation: half-warp = level of the pass, lane = window column

    asm volatile("griddepcontrol.wait;" ::: "memory");  // PDL: see lookup_fwd_tma_kernel
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    auto tile_origin = [&](int it, int& b, int& p0) {  // integer division: only where a cursor is set up
        const int t = first + it * stride;
        b = t / tiles_per_img;
        p0 = (t - b * tiles_per_img) * TPX;
    };
    const int img_span = tiles_per_img * TPX, step_span = stride * TPX;
    auto tile_advance = [&](int& b, int& p0) {  // (b, p0) of the cursor's next tile without dividing
        p0 += step_span;
        while (p0 >= img_span) { p0 -= img_span; ++b; }
    };
    // lane n < PIX_PER_WARP: coordinates of the warp's pixel n (tile pixel q = warp + kWarps * n) of the tile being
    // consumed; lanes PIX_PER_WARP .. 2 PIX_PER_WARP - 1: the same for the next tile; (nx, ny): the tile after that, on
    // its way from memory (merged in at the next tile switch, a whole tile after the load was issued)
    auto load_coords_at = [&](int it, int b, int p0, float& x, float& y) {
        x = y = 0.f;
        if (it < my_tiles && lane < 2 * PIX_PER_WARP) {
            const int p = p0 + warp + kWarps * (lane & (PIX_PER_WARP - 1));
            if (p < N) {
                x = clamp_coord(__ldg(coords + ((int64_t)b * 2 + 0) * N + p));
                y = clamp_coord(__ldg(coords + ((int64_t)b * 2 + 1) * N + p));
            }
        }
    };
    float cx, cy, nx, ny;
    int l_b = 0, l_p0 = 0;  // coordinates cursor: tile it + 2 while tile it is consumed
    {
        int b0 = 0, p00 = 0, b1, p01;
        if (my_tiles > 0) tile_origin(0, b0, p00);
        b1 = b0; p01 = p00;
        tile_advance(b1, p01);
        load_coords_at(lane < PIX_PER_WARP ? 0 : 1, lane < PIX_PER_WARP ? b0 : b1, lane < PIX_PER_WARP ? p00 : p01, cx, cy);
        l_b = b1; l_p0 = p01;
        tile_advance(l_b, l_p0);
        load_coords_at(2, l_b, l_p0, nx, ny);
    }
    const float my_inv = g.inv_scale[lane < LG ? lane : 0];          // issue role: lane = level
    const char* map_base = reinterpret_cast<const char*>(&maps);     // m[l] at 128 * l, ms[l] at 512 + 128 * l

    // issue cursor: next (tile, pixel) whose boxes have not been requested; ring position of the next request
    int i_it = 0, i_n = 0, i_buf = 0, i_b = 0, i_p0 = 0;
    if (my_tiles > 0) tile_origin(0, i_b, i_p0);
    auto issue_next = [&](int c_it) {  // whole warp; lanes < LG each issue one level's box of the next valid pixel
        // never more than one tile ahead of the tile being consumed (c_it): that is as far as the coordinates held in
        // lanes PIX_PER_WARP.. reach (sparse last tiles of an image could otherwise let the cursor run further)
        while (i_it < my_tiles && i_it <= c_it + 1) {
            const int b = i_b;
            const int p = i_p0 + warp + kWarps * i_n;
            const int src = i_n + (i_it != c_it ? PIX_PER_WARP : 0);
            if (++i_n == PIX_PER_WARP) {
                i_n = 0;
                ++i_it;
                tile_advance(i_b, i_p0);
            }
            if (p >= N) continue;  // past the end of an image (last tile only)
            const float x = __shfl_sync(0xffffffffu, cx, src), y = __shfl_sync(0xffffffffu, cy, src);
            if (lane < LG) {
                const int x0 = __float2int_rd(x * my_inv) - R;
                const int y0 = __float2int_rd(y * my_inv) - R;
                const bool quad = true;  // layout 2: every level is stored as row quads
                const int nq = Q::quads(y0 & 3);  // the window spans GMAX or GMAX-1 row quads
                const bool big = nq == Q::GMAX;
                const uint32_t bytes = nq * Q::ROW_BYTES;
                const uint32_t bar = my_bar + 8 * i_buf;
                fence_proxy_async();  // the buffer was read through the generic proxy when it was last evaluated
                if (EXP != 2) {
                    mbar_expect_tx(bar, bytes);
                    tma_load_3d(my_patch + (i_buf * LG + lane) * SLOT,
                                reinterpret_cast<const CUtensorMap*>(map_base + (big ? 0 : 512) + 128 * lane), bar,
                                quad ? 4 * x0 : 2 * (x0 & ~1), quad ? y0 >> 2 : y0 >> 1, b * N + p);
                }
            }
            if (++i_buf == kBufs) i_buf = 0;
            return;
        }
    };
#pragma unroll
    for (int i = 0; i < kBufs - 1; ++i) issue_next(0);

    int c_buf = 0;
    uint32_t c_par = 0;
    int b = 0, p0 = 0;
    if (my_tiles > 0) tile_origin(0, b, p0);
    for (int it = 0; it < my_tiles; ++it) {
#pragma unroll
        for (int n = 0; n < PIX_PER_WARP; ++n) {
            const int q = warp + kWarps * n;
            if (p0 + q >= N) continue;  // warp-uniform
            issue_next(it);
            const float xq = __shfl_sync(0xffffffffu, cx, n), yq = __shfl_sync(0xffffffffu, cy, n);
            if (EXP != 2) mbar_wait_polls(my_bar + 8 * c_buf, c_par);
#pragma unroll
            for (int pass = 0; pass < (LG + 1) / 2; ++pass) {
                const int l = 2 * pass + half;
                if (EXP != 1 && col < Q::K && l < LG) {
                    const float inv = g.inv_scale[l];
                    const float x = xq * inv, y = yq * inv;
                    const float xf = floorf(x), yf = floorf(y);
                    const float ax = x - xf, ay = y - yf;
                    const float* pb = my_patch_gen + (c_buf * LG + l) * (SLOT / 4);
                    {
                        // ---- row quads: column chunk = 16 bytes (rows 4g .. 4g+3)
                        const int sy = ((int)yf - R) & 3;
                        const float4* cp = reinterpret_cast<const float4*>(pb) + col;
                        float hh[4 * Q::GMAX];
                        // the last quad only belongs to windows that span GMAX quads (1 in 4 at r = 4: sy == 3); for the
                        // others its slot holds stale bytes that no valid output reads.  Skipping it matters: only its first
                        // row is ever used, so ptxas turns these loads into 4-byte LDS at a 16-byte lane stride -- 4-way bank
                        // conflicts, 12 % of the kernel's shared-memory wavefronts (ncu source view, round 2)
                        const bool last_quad = Q::quads(sy) == Q::GMAX;
#pragma unroll
                        for (int gq = 0; gq < Q::GMAX; ++gq) {
                            if (gq == Q::GMAX - 1 && !last_quad) {
                                hh[4 * gq + 0] = hh[4 * gq + 1] = hh[4 * gq + 2] = hh[4 * gq + 3] = 0.f;
                                continue;
                            }
                            const float4 a = cp[gq * Q::K1], c = cp[gq * Q::K1 + 1];
                            hh[4 * gq + 0] = a.x + ax * (c.x - a.x);
                            hh[4 * gq + 1] = a.y + ax * (c.y - a.y);
                            hh[4 * gq + 2] = a.z + ax * (c.z - a.z);
                            hh[4 * gq + 3] = a.w + ax * (c.w - a.w);
                        }
                        // output row j = Y - sy of column `col` is channel col * K + j (x offset major, corr.py:66-78)
                        float* o = out_s + ((l * Q::KK + col * Q::K - sy) * QP + q);
#pragma unroll
                        for (int Y = 0; Y < Q::NV; ++Y) {
                            const float v = hh[Y] + ay * (hh[Y + 1] - hh[Y]);
                            const bool ok = Y < 3 ? (sy <= Y) : (Y >= Q::K ? (sy > Y - Q::K) : true);
                            if (ok) o[Y * QP] = v;
                        }
                    }
                }
            }
            if (++c_buf == kBufs) { c_buf = 0; c_par ^= 1u; }
            __syncwarp();  // all lanes are done with the buffer before it is refilled
        }
        __syncthreads();
        if (EXP == 1) {
            // gather alone: nothing was evaluated, nothing leaves
        } else if ((N & 3) == 0 && p0 + TPX <= N) {
            // 8 lanes x 4 pixels per channel row: 128-byte lines leave as 128-bit stores
            float* dst = out + ((int64_t)b * g.ctot + g.ch0) * N + p0;
            for (int idx = threadIdx.x; idx < CH * (TPX / 4); idx += kWarps * 32) {
                const int ch = idx / (TPX / 4), qd = idx - ch * (TPX / 4);
                const float* sp = out_s + ch * QP + 4 * qd;
                *reinterpret_cast<float4*>(dst + (int64_t)ch * N + 4 * qd) = make_float4(sp[0], sp[1], sp[2], sp[3]);
            }
        } else {
            const int p = p0 + lane;
            if (lane < TPX && p < N) {
                float* dst = out + ((int64_t)b * g.ctot + g.ch0) * N + p;
                for (int ch = warp; ch < CH; ch += kWarps) dst[(int64_t)ch * N] = out_s[ch * QP + lane];
            }
        }
        // tile switch: next tile's coordinates move to lanes 0 .. PIX_PER_WARP-1, the pending ones behind them
        {
            const float sx = __shfl_sync(0xffffffffu, cx, (lane & (PIX_PER_WARP - 1)) + PIX_PER_WARP);
            const float sy2 = __shfl_sync(0xffffffffu, cy, (lane & (PIX_PER_WARP - 1)) + PIX_PER_WARP);
            cx = lane < PIX_PER_WARP ? sx : nx;
            cy = lane < PIX_PER_WARP ? sy2 : ny;
            tile_advance(l_b, l_p0);
            load_coords_at(it + 3, l_b, l_p0, nx, ny);
            tile_advance(b, p0);
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
    constexpr int kBufs = 3;
    constexpr int SLOT = Q::SLOT;
    // Tile size.  32 pixels (one 128-byte output line per channel, 3-deep patch ring, two CTAs per SM) or 16 pixels with a
    // 2-deep ring (63 KB per CTA: three CTAs = 24 warps per SM, twice as many tiles to balance).  Measured per launch,
    // 16 vs 32 (profiles/README.md, round 2): config 1 B=1 4.5 / 5.4 us, config 2 B=1 7.1 / 8.1, B=4 23.1 / 21.6,
    // B=8 43.1 / 41.8, config 3 23.3 / 27.3, config 4 91.3 / 103.5 -- small grids and ragged images (N not a multiple of
    // 32) prefer 16, the large aligned ones 32.  RAFTCORR_LOOKUP_TILE=16|32 overrides.
    const char* tile_env = std::getenv("RAFTCORR_LOOKUP_TILE");  // read per launch: the tests drive both variants
    const int forced_tile = (tile_env && tile_env[0] == '1' && tile_env[1] == '6') ? 16
                          : (tile_env && tile_env[0] == '3' && tile_env[1] == '2') ? 32 : 0;
    int dev = 0;
    RC_CUDA(cudaGetDevice(&dev));
    static int sm_count[64] = {};
    int sms = (dev >= 0 && dev < 64) ? sm_count[dev] : 0;
    if (!sms) {
        RC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        if (dev >= 0 && dev < 64) sm_count[dev] = sms;
    }
    // RAFTCORR_LOOKUP_EXP=1|2: timing-only ablations of the bench shape, 32-pixel tiles (the results are NOT a lookup)
    const char* exp_env = (R == 4 && LG == 4) ? std::getenv("RAFTCORR_LOOKUP_EXP") : nullptr;
    const int exp_mode = (exp_env && (exp_env[0] == '1' || exp_env[0] == '2')) ? exp_env[0] - '0' : 0;
    const int64_t tiles32 = (int64_t)ceil_div(g.N, kPix) * g.B;
    const bool small_tile = !exp_mode && (forced_tile ? forced_tile == 16 : ((g.N % kPix) != 0 || tiles32 <= 4 * (int64_t)sms));
    const int bufs = small_tile ? 2 : kBufs, tile = small_tile ? 16 : kPix, per_sm = small_tile ? 3 : 2;
    const size_t smem = 128 + (size_t)kWarps * bufs * LG * SLOT + kWarps * kBarBytes + sizeof(float) * LG * Q::KK * (tile + 1);
    auto kern = small_tile ? lookup_fwd_quad_kernel<R, LG, 2, 0, 16, 3> : lookup_fwd_quad_kernel<R, LG, kBufs>;
    // the shared-memory opt-in is per function AND per device (nn.DataParallel drives several devices from one process)
    static bool attr_done[2][64] = {};
    if (dev < 0 || dev >= 64 || !attr_done[small_tile][dev]) {
        RC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (dev >= 0 && dev < 64) attr_done[small_tile][dev] = true;
    }
    if (exp_mode) {
        kern = exp_mode == 1 ? lookup_fwd_quad_kernel<R, LG, kBufs, (R == 4 && LG == 4) ? 1 : 0>
                             : lookup_fwd_quad_kernel<R, LG, kBufs, (R == 4 && LG == 4) ? 2 : 0>;
        RC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    const int tiles_per_img = ceil_div(g.N, tile);
    const int64_t total = (int64_t)tiles_per_img * g.B;
    RC_REQUIRE(total < (1LL << 31), "lookup: too many tiles");
    const int ctas = per_sm * sms;  // resident CTAs per SM (shared memory), each walking total / ctas tiles
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

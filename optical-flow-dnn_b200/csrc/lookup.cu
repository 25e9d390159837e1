// Correlation lookup (CorrBlock.__call__, raft/core/corr.py:45-91 + bilinear_sampler,
// raft/core/utils/utils.py:57-85) and its adjoint, for sm_100a.
//
// Forward, one launch for all levels (the reference issues ~60 launches per iteration):
//   out[b, l*K*K + i*K + j, h, w] = bilinear_zero_pad(V_l[b,(h,w)], x/2^l + i - r, y/2^l + j - r)
// HBM-bound gather.  The 4*K*K taps of one (pixel, level) are one (K+1)x(K+1) lattice patch with a
// single fractional offset, so a warp fetches the patch once (lane = lattice element, rows are
// contiguous 40-byte runs), parks it in shared memory, and evaluates the K*K outputs from it.
// Results are transposed through a [channels][32 pixels] shared tile so that global stores are
// full 128-byte lines along W of the [B, C, H, W] output (no NHWC->NCHW pass as in corr.py:91).
#include <cstdlib>
#include <cstring>

#include "lookup_common.cuh"

namespace raftcorr {

namespace {

template <int R>
struct Geo {
    static constexpr int K = 2 * R + 1;
    static constexpr int K1 = K + 1;
    static constexpr int NP = K1 * K1;            // lattice points per (pixel, level)
    static constexpr int NK = (NP + 31) / 32;     // lattice elements per lane
    static constexpr int KK = K * K;              // outputs per (pixel, level)
    static constexpr int NT = (KK + 31) / 32;     // outputs per lane
};

template <int R, int LG>
struct LaneGeo {
    int erow[Geo<R>::NK], ecol[Geo<R>::NK];
    __device__ __forceinline__ void init(int lane) {
#pragma unroll
        for (int k = 0; k < Geo<R>::NK; ++k) {
            int e = lane + 32 * k;
            erow[k] = e / Geo<R>::K1;
            ecol[k] = e - erow[k] * Geo<R>::K1;
        }
    }
};

// Fetch the zero-padded lattice patches of one source pixel for all levels of the group.
template <int R, int LG>
__device__ __forceinline__ void fetch_patches(const GroupView& g, int b, int p, float cx, float cy, int lane,
                                              const LaneGeo<R, LG>& lg, float (&v)[LG][Geo<R>::NK],
                                              float (&fx)[LG], float (&fy)[LG], int (&x0)[LG], int (&y0)[LG]) {
    using G = Geo<R>;
#pragma unroll
    for (int l = 0; l < LG; ++l) {
        const LevelView& lv = g.lvl[l];
        const float x = cx * g.inv_scale[l], y = cy * g.inv_scale[l];
        const float xf = floorf(x), yf = floorf(y);
        fx[l] = x - xf;
        fy[l] = y - yf;
        x0[l] = (int)clamp_coord(xf) - R;
        y0[l] = (int)clamp_coord(yf) - R;
        const float* img = lv.base + ((int64_t)b * g.N + p) * lv.img_stride;
#pragma unroll
        for (int k = 0; k < G::NK; ++k) {
            const int yy = y0[l] + lg.erow[k], xx = x0[l] + lg.ecol[k];
            const bool ok = (lane + 32 * k < G::NP) && yy >= 0 && yy < lv.h && xx >= 0 && xx < lv.w;
            v[l][k] = ok ? __ldg(img + lv.at(yy, xx)) : 0.f;
        }
    }
}

// ---- forward ---------------------------------------------------------------------------------
// Patch rows start on 32-byte sectors (the pitch is a multiple of 8 floats), so a (pixel, level)
// patch is K+1 rows x NS sectors.  Each lane copies 16-byte chunks of those sectors straight from
// global to shared memory with cp.async (no register staging; out-of-range chunks and row tails are
// zero-filled by the copy's src-size operand = grid_sample's zero padding).  Every warp keeps the
// copies of its NEXT pixel in flight while it evaluates the current one (double-buffered patches),
// which is what keeps enough bytes in flight to cover HBM latency at 2 CTAs/SM.  The sub-sector
// offset ox = x0 & 7 is applied when the K*K outputs read their four taps.
template <int R>
struct FwdGeo {
    static constexpr int K = 2 * R + 1;
    static constexpr int K1 = K + 1;
    static constexpr int NS = (7 + K1 + 7) / 8;        // 32-byte sectors a patch row can straddle
    static constexpr int PITCH = NS * 8;               // floats per staged patch row
    static constexpr int CHUNKS = NS * 2;              // 16-byte chunks per row
    static constexpr int TASKS = K1 * CHUNKS;          // cp.async tasks per (pixel, level)
    static constexpr int NTK = (TASKS + 31) / 32;      // per lane
    static constexpr int ROWS = (NTK * 32 + CHUNKS - 1) / CHUNKS;  // K1 rows + padding rows hit by idle tasks
    static constexpr int PATCH = ROWS * PITCH;
    static constexpr int KK = K * K;
    static constexpr int NT = (KK + 31) / 32;
};

__device__ __forceinline__ void cp_async16_zfill(uint32_t dst_smem, const void* src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Per-lane constants: which 16-byte chunks of a patch this lane copies.  Tasks past the end of the
// patch (the last partial warp-load) land in one padding row per level with a zero-byte copy, so the
// copy loop has no divergent branch.
template <int R, int LG>
struct FwdLane {
    int row[FwdGeo<R>::NTK], col[FwdGeo<R>::NTK];
    int goff[LG][FwdGeo<R>::NTK];   // row * pitch_l + col: float offset from the patch's first sector
    uint32_t dst[FwdGeo<R>::NTK];   // byte offset inside one level's staged patch
    __device__ __forceinline__ void init(int lane, const GroupView& g) {
        using G = FwdGeo<R>;
#pragma unroll
        for (int k = 0; k < G::NTK; ++k) {
            const int t = lane + 32 * k;
            row[k] = t / G::CHUNKS;
            col[k] = (t - row[k] * G::CHUNKS) * 4;
            dst[k] = (uint32_t)(row[k] * G::PITCH + col[k]) * 4u;
#pragma unroll
            for (int l = 0; l < LG; ++l) goff[l][k] = row[k] * g.lvl[l].pitch + col[k];
        }
    }
};

struct PixelGeo {  // per (pixel, level): fractional offsets and the sub-sector shift of the staged patch
    float fx, fy;
    int ox;
};

// issue the copies of one source pixel's patches (all levels of the group) into `patch_smem`
template <int R, int LG>
__device__ __forceinline__ void prefetch_pixel(const GroupView& g, int64_t bp, float cx, float cy,
                                               const FwdLane<R, LG>& ln, uint32_t patch_smem, PixelGeo (&pg)[LG]) {
    using G = FwdGeo<R>;
#pragma unroll
    for (int l = 0; l < LG; ++l) {
        const LevelView& lv = g.lvl[l];
        const float x = cx * g.inv_scale[l], y = cy * g.inv_scale[l];
        const int xi = __float2int_rd(x), yi = __float2int_rd(y);
        pg[l].fx = x - (float)xi;
        pg[l].fy = y - (float)yi;
        const int x0 = xi - R, y0 = yi - R;
        pg[l].ox = x0 & 7;
        const int xs = x0 & ~7;
        const float* rb = lv.base + bp * lv.img_stride + ((int64_t)y0 * lv.pitch + xs);
#pragma unroll
        for (int k = 0; k < G::NTK; ++k) {
            const int yy = y0 + ln.row[k], cc = xs + ln.col[k];
            const bool ok = ln.row[k] < G::K1 && (unsigned)yy < (unsigned)lv.h && (unsigned)cc < (unsigned)lv.w;
            int bytes = (lv.w - cc) * 4;  // row tail inside the pitch padding -> partial copy, rest zero-filled
            bytes = ok ? (bytes < 16 ? bytes : 16) : 0;
            cp_async16_zfill(patch_smem + (uint32_t)(l * G::PATCH) * 4u + ln.dst[k], ok ? rb + ln.goff[l][k] : lv.base,
                             bytes);
        }
    }
}

template <int R, int LG>
__global__ void __launch_bounds__(kWarps * 32, 2)
lookup_fwd_kernel(const GroupView g, const float* __restrict__ coords, float* __restrict__ out) {
    using G = FwdGeo<R>;
    constexpr int CH = LG * G::KK;
    constexpr int PIX_PER_WARP = kPix / kWarps;
    extern __shared__ __align__(16) float smem[];
    float* out_s = smem;                                       // [CH][kTilePitch]
    float* patch_s = smem + ((CH * kTilePitch + 3) & ~3);      // [kWarps][2][LG][PATCH], 16-byte aligned

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.y, p0 = blockIdx.x * kPix;
    const int N = g.N;

    FwdLane<R, LG> ln;
    ln.init(lane, g);
    int ooff[G::NT];
#pragma unroll
    for (int t = 0; t < G::NT; ++t) {
        const int c = lane + 32 * t;
        const int i = c / G::K, j = c - i * G::K;  // channel c = i*K + j: i along x, j along y
        ooff[t] = j * G::PITCH + i;
    }
    float* patch = patch_s + warp * (2 * LG * G::PATCH);
    const uint32_t patch_u32 = (uint32_t)__cvta_generic_to_shared(patch);

    // this warp's pixels: q = warp + kWarps * n.  Coordinates are clamped once: anything beyond
    // +-2^20 px is outside every level anyway and float->int stays defined (NaN -> -2^20).
    float cxs[PIX_PER_WARP], cys[PIX_PER_WARP];
#pragma unroll
    for (int n = 0; n < PIX_PER_WARP; ++n) {
        const int p = p0 + warp + kWarps * n;
        const bool v = p < N;
        cxs[n] = clamp_coord(v ? __ldg(coords + ((int64_t)b * 2 + 0) * N + p) : 0.f);
        cys[n] = clamp_coord(v ? __ldg(coords + ((int64_t)b * 2 + 1) * N + p) : 0.f);
    }
    PixelGeo cur[LG], nxt[LG];
    const int64_t bp0 = (int64_t)b * N + p0 + warp;
    if (p0 + warp < N) prefetch_pixel<R, LG>(g, bp0, cxs[0], cys[0], ln, patch_u32, nxt);
    cp_async_commit();

#pragma unroll
    for (int n = 0; n < PIX_PER_WARP; ++n) {
        const int q = warp + kWarps * n;
        const int p = p0 + q;
        if (p >= N) break;
        const int buf = n & 1;
#pragma unroll
        for (int l = 0; l < LG; ++l) cur[l] = nxt[l];
        if (n + 1 < PIX_PER_WARP && p + kWarps < N)
            prefetch_pixel<R, LG>(g, bp0 + kWarps * (n + 1), cxs[n + 1 < PIX_PER_WARP ? n + 1 : n],
                                  cys[n + 1 < PIX_PER_WARP ? n + 1 : n], ln,
                                  patch_u32 + (uint32_t)((buf ^ 1) * LG * G::PATCH) * 4u, nxt);
        cp_async_commit();
        cp_async_wait<1>();  // everything but the group just committed has landed
        __syncwarp();
        const float* pbuf = patch + buf * (LG * G::PATCH);
#pragma unroll
        for (int l = 0; l < LG; ++l) {
            const float ax = cur[l].fx, ay = cur[l].fy;
            const float* P = pbuf + l * G::PATCH + cur[l].ox;
#pragma unroll
            for (int t = 0; t < G::NT; ++t) {
                const int c = lane + 32 * t;
                if (c < G::KK) {
                    const int o = ooff[t];
                    const float top = P[o] + ax * (P[o + 1] - P[o]);
                    const float bot = P[o + G::PITCH] + ax * (P[o + G::PITCH + 1] - P[o + G::PITCH]);
                    out_s[(l * G::KK + c) * kTilePitch + q] = top + ay * (bot - top);
                }
            }
        }
        __syncwarp();  // the buffer is refilled by the prefetch issued in the next iteration
    }
    cp_async_wait<0>();
    __syncthreads();
    // channel-major write-out: one warp store = 32 consecutive pixels of one channel
    const int p = p0 + lane;
    if (p < N) {
        float* dst = out + ((int64_t)b * g.ctot + g.ch0) * N + p;
        for (int ch = warp; ch < CH; ch += kWarps) dst[(int64_t)ch * N] = out_s[ch * kTilePitch + lane];
    }
}

// Adjoint: dV_l[lattice] += sum of the <=4 outputs that read it, weighted.  Every (pixel, level)
// patch is owned by exactly one warp, so distinct lanes never hit the same address inside a
// launch; red.global.add keeps successive launches (one per refinement iteration) accumulating
// into the single persistent gradient pyramid without a read-modify-write round trip.
template <int R, int LG, bool DCOORDS>
__global__ void __launch_bounds__(kWarps * 32, 4)
lookup_bwd_kernel(const GroupView dg, const GroupView fg, const float* __restrict__ coords,
                  const float* __restrict__ grad_out, float* __restrict__ dcoords) {
    using G = Geo<R>;
    constexpr int CH = LG * G::KK;
    extern __shared__ float smem[];
    float* g_s = smem;                          // [CH][kTilePitch]
    float* patch_s = smem + CH * kTilePitch;    // [kWarps][LG][NP] (DCOORDS only)

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.y, p0 = blockIdx.x * kPix;
    const int N = dg.N;

    {
        const int p = p0 + lane;
        const float* src = grad_out + ((int64_t)b * dg.ctot + dg.ch0) * N + p;
        for (int ch = warp; ch < CH; ch += kWarps)
            g_s[ch * kTilePitch + lane] = (p < N) ? __ldg(src + (int64_t)ch * N) : 0.f;
    }
    __syncthreads();

    LaneGeo<R, LG> lg;
    lg.init(lane);

    for (int q = warp; q < kPix; q += kWarps) {
        const int p = p0 + q;
        if (p >= N) break;
        const float cx = __ldg(coords + ((int64_t)b * 2 + 0) * N + p);
        const float cy = __ldg(coords + ((int64_t)b * 2 + 1) * N + p);
        float accx = 0.f, accy = 0.f;
        if (DCOORDS) {
            float v[LG][G::NK], fx[LG], fy[LG];
            int x0[LG], y0[LG];
            fetch_patches<R, LG>(fg, b, p, cx, cy, lane, lg, v, fx, fy, x0, y0);
            float* patch = patch_s + warp * (LG * G::NP);
#pragma unroll
            for (int l = 0; l < LG; ++l)
#pragma unroll
                for (int k = 0; k < G::NK; ++k)
                    if (lane + 32 * k < G::NP) patch[l * G::NP + lane + 32 * k] = v[l][k];
            __syncwarp();
#pragma unroll
            for (int l = 0; l < LG; ++l) {
                const float* P = patch + l * G::NP;
#pragma unroll
                for (int t = 0; t < G::NT; ++t) {
                    const int c = lane + 32 * t;
                    if (c < G::KK) {
                        const int i = c / G::K, j = c - i * G::K;
                        const int o = j * G::K1 + i;
                        const float gg = g_s[(l * G::KK + c) * kTilePitch + q] * fg.inv_scale[l];
                        const float ddx = (1.f - fy[l]) * (P[o + 1] - P[o]) + fy[l] * (P[o + G::K1 + 1] - P[o + G::K1]);
                        const float ddy = (1.f - fx[l]) * (P[o + G::K1] - P[o]) + fx[l] * (P[o + G::K1 + 1] - P[o + 1]);
                        accx += gg * ddx;
                        accy += gg * ddy;
                    }
                }
            }
            __syncwarp();
        }
#pragma unroll
        for (int l = 0; l < LG; ++l) {
            const LevelView& lv = dg.lvl[l];
            const float x = cx * dg.inv_scale[l], y = cy * dg.inv_scale[l];
            const float xf = floorf(x), yf = floorf(y);
            const float fx = x - xf, fy = y - yf;
            const int x0 = (int)clamp_coord(xf) - R, y0 = (int)clamp_coord(yf) - R;
            float* img = lv.base + ((int64_t)b * N + p) * lv.img_stride;
            const float* gl = g_s + l * G::KK * kTilePitch + q;
#pragma unroll
            for (int k = 0; k < G::NK; ++k) {
                const int row = lg.erow[k], col = lg.ecol[k];
                const int yy = y0 + row, xx = x0 + col;
                if ((lane + 32 * k < G::NP) && yy >= 0 && yy < lv.h && xx >= 0 && xx < lv.w) {
                    float acc = 0.f;
                    if (col < G::K && row < G::K) acc += (1.f - fx) * (1.f - fy) * gl[(col * G::K + row) * kTilePitch];
                    if (col >= 1 && row < G::K) acc += fx * (1.f - fy) * gl[((col - 1) * G::K + row) * kTilePitch];
                    if (col < G::K && row >= 1) acc += (1.f - fx) * fy * gl[(col * G::K + row - 1) * kTilePitch];
                    if (col >= 1 && row >= 1) acc += fx * fy * gl[((col - 1) * G::K + row - 1) * kTilePitch];
                    atomicAdd(img + lv.at(yy, xx), acc);
                }
            }
        }
        if (DCOORDS) {
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) {
                accx += __shfl_xor_sync(0xffffffffu, accx, s);
                accy += __shfl_xor_sync(0xffffffffu, accy, s);
            }
            if (lane == 0) {
                dcoords[((int64_t)b * 2 + 0) * N + p] += accx;
                dcoords[((int64_t)b * 2 + 1) * N + p] += accy;
            }
        }
    }
}

template <int R, int LG>
int run_fwd(const GroupView& g, const float* coords, float* out, cudaStream_t s) {
    using G = FwdGeo<R>;
    const size_t smem = sizeof(float) * (((LG * G::KK * kTilePitch + 3) & ~3) + kWarps * 2 * LG * G::PATCH);
    auto kern = lookup_fwd_kernel<R, LG>;
    RC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(ceil_div(g.N, kPix), g.B);
    kern<<<grid, kWarps * 32, smem, s>>>(g, coords, out);
    RC_CUDA(cudaGetLastError());
    return 0;
}

template <int R, int LG>
int run_bwd(const GroupView& dg, const GroupView* fg, const float* coords, const float* grad_out, float* dcoords,
            cudaStream_t s) {
    using G = Geo<R>;
    dim3 grid(ceil_div(dg.N, kPix), dg.B);
    if (dcoords) {
        const size_t smem = sizeof(float) * (LG * G::KK * kTilePitch + kWarps * LG * G::NP);
        auto kern = lookup_bwd_kernel<R, LG, true>;
        RC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, kWarps * 32, smem, s>>>(dg, *fg, coords, grad_out, dcoords);
    } else {
        const size_t smem = sizeof(float) * (LG * G::KK * kTilePitch);
        auto kern = lookup_bwd_kernel<R, LG, false>;
        RC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, kWarps * 32, smem, s>>>(dg, dg, coords, grad_out, nullptr);
    }
    RC_CUDA(cudaGetLastError());
    return 0;
}

template <int R>
int dispatch_fwd(int lg, const GroupView& g, const float* coords, float* out, cudaStream_t s) {
    switch (lg) {
        case 1: return run_fwd<R, 1>(g, coords, out, s);
        case 2: return run_fwd<R, 2>(g, coords, out, s);
        case 3: return run_fwd<R, 3>(g, coords, out, s);
        default: return run_fwd<R, 4>(g, coords, out, s);
    }
}

template <int R>
int dispatch_bwd(int lg, const GroupView& dg, const GroupView* fg, const float* coords, const float* go, float* dc,
                 cudaStream_t s) {
    switch (lg) {
        case 1: return run_bwd<R, 1>(dg, fg, coords, go, dc, s);
        case 2: return run_bwd<R, 2>(dg, fg, coords, go, dc, s);
        case 3: return run_bwd<R, 3>(dg, fg, coords, go, dc, s);
        default: return run_bwd<R, 4>(dg, fg, coords, go, dc, s);
    }
}

}  // namespace

GroupView make_group(const PyramidView& pv, int l0, int lg, int radius) {
    GroupView g{};
    const int KK = (2 * radius + 1) * (2 * radius + 1);
    for (int l = 0; l < lg; ++l) {
        g.lvl[l] = pv.lvl[l0 + l];
        g.inv_scale[l] = 1.0f / (float)(1 << (l0 + l));
    }
    g.B = pv.B;
    g.N = pv.N;
    g.ch0 = l0 * KK;
    g.ctot = pv.L * KK;
    return g;
}

// RAFTCORR_LOOKUP=tma|cpasync selects the forward implementation (read once per call; default: tma).
bool use_tma_lookup() {
    const char* e = std::getenv("RAFTCORR_LOOKUP");
    return !(e && std::strcmp(e, "cpasync") == 0);
}

int launch_lookup_fwd_group_cpasync(int radius, int lg, const GroupView& g, const float* coords, float* out,
                                    cudaStream_t s) {
    RC_REQUIRE(!g.lvl[0].il, "lookup: RAFTCORR_LOOKUP=cpasync reads the row-major pyramid layout only (fp32 build mode)");
    switch (radius) {
        case 1: return dispatch_fwd<1>(lg, g, coords, out, s);
        case 2: return dispatch_fwd<2>(lg, g, coords, out, s);
        case 3: return dispatch_fwd<3>(lg, g, coords, out, s);
        case 4: return dispatch_fwd<4>(lg, g, coords, out, s);
        default: return fail("lookup: radius %d not supported (1..4)", radius);
    }
}

int launch_lookup_fwd(const PyramidView& pv, const float* coords, int radius, float* out, cudaStream_t s) {
    if (use_tma_lookup()) return launch_lookup_fwd_tma(pv, coords, radius, out, s);
    RC_REQUIRE(!pv.lvl[0].il, "lookup: RAFTCORR_LOOKUP=cpasync reads the row-major pyramid layout only (fp32 build mode)");
    for (int l0 = 0; l0 < pv.L; l0 += 4) {
        const int lg = pv.L - l0 < 4 ? pv.L - l0 : 4;
        GroupView g = make_group(pv, l0, lg, radius);
        int rc;
        switch (radius) {
            case 1: rc = dispatch_fwd<1>(lg, g, coords, out, s); break;
            case 2: rc = dispatch_fwd<2>(lg, g, coords, out, s); break;
            case 3: rc = dispatch_fwd<3>(lg, g, coords, out, s); break;
            case 4: rc = dispatch_fwd<4>(lg, g, coords, out, s); break;
            default: return fail("lookup: radius %d not supported (1..4)", radius);
        }
        if (rc) return rc;
    }
    return 0;
}

int launch_lookup_bwd(const float* grad_out, const float* coords, const PyramidView* pv_fwd,
                      const PyramidView& dpv, int radius, float* dcoords, cudaStream_t s) {
    for (int l0 = 0; l0 < dpv.L; l0 += 4) {
        const int lg = dpv.L - l0 < 4 ? dpv.L - l0 : 4;
        GroupView dg = make_group(dpv, l0, lg, radius);
        GroupView fg{};
        if (dcoords) fg = make_group(*pv_fwd, l0, lg, radius);
        int rc;
        switch (radius) {
            case 1: rc = dispatch_bwd<1>(lg, dg, &fg, coords, grad_out, dcoords, s); break;
            case 2: rc = dispatch_bwd<2>(lg, dg, &fg, coords, grad_out, dcoords, s); break;
            case 3: rc = dispatch_bwd<3>(lg, dg, &fg, coords, grad_out, dcoords, s); break;
            case 4: rc = dispatch_bwd<4>(lg, dg, &fg, coords, grad_out, dcoords, s); break;
            default: return fail("lookup_bwd: radius %d not supported (1..4)", radius);
        }
        if (rc) return rc;
    }
    return 0;
}

}  // namespace raftcorr

// On-the-fly correlation (AlternateCorrBlock, raft/core/corr.py:119-167, and the native
// alt_cuda_corr extension, raft/alt_cuda_corr/correlation_kernel.cu:18-256) for sm_100a.
//
// No volume is materialised: for every source pixel and level the (K+1)x(K+1) lattice of dot
// products <f1[p], f2_l[floor(y)-r+iy, floor(x)-r+ix]> is computed and bilinearly splatted into
// K*K outputs (channel iy + K*ix, correlation_kernel.cu:92-95 == CorrBlock's i*K+j order).
//
// Differences from the reference kernel (same numbers, different machine mapping):
//   * all L levels in ONE launch, NHWC feature tensors prepared once (not per call per level);
//   * a warp owns a source pixel: its f1 vector lives in registers, each lattice point is one
//     coalesced 128-bit-per-lane read of the f2 vector, 32 partial dots are reduced with a
//     31-shuffle butterfly transpose instead of 32 x 5 shuffles;
//   * outputs are accumulated on chip and written once, channel-major, through a shared tile
//     (the reference read-modify-writes global memory once per 32-channel chunk, :102-114);
//   * launched on the caller's stream with the caller's device, zero padding by predication.
#include <climits>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace raftcorr {

namespace {

constexpr int kWarps = 8;
constexpr int kPix = 32;
constexpr int kTilePitch = 33;
constexpr int kMaxVec = 4;  // float4 per lane -> C <= 512

__device__ __forceinline__ float clamp_coord(float v) { return fminf(fmaxf(v, -1048576.f), 1048576.f); }

struct AltParams {
    const float* f1;  // [B][N1][C]
    AltLevels lv;
    CoordsView cv;
    int B, NC, N1, C;  // NC: coordinate sets per pixel (the reference's N, always 1 from Python)
    float scale;       // 1/sqrt(C) for the fused path, 1 for the raw alt_cuda_corr surface
    float inv_scale[RAFTCORR_MAX_LEVELS];
};

// Sum 32 per-lane partials across the warp; lane i ends up with the total of element i.
__device__ __forceinline__ float transpose_reduce32(float (&p)[32], int lane) {
#pragma unroll
    for (int half = 16; half >= 1; half >>= 1) {
        const bool hi = lane & half;
#pragma unroll
        for (int i = 0; i < half; ++i) {
            const float send = hi ? p[i] : p[i + half];
            const float keep = hi ? p[i + half] : p[i];
            p[i] = keep + __shfl_xor_sync(0xffffffffu, send, half);
        }
    }
    return p[0];
}

template <int R, int NV>
__global__ void __launch_bounds__(kWarps * 32, 2)
alt_fwd_kernel(const AltParams P, float* __restrict__ out) {
    const int LG = P.lv.L;
    constexpr int K = 2 * R + 1, K1 = K + 1, NP = K1 * K1, KK = K * K, NT = (KK + 31) / 32;
    constexpr int NCH = (NP + 31) / 32;
    const int CH = LG * KK;
    extern __shared__ float smem[];
    float* out_s = smem;                       // [CH][kTilePitch]
    float* patch_s = smem + CH * kTilePitch;   // [kWarps][LG][NCH*32]

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int bn = blockIdx.y, b = bn / P.NC, p0 = blockIdx.x * kPix;
    const int nvec = P.C >> 2;
    float* patch = patch_s + warp * (LG * NCH * 32);

    int ooff[NT];
#pragma unroll
    for (int t = 0; t < NT; ++t) {
        const int c = lane + 32 * t;
        const int ix = c / K, iy = c - ix * K;  // channel = iy + K*ix
        ooff[t] = iy * K1 + ix;
    }

    for (int q = warp; q < kPix; q += kWarps) {
        const int p = p0 + q;
        if (p >= P.N1) break;
        const float cx = __ldg(P.cv.ptr + bn * P.cv.batch_stride + p * P.cv.pix_stride);
        const float cy = __ldg(P.cv.ptr + bn * P.cv.batch_stride + P.cv.comp_stride + p * P.cv.pix_stride);
        float4 a[NV];
        const float4* f1v = reinterpret_cast<const float4*>(P.f1 + ((int64_t)b * P.N1 + p) * P.C);
#pragma unroll
        for (int j = 0; j < NV; ++j)
            a[j] = (lane + 32 * j < nvec) ? __ldg(f1v + lane + 32 * j) : make_float4(0.f, 0.f, 0.f, 0.f);

#pragma unroll 1
        for (int l = 0; l < LG; ++l) {
            const int Hl = P.lv.h[l], Wl = P.lv.w[l];
            const float x = cx * P.inv_scale[l], y = cy * P.inv_scale[l];
            const float xf = floorf(x), yf = floorf(y);
            const int x0 = (int)clamp_coord(xf) - R, y0 = (int)clamp_coord(yf) - R;
            const float4* f2b = reinterpret_cast<const float4*>(P.lv.f2[l] + (int64_t)b * Hl * Wl * P.C);
#pragma unroll 1
            for (int ch = 0; ch < NCH; ++ch) {
                float part[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const int e = ch * 32 + i;
                    const int row = e / K1, col = e - row * K1;
                    const int yy = y0 + row, xx = x0 + col;
                    const bool ok = e < NP && yy >= 0 && yy < Hl && xx >= 0 && xx < Wl;
                    const float4* src = f2b + ((int64_t)yy * Wl + xx) * nvec;
                    float s = 0.f;
#pragma unroll
                    for (int j = 0; j < NV; ++j) {
                        if (ok && lane + 32 * j < nvec) {
                            const float4 v = __ldg(src + lane + 32 * j);
                            s = fmaf(a[j].x, v.x, s);
                            s = fmaf(a[j].y, v.y, s);
                            s = fmaf(a[j].z, v.z, s);
                            s = fmaf(a[j].w, v.w, s);
                        }
                    }
                    part[i] = s;
                }
                patch[l * NCH * 32 + ch * 32 + lane] = transpose_reduce32(part, lane);
            }
        }
        __syncwarp();
#pragma unroll 1
        for (int l = 0; l < LG; ++l) {
            const float* S = patch + l * NCH * 32;
            const float xs = cx * P.inv_scale[l], ys = cy * P.inv_scale[l];
            const float dx = xs - floorf(xs), dy = ys - floorf(ys);
#pragma unroll
            for (int t = 0; t < NT; ++t) {
                const int c = lane + 32 * t;
                if (c < KK) {
                    const int o = ooff[t];
                    // se*(1-dy)(1-dx) + sw*(1-dy)dx + ne*dy(1-dx) + nw*dy*dx  (correlation_kernel.cu:97-100)
                    const float top = S[o] + dx * (S[o + 1] - S[o]);
                    const float bot = S[o + K1] + dx * (S[o + K1 + 1] - S[o + K1]);
                    out_s[(l * KK + c) * kTilePitch + q] = (top + dy * (bot - top)) * P.scale;
                }
            }
        }
        __syncwarp();
    }
    __syncthreads();
    const int p = p0 + lane;
    if (p < P.N1) {
        float* dst = out + (int64_t)bn * CH * P.N1 + p;
        for (int ch = warp; ch < CH; ch += kWarps) dst[(int64_t)ch * P.N1] = out_s[ch * kTilePitch + lane];
    }
}

// Sum 64 per-lane partials (index c = 0..63) across the warp; lane i ends up with the totals of
// elements 2i (returned in p[0]) and 2i+1 (p[1]).  62 shuffles for 64 sums.
__device__ __forceinline__ void transpose_reduce64(float (&p)[64], int lane) {
#pragma unroll
    for (int off = 16, n = 32; off >= 1; off >>= 1, n >>= 1) {
        const bool hi = lane & off;
#pragma unroll
        for (int j = 0; j < n; ++j) {
            const float send = hi ? p[j] : p[j + n];
            const float keep = hi ? p[j + n] : p[j];
            p[j] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
}

// Grouped forward: a warp owns FOUR consecutive source pixels.  When the flow is locally smooth their
// (K+1)^2 windows overlap almost completely, so the warp walks the UNION box once: every f2 vector is
// fetched once and dotted with all four f1 vectors held in registers (4x fewer L1/L2 bytes -- the
// single-pixel kernel above is bound by exactly that traffic).  Partial dots of 8 lattice points x 4
// pixels are reduced with a 31-shuffle transpose; each lane then drops its sum into the per-pixel
// patch it belongs to.  If the windows are far apart (union > 2x one window) the pixels are walked
// one by one through the same code path, so arbitrary coordinates stay correct.
template <int R, int NV>
__global__ void __launch_bounds__(kWarps * 32, 2)
alt_fwd_group_kernel(const AltParams P, float* __restrict__ out) {
    const int LG = P.lv.L;
    constexpr int K = 2 * R + 1, K1 = K + 1, NP = K1 * K1, KK = K * K, NT = (KK + 31) / 32;
    constexpr int G = 4;           // pixels per warp: kPix == kWarps * G
    constexpr int UMAX = 2 * NP;   // largest union box walked in one go
    const int CH = LG * KK;
    extern __shared__ float smem[];
    float* out_s = smem;                       // [CH][kTilePitch]
    float* patch_s = smem + CH * kTilePitch;   // [kWarps][G][LG][NP]

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int bn = blockIdx.y, b = bn / P.NC, p0 = blockIdx.x * kPix + warp * G;
    const int nvec = P.C >> 2;
    float* patch = patch_s + warp * (G * LG * NP);

    int ooff[NT];
#pragma unroll
    for (int t = 0; t < NT; ++t) {
        const int c = lane + 32 * t;
        const int ix = c / K, iy = c - ix * K;  // channel = iy + K*ix
        ooff[t] = iy * K1 + ix;
    }

    float cx[G], cy[G];
    float4 a[G][NV];
    bool valid[G];
#pragma unroll
    for (int g = 0; g < G; ++g) {
        const int p = p0 + g;
        valid[g] = p < P.N1;
        const int pc = valid[g] ? p : P.N1 - 1;
        cx[g] = clamp_coord(__ldg(P.cv.ptr + bn * P.cv.batch_stride + pc * P.cv.pix_stride));
        cy[g] = clamp_coord(__ldg(P.cv.ptr + bn * P.cv.batch_stride + P.cv.comp_stride + pc * P.cv.pix_stride));
        const float4* f1v = reinterpret_cast<const float4*>(P.f1 + ((int64_t)b * P.N1 + pc) * P.C);
#pragma unroll
        for (int j = 0; j < NV; ++j)
            a[g][j] = (lane + 32 * j < nvec) ? __ldg(f1v + lane + 32 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (p0 < P.N1) {
#pragma unroll 1
        for (int l = 0; l < LG; ++l) {
            const int Hl = P.lv.h[l], Wl = P.lv.w[l];
            int x0[G], y0[G];
            int ux0 = INT_MAX, uy0 = INT_MAX, ux1 = INT_MIN, uy1 = INT_MIN;
#pragma unroll
            for (int g = 0; g < G; ++g) {
                x0[g] = __float2int_rd(cx[g] * P.inv_scale[l]) - R;
                y0[g] = __float2int_rd(cy[g] * P.inv_scale[l]) - R;
                ux0 = min(ux0, x0[g]); uy0 = min(uy0, y0[g]);
                ux1 = max(ux1, x0[g] + K1); uy1 = max(uy1, y0[g] + K1);
            }
            const bool joint = (int64_t)(ux1 - ux0) * (uy1 - uy0) <= UMAX;
            const float4* f2b = reinterpret_cast<const float4*>(P.lv.f2[l] + (int64_t)b * Hl * Wl * P.C);
            const int qsel = lane >> 3;  // pixel whose sums this lane receives after the transpose
            const int qx0 = qsel == 0 ? x0[0] : qsel == 1 ? x0[1] : qsel == 2 ? x0[2] : x0[3];
            const int qy0 = qsel == 0 ? y0[0] : qsel == 1 ? y0[1] : qsel == 2 ? y0[2] : y0[3];
            float* qpatch = patch + (qsel * LG + l) * NP;
            const int nbox = joint ? 1 : G;
#pragma unroll 1
            for (int bx = 0; bx < nbox; ++bx) {
                int bx0 = ux0, by0 = uy0, bw = ux1 - ux0, bh = uy1 - uy0;
                if (!joint) {
                    bx0 = bx == 0 ? x0[0] : bx == 1 ? x0[1] : bx == 2 ? x0[2] : x0[3];
                    by0 = bx == 0 ? y0[0] : bx == 1 ? y0[1] : bx == 2 ? y0[2] : y0[3];
                    bw = K1; bh = K1;
                }
                const int cnt = bw * bh;
                int row = 0, col = 0;  // walking position inside the box (warp-uniform)
#pragma unroll 1
                for (int e0 = 0; e0 < cnt; e0 += 8) {
                    float part[32];  // [pixel g][point i]: g * 8 + i
#pragma unroll
                    for (int hb = 0; hb < 2; ++hb) {
                        // four points at a time: all their loads are issued before the FMAs that use them
                        float4 v[4][NV];
                        bool okp[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int yy = by0 + row, xx = bx0 + col;
                            okp[i] = e0 + hb * 4 + i < cnt && yy >= 0 && yy < Hl && xx >= 0 && xx < Wl;
                            const float4* src = okp[i] ? f2b + ((int64_t)yy * Wl + xx) * nvec : f2b;
#pragma unroll
                            for (int j = 0; j < NV; ++j)
                                v[i][j] = (lane + 32 * j < nvec) ? __ldg(src + lane + 32 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
                            if (++col == bw) { col = 0; ++row; }
                        }
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
#pragma unroll
                            for (int g = 0; g < G; ++g) {
                                float sg = 0.f;
#pragma unroll
                                for (int j = 0; j < NV; ++j) {
                                    sg = fmaf(a[g][j].x, v[i][j].x, sg);
                                    sg = fmaf(a[g][j].y, v[i][j].y, sg);
                                    sg = fmaf(a[g][j].z, v[i][j].z, sg);
                                    sg = fmaf(a[g][j].w, v[i][j].w, sg);
                                }
                                part[g * 8 + hb * 4 + i] = okp[i] ? sg : 0.f;
                            }
                        }
                    }
                    // lane -> (pixel lane>>3, point e0 + (lane&7)) after the 31-shuffle transpose
                    const float sum = transpose_reduce32(part, lane);
                    const int e = e0 + (lane & 7);
                    if (e < cnt) {
                        const int er = e / bw, ec = e - er * bw;
                        const int wr = by0 + er - qy0, wc = bx0 + ec - qx0;  // position inside pixel qsel's window
                        if (wr >= 0 && wr < K1 && wc >= 0 && wc < K1) qpatch[wr * K1 + wc] = sum;
                    }
                }
            }
        }
    }
    __syncwarp();
#pragma unroll
    for (int g = 0; g < G; ++g) {
        if (!valid[g]) continue;
        const int q = warp * G + g;
#pragma unroll 1
        for (int l = 0; l < LG; ++l) {
            const float* S = patch + (g * LG + l) * NP;
            const float xs = cx[g] * P.inv_scale[l], ys = cy[g] * P.inv_scale[l];
            const float dx = xs - floorf(xs), dy = ys - floorf(ys);
#pragma unroll
            for (int t = 0; t < NT; ++t) {
                const int c = lane + 32 * t;
                if (c < KK) {
                    const int o = ooff[t];
                    const float top = S[o] + dx * (S[o + 1] - S[o]);
                    const float bot = S[o + K1] + dx * (S[o + K1 + 1] - S[o + K1]);
                    out_s[(l * KK + c) * kTilePitch + q] = (top + dy * (bot - top)) * P.scale;
                }
            }
        }
    }
    __syncthreads();
    const int p = blockIdx.x * kPix + lane;
    if (p < P.N1) {
        float* dst = out + (int64_t)bn * CH * P.N1 + p;
        for (int ch = warp; ch < CH; ch += kWarps) dst[(int64_t)ch * P.N1] = out_s[ch * kTilePitch + lane];
    }
}

// Backward (correlation_kernel.cu:122-256 restated): g_e = sum of the <=4 output gradients that a
// lattice point feeds, then df1[p] += g_e * f2[e] (registers, one vector add per pixel) and
// df2[e] += g_e * f1[p] (vector red.global.add.v4.f32 instead of one scalar atomic per channel).
template <int R, int NV>
__global__ void __launch_bounds__(kWarps * 32, 2)
alt_bwd_kernel(const AltParams P, const float* __restrict__ grad_out, float* __restrict__ df1) {
    const int LG = P.lv.L;
    constexpr int K = 2 * R + 1, K1 = K + 1, NP = K1 * K1, KK = K * K;
    constexpr int NCH = (NP + 31) / 32;
    const int CH = LG * KK;
    extern __shared__ float smem[];
    float* g_s = smem;  // [CH][kTilePitch]

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int bn = blockIdx.y, b = bn / P.NC, p0 = blockIdx.x * kPix;
    const int nvec = P.C >> 2;
    {
        const int p = p0 + lane;
        const float* src = grad_out + (int64_t)bn * CH * P.N1 + p;
        for (int ch = warp; ch < CH; ch += kWarps)
            g_s[ch * kTilePitch + lane] = (p < P.N1) ? __ldg(src + (int64_t)ch * P.N1) * P.scale : 0.f;
    }
    __syncthreads();

    for (int q = warp; q < kPix; q += kWarps) {
        const int p = p0 + q;
        if (p >= P.N1) break;
        const float cx = __ldg(P.cv.ptr + bn * P.cv.batch_stride + p * P.cv.pix_stride);
        const float cy = __ldg(P.cv.ptr + bn * P.cv.batch_stride + P.cv.comp_stride + p * P.cv.pix_stride);
        float4 a[NV], da[NV];
        const float4* f1v = reinterpret_cast<const float4*>(P.f1 + ((int64_t)b * P.N1 + p) * P.C);
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            a[j] = (lane + 32 * j < nvec) ? __ldg(f1v + lane + 32 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
            da[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll 1
        for (int l = 0; l < LG; ++l) {
            const int Hl = P.lv.h[l], Wl = P.lv.w[l];
            const float x = cx * P.inv_scale[l], y = cy * P.inv_scale[l];
            const float xf = floorf(x), yf = floorf(y);
            const float dx = x - xf, dy = y - yf;
            const int x0 = (int)clamp_coord(xf) - R, y0 = (int)clamp_coord(yf) - R;
            const float4* f2b = reinterpret_cast<const float4*>(P.lv.f2[l] + (int64_t)b * Hl * Wl * P.C);
            float4* d2b = reinterpret_cast<float4*>(P.lv.df2[l] + (int64_t)b * Hl * Wl * P.C);
            const float* gl = g_s + l * KK * kTilePitch + q;
#pragma unroll 1
            for (int ch = 0; ch < NCH; ++ch) {
                // lane e: gradient reaching lattice point e
                const int e = ch * 32 + lane;
                const int row = e / K1, col = e - row * K1;  // row = iy, col = ix
                float ge = 0.f;
                if (e < NP) {
                    if (row < K && col < K) ge += (1.f - dy) * (1.f - dx) * gl[(row + K * col) * kTilePitch];          // se
                    if (row < K && col >= 1) ge += (1.f - dy) * dx * gl[(row + K * (col - 1)) * kTilePitch];          // sw
                    if (row >= 1 && col < K) ge += dy * (1.f - dx) * gl[(row - 1 + K * col) * kTilePitch];            // ne
                    if (row >= 1 && col >= 1) ge += dy * dx * gl[(row - 1 + K * (col - 1)) * kTilePitch];             // nw
                }
#pragma unroll 4
                for (int i = 0; i < 32; ++i) {
                    const int ee = ch * 32 + i;
                    if (ee >= NP) break;
                    const float g = __shfl_sync(0xffffffffu, ge, i);
                    const int r2 = ee / K1, c2 = ee - r2 * K1;
                    const int yy = y0 + r2, xx = x0 + c2;
                    if (yy < 0 || yy >= Hl || xx < 0 || xx >= Wl || g == 0.f) continue;  // warp-uniform
                    const int64_t off = ((int64_t)yy * Wl + xx) * nvec;
#pragma unroll
                    for (int j = 0; j < NV; ++j) {
                        if (lane + 32 * j < nvec) {
                            const float4 v = __ldg(f2b + off + lane + 32 * j);
                            da[j].x = fmaf(g, v.x, da[j].x);
                            da[j].y = fmaf(g, v.y, da[j].y);
                            da[j].z = fmaf(g, v.z, da[j].z);
                            da[j].w = fmaf(g, v.w, da[j].w);
                            atomicAdd(d2b + off + lane + 32 * j,
                                      make_float4(g * a[j].x, g * a[j].y, g * a[j].z, g * a[j].w));
                        }
                    }
                }
            }
        }
        float4* d1 = reinterpret_cast<float4*>(df1 + ((int64_t)b * P.N1 + p) * P.C);
#pragma unroll
        for (int j = 0; j < NV; ++j)
            if (lane + 32 * j < nvec) atomicAdd(d1 + lane + 32 * j, da[j]);
    }
}

template <int R, int NV>
int run_alt_fwd(const AltParams& P, float* out, cudaStream_t s) {
    constexpr int K = 2 * R + 1, NP = (K + 1) * (K + 1), NCH = (NP + 31) / 32;
    const int LG = P.lv.L;
    // RAFTCORR_ALT=single selects the one-pixel-per-warp kernel (A/B switch); default: grouped
    const char* e = std::getenv("RAFTCORR_ALT");
    const bool single = e && std::strcmp(e, "single") == 0;
    const size_t smem = single ? sizeof(float) * (LG * K * K * kTilePitch + kWarps * LG * NCH * 32)
                               : sizeof(float) * (LG * K * K * kTilePitch + kWarps * 4 * LG * NP);
    auto kern = single ? alt_fwd_kernel<R, NV> : alt_fwd_group_kernel<R, NV>;
    RC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(ceil_div(P.N1, kPix), P.B * P.NC);
    kern<<<grid, kWarps * 32, smem, s>>>(P, out);
    RC_CUDA(cudaGetLastError());
    return 0;
}

template <int R, int NV>
int run_alt_bwd(const AltParams& P, const float* go, float* df1, cudaStream_t s) {
    constexpr int K = 2 * R + 1;
    const size_t smem = sizeof(float) * (P.lv.L * K * K * kTilePitch);
    auto kern = alt_bwd_kernel<R, NV>;
    RC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(ceil_div(P.N1, kPix), P.B * P.NC);
    kern<<<grid, kWarps * 32, smem, s>>>(P, go, df1);
    RC_CUDA(cudaGetLastError());
    return 0;
}

AltParams make_params(const float* f1, const AltLevels& lv, const CoordsView& cv, int B, int NC, int N1, int C,
                      float scale) {
    AltParams P{};
    P.f1 = f1;
    P.cv = cv;
    P.B = B;
    P.NC = NC;
    P.N1 = N1;
    P.C = C;
    P.scale = scale;
    P.lv = lv;
    for (int l = 0; l < lv.L; ++l) P.inv_scale[l] = 1.0f / (float)(1 << l);
    return P;
}

int check_alt(const AltLevels& lv, int C, int radius) {
    if (lv.L < 1 || lv.L > 4) return fail("alt: 1..4 levels per call supported (got %d)", lv.L);
    if (C % 4 != 0 || C > 128 * kMaxVec || C < 4) return fail("alt: C must be a multiple of 4 in [4, 512] (got %d)", C);
    if (radius < 1 || radius > 4) return fail("alt: radius %d not supported (1..4)", radius);
    return 0;
}

}  // namespace

// out: [B*NC][L*K*K][N1]
int launch_alt_fwd(const float* f1_nhwc, const AltLevels& lv, const CoordsView& cv, int B, int NC, int H1, int W1,
                   int C, int radius, float scale, float* out, cudaStream_t s) {
    if (int rc = check_alt(lv, C, radius)) return rc;
    AltParams P = make_params(f1_nhwc, lv, cv, B, NC, H1 * W1, C, scale);
    const int nv = C <= 128 ? 1 : (C <= 256 ? 2 : 4);
#define RC_ALT_CASE(R_)                                            \
    case R_:                                                       \
        if (nv == 1) return run_alt_fwd<R_, 1>(P, out, s);         \
        if (nv == 2) return run_alt_fwd<R_, 2>(P, out, s);         \
        return run_alt_fwd<R_, 4>(P, out, s);
    switch (radius) {
        RC_ALT_CASE(1) RC_ALT_CASE(2) RC_ALT_CASE(3) RC_ALT_CASE(4)
    }
#undef RC_ALT_CASE
    return fail("alt fwd: unreachable");
}

int launch_alt_bwd(const float* f1_nhwc, const AltLevels& lv, const CoordsView& cv, const float* grad_out, int B,
                   int NC, int H1, int W1, int C, int radius, float scale, float* df1_nhwc, cudaStream_t s) {
    if (int rc = check_alt(lv, C, radius)) return rc;
    AltParams P = make_params(f1_nhwc, lv, cv, B, NC, H1 * W1, C, scale);
    const int nv = C <= 128 ? 1 : (C <= 256 ? 2 : 4);
#define RC_ALT_CASE(R_)                                                       \
    case R_:                                                                  \
        if (nv == 1) return run_alt_bwd<R_, 1>(P, grad_out, df1_nhwc, s);     \
        if (nv == 2) return run_alt_bwd<R_, 2>(P, grad_out, df1_nhwc, s);     \
        return run_alt_bwd<R_, 4>(P, grad_out, df1_nhwc, s);
    switch (radius) {
        RC_ALT_CASE(1) RC_ALT_CASE(2) RC_ALT_CASE(3) RC_ALT_CASE(4)
    }
#undef RC_ALT_CASE
    return fail("alt bwd: unreachable");
}

}  // namespace raftcorr

// Convex 8x upsampling of the flow: RAFT.upsample_flow, raft/core/raft.py:70-81, called after
// EVERY refinement iteration (raft.py:161-176; twice per iteration by the uncertainty model, with the same mask).
//
//   out[b, c, 8h+i, 8w+j] = sum_k softmax_k(mask[b, k*64 + i*8 + j, h, w]) * 8 * flow[b, c, h + k/3 - 1, w + k%3 - 1]
//   (k = 0..8 over the 3x3 neighbourhood in F.unfold's order, zeros outside the image)
//
// The reference materialises softmax(mask), unfold(8*flow), their product and the permuted result: ~5 passes over a
// [B, 576, H, W] tensor (16 MB per frame pair at 55x128 -- as many bytes per iteration as the correlation lookup).
// Here the mask is read exactly once: one CTA owns 32 consecutive pixels of one feature row; warp i handles sub-row
// i, lanes run along the pixels (every mask load is one 128-byte line), the 9-way softmax and the weighted sum stay
// in registers, and the [C][8][256] output tile is transposed through shared memory so that the eight output rows
// leave as contiguous 1 KB runs.  Up to 4 channels share the pass (flow + variance of the uncertainty model).
// The backward recomputes the softmax from the mask (one read) and writes d_mask once; d_flow is accumulated in
// shared memory and leaves with one red.global.add per neighbourhood element.
#include <cstdlib>

#include "common.cuh"

namespace raftcorr {

namespace {

constexpr int kUpPix = 32;      // pixels per CTA (one warp lane each)
constexpr int kUpMaxC = 4;     // flow (2) + flow variance (2) of the uncertainty model share one pass

__device__ __forceinline__ float ld_stream(const float* p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

// element `off` (32-bit) behind a 64-bit base: ONE IMAD.WIDE.U32 per address (the compiler's own code for
// base + 64-bit index was 5-10 integer instructions per load)
__device__ __forceinline__ float ld_stream_off(const float* base, unsigned off) {
    float v;
    asm volatile("{ .reg .u64 a; mad.wide.u32 a, %2, 4, %1; ld.global.nc.L1::no_allocate.f32 %0, [a]; }"
                 : "=f"(v) : "l"(base), "r"(off));
    return v;
}
__device__ __forceinline__ void st_off(float* base, unsigned off, float v) {
    asm volatile("{ .reg .u64 a; mad.wide.u32 a, %1, 4, %0; st.global.f32 [a], %2; }" :: "l"(base), "r"(off), "f"(v) : "memory");
}

__device__ __forceinline__ void st_v4_off(float* base, unsigned off, float4 v) {
    asm volatile("{ .reg .u64 a; mad.wide.u32 a, %1, 4, %0; st.global.v4.f32 [a], {%2, %3, %4, %5}; }"
                 :: "l"(base), "r"(off), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 ld_v4_off(const float* base, unsigned off) {
    float4 v;
    asm volatile("{ .reg .u64 a; mad.wide.u32 a, %5, 4, %4; ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [a]; }"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(base), "r"(off));
    return v;
}

// exp2 / reciprocal on the SFU (one MUFU each).  ex2.approx is good to 2^-22 relative; the softmax argument is formed as
// fma(m, log2e, -max * log2e), whose rounding of max * log2e is common to all nine terms and cancels in the normalisation.
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
constexpr float kLog2e = 1.4426950408889634f;

// flow neighbourhood of the CTA's pixels: rows h-1..h+1, columns w0-1..w0+32, times 8, zero outside the image
template <int C>
__device__ __forceinline__ void load_flow_tile(float (*fs)[3][kUpPix + 2], const float* __restrict__ flow, int b, int h, int w0,
                                               int H, int W) {
    for (int idx = threadIdx.x; idx < C * 3 * (kUpPix + 2); idx += blockDim.x) {
        const int x = idx % (kUpPix + 2);
        const int r = (idx / (kUpPix + 2)) % 3;
        const int c = idx / (3 * (kUpPix + 2));
        const int yy = h + r - 1, xx = w0 + x - 1;
        fs[c][r][x] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? 8.0f * __ldg(flow + (((int64_t)b * C + c) * H + yy) * W + xx) : 0.f;
    }
}

template <int C>
__global__ void __launch_bounds__(256, 2)
upsample_fwd_kernel(const float* __restrict__ flow, const float* __restrict__ mask, float* __restrict__ out, int H, int W) {
    __shared__ float fs[C][3][kUpPix + 2];
    __shared__ __align__(16) float os[C][8][8 * kUpPix + 4];  // [c][i][8 * pixel + j]; rows are read back as float4
    const int b = blockIdx.z, h = blockIdx.y, w0 = blockIdx.x * kUpPix;
    const int lane = threadIdx.x & 31, i = threadIdx.x >> 5;
    const int w = w0 + lane;
    const int64_t HW = (int64_t)H * W;
    // all 72 mask values of this thread are requested first (they do not depend on the flow tile): one DRAM round
    // trip per CTA, overlapped with the flow-tile load
    float mm[8][9];
    if (w < W) {
        const float* mp = mask + ((int64_t)b * 576 + i * 8) * HW + (int64_t)h * W + w;  // channel k*64 + i*8 + j
        // channel offsets as 32-bit element counts (576 * H * W < 2^31 is checked by the launcher): one IMAD + one
        // IMAD.WIDE (ld_stream_off) per load instead of a 64-bit multiply-add chain -- address arithmetic was 42 % of this kernel's
        // instructions (ncu: issue slots 55 % busy at 25 % occupancy)
        const unsigned hw = (unsigned)HW;
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int k = 0; k < 9; ++k) mm[j][k] = ld_stream_off(mp, (unsigned)(k * 64 + j) * hw);
    }
    load_flow_tile<C>(fs, flow, b, h, w0, H, W);
    __syncthreads();
    if (w < W) {
        float nb[C][9];
#pragma unroll
        for (int c = 0; c < C; ++c)
#pragma unroll
            for (int k = 0; k < 9; ++k) nb[c][k] = fs[c][k / 3][lane + k % 3];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float* m = mm[j];
            float mx = m[0];
#pragma unroll
            for (int k = 1; k < 9; ++k) mx = fmaxf(mx, m[k]);
            float den = 0.f;
            const float mxl = mx * kLog2e;
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                m[k] = ex2_approx(fmaf(m[k], kLog2e, -mxl));
                den += m[k];
            }
            const float inv = rcp_approx(den);
#pragma unroll
            for (int c = 0; c < C; ++c) {
                float acc = 0.f;
#pragma unroll
                for (int k = 0; k < 9; ++k) acc += m[k] * nb[c][k];
                os[c][i][8 * lane + j] = acc * inv;
            }
        }
    }
    __syncthreads();
    // eight output rows per channel, each 8 * (pixels of this CTA) contiguous floats
    const int npx = min(kUpPix, W - w0);
    if (npx == kUpPix) {
        // full tile: 64 float4 per output row, C * 8 rows; a thread keeps its column q and walks the rows in steps of 4
        // (no runtime division, one 64-bit base + 32-bit row offsets)
        const int q = threadIdx.x & 63, r0 = threadIdx.x >> 6;
        float* ob = out + ((int64_t)b * C * 8 * H + 8 * h + r0) * (8 * (int64_t)W) + 8 * w0 + 4 * q;
        const unsigned row = 8u * (unsigned)W;
#pragma unroll
        for (int it = 0; it < 2 * C; ++it) {  // row r0 + 4 * it = (c = it / 2, ii = 4 * (it & 1) + r0)
            const float4 v = *reinterpret_cast<const float4*>(&os[it >> 1][4 * (it & 1) + r0][4 * q]);
            st_v4_off(ob, ((unsigned)(it >> 1) * 8u * (unsigned)H + 4u * (it & 1)) * row, v);
        }
        return;
    }
    const int row_len = 8 * npx;
    for (int idx = threadIdx.x; idx < C * 8 * (row_len / 4); idx += blockDim.x) {
        const int q = idx % (row_len / 4);
        const int ii = (idx / (row_len / 4)) % 8;
        const int c = idx / ((row_len / 4) * 8);
        const float4 v = *reinterpret_cast<const float4*>(&os[c][ii][4 * q]);
        *reinterpret_cast<float4*>(out + (((int64_t)b * C + c) * 8 * H + 8 * h + ii) * (8 * (int64_t)W) + 8 * w0 + 4 * q) = v;
    }
}

// gout [B, C, 8H, 8W] -> dmask [B, 576, H, W] (overwritten), dflow [B, C, H, W] (ACCUMULATED into: zero it first)
template <int C, int U>
__global__ void __launch_bounds__(256, 2)
upsample_bwd_kernel(const float* __restrict__ flow, const float* __restrict__ mask, const float* __restrict__ gout,
                    float* __restrict__ dmask, float* __restrict__ dflow, int H, int W) {
    __shared__ float fs[C][3][kUpPix + 2];
    __shared__ __align__(16) float gs[C][8][8 * kUpPix + 4];
    __shared__ float dfs[C][3][kUpPix + 2];
    const int b = blockIdx.z, h = blockIdx.y, w0 = blockIdx.x * kUpPix;
    const int lane = threadIdx.x & 31, i = threadIdx.x >> 5;
    const int w = w0 + lane;
    const int64_t HW = (int64_t)H * W;
    load_flow_tile<C>(fs, flow, b, h, w0, H, W);
    for (int idx = threadIdx.x; idx < C * 3 * (kUpPix + 2); idx += blockDim.x) (&dfs[0][0][0])[idx] = 0.f;
    const int npx = min(kUpPix, W - w0);
    const int row_len = 8 * npx;
    if (npx == kUpPix) {  // full tile: same walk as the forward kernel's write-out
        const int q = threadIdx.x & 63, r0 = threadIdx.x >> 6;
        const float* gb = gout + ((int64_t)b * C * 8 * H + 8 * h + r0) * (8 * (int64_t)W) + 8 * w0 + 4 * q;
        const unsigned row = 8u * (unsigned)W;
#pragma unroll
        for (int it = 0; it < 2 * C; ++it)
            *reinterpret_cast<float4*>(&gs[it >> 1][4 * (it & 1) + r0][4 * q]) =
                ld_v4_off(gb, ((unsigned)(it >> 1) * 8u * (unsigned)H + 4u * (it & 1)) * row);
    } else {
        for (int idx = threadIdx.x; idx < C * 8 * (row_len / 4); idx += blockDim.x) {
            const int q = idx % (row_len / 4);
            const int ii = (idx / (row_len / 4)) % 8;
            const int c = idx / ((row_len / 4) * 8);
            *reinterpret_cast<float4*>(&gs[c][ii][4 * q]) =
                *reinterpret_cast<const float4*>(gout + (((int64_t)b * C + c) * 8 * H + 8 * h + ii) * (8 * (int64_t)W) + 8 * w0 + 4 * q);
        }
    }
    __syncthreads();
    if (w < W) {
        const int64_t moff = ((int64_t)b * 576 + i * 8) * HW + (int64_t)h * W + w;
        const float* mp = mask + moff;
        float* dp = dmask + moff;
        const unsigned hw = (unsigned)HW;  // 32-bit channel offsets, see the forward kernel
        float nb[C][9], dn[C][9];
#pragma unroll
        for (int c = 0; c < C; ++c)
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                nb[c][k] = fs[c][k / 3][lane + k % 3];
                dn[c][k] = 0.f;
            }
        // the mask rows of U sub-columns are requested together (9 * U loads in flight per thread); the asm loads and
        // stores are volatile, i.e. kept in program order, so the batching has to be explicit
#pragma unroll 1
        for (int jb = 0; jb < 8; jb += U) {
          float mblk[U][9];
          const unsigned ob = (unsigned)jb * hw;
#pragma unroll
          for (int u = 0; u < U; ++u)
#pragma unroll
              for (int k = 0; k < 9; ++k) mblk[u][k] = ld_stream_off(mp, (unsigned)(k * 64 + u) * hw + ob);
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int j = jb + u;
            float* m = mblk[u];
            float mx = m[0];
#pragma unroll
            for (int k = 1; k < 9; ++k) mx = fmaxf(mx, m[k]);
            float den = 0.f;
            const float mxl = mx * kLog2e;
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                m[k] = ex2_approx(fmaf(m[k], kLog2e, -mxl));
                den += m[k];
            }
            const float inv = rcp_approx(den);
            float g[C];
#pragma unroll
            for (int c = 0; c < C; ++c) g[c] = gs[c][i][8 * lane + j];
            // s_k = d out / d p_k;  d mask_k = p_k (s_k - sum_k' p_k' s_k')
            float s[9], dot = 0.f;
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                m[k] *= inv;
                float a = 0.f;
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    a += g[c] * nb[c][k];
                    dn[c][k] += m[k] * g[c];
                }
                s[k] = a;
                dot += m[k] * a;
            }
#pragma unroll
            for (int k = 0; k < 9; ++k) st_off(dp, (unsigned)(k * 64 + u) * hw + ob, m[k] * (s[k] - dot));
          }
        }
#pragma unroll
        for (int c = 0; c < C; ++c)
#pragma unroll
            for (int k = 0; k < 9; ++k) atomicAdd(&dfs[c][k / 3][lane + k % 3], 8.0f * dn[c][k]);
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < C * 3 * (kUpPix + 2); idx += blockDim.x) {
        const int x = idx % (kUpPix + 2);
        const int r = (idx / (kUpPix + 2)) % 3;
        const int c = idx / (3 * (kUpPix + 2));
        const int yy = h + r - 1, xx = w0 + x - 1;
        if (yy >= 0 && yy < H && xx >= 0 && xx < W && xx <= w0 + npx)
            atomicAdd(dflow + (((int64_t)b * C + c) * H + yy) * W + xx, dfs[c][r][x]);
    }
}

}  // namespace

int launch_upsample_fwd(const float* flow, const float* mask, int B, int C, int H, int W, float* out, cudaStream_t s) {
    RC_REQUIRE(C >= 1 && C <= kUpMaxC, "upsample_flow: 1..%d channels (got %d)", kUpMaxC, C);
    RC_REQUIRE(B <= 65535 && H <= 65535, "upsample_flow: batch / height too large for one launch");
    RC_REQUIRE(576 * (int64_t)H * W < (1ll << 31), "upsample_flow: %d x %d map too large (576 * H * W must stay below 2^31)", H, W);
    dim3 grid(ceil_div(W, kUpPix), H, B);
#define RAFTCORR_UP_CASE(C_) case C_: upsample_fwd_kernel<C_><<<grid, 256, 0, s>>>(flow, mask, out, H, W); break;
    switch (C) {
        RAFTCORR_UP_CASE(1) RAFTCORR_UP_CASE(2) RAFTCORR_UP_CASE(3) RAFTCORR_UP_CASE(4)
    }
#undef RAFTCORR_UP_CASE
    RC_CUDA(cudaGetLastError());
    return 0;
}

int launch_upsample_bwd(const float* flow, const float* mask, const float* gout, int B, int C, int H, int W, float* dmask,
                        float* dflow, cudaStream_t s) {
    RC_REQUIRE(C >= 1 && C <= 4, "upsample_flow backward: 1..4 channels (got %d)", C);
    RC_REQUIRE(B <= 65535 && H <= 65535, "upsample_flow: batch / height too large for one launch");
    RC_REQUIRE(576 * (int64_t)H * W < (1ll << 31), "upsample_flow: %d x %d map too large (576 * H * W must stay below 2^31)", H, W);
    dim3 grid(ceil_div(W, kUpPix), H, B);
    static const int unroll = [] {  // RAFTCORR_UP_BWD_UNROLL=2|4|8: mask rows in flight per thread (A/B switch)
        const char* e = std::getenv("RAFTCORR_UP_BWD_UNROLL");
        return (e && e[0] == '4') ? 4 : (e && e[0] == '8') ? 8 : 2;
    }();
#define RAFTCORR_UP_CASE(C_)                                                                                            \
    case C_:                                                                                                            \
        if (unroll == 4) upsample_bwd_kernel<C_, 4><<<grid, 256, 0, s>>>(flow, mask, gout, dmask, dflow, H, W);          \
        else if (unroll == 8) upsample_bwd_kernel<C_, 8><<<grid, 256, 0, s>>>(flow, mask, gout, dmask, dflow, H, W);     \
        else upsample_bwd_kernel<C_, 2><<<grid, 256, 0, s>>>(flow, mask, gout, dmask, dflow, H, W);                      \
        break;
    switch (C) { RAFTCORR_UP_CASE(1) RAFTCORR_UP_CASE(2) RAFTCORR_UP_CASE(3) RAFTCORR_UP_CASE(4) }
#undef RAFTCORR_UP_CASE
    RC_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace raftcorr

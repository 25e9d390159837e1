// Layout / pooling helpers around the hot kernels (all tiny next to the volume):
//   NCHW <-> NHWC transposes of the feature maps (the tcgen05 build and the on-the-fly path want
//   the feature dim contiguous = K-major; the reference re-does this permute on every call,
//   raft/core/corr.py:158-159), 2x2 average pooling (corr.py:42,138-139) and its adjoint.
#include <cuda_fp16.h>

#include <algorithm>

#include "common.cuh"

namespace raftcorr {

namespace {

// ---- block-scaled fp16 operands for the kind::f16 build (RAFTCORR_BUILD_F16_TC) ----------------------------------
// fp16 carries the same 11 significant bits as TF32 but only 5 exponent bits, so every operand BLOCK of 64 pixels (all
// channels) gets its own power-of-two scale that puts the block's largest magnitude into [2^14, 2^15): anything down to
// 2^-29 of the block maximum keeps full precision (below that it fades out through the subnormals, far below what 1e-3
// of max|V| can see), nothing overflows, and powers of two make the scaling itself exact.  Blocks follow the build's
// tiles -- fmap1: 64 consecutive source pixels (half an m-tile = one half of the TMEM lanes); fmap2: one ROW PAIR of an
// 8 x 32 target tile (2 rows x 32 columns = one epilogue phase) -- so undoing the scales costs the epilogue one scalar
// per thread and one per phase.  Both kinds of block are two segments of 32 contiguous pixels.
struct F16PrepParams {
    const float* f1;
    const float* f2;
    __half* o1;
    __half* o2;
    float* inv_a;   // [B][2 * m_tiles]
    float* inv_b;   // [B][4 * bands][w_tiles]
    int B, D, Dp, H, W, N, a_blocks, bands4, w_tiles;   // a_blocks = 2 * m_tiles per image
    // split = 1 (RAFTCORR_BUILD_F16X3_TC): every value is written as hi = fp16(x s) and lo = fp16(x s - hi); the K axis
    // of a pixel's row is [hi | lo | hi] for fmap1 and [hi | hi | lo] for fmap2 (3 * Dk halves, Dk = the single-term
    // pitch), so that ONE accumulation computes hi*hi + lo*hi + hi*lo: the product to ~2^-21 instead of 2^-11.
    int split, Dk;
};

struct F16Block {
    const float* src;   // image base [D][N]
    __half* dst;        // image base [N][Dp]
    float* inv;
    int p0, p1, len0, len1;   // the two segments: first pixel, valid pixels (0..32)
    int is_b;                 // block of fmap2
};

__device__ __forceinline__ F16Block f16_block_of(const F16PrepParams& P, int id) {
    F16Block k;
    if (id < P.B * P.a_blocks) {
        const int b = id / P.a_blocks, blk = id - b * P.a_blocks;
        k.src = P.f1 + (int64_t)b * P.D * P.N;
        k.dst = P.o1 + (int64_t)b * P.N * P.Dp;
        k.inv = P.inv_a + id;
        k.p0 = blk * 64;
        k.p1 = k.p0 + 32;
        k.len0 = max(0, min(32, P.N - k.p0));
        k.len1 = max(0, min(32, P.N - k.p1));
        k.is_b = 0;
    } else {
        id -= P.B * P.a_blocks;
        const int per_b = P.bands4 * P.w_tiles;
        const int b = id / per_b, r = id - b * per_b, hp = r / P.w_tiles, wt = r - hp * P.w_tiles;
        k.src = P.f2 + (int64_t)b * P.D * P.N;
        k.dst = P.o2 + (int64_t)b * P.N * P.Dp;
        k.inv = P.inv_b + id;
        const int cols = max(0, min(32, P.W - wt * 32));
        k.p0 = (2 * hp) * P.W + wt * 32;
        k.p1 = k.p0 + P.W;
        k.len0 = 2 * hp < P.H ? cols : 0;
        k.len1 = 2 * hp + 1 < P.H ? cols : 0;
        k.is_b = 1;
    }
    return k;
}

// One CTA per block, ONE pass over HBM: the block (64 pixels x up to 256 channels) sits in registers (64 values per
// thread, coalesced reads along the pixels) while its absmax is reduced, then it is scaled, rounded to fp16 and
// transposed through shared memory; every pixel leaves as 128-byte runs of 64 channels.  Channels D..Dp are zeros.
// NCH = ceil(Dp / 64) channel chunks of 64 (compile-time so that the register block stays in registers).
template <int NCH>
__global__ void __launch_bounds__(256, 2) f16_operand_kernel(const F16PrepParams P) {
    __shared__ __half tile[64][NCH * 64 + 2];   // odd word stride: conflict-free both ways
    __shared__ float red[8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const F16Block k = f16_block_of(P, blockIdx.x);
    float v[NCH][8][2];
    float mx = 0.f;
#pragma unroll
    for (int q = 0; q < NCH; ++q)
#pragma unroll
        for (int c = 0; c < 8; ++c) {              // warp: channel rows q*64 + warp*8 + c, lane = pixel of a segment
            const int ch = q * 64 + warp * 8 + c;
            const float* row = k.src + (int64_t)ch * P.N;
            v[q][c][0] = (ch < P.D && lane < k.len0) ? __ldg(row + k.p0 + lane) : 0.f;
            v[q][c][1] = (ch < P.D && lane < k.len1) ? __ldg(row + k.p1 + lane) : 0.f;
        }
#pragma unroll
    for (int q = 0; q < NCH; ++q)
#pragma unroll
        for (int c = 0; c < 8; ++c) mx = fmaxf(mx, fmaxf(fabsf(v[q][c][0]), fabsf(v[q][c][1])));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    mx = red[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) mx = fmaxf(mx, red[i]);
    float scale = 1.f, inv = 1.f;
    if (mx > 0.f && mx < __int_as_float(0x7f800000)) {   // zero blocks and inf/NaN blocks keep scale 1 (inf/NaN propagate)
        int e;
        frexpf(mx, &e);                                  // mx = f * 2^e, f in [0.5, 1)
        const int kk = max(-126, min(126, 15 - e));
        scale = ldexpf(1.f, kk);
        inv = ldexpf(1.f, -kk);
    }
    if (threadIdx.x == 0) *k.inv = inv;
    const int parts = P.split ? 2 : 1;
    for (int part = 0; part < parts; ++part) {     // part 0: hi (or the only term), part 1: lo = x s - hi
#pragma unroll
        for (int q = 0; q < NCH; ++q)
#pragma unroll
            for (int c = 0; c < 8; c += 2) {
                const int col = q * 64 + warp * 8 + c;
                float a0 = v[q][c][0] * scale, a1 = v[q][c + 1][0] * scale, b0 = v[q][c][1] * scale, b1 = v[q][c + 1][1] * scale;
                if (part) {
                    a0 -= __half2float(__float2half_rn(a0)); a1 -= __half2float(__float2half_rn(a1));
                    b0 -= __half2float(__float2half_rn(b0)); b1 -= __half2float(__float2half_rn(b1));
                }
                *reinterpret_cast<__half2*>(&tile[lane][col]) = __floats2half2_rn(a0, a1);
                *reinterpret_cast<__half2*>(&tile[32 + lane][col]) = __floats2half2_rn(b0, b1);
            }
        __syncthreads();
        // K-axis slots this part goes to: single term: 0; split: hi -> 0 and (fmap1: 2, fmap2: 1), lo -> (fmap1: 1, fmap2: 2)
        const int slot_a = !P.split ? 0 : (part == 0 ? 0 : (k.is_b ? 2 : 1));
        const int slot_b = !P.split ? -1 : (part == 0 ? (k.is_b ? 1 : 2) : -1);
#pragma unroll
        for (int i = 0; i < 8; ++i) {              // warp: 8 of the 64 pixels, lane = channel pair of every chunk
            const int px = warp * 8 + i, seg = px >> 5, off = px & 31;
            if (off < (seg ? k.len1 : k.len0)) {
                __half* out = k.dst + (int64_t)((seg ? k.p1 : k.p0) + off) * P.Dp;
#pragma unroll
                for (int q = 0; q < NCH; ++q) {
                    const __half2 h = *reinterpret_cast<const __half2*>(&tile[px][q * 64 + 2 * lane]);
                    *reinterpret_cast<__half2*>(out + slot_a * P.Dk + q * 64 + 2 * lane) = h;
                    if (slot_b >= 0) *reinterpret_cast<__half2*>(out + slot_b * P.Dk + q * 64 + 2 * lane) = h;
                }
            }
        }
        __syncthreads();
    }
}

template <int NCH>
__global__ void __launch_bounds__(256, 2) f16_from_nhwc_kernel(const F16PrepParams P, int Dp32) {
    __shared__ float red[8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    F16Block k = f16_block_of(P, blockIdx.x);
    const float* src = k.src;                   // image base [N][Dp32] (P.D carries Dp32, see the launcher)
    constexpr int Q = NCH * 16;                 // float4 per pixel row of NCH*64 channels
    float4 v[NCH * 4];                          // 64 px * Q float4 / 256 threads
    float mx = 0.f;
#pragma unroll
    for (int i = 0; i < NCH * 4; ++i) {
        const int idx = threadIdx.x + 256 * i, px = idx / Q, c4 = idx - px * Q;
        const int seg = px >> 5, off = px & 31;
        const bool ok = off < (seg ? k.len1 : k.len0) && 4 * c4 < Dp32;
        v[i] = ok ? __ldg(reinterpret_cast<const float4*>(src + (int64_t)((seg ? k.p1 : k.p0) + off) * Dp32) + c4)
                  : make_float4(0.f, 0.f, 0.f, 0.f);
        mx = fmaxf(mx, fmaxf(fmaxf(fabsf(v[i].x), fabsf(v[i].y)), fmaxf(fabsf(v[i].z), fabsf(v[i].w))));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    mx = red[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) mx = fmaxf(mx, red[i]);
    float scale = 1.f, inv = 1.f;
    if (mx > 0.f && mx < __int_as_float(0x7f800000)) {
        int e;
        frexpf(mx, &e);
        const int kk = max(-126, min(126, 15 - e));
        scale = ldexpf(1.f, kk);
        inv = ldexpf(1.f, -kk);
    }
    if (threadIdx.x == 0) *k.inv = inv;
#pragma unroll
    for (int i = 0; i < NCH * 4; ++i) {
        const int idx = threadIdx.x + 256 * i, px = idx / Q, c4 = idx - px * Q;
        const int seg = px >> 5, off = px & 31;
        if (off < (seg ? k.len1 : k.len0)) {
            const __half2 lo = __floats2half2_rn(v[i].x * scale, v[i].y * scale), hi = __floats2half2_rn(v[i].z * scale, v[i].w * scale);
            uint2 o;
            o.x = *reinterpret_cast<const unsigned*>(&lo);
            o.y = *reinterpret_cast<const unsigned*>(&hi);
            *(reinterpret_cast<uint2*>(k.dst + (int64_t)((seg ? k.p1 : k.p0) + off) * P.Dp) + c4) = o;
        }
    }
}

// ---- fp16 operands for the on-the-fly tensor-core lookup: one power-of-two scale per image and tensor --------------
// The lattice tiles of that kernel sit at arbitrary offsets (bounding boxes of the windows), so per-tile scales do not
// exist; one scale per (tensor, image) leaves 2^29 of dynamic range inside an image at full 11-bit precision.
struct AltF16Tensors {
    const float* src[5];
    __half* dst[5];
    int64_t per_img[5];
    int count, B;
    float* inv;          // [count][B]
    unsigned* maxbits;   // [count][B], zeroed before the first kernel
};

__global__ void __launch_bounds__(256) alt_f16_absmax_kernel(const AltF16Tensors T) {
    const int t = blockIdx.z, b = blockIdx.y;
    const int64_t n4 = T.per_img[t] >> 2;   // per-image sizes are multiples of 64 elements
    const float4* src = reinterpret_cast<const float4*>(T.src[t] + (int64_t)b * T.per_img[t]);
    float mx = 0.f;
    for (int64_t i = blockIdx.x * 256 + threadIdx.x; i < n4; i += (int64_t)gridDim.x * 256) {
        const float4 v = __ldg(src + i);
        mx = fmaxf(mx, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    // non-negative floats order like their bit patterns (NaN compares above everything: the image keeps scale 1)
    if ((threadIdx.x & 31) == 0 && mx > 0.f) atomicMax(T.maxbits + t * T.B + b, __float_as_uint(mx));
}

__global__ void __launch_bounds__(256) alt_f16_convert_kernel(const AltF16Tensors T) {
    const int t = blockIdx.z, b = blockIdx.y;
    const float mx = __uint_as_float(T.maxbits[t * T.B + b]);
    float scale = 1.f, inv = 1.f;
    if (mx > 0.f && mx < __int_as_float(0x7f800000)) {
        int e;
        frexpf(mx, &e);
        const int kk = max(-126, min(126, 15 - e));
        scale = ldexpf(1.f, kk);
        inv = ldexpf(1.f, -kk);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) T.inv[t * T.B + b] = inv;
    const int64_t n4 = T.per_img[t] >> 2;
    const float4* src = reinterpret_cast<const float4*>(T.src[t] + (int64_t)b * T.per_img[t]);
    uint2* dst = reinterpret_cast<uint2*>(T.dst[t] + (int64_t)b * T.per_img[t]);
    for (int64_t i = blockIdx.x * 256 + threadIdx.x; i < n4; i += (int64_t)gridDim.x * 256) {
        const float4 v = __ldg(src + i);
        const __half2 lo = __floats2half2_rn(v.x * scale, v.y * scale), hi = __floats2half2_rn(v.z * scale, v.w * scale);
        uint2 o;
        o.x = *reinterpret_cast<const unsigned*>(&lo);
        o.y = *reinterpret_cast<const unsigned*>(&hi);
        dst[i] = o;
    }
}

__device__ __forceinline__ float round_tf32(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}

// in [B][D][HW] -> out [B][HW][D]; 32x32 tiles through shared memory, both sides coalesced.
// Dout >= D is the output row pitch; channels [D, Dout) are zero-filled (K padding for the tensor-core path).
template <bool ROUND>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ in, float* __restrict__ out, int D, int HW, int Dout) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const int p0 = blockIdx.x * 32, d0 = blockIdx.y * 32;
    const float* src = in + (int64_t)b * D * HW;
    float* dst = out + (int64_t)b * Dout * HW;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int d = d0 + i, p = p0 + threadIdx.x;
        tile[i][threadIdx.x] = (d < D && p < HW) ? src[(int64_t)d * HW + p] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int p = p0 + i, d = d0 + threadIdx.x;
        if (p < HW && d < Dout) {
            float v = tile[threadIdx.x][i];
            dst[(int64_t)p * Dout + d] = ROUND ? round_tf32(v) : v;
        }
    }
}

// Same transpose with 4x the bytes in flight per thread (32 channels x 128 pixels per block): the
// tensor-core build calls it on both feature maps before every volume, so it should run near copy speed.
template <bool ROUND>
__global__ void __launch_bounds__(256)
nchw_to_nhwc_wide_kernel(const float* __restrict__ in, float* __restrict__ out, const float* __restrict__ in2,
                         float* __restrict__ out2, int B, int D, int HW, int Dout) {
    __shared__ float tile[32][129];
    // blockIdx.z in [B, 2B) handles the second map of a pair (both feature maps in one launch)
    const bool second = (int)blockIdx.z >= B;
    const int b = second ? blockIdx.z - B : blockIdx.z;
    const int p0 = blockIdx.x * 128, d0 = blockIdx.y * 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float* src = (second ? in2 : in) + (int64_t)b * D * HW;
    float* dst = (second ? out2 : out) + (int64_t)b * Dout * HW;
    float r[16];
#pragma unroll
    for (int i = 0; i < 4; ++i) {      // channel rows warp, warp+8, ...
        const int d = d0 + warp + 8 * i;
#pragma unroll
        for (int j = 0; j < 4; ++j) {  // 4 x 32 consecutive pixels
            const int p = p0 + lane + 32 * j;
            r[i * 4 + j] = (d < D && p < HW) ? __ldg(src + (int64_t)d * HW + p) : 0.f;
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) tile[warp + 8 * i][lane + 32 * j] = r[i * 4 + j];
    __syncthreads();
    const int d = d0 + lane;
    if (d < Dout) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int pl = warp + 8 * i, p = p0 + pl;
            if (p < HW) {
                const float v = tile[lane][pl];
                dst[(int64_t)p * Dout + d] = ROUND ? round_tf32(v) : v;
            }
        }
    }
}

// NHWC 2x2 mean, floor mode: in [B][Hi][Wi][D] -> out [B][Hi/2][Wi/2][D]
template <bool ROUND>
__global__ void pool_nhwc_kernel(const float* __restrict__ in, float* __restrict__ out, int Hi, int Wi, int D,
                                 int64_t total) {
    const int Ho = Hi / 2, Wo = Wi / 2;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int d = idx % D;
        int64_t t = idx / D;
        const int x = t % Wo;
        t /= Wo;
        const int y = t % Ho;
        const int b = t / Ho;
        const float* s = in + (((int64_t)b * Hi + 2 * y) * Wi + 2 * x) * D + d;
        const float v = (s[0] + s[D] + s[(int64_t)Wi * D] + s[(int64_t)Wi * D + D]) * 0.25f;
        out[idx] = ROUND ? round_tf32(v) : v;
    }
}

// adjoint: gin [B][Hi][Wi][D] += 0.25 * gout[b][y/2][x/2][d] for (y,x) inside the pooled footprint
__global__ void pool_adjoint_nhwc_kernel(const float* __restrict__ gout, float* __restrict__ gin, int Hi, int Wi,
                                         int D, int64_t total) {
    const int Ho = Hi / 2, Wo = Wi / 2;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int d = idx % D;
        int64_t t = idx / D;
        const int x = t % Wi;
        t /= Wi;
        const int y = t % Hi;
        const int b = t / Hi;
        if ((y >> 1) < Ho && (x >> 1) < Wo)
            gin[idx] += 0.25f * gout[(((int64_t)b * Ho + (y >> 1)) * Wo + (x >> 1)) * D + d];
    }
}

// Pyramid-level pooling on the pitched volume layout: one thread per output element.
__global__ void pool_level_kernel(const LevelView in, const LevelView out, int64_t total) {
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int x = idx % out.w;
        int64_t t = idx / out.w;
        const int y = t % out.h;
        const int64_t img = t / out.h;
        const float* s = in.base + img * in.img_stride;
        out.base[img * out.img_stride + out.at(y, x)] =
            (s[in.at(2 * y, 2 * x)] + s[in.at(2 * y, 2 * x + 1)] + s[in.at(2 * y + 1, 2 * x)] + s[in.at(2 * y + 1, 2 * x + 1)]) * 0.25f;
        // interleaved layout: the partner of an odd last row must read as zero
        if (out.il && y == out.h - 1)  // the rows that complete the last group must read as zero
            for (int yz = y + 1; yz % out.group_rows(); ++yz) out.base[img * out.img_stride + out.at(yz, x)] = 0.f;
    }
}

// Adjoint of one pooling level on the pitched layout: fine[y][x] += coarse[y/2][x/2] / 4 inside the pooled
// footprint.  One thread per 4 consecutive fine elements (rows are 32-byte aligned: pitch % 8 == 0), so
// the pass streams float4s; values written into the row padding are harmless (nothing reads it).
template <bool ROUND>
__global__ void __launch_bounds__(256)
pool_adjoint_level_kernel(const LevelView coarse, const LevelView fine, int64_t imgs) {
    // One image per CTA step (grid-stride over the B*N images), threads laid out as (row, float4 column) ONCE: the first
    // version split a flat 64-bit element index with three 64-bit divisions per float4 -- 123 instructions per element,
    // 78 M warp instructions and issue slots 58 % busy for a pass that should only stream (ncu, fourth session).
    const int q4 = fine.pitch >> 2;  // float4 chunks per fine row
    const int per = q4 * fine.h;     // float4 chunks per image
    if (per <= 128) {
        // small images (the coarse folds): several images per CTA step, one chunk per thread and step
        const int ipp = 256 / per;
        const int sub = (int)threadIdx.x / per, r = (int)threadIdx.x - sub * per;
        const int y = r / q4, x = (r - y * q4) * 4, yc = y >> 1, xc = x >> 1;
        if (sub >= ipp || x >= fine.w || (!ROUND && yc >= coarse.h)) return;
        const int foff = y * fine.pitch + x, coff = yc * coarse.pitch + xc;
        const bool has_c = yc < coarse.h, in0 = xc < coarse.w, in1 = xc + 1 < coarse.w;
        for (int64_t img = (int64_t)blockIdx.x * ipp + sub; img < imgs; img += (int64_t)gridDim.x * ipp) {
            float4* dst = reinterpret_cast<float4*>(fine.base + img * fine.img_stride + foff);
            float4 v = *dst;
            if (has_c) {
                const float* c = coarse.base + img * coarse.img_stride + coff;
                const float c0 = in0 ? 0.25f * c[0] : 0.f;
                const float c1 = in1 ? 0.25f * c[1] : 0.f;
                v.x += c0; v.y += c0; v.z += c1; v.w += c1;
            }
            if (ROUND) {
                v.x = round_tf32(v.x); v.y = round_tf32(v.y); v.z = round_tf32(v.z); v.w = round_tf32(v.w);
            }
            *dst = v;
        }
        return;
    }
    const int tx = q4 >= 256 ? (int)threadIdx.x : (int)threadIdx.x % q4;
    const int ty = q4 >= 256 ? 0 : (int)threadIdx.x / q4;
    const int xstep = q4 >= 256 ? 256 : q4, ystep = q4 >= 256 ? 1 : 256 / q4;
    if (ty >= ystep) return;
    for (int64_t img = blockIdx.x; img < imgs; img += gridDim.x) {
        float* fb = fine.base + img * fine.img_stride;
        const float* cb = coarse.base + img * coarse.img_stride;
        for (int y = ty; y < fine.h; y += ystep) {
            const int yc = y >> 1;
            if (!ROUND && yc >= coarse.h) continue;
            for (int xq = tx; xq < q4; xq += xstep) {
                const int x = xq * 4;
                if (x >= fine.w) continue;
                float4* dst = reinterpret_cast<float4*>(fb + y * fine.pitch + x);
                float4 v = *dst;
                if (yc < coarse.h) {
                    const int xc = x >> 1;
                    const float* c = cb + yc * coarse.pitch + xc;
                    const float c0 = xc < coarse.w ? 0.25f * c[0] : 0.f;
                    const float c1 = xc + 1 < coarse.w ? 0.25f * c[1] : 0.f;
                    v.x += c0; v.y += c0; v.z += c1; v.w += c1;
                }
                if (ROUND) {  // the folded level 0 feeds the tensor-core backward GEMMs
                    v.x = round_tf32(v.x); v.y = round_tf32(v.y); v.z = round_tf32(v.z); v.w = round_tf32(v.w);
                }
                *dst = v;
            }
        }
    }
}

inline int grid_for(int64_t total, int threads) {
    int64_t g = ceil_div64(total, threads);
    const int64_t cap = 148 * 16;  // persistent-ish grid-stride: 16 CTAs per SM
    return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

}  // namespace

size_t f16_scale_floats(int B, int H, int W) {   // inv_a [B][2 m_tiles] followed by inv_b [B][4 bands][w_tiles]
    return (size_t)B * (2 * ceil_div(H * W, 128) + 4 * ceil_div(H, 8) * ceil_div(W, 32));
}

int launch_f16_operands(const float* f1, const float* f2, int B, int D, int H, int W, void* f1h, void* f2h, float* inv_a,
                        float* inv_b, cudaStream_t s, bool split) {
    F16PrepParams P{};
    P.f1 = f1; P.f2 = f2; P.o1 = (__half*)f1h; P.o2 = (__half*)f2h; P.inv_a = inv_a; P.inv_b = inv_b;
    P.split = split ? 1 : 0; P.Dk = tc_feature_pitch_f16(D);
    P.B = B; P.D = D; P.Dp = (split ? 3 : 1) * P.Dk; P.H = H; P.W = W; P.N = H * W;
    P.a_blocks = 2 * ceil_div(P.N, 128); P.bands4 = 4 * ceil_div(H, 8); P.w_tiles = ceil_div(W, 32);
    const int blocks = B * (P.a_blocks + P.bands4 * P.w_tiles);
    switch (P.Dk / 64) {
        case 1: f16_operand_kernel<1><<<blocks, 256, 0, s>>>(P); break;
        case 2: f16_operand_kernel<2><<<blocks, 256, 0, s>>>(P); break;
        case 3: f16_operand_kernel<3><<<blocks, 256, 0, s>>>(P); break;
        case 4: f16_operand_kernel<4><<<blocks, 256, 0, s>>>(P); break;
        default: return fail("f16 build: feature dim %d > 256 is not supported (use the tf32 build)", D);
    }
    RC_CUDA(cudaGetLastError());
    return 0;
}

// f1n/f2n: [B][N][Dp32] fp32 (fnet_head_kernel's output) -> block-scaled fp16 operands, same blocks as launch_f16_operands
int launch_f16_operands_from_nhwc(const float* f1n, const float* f2n, int B, int D, int Dp32, int H, int W, void* f1h,
                                  void* f2h, float* inv_a, float* inv_b, cudaStream_t s) {
    F16PrepParams P{};
    P.f1 = f1n; P.f2 = f2n; P.o1 = (__half*)f1h; P.o2 = (__half*)f2h; P.inv_a = inv_a; P.inv_b = inv_b;
    // f16_block_of strides images by D * N floats: the pixel-major source has Dp32 * N per image
    P.B = B; P.D = Dp32; P.Dp = tc_feature_pitch_f16(D); P.H = H; P.W = W; P.N = H * W;
    P.a_blocks = 2 * ceil_div(P.N, 128); P.bands4 = 4 * ceil_div(H, 8); P.w_tiles = ceil_div(W, 32);
    const int blocks = B * (P.a_blocks + P.bands4 * P.w_tiles);
    switch (P.Dp / 64) {
        case 1: f16_from_nhwc_kernel<1><<<blocks, 256, 0, s>>>(P, Dp32); break;
        case 2: f16_from_nhwc_kernel<2><<<blocks, 256, 0, s>>>(P, Dp32); break;
        case 3: f16_from_nhwc_kernel<3><<<blocks, 256, 0, s>>>(P, Dp32); break;
        case 4: f16_from_nhwc_kernel<4><<<blocks, 256, 0, s>>>(P, Dp32); break;
        default: return fail("f16 build: feature dim %d > 256 is not supported (use the tf32 build)", D);
    }
    RC_CUDA(cudaGetLastError());
    return 0;
}

int launch_alt_f16_operands(const float* src_f1, const float* src_f2, void* dst_f1, void* dst_f2, const int64_t* off,
                            const int64_t* per_img, int count, int B, float* inv, unsigned* maxbits, cudaStream_t s) {
    RC_REQUIRE(count >= 2 && count <= 5, "alt f16 operands: 1..4 pyramid levels");
    AltF16Tensors T{};
    T.count = count; T.B = B; T.inv = inv; T.maxbits = maxbits;
    int64_t biggest = 0;
    for (int t = 0; t < count; ++t) {   // tensor 0 = fmap1, 1.. = fmap2 levels (off[] = element offset inside the pyramid)
        T.src[t] = t == 0 ? src_f1 : src_f2 + off[t];
        T.dst[t] = t == 0 ? (__half*)dst_f1 : (__half*)dst_f2 + off[t];
        T.per_img[t] = per_img[t];
        RC_REQUIRE(per_img[t] % 4 == 0, "alt f16 operands: per-image size must be a multiple of 4");
        biggest = per_img[t] > biggest ? per_img[t] : biggest;
    }
    RC_CUDA(cudaMemsetAsync(maxbits, 0, sizeof(unsigned) * count * B, s));
    const int chunks = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div64(biggest / 4, 256 * 8), 148 * 4));
    dim3 grid(chunks, B, count);
    alt_f16_absmax_kernel<<<grid, 256, 0, s>>>(T);
    alt_f16_convert_kernel<<<grid, 256, 0, s>>>(T);
    RC_CUDA(cudaGetLastError());
    return 0;
}

int launch_nchw_to_nhwc_pair(const float* in1, float* out1, const float* in2, float* out2, int B, int D, int HW,
                             int Dout, bool round, cudaStream_t s) {
    dim3 grid(ceil_div(HW, 128), ceil_div(Dout, 32), in2 ? 2 * B : B);
    if (round)
        nchw_to_nhwc_wide_kernel<true><<<grid, 256, 0, s>>>(in1, out1, in2, out2, B, D, HW, Dout);
    else
        nchw_to_nhwc_wide_kernel<false><<<grid, 256, 0, s>>>(in1, out1, in2, out2, B, D, HW, Dout);
    RC_CUDA(cudaGetLastError());
    return 0;
}

int launch_nchw_to_nhwc(const float* in, float* out, int B, int D, int HW, int Dout, bool round, cudaStream_t s) {
    return launch_nchw_to_nhwc_pair(in, out, nullptr, nullptr, B, D, HW, Dout, round, s);
}

int launch_nhwc_to_nchw(const float* in, float* out, int B, int D, int HW, cudaStream_t s) {
    // the same tile transpose with the roles of D and HW swapped
    dim3 grid(ceil_div(D, 32), ceil_div(HW, 32), B), block(32, 8);
    nchw_to_nhwc_kernel<false><<<grid, block, 0, s>>>(in, out, HW, D, HW);
    RC_CUDA(cudaGetLastError());
    return 0;
}

int launch_pool_nhwc(const float* in, float* out, int B, int Hi, int Wi, int D, bool round, cudaStream_t s) {
    const int64_t total = (int64_t)B * (Hi / 2) * (Wi / 2) * D;
    if (total == 0) return 0;
    if (round)
        pool_nhwc_kernel<true><<<grid_for(total, 256), 256, 0, s>>>(in, out, Hi, Wi, D, total);
    else
        pool_nhwc_kernel<false><<<grid_for(total, 256), 256, 0, s>>>(in, out, Hi, Wi, D, total);
    RC_CUDA(cudaGetLastError());
    return 0;
}

int launch_pool_adjoint_nhwc(const float* gout, float* gin, int B, int Hi, int Wi, int D, cudaStream_t s) {
    const int64_t total = (int64_t)B * Hi * Wi * D;
    if (total == 0) return 0;
    pool_adjoint_nhwc_kernel<<<grid_for(total, 256), 256, 0, s>>>(gout, gin, Hi, Wi, D, total);
    RC_CUDA(cudaGetLastError());
    return 0;
}

int launch_pool_level(const LevelView& in, const LevelView& out, int64_t imgs, cudaStream_t s) {
    const int64_t total = imgs * out.h * out.w;
    if (total == 0) return 0;
    pool_level_kernel<<<grid_for(total, 256), 256, 0, s>>>(in, out, total);
    RC_CUDA(cudaGetLastError());
    return 0;
}

int launch_pool_adjoint_level(const LevelView& coarse, const LevelView& fine, int64_t imgs, bool round_tf32_out,
                              cudaStream_t s) {
    const int64_t total = imgs * fine.h * (fine.pitch >> 2);
    if (total == 0) return 0;
    const int per = (fine.pitch >> 2) * fine.h;
    const int64_t steps = per <= 128 ? ceil_div64(imgs, 256 / per) : imgs;  // CTA steps (several small images per step)
    const int grid = (int)(steps < 148 * 32 ? steps : 148 * 32);
    if (round_tf32_out)
        pool_adjoint_level_kernel<true><<<grid, 256, 0, s>>>(coarse, fine, imgs);
    else
        pool_adjoint_level_kernel<false><<<grid, 256, 0, s>>>(coarse, fine, imgs);
    RC_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace raftcorr

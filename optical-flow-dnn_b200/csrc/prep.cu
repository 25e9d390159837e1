// Layout / pooling helpers around the hot kernels (all tiny next to the volume):
//   NCHW <-> NHWC transposes of the feature maps (the tcgen05 build and the on-the-fly path want
//   the feature dim contiguous = K-major; the reference re-does this permute on every call,
//   raft/core/corr.py:158-159), 2x2 average pooling (corr.py:42,138-139) and its adjoint.
#include "common.cuh"

namespace raftcorr {

namespace {

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
        if (out.il && y == out.h - 1 && !(y & 1)) out.base[img * out.img_stride + out.at(y + 1, x)] = 0.f;
    }
}

// Adjoint of one pooling level on the pitched layout: fine[y][x] += coarse[y/2][x/2] / 4 inside the pooled
// footprint.  One thread per 4 consecutive fine elements (rows are 32-byte aligned: pitch % 8 == 0), so
// the pass streams float4s; values written into the row padding are harmless (nothing reads it).
template <bool ROUND>
__global__ void pool_adjoint_level_kernel(const LevelView coarse, const LevelView fine, int64_t total4) {
    const int q4 = fine.pitch >> 2;  // float4 chunks per fine row
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total4;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int xq = (int)(idx % q4);
        int64_t t = idx / q4;
        const int y = (int)(t % fine.h);
        const int64_t img = t / fine.h;
        const int x = xq * 4;
        if (x >= fine.w) continue;
        float4* dst = reinterpret_cast<float4*>(fine.base + img * fine.img_stride + (int64_t)y * fine.pitch + x);
        float4 v = *dst;
        const int yc = y >> 1, xc = x >> 1;
        if (yc < coarse.h) {
            const float* c = coarse.base + img * coarse.img_stride + (int64_t)yc * coarse.pitch + xc;
            const float c0 = xc < coarse.w ? 0.25f * c[0] : 0.f;
            const float c1 = xc + 1 < coarse.w ? 0.25f * c[1] : 0.f;
            v.x += c0; v.y += c0; v.z += c1; v.w += c1;
        } else if (!ROUND) {
            continue;
        }
        if (ROUND) {  // the folded level 0 feeds the tensor-core backward GEMMs
            v.x = round_tf32(v.x); v.y = round_tf32(v.y); v.z = round_tf32(v.z); v.w = round_tf32(v.w);
        }
        *dst = v;
    }
}

inline int grid_for(int64_t total, int threads) {
    int64_t g = ceil_div64(total, threads);
    const int64_t cap = 148 * 16;  // persistent-ish grid-stride: 16 CTAs per SM
    return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

}  // namespace

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
    if (round_tf32_out)
        pool_adjoint_level_kernel<true><<<grid_for(total, 256), 256, 0, s>>>(coarse, fine, total);
    else
        pool_adjoint_level_kernel<false><<<grid_for(total, 256), 256, 0, s>>>(coarse, fine, total);
    RC_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace raftcorr

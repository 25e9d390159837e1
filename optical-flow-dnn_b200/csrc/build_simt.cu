// Exact-fp32 correlation volume build on CUDA cores (RAFTCORR_BUILD_FP32_SIMT) and the SIMT
// GEMMs of the build backward.  CorrBlock.corr, raft/core/corr.py:93-116:
//     V0[b][m][n] = (sum_k f1[b][k][m] * f2[b][k][n]) / sqrt(D)
// This is the bit-for-bit-fp32 mode (FFMA accumulation, like the reference's cuBLAS SGEMM with
// TF32 off) and the on-device cross-check for the tcgen05 kernel in build_tc.cu, which is the
// fast path.  Both feature maps are NCHW, i.e. [k][m] / [k][n] with the pixel index contiguous --
// exactly the [BK][BM] shared-memory arrangement a register-tiled SGEMM wants.
#include "common.cuh"

namespace raftcorr {

namespace {

constexpr int BM = 128, BN = 128, BK = 8, TM = 8, TN = 8;

// C[i][j] = scale * sum_k A(k,i) * Bm(k,j), with generic element strides so the same kernel
// serves the forward (A=f1, B=f2) and both backward products.
struct GemmOperand {
    const float* ptr;
    int64_t stride_k, stride_mn, batch_stride;
    // optional 2-D pitched addressing of the mn index: off = (mn / w) * pitch + mn % w   (w == 0: linear)
    int w, pitch;
    __device__ __forceinline__ int64_t mn_off(int mn) const {
        return w ? (int64_t)(mn / w) * pitch + (mn % w) : (int64_t)mn;
    }
    // optional 2-D pitched addressing of the k index (used when K runs over pitched target pixels)
    int kw, kpitch;
    __device__ __forceinline__ int64_t k_off(int k) const {
        return kw ? (int64_t)(k / kw) * kpitch + (k % kw) : (int64_t)k;
    }
};

struct GemmOut {
    float* ptr;
    int64_t stride_i, batch_stride;  // row stride (floats); column index may be pitched:
    int w, pitch;                    // off_j = (j / w) * pitch + j % w  (w == 0: linear)
};

__global__ void __launch_bounds__(256, 2)
sgemm_kernel(GemmOperand A, GemmOperand Bo, GemmOut C, int M, int N, int K, float scale) {
    __shared__ float As[2][BK][BM];
    __shared__ float Bs[2][BK][BN];
    const int t = threadIdx.x;
    const int b = blockIdx.z;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const float* Ab = A.ptr + (int64_t)b * A.batch_stride;
    const float* Bb = Bo.ptr + (int64_t)b * Bo.batch_stride;

    const int lk = t >> 5, lm = t & 31;  // loader: k row, base column
    const int ty = t >> 4, tx = t & 15;

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    float ra[4], rb[4];
    auto gload = [&](int k0) {
        const int k = k0 + lk;
        const int64_t ka = A.k_off(k) * A.stride_k, kb = Bo.k_off(k) * Bo.stride_k;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int m = m0 + lm + 32 * j, n = n0 + lm + 32 * j;
            ra[j] = (k < K && m < M) ? __ldg(Ab + ka + A.mn_off(m) * A.stride_mn) : 0.f;
            rb[j] = (k < K && n < N) ? __ldg(Bb + kb + Bo.mn_off(n) * Bo.stride_mn) : 0.f;
        }
    };
    auto sstore = [&](int buf) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            As[buf][lk][lm + 32 * j] = ra[j];
            Bs[buf][lk][lm + 32 * j] = rb[j];
        }
    };

    gload(0);
    sstore(0);
    __syncthreads();
    int buf = 0;
    for (int k0 = 0; k0 < K; k0 += BK) {
        if (k0 + BK < K) gload(k0 + BK);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[TM], bb[TN];
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
            a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
            bb[0] = b0.x; bb[1] = b0.y; bb[2] = b0.z; bb[3] = b0.w; bb[4] = b1.x; bb[5] = b1.y; bb[6] = b1.z; bb[7] = b1.w;
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
        }
        if (k0 + BK < K) {
            sstore(buf ^ 1);
            __syncthreads();
            buf ^= 1;
        }
    }

    float* Cb = C.ptr + (int64_t)b * C.batch_stride;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + i - 4);
        if (m >= M) continue;
        float* row = Cb + (int64_t)m * C.stride_i;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + j - 4);
            if (n < N) {
                const int64_t off = C.w ? (int64_t)(n / C.w) * C.pitch + (n % C.w) : (int64_t)n;
                row[off] = acc[i][j] * scale;
            }
        }
    }
}

int run_sgemm(const GemmOperand& A, const GemmOperand& B, const GemmOut& C, int M, int N, int K, int batch,
              float scale, cudaStream_t s) {
    dim3 grid(ceil_div(N, BN), ceil_div(M, BM), batch);
    sgemm_kernel<<<grid, 256, 0, s>>>(A, B, C, M, N, K, scale);
    RC_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace

int launch_build_simt(const float* f1, const float* f2, int D, const PyramidView& pv, cudaStream_t s) {
    const int N = pv.N;
    const LevelView& l0 = pv.lvl[0];
    GemmOperand A{f1, (int64_t)N, 1, (int64_t)D * N, 0, 0, 0, 0};
    GemmOperand B{f2, (int64_t)N, 1, (int64_t)D * N, 0, 0, 0, 0};
    GemmOut C{l0.base, l0.img_stride, (int64_t)N * l0.img_stride, l0.pitch == l0.w ? 0 : l0.w, l0.pitch};
    const float scale = 1.0f / sqrtf((float)D);  // corr.py:116
    int rc = run_sgemm(A, B, C, N, N, D, pv.B, scale, s);
    if (rc) return rc;
    for (int l = 1; l < pv.L; ++l) {  // corr.py:41-43, cascaded like the reference
        rc = launch_pool_level(pv.lvl[l - 1], pv.lvl[l], (int64_t)pv.B * N, s);
        if (rc) return rc;
    }
    return 0;
}

// dF1[b][d][m] = scale * sum_n dV0[b][m][n] * f2[b][d][n]      (corr.py:112 backward, first operand)
// dF2[b][d][n] = scale * sum_m dV0[b][m][n] * f1[b][d][m]      (second operand)
int launch_build_bwd_gemms(const PyramidView& dpv, const float* f1, const float* f2, int D, float* df1,
                           float* df2, cudaStream_t s) {
    const int N = dpv.N;
    const LevelView& g0 = dpv.lvl[0];
    const float scale = 1.0f / sqrtf((float)D);
    const int pw = g0.pitch == g0.w ? 0 : g0.w;
    {   // out[i=d][j=m] = sum_{k=n} f2[d][n] * dV0[m][n]
        GemmOperand A{f2, 1, (int64_t)N, (int64_t)D * N, 0, 0, 0, 0};                       // A(k=n, i=d)
        GemmOperand B{g0.base, 1, g0.img_stride, (int64_t)N * g0.img_stride, 0, 0, pw, g0.pitch};  // B(k=n, j=m)
        GemmOut C{df1, (int64_t)N, (int64_t)D * N, 0, 0};
        int rc = run_sgemm(A, B, C, D, N, N, dpv.B, scale, s);
        if (rc) return rc;
    }
    {   // out[i=d][j=n] = sum_{k=m} f1[d][m] * dV0[m][n]
        GemmOperand A{f1, 1, (int64_t)N, (int64_t)D * N, 0, 0, 0, 0};                       // A(k=m, i=d)
        GemmOperand B{g0.base, g0.img_stride, 1, (int64_t)N * g0.img_stride, pw, g0.pitch, 0, 0};  // B(k=m, j=n)
        GemmOut C{df2, (int64_t)N, (int64_t)D * N, 0, 0};
        int rc = run_sgemm(A, B, C, D, N, N, dpv.B, scale, s);
        if (rc) return rc;
    }
    return 0;
}

}  // namespace raftcorr

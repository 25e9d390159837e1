// Build backward on the tensor cores: the two products of the volume's adjoint (autograd of
// raft/core/corr.py:112-116) as TMA-fed tcgen05 TF32 GEMMs reading the (folded) level-0 gradient dV0.
//
//   GEMM 1:  dF1[b][d][m] = s * sum_n dV0[b][m][n] * F2[b][d][n]      M = m (source px), N = d, K = n
//            A = dV0 rows: K-major straight from the gradient pyramid  ([128 m][32 w] boxes, one target row)
//            B = F2 [d][n]: K-major from a TF32-rounded, row-padded NCHW copy
//   GEMM 2:  dF2[b][d][n] = s * sum_m dV0[b][m][n] * F1[b][d][m]      M = n (target px), N = d, K = m
//            A = dV0^T: the reduction index m is the STRIDED one -> MN-major UMMA operand (TF32 needs the
//                SWIZZLE_128B_BASE32B layout, TMA mode 128B_ATOM_32B, see tc_util.cuh); a 128-pixel
//                target tile is 4 rows x 32 cols, fetched as four [32 m][32 w] boxes (one per row)
//            B = F1 [d][m]: K-major from a TF32-rounded, padded copy
//
// dV0 is rounded to TF32 (nearest) by the fold pass that precedes these kernels, the feature maps by
// the small copy kernels below, so the MMA's operand truncation adds no bias; accumulation is fp32
// in TMEM.  Same warp roles as build_tc.cu (TMA producer / MMA issuer / TMEM allocator / 4 epilogue
// warps), 4-stage smem ring, two TMEM accumulator stages; the epilogue scales by 1/sqrt(D) and writes
// the NCHW gradient with 128-byte coalesced stores.  Both kernels are bound by streaming dV0 once each.
#include <algorithm>

#include "tc_util.cuh"

namespace raftcorr {

namespace {

constexpr int kTM = 128;        // MMA M: source pixels (GEMM 1) / target pixels 4x32 (GEMM 2)
constexpr int kBK = 32;         // reduction elements per stage (one 128-byte swizzle row)
constexpr int kStagesB = 4;
constexpr int kABytesB = kTM * kBK * 4;       // 16 KB
constexpr int kBMaxBytes = 256 * kBK * 4;     // 32 KB (D <= 256)
constexpr int kStageBytesB = kABytesB + kBMaxBytes;
constexpr int kSmemBytesB = kStagesB * kStageBytesB + 1024 + 256;
constexpr int kThreadsB = 256;

struct BwdParams {
    float* out;          // dF [B][D][...]
    int B, D, Nmma;      // Nmma = D rounded up to 16 (UMMA N)
    int N1, H2, W2;      // source pixels; target map size
    int m_tiles;         // GEMM 1: ceil(N1/128)
    int bands4, w_tiles; // GEMM 2: target tiles of 4 rows x 32 cols
    int wch;             // GEMM 1: 32-wide column chunks per target row
    int k_blocks;
    int total_tiles;
    float scale;
};

template <int MODE>  // 1: dF1 (A K-major), 2: dF2 (A MN-major)
__global__ void __launch_bounds__(kThreadsB, 1)
build_bwd_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                    const BwdParams P) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + kStagesB * kStageBytesB;
    const uint32_t full_bar = bar_base, empty_bar = bar_base + 8 * kStagesB;
    const uint32_t tfull_bar = bar_base + 16 * kStagesB, tempty_bar = tfull_bar + 16, tmem_slot = tempty_bar + 16;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t stage_tx = kABytesB + (uint32_t)P.D * kBK * 4;  // bytes TMA delivers per stage

    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&map_a);
        prefetch_tensormap(&map_b);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStagesB; ++s) {
            mbar_init(full_bar + 8 * s, 1);
            mbar_init(empty_bar + 8 * s, 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull_bar + 8 * a, 1);
            mbar_init(tempty_bar + 8 * a, 128);
        }
        fence_mbar_init();
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));
    const int tiles_per_b = MODE == 1 ? P.m_tiles : P.bands4 * P.w_tiles;

    if (warp == 0) {
        if (lane == 0) {  // ===================== TMA producer =====================
            int stage = 0;
            uint32_t phase = 0;
            for (int t = blockIdx.x; t < P.total_tiles; t += gridDim.x) {
                const int b = t / tiles_per_b, tt = t - b * tiles_per_b;
                for (int kb = 0; kb < P.k_blocks; ++kb) {
                    mbar_wait(empty_bar + 8 * stage, phase ^ 1);
                    const uint32_t sa = smem_base + stage * kStageBytesB, sb = sa + kABytesB;
                    mbar_expect_tx(full_bar + 8 * stage, stage_tx);
                    if (MODE == 1) {
                        const int h = kb / P.wch, w0 = (kb - h * P.wch) * kBK;
                        tma_load_4d(sa, &map_a, full_bar + 8 * stage, w0, h, tt * kTM, b);   // dV0[m0..+128][h][w0..+32]
                        tma_load_4d(sb, &map_b, full_bar + 8 * stage, w0, h, 0, b);          // F2[0..D][h][w0..+32]
                    } else {
                        const int band = tt / P.w_tiles, w0 = (tt - band * P.w_tiles) * 32, m0 = kb * kBK;
#pragma unroll
                        for (int j = 0; j < 4; ++j)  // one [32 m][32 w] box per target row of the 4x32 tile
                            tma_load_4d(sa + j * 4096, &map_a, full_bar + 8 * stage, w0, band * 4 + j, m0, b);
                        tma_load_3d(sb, &map_b, full_bar + 8 * stage, m0, 0, b);              // F1[0..D][m0..+32]
                    }
                    if (++stage == kStagesB) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {  // ===================== MMA issuer =====================
            const uint32_t idesc = make_idesc_tf32(kTM, P.Nmma, MODE == 2, false);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int t = blockIdx.x; t < P.total_tiles; t += gridDim.x, ++it) {
                const int acc = it & 1;
                mbar_wait(tempty_bar + 8 * acc, ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * 256;
                for (int kb = 0; kb < P.k_blocks; ++kb) {
                    mbar_wait(full_bar + 8 * stage, phase);
                    tc_fence_after();
                    const uint32_t sa = smem_base + stage * kStageBytesB;
                    const uint64_t bdesc = make_smem_desc(sa + kABytesB);
#pragma unroll
                    for (int k = 0; k < kBK / 8; ++k) {
                        // K-major: +32 bytes inside the swizzle row; MN-major: next 8-row atom (+1024 bytes)
                        const uint64_t adesc = MODE == 1 ? make_smem_desc(sa) + (uint64_t)(2 * k)
                                                         : make_smem_desc_mn_tf32(sa + k * 1024, 4096);
                        tc_mma_tf32(d_tmem, adesc, bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    tc_commit(empty_bar + 8 * stage);
                    if (++stage == kStagesB) { stage = 0; phase ^= 1; }
                }
                tc_commit(tfull_bar + 8 * acc);
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue: TMEM -> scale -> NCHW gradient =====================
        const int ew = warp - 4, ml = ew * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(ew * 32) << 16;
        int it = 0;
        for (int t = blockIdx.x; t < P.total_tiles; t += gridDim.x, ++it) {
            const int b = t / tiles_per_b, tt = t - b * tiles_per_b;
            const int acc = it & 1;
            int64_t row_off;  // offset of this thread's pixel inside one [d] plane of the output
            bool ok;
            if (MODE == 1) {
                const int m = tt * kTM + ml;
                ok = m < P.N1;
                row_off = m;
            } else {
                const int band = tt / P.w_tiles, w = (tt - band * P.w_tiles) * 32 + (ml & 31), h = band * 4 + (ml >> 5);
                ok = h < P.H2 && w < P.W2;
                row_off = (int64_t)h * P.W2 + w;
            }
            const int64_t plane = MODE == 1 ? (int64_t)P.N1 : (int64_t)P.H2 * P.W2;
            float* dst = P.out + (int64_t)b * P.D * plane + row_off;
            mbar_wait(tfull_bar + 8 * acc, (it >> 1) & 1);
            tc_fence_after();
            const uint32_t t_acc = tmem_base + lane_addr + acc * 256;
            for (int c = 0; c < P.Nmma; c += 32) {
                float v[32];
                tmem_ld32(t_acc + c, v);
                tmem_ld_wait();
                if (ok) {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (c + j < P.D) dst[(int64_t)(c + j) * plane] = v[j] * P.scale;
                }
            }
            tc_fence_before();
            mbar_arrive(tempty_bar + 8 * acc);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

__device__ __forceinline__ float round_tf32_dev(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}

// dst[r][0..wp) = tf32(src[r][0..w)), zero padded
__global__ void round_pad_rows_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t rows, int w, int wp) {
    // (row, column) thread layout fixed once per thread -- the flat version paid a 64-bit division per ELEMENT (2.5 TB/s)
    if (wp <= 256) {  // short rows: several rows per CTA step
        const int rpp = 256 / wp, ty = (int)threadIdx.x / wp, tx = (int)threadIdx.x - ty * wp;
        if (ty >= rpp) return;
        for (int64_t r = (int64_t)blockIdx.x * rpp + ty; r < rows; r += (int64_t)gridDim.x * rpp)
            dst[r * wp + tx] = tx < w ? round_tf32_dev(__ldg(src + r * w + tx)) : 0.f;
    } else {
        for (int64_t r = blockIdx.x; r < rows; r += gridDim.x)
            for (int c = threadIdx.x; c < wp; c += 256) dst[r * wp + c] = c < w ? round_tf32_dev(__ldg(src + r * w + c)) : 0.f;
    }
}

__global__ void round_level_kernel(const LevelView lv, int64_t total) {
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int x = idx % lv.w;
        int64_t t = idx / lv.w;
        const int y = t % lv.h;
        const int64_t img = t / lv.h;
        float* p = lv.base + img * lv.img_stride + (int64_t)y * lv.pitch + x;
        *p = round_tf32_dev(*p);
    }
}

inline size_t align_up_sz(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace

bool build_bwd_tc_supported(int D) { return D >= 8 && D <= 256; }

size_t build_bwd_tc_workspace_bytes(int B, int D, int H, int W) {
    const size_t n1p = align_up_sz((size_t)H * W, 4), wp = align_up_sz((size_t)W, 8);
    return align_up_sz((size_t)B * D * n1p * 4, 1024) + align_up_sz((size_t)B * D * H * wp * 4, 1024);
}

int launch_round_level(const LevelView& lv, int64_t imgs, cudaStream_t s) {
    const int64_t total = imgs * lv.h * lv.w;
    if (total == 0) return 0;
    const int64_t g = ceil_div64(total, 256);
    round_level_kernel<<<(int)(g < 148 * 16 ? g : 148 * 16), 256, 0, s>>>(lv, total);
    RC_CUDA(cudaGetLastError());
    return 0;
}

// dpv level 0 must already be folded and TF32-rounded.
int launch_build_bwd_tc(const PyramidView& dpv, const float* f1, const float* f2, int D, float* df1, float* df2,
                        void* workspace, int device, cudaStream_t s) {
    EncodeTiledFn fn;
    if (int rc = get_encode_fn(&fn)) return rc;
    const LevelView& g0 = dpv.lvl[0];
    const int B = dpv.B, N1 = dpv.N, H2 = g0.h, W2 = g0.w;
    const int N1p = (int)align_up_sz((size_t)N1, 4), Wp = (int)align_up_sz((size_t)W2, 8);
    float* f1r = (float*)workspace;
    float* f2r = (float*)((char*)workspace + align_up_sz((size_t)B * D * N1p * 4, 1024));
    {
        const int64_t r1 = (int64_t)B * D, r2 = (int64_t)B * D * H2;
        round_pad_rows_kernel<<<(int)std::min<int64_t>(ceil_div64(r1 * N1p, 256), 148 * 16), 256, 0, s>>>(f1, f1r, r1, N1, N1p);
        round_pad_rows_kernel<<<(int)std::min<int64_t>(ceil_div64(r2 * Wp, 256), 148 * 16), 256, 0, s>>>(f2, f2r, r2, W2, Wp);
        RC_CUDA(cudaGetLastError());
    }
    int sms = 0;
    RC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));

    BwdParams P{};
    P.B = B; P.D = D; P.Nmma = (D + 15) / 16 * 16; P.N1 = N1; P.H2 = H2; P.W2 = W2;
    P.scale = 1.0f / sqrtf((float)D);
    const cuuint64_t img = (cuuint64_t)g0.img_stride;
    cuuint64_t gdims[4] = {(cuuint64_t)W2, (cuuint64_t)H2, (cuuint64_t)N1, (cuuint64_t)B};
    cuuint64_t gstr[3] = {(cuuint64_t)g0.pitch * 4, img * 4, (cuuint64_t)N1 * img * 4};
    {   // ---- GEMM 1: dF1
        CUtensorMap map_a, map_b;
        cuuint32_t boxa[4] = {kBK, 1, kTM, 1};
        if (int rc = encode_f32_map(fn, &map_a, (void*)g0.base, 4, gdims, gstr, boxa, CU_TENSOR_MAP_SWIZZLE_128B,
                                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, "dV0 (rows)"))
            return rc;
        cuuint64_t bd[4] = {(cuuint64_t)Wp, (cuuint64_t)H2, (cuuint64_t)D, (cuuint64_t)B};
        cuuint64_t bs[3] = {(cuuint64_t)Wp * 4, (cuuint64_t)H2 * Wp * 4, (cuuint64_t)D * H2 * Wp * 4};
        cuuint32_t boxb[4] = {kBK, 1, (cuuint32_t)D, 1};
        if (int rc = encode_f32_map(fn, &map_b, (void*)f2r, 4, bd, bs, boxb, CU_TENSOR_MAP_SWIZZLE_128B,
                                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, "fmap2 (padded)"))
            return rc;
        P.out = df1;
        P.m_tiles = ceil_div(N1, kTM);
        P.wch = ceil_div(W2, kBK);
        P.k_blocks = H2 * P.wch;
        P.total_tiles = B * P.m_tiles;
        RC_CUDA(cudaFuncSetAttribute(build_bwd_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytesB));
        build_bwd_tc_kernel<1><<<P.total_tiles < sms ? P.total_tiles : sms, kThreadsB, kSmemBytesB, s>>>(map_a, map_b, P);
        RC_CUDA(cudaGetLastError());
    }
    {   // ---- GEMM 2: dF2
        CUtensorMap map_a, map_b;
        cuuint32_t boxa[4] = {32, 1, kBK, 1};
        if (int rc = encode_f32_map(fn, &map_a, (void*)g0.base, 4, gdims, gstr, boxa, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, "dV0 (columns)"))
            return rc;
        cuuint64_t bd[3] = {(cuuint64_t)N1p, (cuuint64_t)D, (cuuint64_t)B};
        cuuint64_t bs[2] = {(cuuint64_t)N1p * 4, (cuuint64_t)D * N1p * 4};
        cuuint32_t boxb[3] = {kBK, (cuuint32_t)D, 1};
        if (int rc = encode_f32_map(fn, &map_b, (void*)f1r, 3, bd, bs, boxb, CU_TENSOR_MAP_SWIZZLE_128B,
                                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, "fmap1 (padded)"))
            return rc;
        P.out = df2;
        P.bands4 = ceil_div(H2, 4);
        P.w_tiles = ceil_div(W2, 32);
        P.k_blocks = ceil_div(N1, kBK);
        P.total_tiles = B * P.bands4 * P.w_tiles;
        RC_CUDA(cudaFuncSetAttribute(build_bwd_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytesB));
        build_bwd_tc_kernel<2><<<P.total_tiles < sms ? P.total_tiles : sms, kThreadsB, kSmemBytesB, s>>>(map_a, map_b, P);
        RC_CUDA(cudaGetLastError());
    }
    return 0;
}

}  // namespace raftcorr

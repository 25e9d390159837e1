// the feature encoder's output head in front of the volume build.
//
// The reference finishes both encoders with a 1x1 convolution (raft/core/extractor.py:172,231 BasicEncoder:
// conv2 = Conv2d(128, 256, 1); :287,341 SmallEncoder: Conv2d(96, 128, 1)), splits the batch into the two feature
// maps (:239), casts them to fp32 (raft/core/raft.py:114-115) and hands them, NCHW, to CorrBlock.  The tensor-core
// build (build_tc.cu) wants them pixel-major (feature dim contiguous = K-major), TF32-rounded and K-padded, so the
// stand-alone build first transposes both maps.  This kernel produces the build's operand buffers DIRECTLY from the
// head's input x [2B][Cin][H*W] (the output of layer3): one read of x, one write of the operands -- the NCHW
// feature maps (write + read + transposed write again) never exist.
//
//   out[img][p][d] = tf32( bias[d] + sum_c x[img][c][p] * W[d][c] )        img < B -> fmap1 side, else fmap2 side
//
// A TMA-fed tcgen05 TF32 GEMM per 128-pixel tile: M = pixels, N = D, K = Cin.
//   A = x tile: the reduction index c is the STRIDED one (NCHW) -> MN-major UMMA operand, four [32 c][32 px] boxes per
//       stage in the SWIZZLE_128B_BASE32B layout (tc_util.cuh).  x arrives unrounded; kind::tf32 would truncate it,
//       which biases every product the same way, so four warps round each stage to nearest IN PLACE in shared memory
//       (generic proxy -> fence.proxy.async -> mbarrier) before the MMA warp may read it.
//   B = W [D][Cin], K-major, TF32-rounded (nearest) by a tiny copy kernel, loaded ONCE per CTA and kept resident
//       in shared memory (128 KB at 256 x 128), so the steady state ingests only the 64 KB x tile per 128 KB of output.
//   Epilogue: TMEM lane = pixel; + bias, round to TF32, 32 columns at a time into a 128B-swizzled shared-memory tile
//       that leaves as ONE [128 px][128 B] TMA store (double-buffered).  The first version stored straight from
//       registers -- 32 lanes x 32 bytes at a 1 KB stride per instruction -- and spent 17 of its 45 us there
//       (timing ablation without the stores: 28 us).
// Bound: HBM (57.7 MB read + 115.3 MB written at config 2 against 7.4 GFLOP).
#include <algorithm>

#include "tc_util.cuh"

namespace raftcorr {

namespace {

constexpr int kHM = 128;                        // pixels per tile (MMA M)
constexpr int kHK = 32;                         // input channels per stage (one 128-byte swizzle row of W)
constexpr int kHStageBytes = kHM * kHK * 4;     // 16 KB: four [32 c][32 px] boxes
constexpr int kHStages = 4;      // one whole x tile in flight (3 stages + 3 store buffers measured the same: 40.5 vs 39.5 us)
constexpr int kHStoreBufs = 2;
constexpr int kHStoreBytes = kHM * 128;          // 16 KB: 128 pixels x 32 columns
constexpr int kHThreads = 384;
constexpr int kHMaxWBytes = 128 * 1024;         // resident weight: ceil(Cin/32) * Dp * 128 bytes

struct HeadParams {
    float* out1;        // [B][HW][Dp]  operand buffer of fmap1
    float* out2;        // [B][HW][Dp]  operand buffer of fmap2
    const float* bias;  // [D] or nullptr
    int B, D, Dp, HW;
    int k_blocks, tiles_per_img, total_tiles;
};

__device__ __forceinline__ float round_tf32_h(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}

__global__ void __launch_bounds__(kHThreads, 1)
fnet_head_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                 const __grid_constant__ CUtensorMap map_o1, const __grid_constant__ CUtensorMap map_o2, const HeadParams P) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ float sbias[256];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t w_base = smem_base + kHStages * kHStageBytes;
    const uint32_t w_kb_bytes = (uint32_t)P.Dp * 128u;
    const uint32_t store_base = w_base + (uint32_t)P.k_blocks * w_kb_bytes;   // kHStoreBufs x [128 px][128 B], 128B-swizzled
    const uint32_t bar_base = store_base + kHStoreBufs * kHStoreBytes;
    const uint32_t full_bar = bar_base, rounded_bar = bar_base + 8 * kHStages, empty_bar = bar_base + 16 * kHStages;
    const uint32_t wfull_bar = bar_base + 24 * kHStages;
    const uint32_t tfull_bar = wfull_bar + 8, tempty_bar = tfull_bar + 16, tmem_slot = tempty_bar + 16;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&map_x);
        prefetch_tensormap(&map_w);
        prefetch_tensormap(&map_o1);
        prefetch_tensormap(&map_o2);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kHStages; ++s) {
            mbar_init(full_bar + 8 * s, 1);
            mbar_init(rounded_bar + 8 * s, 128);
            mbar_init(empty_bar + 8 * s, 1);
        }
        mbar_init(wfull_bar, 1);
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
    for (int i = threadIdx.x; i < 256; i += kHThreads) sbias[i] = (P.bias && i < P.D) ? P.bias[i] : 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

    if (warp == 0) {
        if (lane == 0) {  // ===================== TMA producer =====================
            mbar_expect_tx(wfull_bar, (uint32_t)P.k_blocks * w_kb_bytes);
            for (int kb = 0; kb < P.k_blocks; ++kb)  // the whole weight, once: [Dp rows][32 c] per k-block
                tma_load_3d(w_base + kb * w_kb_bytes, &map_w, wfull_bar, kb * kHK, 0, 0);
            int stage = 0;
            uint32_t phase = 0;
            for (int t = blockIdx.x; t < P.total_tiles; t += gridDim.x) {
                const int img = t / P.tiles_per_img, p0 = (t - img * P.tiles_per_img) * kHM;
                for (int kb = 0; kb < P.k_blocks; ++kb) {
                    mbar_wait(empty_bar + 8 * stage, phase ^ 1);
                    const uint32_t sa = smem_base + stage * kHStageBytes;
                    mbar_expect_tx(full_bar + 8 * stage, kHStageBytes);
#pragma unroll
                    for (int j = 0; j < 4; ++j)  // x[img][kb*32 .. +32][p0 + 32j .. +32]; pixels/channels past the end read as zero
                        tma_load_3d(sa + j * 4096, &map_x, full_bar + 8 * stage, p0 + 32 * j, kb * kHK, img);
                    if (++stage == kHStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {  // ===================== MMA issuer =====================
            const uint32_t idesc = make_idesc_tf32(kHM, P.Dp, /*a_mn=*/true, /*b_mn=*/false);
            mbar_wait(wfull_bar, 0);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int t = blockIdx.x; t < P.total_tiles; t += gridDim.x, ++it) {
                const int acc = it & 1;
                mbar_wait(tempty_bar + 8 * acc, ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * 256;
                for (int kb = 0; kb < P.k_blocks; ++kb) {
                    mbar_wait(rounded_bar + 8 * stage, phase);
                    tc_fence_after();
                    const uint32_t sa = smem_base + stage * kHStageBytes;
                    const uint64_t bdesc = make_smem_desc(w_base + kb * w_kb_bytes);
#pragma unroll
                    for (int k = 0; k < kHK / 8; ++k)  // A: next 8-channel atom (+1024 bytes); B: +32 bytes inside the swizzle row
                        tc_mma_tf32(d_tmem, make_smem_desc_mn_tf32(sa + k * 1024, 4096), bdesc + (uint64_t)(2 * k), idesc,
                                    (kb | k) != 0 ? 1u : 0u);
                    tc_commit(empty_bar + 8 * stage);
                    if (++stage == kHStages) { stage = 0; phase ^= 1; }
                }
                tc_commit(tfull_bar + 8 * acc);
            }
        }
    } else if (warp >= 8) {
        // ===================== rounders: x stage -> nearest TF32, in place =====================
        const int rt = threadIdx.x - 256;
        int stage = 0;
        uint32_t phase = 0;
        for (int t = blockIdx.x; t < P.total_tiles; t += gridDim.x) {
            for (int kb = 0; kb < P.k_blocks; ++kb) {
                mbar_wait(full_bar + 8 * stage, phase);
                float4* tile = reinterpret_cast<float4*>(smem_gen + stage * kHStageBytes);
#pragma unroll
                for (int i = 0; i < kHStageBytes / 16 / 128; ++i) {  // elementwise: the swizzle does not matter
                    float4 v = tile[rt + 128 * i];
                    v.x = round_tf32_h(v.x); v.y = round_tf32_h(v.y); v.z = round_tf32_h(v.z); v.w = round_tf32_h(v.w);
                    tile[rt + 128 * i] = v;
                }
                fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
                mbar_arrive(rounded_bar + 8 * stage);
                if (++stage == kHStages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue: TMEM -> + bias -> TF32 -> swizzled tile -> TMA store =====================
        const int ew = warp - 4, ml = ew * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(ew * 32) << 16;
        const uint32_t swz = (uint32_t)(ml & 7);
        const bool issuer = threadIdx.x == 4 * 32;
        int it = 0, buf = 0;
        for (int t = blockIdx.x; t < P.total_tiles; t += gridDim.x, ++it) {
            const int img = t / P.tiles_per_img, p0 = (t - img * P.tiles_per_img) * kHM;
            const int acc = it & 1;
            mbar_wait(tfull_bar + 8 * acc, (it >> 1) & 1);
            tc_fence_after();
            const uint32_t t_acc = tmem_base + lane_addr + acc * 256;
            for (int c = 0; c < P.Dp; c += 32) {
                float v[32];
                tmem_ld32(t_acc + c, v);
                tmem_ld_wait();
                if (c + 32 >= P.Dp) {   // this warp has read all of the accumulator
                    tc_fence_before();
                    mbar_arrive(tempty_bar + 8 * acc);
                }
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = round_tf32_h(v[j] + sbias[c + j]);
                const uint32_t row = store_base + buf * kHStoreBytes + ml * 128;
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(row + (((uint32_t)j ^ swz) << 4)), "f"(v[4 * j]),
                                 "f"(v[4 * j + 1]), "f"(v[4 * j + 2]), "f"(v[4 * j + 3])
                                 : "memory");
                fence_proxy_async();
                // the buffer the NEXT chunk overwrites (stored kHStoreBufs-1 chunks ago) must have been read out: waiting
                // for all but the newest kHStoreBufs-2 groups keeps kHStoreBufs-1 stores in flight
                if (issuer) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kHStoreBufs - 2) : "memory");
                named_bar_sync(1, 128);
                if (issuer) {   // pixels past the end of the image are clipped by the tensor map
                    if (img < P.B) tma_store_3d(&map_o1, store_base + buf * kHStoreBytes, c, p0, img);
                    else tma_store_3d(&map_o2, store_base + buf * kHStoreBytes, c, p0, img - P.B);
                    tma_store_commit();
                }
                if (++buf == kHStoreBufs) buf = 0;
            }
        }
        if (issuer) tma_store_wait_all();   // global writes complete before the CTA exits
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// dst[Dp][Cin] = tf32(src[D][Cin]); rows D..Dp zero
__global__ void round_weight_kernel(const float* __restrict__ src, float* __restrict__ dst, int n_src, int n_dst) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_dst; i += gridDim.x * blockDim.x)
        dst[i] = i < n_src ? round_tf32_h(__ldg(src + i)) : 0.f;
}

inline size_t align_up_h(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace

bool fnet_head_supported(int Cin, int D, int HW) {
    if (Cin < 4 || Cin % 4 != 0 || D < 1 || D > 256 || HW < 1 || HW % 4 != 0) return false;  // TMA: 16-byte strides
    return (size_t)ceil_div(Cin, kHK) * tc_feature_pitch(D) * 128 <= (size_t)kHMaxWBytes;
}

size_t fnet_head_weight_bytes(int Cin, int D) { return align_up_h((size_t)tc_feature_pitch(D) * Cin * sizeof(float), 1024); }

// x [2B][Cin][HW] -> f1n, f2n [B][HW][Dp] (TF32-rounded, channels D..Dp zero); w_rounded: fnet_head_weight_bytes scratch
int launch_fnet_head(const float* x, const float* weight, const float* bias, int B, int Cin, int D, int HW, float* f1n,
                     float* f2n, float* w_rounded, int device, cudaStream_t s) {
    RC_REQUIRE(fnet_head_supported(Cin, D, HW),
               "fnet_head: needs Cin %% 4 == 0, D <= 256, H*W %% 4 == 0 and ceil(Cin/32)*D <= 1024 (got Cin=%d D=%d HW=%d)", Cin,
               D, HW);
    EncodeTiledFn fn;
    if (int rc = get_encode_fn(&fn)) return rc;
    const int Dp = tc_feature_pitch(D);
    round_weight_kernel<<<std::min(ceil_div(Dp * Cin, 256), 148 * 4), 256, 0, s>>>(weight, w_rounded, D * Cin, Dp * Cin);
    RC_CUDA(cudaGetLastError());

    HeadParams P{};
    P.out1 = f1n; P.out2 = f2n; P.bias = bias;
    P.B = B; P.D = D; P.Dp = Dp; P.HW = HW;
    P.k_blocks = ceil_div(Cin, kHK);
    P.tiles_per_img = ceil_div(HW, kHM);
    P.total_tiles = 2 * B * P.tiles_per_img;

    CUtensorMap map_x, map_w, map_o1, map_o2;
    for (int k = 0; k < 2; ++k) {   // operand buffers [B][HW][Dp], stored as [128 px][32 columns] boxes
        cuuint64_t dims[3] = {(cuuint64_t)Dp, (cuuint64_t)HW, (cuuint64_t)B};
        cuuint64_t str[2] = {(cuuint64_t)Dp * 4, (cuuint64_t)HW * Dp * 4};
        cuuint32_t box[3] = {32, kHM, 1};
        if (int rc = encode_f32_map(fn, k ? &map_o2 : &map_o1, (void*)(k ? f2n : f1n), 3, dims, str, box,
                                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, "encoder head output"))
            return rc;
    }
    {
        cuuint64_t dims[3] = {(cuuint64_t)HW, (cuuint64_t)Cin, (cuuint64_t)(2 * B)};
        cuuint64_t str[2] = {(cuuint64_t)HW * 4, (cuuint64_t)Cin * HW * 4};
        cuuint32_t box[3] = {32, kHK, 1};
        if (int rc = encode_f32_map(fn, &map_x, (void*)x, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, "encoder head input"))
            return rc;
    }
    {
        cuuint64_t dims[3] = {(cuuint64_t)Cin, (cuuint64_t)Dp, 1};
        cuuint64_t str[2] = {(cuuint64_t)Cin * 4, (cuuint64_t)Dp * Cin * 4};
        cuuint32_t box[3] = {kHK, (cuuint32_t)Dp, 1};
        if (int rc = encode_f32_map(fn, &map_w, (void*)w_rounded, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B,
                                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, "encoder head weight"))
            return rc;
    }
    int sms = 0;
    RC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    const int smem = 1024 + kHStages * kHStageBytes + P.k_blocks * Dp * 128 + kHStoreBufs * kHStoreBytes + 256;
    RC_CUDA(cudaFuncSetAttribute(fnet_head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    fnet_head_kernel<<<P.total_tiles < sms ? P.total_tiles : sms, kHThreads, smem, s>>>(map_x, map_w, map_o1, map_o2, P);
    RC_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace raftcorr

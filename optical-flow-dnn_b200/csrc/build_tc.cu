// Correlation volume + pyramid build on the 5th-gen tensor cores (RAFTCORR_BUILD_TF32_TC).
//
//   V0[b][m][n] = <f1[b][m][:], f2[b][n][:]> / sqrt(D)           (CorrBlock.corr, raft/core/corr.py:93-116)
//   V_{l+1}     = avg_pool2d(V_l, 2, stride 2) over the target dims (corr.py:41-43)
//
// One persistent, warp-specialised kernel replaces the reference's SGEMM + scale + 3 pooling passes:
//
//   warp 0  TMA producer : A tile [128 source px][32 k] and B tile [8 x 32 target px][32 k] per k-block,
//                          128B-swizzled, K-major, straight from the NHWC (TF32-rounded) feature maps;
//                          B is fetched as a 4-D box over [b][h][w][d] so that one N tile is a spatially
//                          aligned 8x32 block of target pixels (out-of-range rows/cols are zero-filled).
//   warp 1  MMA issuer   : tcgen05.mma.cta_group::1.kind::tf32, M=128, N=256, K=8 per instruction,
//                          fp32 accumulators in TMEM, two 256-column accumulator stages (all 512 columns)
//                          so the MMAs of tile i+1 overlap the epilogue of tile i.
//   warp 2  TMEM allocator
//   warps 4-7  epilogue A: TMEM lane == source pixel, so one thread owns the whole 8x32 target block of
//                          its source pixel: tcgen05.ld two rows at a time, scale by 1/sqrt(D), stage the
//                          level-0 rows in swizzled shared memory and TMA-store them ([128 px][128 B] boxes,
//                          double-buffered, bounds clipped by the tensor map).
//   warps 8-11 epilogue B: read the same accumulator again and reduce 2x2 / 4x4 / 8x8 means in registers
//                          for levels 1..3 (floor-mode pooling == block means over the floor-cropped region,
//                          since floor(floor(n/2)/2) == floor(n/4)); 256-bit stores, off the critical path
//                          of the level-0 store pipeline.
//
// The kernel is bound by the fp32 pyramid write (174 KB per tile vs 4096 tensor cycles per tile),
// see DESIGN.md; the MMA pipe has slack, the epilogue/TMA-store path is what is tuned.  Ablations on
// B200 (B=8, 55x128, D=256, L=4; profiles/README.md): level-0 stores alone 280 us (6.0 TB/s), + pooled
// levels 418 us, + operand loads 541 us -- the operand ingress (384 KB per tile) competes with the
// stores for the SM<->L2 port (l1tex2xbar 75 % busy), it is not hidden behind them.
#include <cstdlib>

#include "tc_util.cuh"

namespace raftcorr {

namespace {

constexpr int kTileM = 128;       // source pixels per tile (TMEM lanes)
constexpr int kTileH = 8;         // target rows per tile
constexpr int kTileW = 32;        // target cols per tile
constexpr int kTileN = kTileH * kTileW;  // 256 = UMMA N
constexpr int kBlockK = 32;       // floats per k-block = one 128-byte swizzle row
constexpr int kUmmaK = 8;         // tf32: 32 bytes per instruction
constexpr int kABytes = kTileM * kBlockK * 4;   // 16 KB
constexpr int kBBytes = kTileN * kBlockK * 4;   // 32 KB
constexpr int kStageBytes = kABytes + kBBytes;  // 48 KB
constexpr int kRowTileBytes = kTileM * kTileW * 4;  // 16 KB: one target row for 128 source pixels
// pipeline shape: NS operand stages, NB level-0 store buffers of RP target rows each
constexpr int smem_bytes(int ns, int nb, int rp) { return ns * kStageBytes + nb * rp * kRowTileBytes + 1024 /*align*/ + 256 /*barriers*/; }
constexpr int kThreads = 384;          // 4 control warps + 8 epilogue warps
constexpr int kEpilogueThreads = 256;
constexpr uint32_t kTmemCols = 512;
constexpr int kDefaultBuildMode = 0;

// kind::tf32 instruction descriptor: D=f32 (bits 4-5 = 1), A=B=tf32 (bits 7-9, 10-12 = 2), both K-major,
// N>>3 at bits 17-22, M>>4 at bits 24-28.
constexpr uint32_t kInstrDesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kTileN >> 3) << 17) |
                                ((uint32_t)(kTileM >> 4) << 24);

constexpr uint32_t kInstrDesc2 = make_idesc_tf32(2 * kTileM, kTileN, false, false);  // cta_group::2: M = 256 over the pair
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared::cluster address -> the pair's leader

template <bool PAIR>
__device__ __forceinline__ void tempty_arrive(uint32_t bar) {
    if (PAIR)  // the leader's MMA thread waits for the epilogues of BOTH CTAs
        asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar & kPeerBitMask) : "memory");
    else
        mbar_arrive(bar);
}

struct BuildParams {
    float* lvl_base[4];     // pooled levels 1..3 (index by level); [0] unused (TMA path)
    int lvl_h[4], lvl_w[4], lvl_pitch[4];
    int64_t lvl_img_stride[4];
    int B, N1, H2, W2, L;
    int m_tiles, m_groups, bands, w_tiles, k_blocks;  // m_groups = ceil(m_tiles / cluster size)
    int total_tiles;
    // Work units: a CTA takes `unit_bands` vertically consecutive tiles (bands bg * unit_bands ...) of one (m-tile,
    // w-tile) back to back.  1 for the row / row-pair layouts; 4 for the row-quad layout, where the epilogue of the
    // pooled levels carries level-2 / level-3 rows from band to band until a whole quad can be stored.
    int unit_bands, band_groups, total_units;
    int lsu_store;  // A/B: level-0 rows leave through LSU stores instead of TMA
    int store_hint; // pyramid stores carry an L2 evict-first policy (default; RAFTCORR_BUILD_STORE_HINT=0: A/B)
    float scale;
    // RAFTCORR_BUILD_F16_TC: operands are fp16 with one power-of-two scale per 64-pixel operand BLOCK (fmap1: half an
    // m-tile = 64 TMEM lanes; fmap2: one row pair of the 8x32 target tile = one epilogue phase) -- the same 11
    // significant bits as TF32 at half the operand bytes, half the MMA instructions (K = 16) and half the accumulator
    // read-modify-writes; undoing the scales is one scalar per thread times one per phase
    int f16, k_elems;            // k_elems: operand elements per 128-byte k-block (32 tf32 / 64 fp16)
    const float* inv_a;          // [B][2 * m_tiles]
    const float* inv_b;          // [B][4 * bands][w_tiles]
};

// kind::f16 twin of tc_mma_tf32 (A = B = fp16, D = fp32): format fields 0, same shape fields
constexpr uint32_t kInstrDescF16 = (1u << 4) | ((uint32_t)(kTileN >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// MODE 0: one CTA per tile.  MODE 1: CTA pair, each loads half of the B tile and TMA-multicasts it to both
// (every CTA still ingests the whole tile).  MODE 2: CTA pair issuing ONE tcgen05.mma.cta_group::2 (M = 256):
// each CTA keeps only its half of B in shared memory, so the operand ingress per SM drops by a third.
template <int N>
__device__ __forceinline__ void tma_store_wait_read_n() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }

// IL: raftcorr_pyramid_layout::interleaved -- 0 plain rows, 1 row pairs, 2 row quads interleaved; 1 and 2 need RP == 2
// (a store buffer then holds two [128 px][128 B] boxes: a row pair as columns 0..15 | 16..31, or 4 rows x 8 | 8 columns).
template <int MODE, int kStages = 3, int kNumStoreBufs = 2, int RP = 2, int IL = 0>
__global__ void __launch_bounds__(kThreads, 1)
build_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                const __grid_constant__ CUtensorMap map_out, const BuildParams P) {
    constexpr int CL = MODE == 0 ? 1 : 2;
    constexpr bool PAIR_MMA = MODE == 2;
    static_assert(!IL || RP == 2, "the interleaved layouts stage two boxes per phase");
    constexpr int kStoreBufBytes = RP * kRowTileBytes;
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment: required by the 128B swizzle atoms of TMA and UMMA
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t stage_base = smem_base;                                  // kStages x (A | B)
    // pair MMA: a CTA holds its A tile and HALF of the B tile per stage -> 32 KB stages, so more of them fit
    constexpr int kStageStride = PAIR_MMA ? kABytes + kBBytes / 2 : kStageBytes;
    const uint32_t store_base = smem_base + kStages * kStageStride;         // kNumStoreBufs x 2 rows
    const uint32_t bar_base = store_base + kNumStoreBufs * kStoreBufBytes;  // mbarriers (8 B each)
    const uint32_t full_bar = bar_base;                       // [kStages]
    const uint32_t empty_bar = bar_base + 8 * kStages;        // [kStages]
    const uint32_t tfull_bar = bar_base + 16 * kStages;       // [2]
    const uint32_t tempty_bar = tfull_bar + 16;               // [2]
    const uint32_t tmem_slot = tempty_bar + 16;               // u32 written by tcgen05.alloc
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&map_a);
        prefetch_tensormap(&map_b);
        prefetch_tensormap(&map_out);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(full_bar + 8 * s, 1);
            mbar_init(empty_bar + 8 * s, MODE == 1 ? 2 : 1);  // multicast pair: both CTAs must have consumed the stage
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull_bar + 8 * a, 1);
            mbar_init(tempty_bar + 8 * a, PAIR_MMA ? 2 * kEpilogueThreads : kEpilogueThreads);  // pair MMA: both CTAs' epilogues
        }
        fence_mbar_init();
    }
    if (warp == 2) {
        if (PAIR_MMA) {  // both CTAs of the pair allocate the same columns (same warp id, same smem slot)
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(kTmemCols)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(kTmemCols)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    if (CL > 1) cluster_sync_all();  // peers' barriers are initialised before anyone multicasts / arrives remotely
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));
    const int crank = CL > 1 ? (int)(blockIdx.x % CL) : 0;      // rank inside the cluster (cluster dims = (CL,1,1))
    const int first = (int)(blockIdx.x / CL), stride = (int)(gridDim.x / CL);

    const int units_per_m = P.band_groups * P.w_tiles;
// (w-tile, band group, m-tile, batch item) of work unit u; the unit's tiles are bands bg * unit_bands + bi
#define RAFTCORR_DECODE_UNIT(u)                                      \
    const int wt = (u) % P.w_tiles;                                  \
    const int bg = ((u) / P.w_tiles) % P.band_groups;                \
    const int mt = (((u) / units_per_m) % P.m_groups) * CL + crank;  \
    const int b = (u) / (units_per_m * P.m_groups);

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int u = first; u < P.total_units; u += stride) {
              RAFTCORR_DECODE_UNIT(u)  // mt may be a dummy tile past the end (pair modes)
              for (int bi = 0; bi < P.unit_bands; ++bi) {
                const int band = bg * P.unit_bands + bi;
                if (band >= P.bands) break;
                for (int kb = 0; kb < P.k_blocks; ++kb) {
                    mbar_wait(empty_bar + 8 * stage, phase ^ 1);
                    const uint32_t sa = stage_base + stage * kStageStride;
                    const uint32_t sb = sa + kABytes;
                    if (PAIR_MMA) {
                        // both CTAs' loads complete on the LEADER's barrier (peer bit cleared); it expects both halves
                        const uint32_t lead_bar = (full_bar + 8 * stage) & kPeerBitMask;
                        if (crank == 0) mbar_expect_tx(full_bar + 8 * stage, 2 * (kABytes + kBBytes / 2));
                        tma_load_3d_2sm(sa, &map_a, lead_bar, kb * P.k_elems, mt * kTileM, b);
                        tma_load_4d_2sm(sb, &map_b, lead_bar, kb * P.k_elems, wt * kTileW, band * kTileH + crank * (kTileH / 2), b);
                        if (++stage == kStages) { stage = 0; phase ^= 1; }
                        continue;
                    }
                    mbar_expect_tx(full_bar + 8 * stage, kStageBytes);
                    tma_load_3d(sa, &map_a, full_bar + 8 * stage, kb * P.k_elems, mt * kTileM, b);
                    if (CL == 1) {
                        tma_load_4d(sb, &map_b, full_bar + 8 * stage, kb * P.k_elems, wt * kTileW, band * kTileH, b);
                    } else {
                        // this CTA fetches 1/CL of the B tile (kTileH/CL target rows) and multicasts it into the
                        // same smem offset of every CTA of the cluster; each full barrier collects all CL parts
                        tma_load_4d_mcast(sb + crank * (kBBytes / CL), &map_b, full_bar + 8 * stage, kb * P.k_elems,
                                          wt * kTileW, band * kTileH + crank * (kTileH / CL), b, (uint16_t)((1u << CL) - 1));
                    }
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
              }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (pair MMA: the leader CTA issues for both) =====================
        if (lane == 0 && !(PAIR_MMA && crank != 0)) {
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int u = first; u < P.total_units; u += stride) {
              const int bg_ = (u / P.w_tiles) % P.band_groups;
              for (int bi = 0; bi < P.unit_bands && bg_ * P.unit_bands + bi < P.bands; ++bi, ++it) {
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                mbar_wait(tempty_bar + 8 * acc, acc_phase ^ 1);  // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * kTileN;
                for (int kb = 0; kb < P.k_blocks; ++kb) {
                    mbar_wait(full_bar + 8 * stage, phase);
                    tc_fence_after();
                    const uint32_t sa = stage_base + stage * kStageStride;
                    const uint64_t adesc = make_smem_desc(sa);
                    const uint64_t bdesc = make_smem_desc(sa + kABytes);
#pragma unroll
                    for (int k = 0; k < kBlockK / kUmmaK; ++k) {
                        // advance 32 bytes along K inside the swizzle row: +2 in the (>>4) start-address field
                        if (PAIR_MMA)
                            tc_mma_tf32_2sm(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), kInstrDesc2,
                                            (kb | k) != 0 ? 1u : 0u);
                        else if (P.f16)  // K = 16 halves = the same 32 bytes per instruction
                            tc_mma_f16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), kInstrDescF16,
                                       (kb | k) != 0 ? 1u : 0u);
                        else
                            tc_mma_tf32(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), kInstrDesc,
                                        (kb | k) != 0 ? 1u : 0u);
                    }
                    if (MODE == 0) tc_commit(empty_bar + 8 * stage);  // frees the smem slot when these MMAs retire
                    else if (MODE == 1) tc_commit_mcast(empty_bar + 8 * stage, 3);  // ... in both CTAs of the pair
                    else tc_commit_2sm_mcast(empty_bar + 8 * stage, 3);
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
                if (PAIR_MMA) tc_commit_2sm_mcast(tfull_bar + 8 * acc, 3);  // accumulators complete in both CTAs
                else tc_commit(tfull_bar + 8 * acc);                        // accumulator complete -> epilogue
              }
            }
        }
    } else if (warp >= 4 && warp < 8) {
        // ===================== epilogue A (warps 4-7): level 0 =====================
        // TMEM lane == source pixel: each thread moves the 8x32 block of its source pixel, two rows at a
        // time: tcgen05.ld -> scale -> swizzled smem staging -> TMA store of [128 px][128 B] boxes.
        const int ew = warp & 3;                  // TMEM lane quarter this warp may access (hardware rule)
        const int ml = ew * 32 + lane;
        const bool issuer = (threadIdx.x == 4 * 32);
        const uint32_t lane_addr = (uint32_t)(ew * 32) << 16;
        const uint32_t swz = (uint32_t)(ml & 7);
        int it = 0;
        int buf = 0;
        for (int u = first; u < P.total_units; u += stride) {
          RAFTCORR_DECODE_UNIT(u)
          for (int bi = 0; bi < P.unit_bands && bg * P.unit_bands + bi < P.bands; ++bi, ++it) {
            const int band = bg * P.unit_bands + bi;
            const int acc = it & 1;
            // block-scaled fp16 operands: undo the power-of-two block scales together with 1/sqrt(D); all five scalars
            // are fetched before the accumulator wait so that their latency never sits on the store chain
            float sc4[4] = {P.scale, P.scale, P.scale, P.scale};
            if (P.f16) {
                const float sa = P.scale * __ldg(P.inv_a + (b * P.m_tiles + mt) * 2 + (ml >> 6));
                const float* sb = P.inv_b + (int64_t)(b * P.bands + band) * 4 * P.w_tiles + wt;
#pragma unroll
                for (int p = 0; p < 4; ++p) sc4[p] = sa * __ldg(sb + p * P.w_tiles);
            }
            mbar_wait(tfull_bar + 8 * acc, (it >> 1) & 1);
            tc_fence_after();
            const uint32_t t_acc = tmem_base + lane_addr + acc * kTileN;
            if constexpr (IL == 2) {
                // row quads: a phase is 4 target rows x 16 columns; column x of a quad is one 16-byte chunk (r0 r1 r2 r3),
                // so the phase leaves as two [128 px][128 B] boxes of 8 columns each
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    const int rgp = p >> 1, chh = p & 1;
                    float r0[16], r1[16], r2[16], r3[16];
                    tmem_ld16(t_acc + (4 * rgp + 0) * kTileW + 16 * chh, r0);
                    tmem_ld16(t_acc + (4 * rgp + 1) * kTileW + 16 * chh, r1);
                    tmem_ld16(t_acc + (4 * rgp + 2) * kTileW + 16 * chh, r2);
                    tmem_ld16(t_acc + (4 * rgp + 3) * kTileW + 16 * chh, r3);
                    tmem_ld_wait();
                    if (p == 3) {  // this warp has read all of the accumulator
                        tc_fence_before();
                        tempty_arrive<PAIR_MMA>(tempty_bar + 8 * acc);
                    }
                    const float s01 = sc4[2 * rgp], s23 = sc4[2 * rgp + 1];
                    const uint32_t sa = store_base + buf * kStoreBufBytes + ml * 128;
                    const uint32_t sb = sa + kRowTileBytes;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const uint32_t off = ((uint32_t)j ^ swz) << 4;
                        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(sa + off), "f"(r0[j] * s01), "f"(r1[j] * s01),
                                     "f"(r2[j] * s23), "f"(r3[j] * s23)
                                     : "memory");
                        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(sb + off), "f"(r0[8 + j] * s01),
                                     "f"(r1[8 + j] * s01), "f"(r2[8 + j] * s23), "f"(r3[8 + j] * s23)
                                     : "memory");
                    }
                    fence_proxy_async();
                    if (issuer) tma_store_wait_read_n<kNumStoreBufs - 2>();
                    named_bar_sync(1, 128);
                    if (issuer) {  // map_out: [4 * W interleaved floats][row quads][N1][B]
                        const uint32_t s0 = store_base + buf * kStoreBufBytes;
                        const int x4 = 4 * (wt * kTileW + 16 * chh);
                        if (P.store_hint) {
                            const uint64_t pol = l2_policy_evict_first();
                            tma_store_4d_hint(&map_out, s0, x4, band * 2 + rgp, mt * kTileM, b, pol);
                            tma_store_4d_hint(&map_out, s0 + kRowTileBytes, x4 + 32, band * 2 + rgp, mt * kTileM, b, pol);
                        } else {
                            tma_store_4d(&map_out, s0, x4, band * 2 + rgp, mt * kTileM, b);
                            tma_store_4d(&map_out, s0 + kRowTileBytes, x4 + 32, band * 2 + rgp, mt * kTileM, b);
                        }
                        tma_store_commit();
                    }
                    if (++buf == kNumStoreBufs) buf = 0;
                }
                continue;
            }
#pragma unroll
            for (int p = 0; p < kTileH / RP; ++p) {
                float ra[32], rb[32];
                const float tsc = sc4[RP == 2 ? p : p >> 1];
                tmem_ld32(t_acc + (RP * p) * kTileW, ra);
                if (RP == 2) tmem_ld32(t_acc + (2 * p + 1) * kTileW, rb);
                tmem_ld_wait();
                if (p == kTileH / RP - 1) {  // this warp has read all of the accumulator
                    tc_fence_before();
                    tempty_arrive<PAIR_MMA>(tempty_bar + 8 * acc);
                }
                const uint32_t sa = store_base + buf * kStoreBufBytes + ml * 128;
                const uint32_t sb = sa + kRowTileBytes;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint32_t off = ((uint32_t)j ^ swz) << 4;
                    if (IL == 1) {
                        // the two rows leave as ONE run (a0 b0 a1 b1 ...): first box = target columns 0..15, second = 16..31
                        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(sa + off), "f"(ra[2 * j] * tsc),
                                     "f"(rb[2 * j] * tsc), "f"(ra[2 * j + 1] * tsc), "f"(rb[2 * j + 1] * tsc)
                                     : "memory");
                        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(sb + off), "f"(ra[16 + 2 * j] * tsc),
                                     "f"(rb[16 + 2 * j] * tsc), "f"(ra[17 + 2 * j] * tsc), "f"(rb[17 + 2 * j] * tsc)
                                     : "memory");
                        continue;
                    }
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(sa + off), "f"(ra[4 * j] * tsc),
                                 "f"(ra[4 * j + 1] * tsc), "f"(ra[4 * j + 2] * tsc), "f"(ra[4 * j + 3] * tsc)
                                 : "memory");
                    if (RP == 2)
                        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(sb + off), "f"(rb[4 * j] * tsc),
                                     "f"(rb[4 * j + 1] * tsc), "f"(rb[4 * j + 2] * tsc), "f"(rb[4 * j + 3] * tsc)
                                     : "memory");
                }
                if (P.lsu_store && RP == 2 && !IL) {
                    // A/B variant: drain the staged rows with coalesced 128-bit stores (8 lanes per 128-byte row)
                    // instead of TMA, so that the TMA unit only carries the operand loads
                    named_bar_sync(1, 128);
                    const uint32_t s0 = store_base + buf * kStoreBufBytes;
                    const int tid = threadIdx.x - 128;
#pragma unroll 4
                    for (int k = 0; k < 16; ++k) {
                        const int idx = k * 128 + tid;
                        const int rr = idx >> 3, c16 = idx & 7;       // staged row (0..255), 16-byte chunk in it
                        const int mm = mt * kTileM + (rr & 127), hh = band * kTileH + 2 * p + (rr >> 7);
                        const int ww = wt * kTileW + c16 * 4;
                        float4 v;
                        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                                     : "r"(s0 + rr * 128 + ((c16 ^ (rr & 7)) << 4)));
                        if (mm < P.N1 && hh < P.lvl_h[0] && ww + 4 <= P.lvl_pitch[0])
                            *reinterpret_cast<float4*>(P.lvl_base[0] + ((int64_t)b * P.N1 + mm) * P.lvl_img_stride[0] +
                                                       (int64_t)hh * P.lvl_pitch[0] + ww) = v;
                    }
                } else {
                    fence_proxy_async();
                    // the buffer the NEXT phase overwrites (used kNumStoreBufs-1 phases ago) must have been read
                    // out; waiting here, one phase early, keeps kNumStoreBufs-1 store groups in flight
                    if (issuer) tma_store_wait_read_n<kNumStoreBufs - 2>();
                    named_bar_sync(1, 128);
                    if (issuer) {
                        const uint32_t s0 = store_base + buf * kStoreBufBytes;
                        const int h = band * kTileH + RP * p;
                        if (IL == 1) {  // map_out: [2 * W interleaved floats][pair rows][N1][B]
                            tma_store_4d(&map_out, s0, wt * (2 * kTileW), h >> 1, mt * kTileM, b);
                            tma_store_4d(&map_out, s0 + kRowTileBytes, wt * (2 * kTileW) + kTileW, h >> 1, mt * kTileM, b);
                        } else {
                            tma_store_4d(&map_out, s0, wt * kTileW, h, mt * kTileM, b);
                            if (RP == 2) tma_store_4d(&map_out, s0 + kRowTileBytes, wt * kTileW, h + 1, mt * kTileM, b);
                        }
                        tma_store_commit();
                    }
                }
                if (++buf == kNumStoreBufs) buf = 0;
            }
          }
        }
        if (issuer) tma_store_wait_all();  // global writes complete before the CTA exits
    } else if (warp >= 8) {
        // ===================== epilogue B (warps 8-11): levels 1..3 =====================
        // Re-reads the same accumulator from TMEM (cheap) so that the pooled levels never sit on the
        // critical path of the level-0 store pipeline.  Floor-mode cascaded pooling == block means over
        // the floor-cropped region, so 2x2 / 4x4 / 8x8 means are thread-local register reductions; the
        // 1/sqrt(D) scale is applied to the means.  Full 32-byte sectors per store (levels 1, 2).
        const uint64_t pool_pol = l2_policy_evict_first();  // used when P.store_hint is set
        const int ew = warp & 3;
        const int ml = ew * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(ew * 32) << 16;
        int it = 0;
        for (int u = first; u < P.total_units; u += stride) {
          RAFTCORR_DECODE_UNIT(u)
          // row-quad layout: level-2 rows of an even band wait for the odd band, level-3 rows for the end of the unit
          float l2keep[2][8], l3keep[3][4];
          for (int bi = 0; bi < P.unit_bands && bg * P.unit_bands + bi < P.bands; ++bi, ++it) {
            const int band = bg * P.unit_bands + bi;
            const bool last_band = bi == P.unit_bands - 1 || band == P.bands - 1;
            const int acc = it & 1;
            const int m = mt * kTileM + ml;
            const bool m_ok = m < P.N1;
            const int w0 = wt * kTileW;
            float q4[4] = {0.25f * P.scale, 0.25f * P.scale, 0.25f * P.scale, 0.25f * P.scale};
            if (P.f16) {
                const float sa = 0.25f * P.scale * __ldg(P.inv_a + (b * P.m_tiles + mt) * 2 + (ml >> 6));
                const float* sb = P.inv_b + (int64_t)(b * P.bands + band) * 4 * P.w_tiles + wt;
#pragma unroll
                for (int p = 0; p < 4; ++p) q4[p] = sa * __ldg(sb + p * P.w_tiles);
            }
            mbar_wait(tfull_bar + 8 * acc, (it >> 1) & 1);
            tc_fence_after();
            const uint32_t t_acc = tmem_base + lane_addr + acc * kTileN;
            if (P.L <= 1) {  // nothing to pool: just release the accumulator
                tc_fence_before();
                tempty_arrive<PAIR_MMA>(tempty_bar + 8 * acc);
                continue;
            }
            const int64_t pix = (int64_t)b * P.N1 + m;
            if constexpr (IL == 2) {
                // Row quads.  Level 1: the tile's 4 pooled rows are exactly one quad (group `band`), 16 columns, written
                // as 32-byte sectors (2 columns x 4 rows).  Rows past the level's height inside its last quad are written
                // as zeros (the lookup's tensor maps see whole quads).
                const int hp1 = (P.lvl_h[1] + 3) & ~3;
                float l2row[2][8];
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    float l1h[4][8];
#pragma unroll
                    for (int p = 0; p < 4; ++p) {
                        float ra[16], rb[16];
                        tmem_ld16(t_acc + (2 * p) * kTileW + 16 * half, ra);
                        tmem_ld16(t_acc + (2 * p + 1) * kTileW + 16 * half, rb);
                        tmem_ld_wait();
                        if (half == 1 && p == 3) {
                            tc_fence_before();
                            tempty_arrive<PAIR_MMA>(tempty_bar + 8 * acc);
                        }
                        const float q = (band * 4 + p < P.lvl_h[1]) ? q4[p] : 0.f;
#pragma unroll
                        for (int j = 0; j < 8; ++j) l1h[p][j] = (ra[2 * j] + ra[2 * j + 1] + rb[2 * j] + rb[2 * j + 1]) * q;
                    }
                    if (m_ok && band * 4 < hp1) {
                        const int x1 = (w0 >> 1) + 8 * half;
                        float* dst = P.lvl_base[1] + pix * P.lvl_img_stride[1] + (int64_t)band * (4 * P.lvl_pitch[1]) + 4 * x1;
#pragma unroll
                        for (int c = 0; c < 4; ++c)
                            if (x1 + 2 * c + 2 <= P.lvl_pitch[1]) {
                                const float v[8] = {l1h[0][2 * c], l1h[1][2 * c], l1h[2][2 * c], l1h[3][2 * c],
                                                    l1h[0][2 * c + 1], l1h[1][2 * c + 1], l1h[2][2 * c + 1], l1h[3][2 * c + 1]};
                                st_global_v8(dst + 8 * c, v, pool_pol, P.store_hint != 0);
                            }
                    }
#pragma unroll
                    for (int rr = 0; rr < 2; ++rr)
#pragma unroll
                        for (int c = 0; c < 4; ++c)
                            l2row[rr][4 * half + c] = (l1h[2 * rr][2 * c] + l1h[2 * rr][2 * c + 1] + l1h[2 * rr + 1][2 * c] +
                                                       l1h[2 * rr + 1][2 * c + 1]) * 0.25f;
                }
                // Level 2: the tile's two rows are half of quad band / 2 -- an even band keeps them, the odd band (or the
                // end of the unit) stores the whole quad: 8 columns x 4 rows = 128 contiguous bytes per thread.
                // Level 3: one row per band = one slot of quad bg (a unit is exactly 4 bands): stored at the unit's end.
                // Rows past a level's height inside its last quad are zeros (the lookup's tensor maps see whole quads).
                if (P.L > 2) {
                    const bool ok0 = band * 2 < P.lvl_h[2], ok1 = band * 2 + 1 < P.lvl_h[2];
                    if (!(bi & 1) && !last_band) {
#pragma unroll
                        for (int c = 0; c < 8; ++c) { l2keep[0][c] = ok0 ? l2row[0][c] : 0.f; l2keep[1][c] = ok1 ? l2row[1][c] : 0.f; }
                    } else if (m_ok && (band >> 1) * 4 < P.lvl_h[2]) {
                        const bool odd = bi & 1;  // odd band: slots 2,3 of the quad are this band's rows; else they are zeros
                        const int x2 = w0 >> 2;
                        float* dst = P.lvl_base[2] + pix * P.lvl_img_stride[2] + (int64_t)(band >> 1) * (4 * P.lvl_pitch[2]) + 4 * x2;
#pragma unroll
                        for (int c = 0; c < 4; ++c)
                            if (x2 + 2 * c + 2 <= P.lvl_pitch[2]) {
                                float v[8];
#pragma unroll
                                for (int e = 0; e < 2; ++e) {
                                    const int cc = 2 * c + e;
                                    v[4 * e + 0] = odd ? l2keep[0][cc] : (ok0 ? l2row[0][cc] : 0.f);
                                    v[4 * e + 1] = odd ? l2keep[1][cc] : (ok1 ? l2row[1][cc] : 0.f);
                                    v[4 * e + 2] = odd && ok0 ? l2row[0][cc] : 0.f;
                                    v[4 * e + 3] = odd && ok1 ? l2row[1][cc] : 0.f;
                                }
                                st_global_v8(dst + 8 * c, v, pool_pol, P.store_hint != 0);
                            }
                    }
                }
                if (P.L > 3) {
                    float l3[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        l3[c] = band < P.lvl_h[3] ? (l2row[0][2 * c] + l2row[0][2 * c + 1] + l2row[1][2 * c] + l2row[1][2 * c + 1]) * 0.25f : 0.f;
                    if (!last_band) {
#pragma unroll
                        for (int sl = 0; sl < 3; ++sl)
                            if (sl == bi) {
#pragma unroll
                                for (int c = 0; c < 4; ++c) l3keep[sl][c] = l3[c];
                            }
                    } else if (m_ok && bg * 4 < P.lvl_h[3]) {
                        const int x3 = w0 >> 3;
                        float* dst = P.lvl_base[3] + pix * P.lvl_img_stride[3] + (int64_t)bg * (4 * P.lvl_pitch[3]) + 4 * x3;
#pragma unroll
                        for (int c2 = 0; c2 < 2; ++c2)
                            if (x3 + 2 * c2 + 2 <= P.lvl_pitch[3]) {
                                float v[8];
#pragma unroll
                                for (int e = 0; e < 2; ++e) {
                                    const int cc = 2 * c2 + e;
#pragma unroll
                                    for (int sl = 0; sl < 4; ++sl)
                                        v[4 * e + sl] = sl < bi ? l3keep[sl < 3 ? sl : 0][cc] : (sl == bi ? l3[cc] : 0.f);
                                }
                                st_global_v8(dst + 8 * c2, v, pool_pol, P.store_hint != 0);
                            }
                    }
                }
                continue;
            }
            float l1prev[16], l2prev[8];
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                float ra[32], rb[32];
                tmem_ld32(t_acc + (2 * p) * kTileW, ra);
                tmem_ld32(t_acc + (2 * p + 1) * kTileW, rb);
                tmem_ld_wait();
                if (p == 3) {
                    tc_fence_before();
                    tempty_arrive<PAIR_MMA>(tempty_bar + 8 * acc);
                }
                float l1[16];
                const float q = q4[p];
#pragma unroll
                for (int j = 0; j < 16; ++j) l1[j] = (ra[2 * j] + ra[2 * j + 1] + rb[2 * j] + rb[2 * j + 1]) * q;
                const int h1 = band * 4 + p;
                if (IL) {
                    // interleaved: rows (h1-1, h1) form a pair and leave together at odd p; an odd last row is paired with zeros
                    if ((p & 1) && m_ok && h1 - 1 < P.lvl_h[1]) {
                        const bool odd_ok = h1 < P.lvl_h[1];
                        float* dst = P.lvl_base[1] + pix * P.lvl_img_stride[1] + (int64_t)(h1 >> 1) * (2 * P.lvl_pitch[1]) + w0;
#pragma unroll
                        for (int c = 0; c < 4; ++c)
                            if ((w0 >> 1) + 4 * c + 4 <= P.lvl_pitch[1]) {
                                float v[8];
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    v[2 * e] = l1prev[4 * c + e];
                                    v[2 * e + 1] = odd_ok ? l1[4 * c + e] : 0.f;
                                }
                                st_global_v8(dst + 8 * c, v, pool_pol, P.store_hint != 0);
                            }
                    }
                } else if (m_ok && h1 < P.lvl_h[1]) {
                    float* dst = P.lvl_base[1] + pix * P.lvl_img_stride[1] + (int64_t)h1 * P.lvl_pitch[1] + (w0 >> 1);
#pragma unroll
                    for (int c = 0; c < 2; ++c)
                        if ((w0 >> 1) + 8 * c + 8 <= P.lvl_pitch[1]) st_global_v8(dst + 8 * c, l1 + 8 * c, pool_pol, P.store_hint != 0);
                }
                if ((p & 1) == 0) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) l1prev[j] = l1[j];
                } else if (P.L > 2) {
                    float l2[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        l2[j] = (l1prev[2 * j] + l1prev[2 * j + 1] + l1[2 * j] + l1[2 * j + 1]) * 0.25f;
                    const int h2 = band * 2 + (p >> 1);
                    if (IL) {
                        if (p == 3 && m_ok && h2 - 1 < P.lvl_h[2]) {
                            const bool odd_ok = h2 < P.lvl_h[2];
                            float* dst = P.lvl_base[2] + pix * P.lvl_img_stride[2] + (int64_t)band * (2 * P.lvl_pitch[2]) + (w0 >> 1);
#pragma unroll
                            for (int c = 0; c < 2; ++c)
                                if ((w0 >> 2) + 4 * c + 4 <= P.lvl_pitch[2]) {
                                    float v[8];
#pragma unroll
                                    for (int e = 0; e < 4; ++e) {
                                        v[2 * e] = l2prev[4 * c + e];
                                        v[2 * e + 1] = odd_ok ? l2[4 * c + e] : 0.f;
                                    }
                                    st_global_v8(dst + 8 * c, v, pool_pol, P.store_hint != 0);
                                }
                        }
                    } else if (m_ok && h2 < P.lvl_h[2] && (w0 >> 2) + 8 <= P.lvl_pitch[2]) {
                        st_global_v8(P.lvl_base[2] + pix * P.lvl_img_stride[2] + (int64_t)h2 * P.lvl_pitch[2] + (w0 >> 2), l2, pool_pol, P.store_hint != 0);
                    }
                    if (p == 1) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) l2prev[j] = l2[j];
                    } else if (P.L > 3) {
                        float4 l3;
                        l3.x = (l2prev[0] + l2prev[1] + l2[0] + l2[1]) * 0.25f;
                        l3.y = (l2prev[2] + l2prev[3] + l2[2] + l2[3]) * 0.25f;
                        l3.z = (l2prev[4] + l2prev[5] + l2[4] + l2[5]) * 0.25f;
                        l3.w = (l2prev[6] + l2prev[7] + l2[6] + l2[7]) * 0.25f;
                        if (m_ok && band < P.lvl_h[3] && (w0 >> 3) + 4 <= P.lvl_pitch[3]) {
                            if (IL) {
                                // one level-3 row per tile: its 4 values go to every other slot of the pair row; the tile
                                // that owns the last (even) row of an odd-height level also writes the zero partner row
                                float* dst = P.lvl_base[3] + pix * P.lvl_img_stride[3] + (int64_t)(band >> 1) * (2 * P.lvl_pitch[3]) +
                                             (w0 >> 2) + (band & 1);
                                dst[0] = l3.x; dst[2] = l3.y; dst[4] = l3.z; dst[6] = l3.w;
                                if (!(band & 1) && band + 1 >= P.lvl_h[3]) dst[1] = dst[3] = dst[5] = dst[7] = 0.f;
                            } else {
                                *reinterpret_cast<float4*>(P.lvl_base[3] + pix * P.lvl_img_stride[3] +
                                                           (int64_t)band * P.lvl_pitch[3] + (w0 >> 3)) = l3;
                            }
                        }
                    }
                }
            }
          }
        }
    }
#undef RAFTCORR_DECODE_UNIT

    tc_fence_before();
    if (CL > 1) cluster_sync_all();  // no CTA may exit while a peer can still multicast into / arrive on it
    else __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        if (PAIR_MMA)
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
        else
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

// ---- host side ----------------------------------------------------------------------------------
}  // namespace

int tc_feature_pitch(int D) { return (D + kBlockK - 1) / kBlockK * kBlockK; }
int tc_feature_pitch_f16(int D) { return (D + 63) / 64 * 64; }

// f1n/f2n: [B][H*W][Dp] fp32, TF32-rounded, zero-padded to Dp = tc_feature_pitch(D)
int launch_build_tc(const float* f1n, const float* f2n, int D, float* pyramid, const raftcorr_pyramid_layout& lay,
                    int device, cudaStream_t s) {
    return launch_build_tc_ex(f1n, f2n, D, pyramid, lay, device, s, false, nullptr, nullptr);
}

// f16: f1n/f2n are [B][H*W][Dp] fp16 (Dp = tc_feature_pitch_f16(D)), block-scaled by launch_f16_operands, whose
// inverse scales are inv_a [B][2 ceil(N/128)] and inv_b [B][4 ceil(H/8)][ceil(W/32)]
// split: the K axis holds three fp16 terms per feature ([hi|lo|hi] x [hi|hi|lo], launch_f16_operands): 3 * Dp halves
int launch_build_tc_ex(const void* f1n, const void* f2n, int D, float* pyramid, const raftcorr_pyramid_layout& lay,
                       int device, cudaStream_t s, bool f16, const float* inv_a, const float* inv_b, bool split) {
    const int fusedL = lay.L < 4 ? lay.L : 4;  // levels 0..3 come out of the epilogue; deeper ones are pooled afterwards
    EncodeTiledFn fn;
    if (int rc = get_encode_fn(&fn)) return rc;
    const int B = lay.B, H = lay.H, W = lay.W, N1 = H * W;
    RC_REQUIRE(!split || f16, "build: the three-term split exists for fp16 operands only");
    const int Dp = f16 ? (split ? 3 : 1) * tc_feature_pitch_f16(D) : tc_feature_pitch(D);
    const int esz = f16 ? 2 : 4, k_elems = 128 / esz;   // a k-block is one 128-byte swizzle row
    const CUtensorMapDataType dt = f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    // RAFTCORR_BUILD_CLUSTER: 1 = one CTA per tile, 2 = CTA pair with TMA multicast of B (measured slower),
    // 3 = CTA pair issuing tcgen05.mma.cta_group::2 (each CTA ingests only half of B)
    const char* ce = std::getenv("RAFTCORR_BUILD_CLUSTER");
    const int mode = (ce && ce[0] == '2') ? 1 : (ce && ce[0] == '3') ? 2 : kDefaultBuildMode;
    const int cl = mode == 0 ? 1 : 2;
    RC_REQUIRE(!f16 || mode == 0, "f16 build: the CTA-pair experiments take TF32 operands only");

    CUtensorMap map_a, map_b, map_out;
    {
        cuuint64_t dims[3] = {(cuuint64_t)Dp, (cuuint64_t)N1, (cuuint64_t)B};
        cuuint64_t strides[2] = {(cuuint64_t)Dp * esz, (cuuint64_t)N1 * Dp * esz};
        cuuint32_t box[3] = {(cuuint32_t)k_elems, kTileM, 1};
        if (int rc = encode_map(fn, &map_a, dt, (void*)f1n, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, "fmap1"))
            return rc;
    }
    {
        cuuint64_t dims[4] = {(cuuint64_t)Dp, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
        cuuint64_t strides[3] = {(cuuint64_t)Dp * esz, (cuuint64_t)W * Dp * esz, (cuuint64_t)N1 * Dp * esz};
        cuuint32_t box[4] = {(cuuint32_t)k_elems, kTileW, (cuuint32_t)(kTileH / cl), 1};
        if (int rc = encode_map(fn, &map_b, dt, (void*)f2n, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, "fmap2"))
            return rc;
    }
    const int il = lay.interleaved;
    RC_REQUIRE(!il || mode != 1, "tf32 build: the multicast-pair experiment writes the row-major layout only");
    RC_REQUIRE(il != 2 || mode == 0, "tf32 build: the CTA-pair experiments write the row-major / row-pair layouts only");
    if (il) {  // level 0 as [rg * W interleaved floats][row groups][N1][B], rg = 2 (pairs) or 4 (quads)
        const int rg = layout_group_rows(level_interleave(il, 0));
        const int hp = (lay.h[0] + rg - 1) / rg;
        const int64_t img = (int64_t)hp * rg * lay.pitch[0];
        cuuint64_t dims[4] = {(cuuint64_t)lay.w[0] * rg, (cuuint64_t)hp, (cuuint64_t)N1, (cuuint64_t)B};
        cuuint64_t strides[3] = {(cuuint64_t)lay.pitch[0] * 4 * rg, (cuuint64_t)img * 4, (cuuint64_t)N1 * img * 4};
        cuuint32_t box[4] = {kTileW, 1, kTileM, 1};
        if (int rc = encode_f32_map(fn, &map_out, (void*)(pyramid + lay.offset[0]), 4, dims, strides, box,
                                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, "volume (interleaved rows)"))
            return rc;
    } else {
        const int64_t img = (int64_t)lay.h[0] * lay.pitch[0];
        cuuint64_t dims[4] = {(cuuint64_t)lay.w[0], (cuuint64_t)lay.h[0], (cuuint64_t)N1, (cuuint64_t)B};
        cuuint64_t strides[3] = {(cuuint64_t)lay.pitch[0] * 4, (cuuint64_t)img * 4, (cuuint64_t)N1 * img * 4};
        cuuint32_t box[4] = {kTileW, 1, kTileM, 1};
        if (int rc = encode_f32_map(fn, &map_out, (void*)(pyramid + lay.offset[0]), 4, dims, strides, box,
                                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, "volume"))
            return rc;
    }

    BuildParams P{};
    for (int l = 0; l < fusedL; ++l) {
        P.lvl_base[l] = pyramid + lay.offset[l];
        P.lvl_h[l] = lay.h[l];
        P.lvl_w[l] = lay.w[l];
        P.lvl_pitch[l] = lay.pitch[l];
        const int rg = layout_group_rows(level_interleave(il, l));
        P.lvl_img_stride[l] = (int64_t)((lay.h[l] + rg - 1) / rg * rg) * lay.pitch[l];
    }
    P.B = B; P.N1 = N1; P.H2 = H; P.W2 = W; P.L = fusedL;
    P.m_tiles = ceil_div(N1, kTileM);
    P.m_groups = ceil_div(P.m_tiles, cl);
    P.bands = ceil_div(H, kTileH);
    P.w_tiles = ceil_div(W, kTileW);
    P.k_blocks = Dp / k_elems;
    P.f16 = f16 ? 1 : 0; P.k_elems = k_elems; P.inv_a = inv_a; P.inv_b = inv_b;
    const int64_t total = (int64_t)B * P.m_groups * P.bands * P.w_tiles;  // tiles per cluster rank
    RC_REQUIRE(total < (1LL << 31), "tf32 build: too many tiles");
    P.total_tiles = (int)total;
    P.unit_bands = il == 2 ? 4 : 1;
    P.band_groups = ceil_div(P.bands, P.unit_bands);
    P.total_units = B * P.m_groups * P.band_groups * P.w_tiles;
    P.scale = 1.0f / sqrtf((float)D);  // corr.py:116
    { const char* e = std::getenv("RAFTCORR_BUILD_STORE"); P.lsu_store = (e && e[0] == 'l') ? 1 : 0; }
    // pyramid stores (TMA and 256-bit) carry an L2 evict-first policy: the 2.1 GB volume streams through the L2 while the
    // 58 MB of operands are re-read from it by every tile -- measured 486 -> 450 us per build group (B = 8, 55 x 128,
    // D = 256, same box, alternating); RAFTCORR_BUILD_STORE_HINT=0 switches the policy off (A/B)
    { const char* e = std::getenv("RAFTCORR_BUILD_STORE_HINT"); P.store_hint = (e && e[0] == '0') ? 0 : 1; }

    int sms = 0;
    RC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    const int max_clusters = sms / cl;
    const int clusters = P.total_units < max_clusters ? P.total_units : max_clusters;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(clusters * cl);
    cfg.blockDim = dim3(kThreads);
    constexpr int kSmemBytes = smem_bytes(3, 2, 2);
    cfg.dynamicSmemBytes = kSmemBytes;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cl;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (mode == 2) {
        // 32 KB stages: RAFTCORR_BUILD_PAIR_STAGES = 3..5 operand stages (default 5)
        const char* se = std::getenv("RAFTCORR_BUILD_PAIR_STAGES");
        const int ns = se ? atoi(se) : 5;
        constexpr int kPairStage = kABytes + kBBytes / 2;
#define RAFTCORR_PAIR_CASE(NS)                                                                                      \
    case NS: {                                                                                                      \
        constexpr int smem2 = NS * kPairStage + 2 * 2 * kRowTileBytes + 1024 + 256;                                 \
        static_assert(smem2 <= 232448, "pair pipeline does not fit shared memory");                                 \
        cfg.dynamicSmemBytes = smem2;                                                                               \
        if (il) {                                                                                                   \
            RC_CUDA(cudaFuncSetAttribute(build_tc_kernel<2, NS, 2, 2, 1>,                                           \
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem2));                      \
            RC_CUDA(cudaLaunchKernelEx(&cfg, build_tc_kernel<2, NS, 2, 2, 1>, map_a, map_b, map_out, P));           \
        } else {                                                                                                    \
            RC_CUDA(cudaFuncSetAttribute(build_tc_kernel<2, NS, 2, 2, 0>,                                           \
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem2));                      \
            RC_CUDA(cudaLaunchKernelEx(&cfg, build_tc_kernel<2, NS, 2, 2, 0>, map_a, map_b, map_out, P));           \
        }                                                                                                           \
        break;                                                                                                      \
    }
        switch (ns) {
            RAFTCORR_PAIR_CASE(3)
            RAFTCORR_PAIR_CASE(4)
            RAFTCORR_PAIR_CASE(5)
            default: return fail("RAFTCORR_BUILD_PAIR_STAGES=%d not compiled (3..5)", ns);
        }
#undef RAFTCORR_PAIR_CASE
    } else if (mode == 1) {
        RC_CUDA(cudaFuncSetAttribute(build_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
        RC_CUDA(cudaLaunchKernelEx(&cfg, build_tc_kernel<1>, map_a, map_b, map_out, P));
    } else {
        // RAFTCORR_BUILD_PIPE = <operand stages><store buffers><rows per buffer>.  Measured on B200 (B=8, 55x128,
        // L=4): at D=256 three stages are needed (2: 646 us, 3: 540, 4: 547) and deeper store buffering changes
        // nothing; at D<=128 the MMA side has slack and operand loads running far ahead only steal bandwidth
        // from the store pipeline, so two stages are faster (420 vs 462 us).
        const char* pe = std::getenv("RAFTCORR_BUILD_PIPE");
        // fp16 operands at D = 256 (4 k-blocks of 64): 322 measured 472 us vs 478 us for 222 and 242
        const int pipe = pe ? atoi(pe) : (f16 ? (P.k_blocks >= 4 ? 322 : 222) : (P.k_blocks <= 4 ? 222 : 322));
#define RAFTCORR_PIPE_CASE(NS, NB, RP, IL_)                                                                             \
    case NS * 100 + NB * 10 + RP: {                                                                                     \
        static_assert(smem_bytes(NS, NB, RP) <= 232448, "pipeline does not fit shared memory");                         \
        RC_CUDA(cudaFuncSetAttribute(build_tc_kernel<0, NS, NB, RP, IL_>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                     smem_bytes(NS, NB, RP)));                                                          \
        build_tc_kernel<0, NS, NB, RP, IL_><<<clusters, kThreads, smem_bytes(NS, NB, RP), s>>>(map_a, map_b, map_out, P); \
        break;                                                                                                          \
    }
        if (il == 2) {
            switch (pipe) {
                RAFTCORR_PIPE_CASE(3, 2, 2, 2)
                RAFTCORR_PIPE_CASE(2, 2, 2, 2)
                RAFTCORR_PIPE_CASE(2, 4, 2, 2)
                default: return fail("RAFTCORR_BUILD_PIPE=%d is not a compiled pipeline shape for the row-quad layout", pipe);
            }
        } else if (il) {
            switch (pipe) {
                RAFTCORR_PIPE_CASE(3, 2, 2, 1)
                RAFTCORR_PIPE_CASE(2, 2, 2, 1)
                RAFTCORR_PIPE_CASE(2, 4, 2, 1)
                default: return fail("RAFTCORR_BUILD_PIPE=%d is not a compiled pipeline shape for the interleaved layout", pipe);
            }
        } else {
            switch (pipe) {
                RAFTCORR_PIPE_CASE(3, 2, 2, 0)
                RAFTCORR_PIPE_CASE(2, 2, 2, 0)
                RAFTCORR_PIPE_CASE(2, 4, 2, 0)
                RAFTCORR_PIPE_CASE(3, 4, 1, 0)
                RAFTCORR_PIPE_CASE(4, 2, 1, 0)
                default: return fail("RAFTCORR_BUILD_PIPE=%d is not a compiled pipeline shape", pipe);
            }
        }
#undef RAFTCORR_PIPE_CASE
    }
    RC_CUDA(cudaGetLastError());
    if (lay.L > fusedL) {  // never used by RAFT (num_levels = 4); kept correct rather than rejected
        PyramidView pv = make_view(pyramid, lay);
        for (int l = fusedL; l < lay.L; ++l)
            if (int rc = launch_pool_level(pv.lvl[l - 1], pv.lvl[l], (int64_t)B * N1, s)) return rc;
    }
    return 0;
}

}  // namespace raftcorr

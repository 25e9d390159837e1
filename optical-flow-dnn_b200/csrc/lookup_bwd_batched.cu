// Adjoint of ALL lookups of one backward pass in one launch (autograd of raft/core/corr.py:45-91 for the 12-32
// refinement iterations of raft/core/raft.py:141-144, followed by the adjoint of the avg_pool2d cascade, corr.py:41-43).
//
// The per-iteration adjoint (lookup.cu: lookup_bwd_kernel) issues one red.global.add.f32 per lattice element into a
// zero-filled gradient pyramid -- 100 reductions per (pixel, level) window, bound by the L2's reduction throughput
// (38.7 us per launch at config 3, 12 launches), and the pooled levels are folded into level 0 afterwards by three
// streaming passes over the whole gradient pyramid (154 us, 930 MB) before the build backward's GEMMs can read it.
// But the gradient image of ONE source pixel (all levels: 15 KB at 46x62, 37 KB at 55x128) fits in shared memory, and
// every window of every iteration of that pixel lands in it.  So here a warp owns (source pixel, level): it scatters
// the K*K gradients of each iteration into the pixel's level image in SHARED memory (plain read-modify-write: the lanes
// of a phase hit distinct elements), the CTA then folds the pooled levels into level 0 on chip
// (dV0[y][x] += dV_l[y >> l][x >> l] / 4^l for floor-mode 2x2 means), rounds to TF32 when the tensor-core GEMMs follow
// and writes the level-0 image ONCE with full-line stores.  No atomics, no zero-fill of the gradient pyramid, no fold
// passes: DRAM traffic = one read of the iterations' grad_out tensors + one write of dV0.
#include <cstdlib>
#include <type_traits>

#include "lookup_common.cuh"

namespace raftcorr {

namespace {

constexpr int kMaxBatch = RAFTCORR_LOOKUP_BWD_MAX_BATCH;

struct BatchPtrs {
    const float* gout[kMaxBatch];
    const float* coords[kMaxBatch];
};

struct BatchGeo {
    int h[4], w[4], sp[4], soff[4];  // level size, shared-memory row pitch, offset (floats) of element (0, 0) inside a
                                     // pixel's image area
    int gb[4];                       // guard cells around the level image (0, or K: every window that touches the level
                                     // then lies inside the guarded array and needs no bounds predicates)
    float inv_scale[4];
    int img_floats;                  // shared-memory floats per source pixel (all levels), multiple of 4
    int L, PX;                       // levels (<= 4), source pixels per CTA step
    int B, N, ctot;
    int groups_per_img;              // ceil(N / PX)
    int64_t total_groups;
    int P0;                          // global row pitch of level 0 (floats)
    int64_t img_stride0;             // floats between consecutive source pixels' level-0 images
    int count, accumulate, round_out;
    int es;                          // bytes per cp.async chunk of the gradient staging (16 / 8 / 4)
    int img_total;                   // PX * img_floats: where the staging ring starts in shared memory (floats)
};

constexpr int kStages = 3;  // ring of staged (grad_out, coords) tiles: two iterations in flight behind the one being scattered

template <int ES>
__device__ __forceinline__ void cp_async_chunk(uint32_t dst, const void* src, bool valid) {
    const int sz = valid ? ES : 0;  // src-size 0: the chunk is zero-filled, src is not read
    if constexpr (ES == 16)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
    else if constexpr (ES == 8)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
    else
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ float round_tf32_b(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// PXT: pixels per CTA step as a compile-time constant (4: the config-3 / RAFT-small shapes), or 0 = geo.PX.  With PX known
// the offsets inside a staged tile are immediates: the ncu source view of the runtime-PX kernel charged 24 instructions
// per window to rematerialised offset arithmetic.
template <int R, int ES, int PXT = 0>
__global__ void __launch_bounds__(1024, 1)
lookup_bwd_batched_kernel(const BatchPtrs ptrs, const BatchGeo geo, float* __restrict__ dv0) {
    constexpr int K = 2 * R + 1, KK = K * K;
    constexpr int EF = ES / 4;  // floats per staging chunk
    static_assert(PXT == 0 || (PXT * 4) % ES == 0, "a row of PX pixels is a whole number of staging chunks");
    extern __shared__ float4 smem4[];
    float* smem = reinterpret_cast<float*>(smem4);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nwarps = blockDim.x >> 5;
    const int L = geo.L, N = geo.N, PX = PXT ? PXT : geo.PX;
    const int q = warp / L, l = warp - q * L;  // this warp's source pixel of the CTA step and its pyramid level
    const int lh = geo.h[l], lw = geo.w[l], lsp = geo.sp[l];
    const float inv = geo.inv_scale[l];
    float* simg = smem + q * geo.img_floats + geo.soff[l];
    // staging ring: kStages tiles of [ctot + 2][PX] floats -- the K*K*L gradient channels and the two coordinate
    // channels of the CTA's PX consecutive pixels for one iteration, fetched with cp.async in chunks of PX (or fewer)
    // consecutive pixels.  (One global load per lane and channel instead -- 32 different sectors per warp load -- made
    // the first version of this kernel L1-tag bound: 791 us at config 3 for 780 MB of DRAM traffic.)
    const int rows = geo.ctot + 2, tile_floats = rows * PX;
    float* stage = smem + geo.img_total;
    const uint32_t stage_u32 = smem_u32(stage);
    const int cpr = PX / EF;  // chunks per channel row
    const bool cpr1 = cpr == 1;

    // Lane -> (part, lattice column c) of the warp's window: the (K+1) x (K+1) lattice footprint of a window is split
    // into NP groups of consecutive lattice rows, K+1 lanes per group (30 of 32 lanes at r = 4: rows 0-3 | 4-6 | 7-9).
    // The bilinear adjoint is separable: a lane loads its column of the K x K gradient window for the rows it needs
    // (channel of (x offset i, y offset j) = i * K + j, corr.py:66-78), takes the column to its left from the
    // neighbouring lane (shuffle), forms h[j] = (1 - fx) g[j][c] + fx g[j][c - 1] and then each lattice row
    // v[j'] = (1 - fy) h[j'] + fy h[j' - 1] -- ONE shared-memory read-modify-write per lattice point of the footprint
    // (distinct lanes own distinct points: no phases, no __syncwarp), 100 per window instead of 4 x 81.  (The first
    // version had lane = window element and four read-modify-write phases, one per bilinear tap: ~150 instructions per
    // window, the kernel was bound by its instruction stream.)
    constexpr int KP = K + 1;
    constexpr int NP = (32 / KP) < KP ? (32 / KP) : KP;              // lane groups
    constexpr int RB = KP / NP, RREM = KP % NP, MAXR = RB + (RREM ? 1 : 0);  // lattice rows per group
    const int part = lane / KP, c = lane - part * KP;
    const bool lane_on = part < NP;
    const int r_cnt = lane_on ? RB + (part < RREM ? 1 : 0) : 0;       // this lane's lattice rows: r_start .. r_start + r_cnt - 1
    const int r_start = part * RB + (part < RREM ? part : RREM);
    // gradient rows j = r_start - 1 + k, k = 0 .. r_cnt (the rows above and inside the lane's lattice rows); rows outside
    // 0 .. K-1 and the lattice column c = K (which has no gradient column of its own) read as zero
    bool gon[MAXR + 1];
#pragma unroll
    for (int k = 0; k <= MAXR; ++k) {
        const int j = r_start - 1 + k;
        gon[k] = lane_on && k <= r_cnt && j >= 0 && j < K && c < K;
    }
    // element offset inside a staged tile of gradient row r_start - 1 of this lane's column (lanes without a column of
    // their own read column 0; rows outside 0 .. K-1 are the neighbouring channels' values: loaded, then discarded)
    const int gbase = (l * KK + ((lane_on && c < K) ? c * K : 0) + r_start - 1) * PX + q;
    const int lat_off = r_start * lsp + c;  // the lane's first lattice point relative to the window origin
    // staging: the tile rows this thread copies (fixed for the whole launch) and their element offsets in the source
    constexpr int kRowsPerThread = 3;  // (K*K*L + 2) rows over PX*L*32 threads, PX >= 1
    int my_row[kRowsPerThread], my_src[kRowsPerThread], my_sel[kRowsPerThread];
#pragma unroll
    for (int r = 0; r < kRowsPerThread; ++r) {
        const int row = threadIdx.x + r * blockDim.x;
        my_row[r] = row < rows ? row : -1;
        my_src[r] = (row < geo.ctot ? row : row - geo.ctot) * N;  // ctot * N < 2^31 (checked by the launcher)
        my_sel[r] = row < geo.ctot ? 0 : kMaxBatch;               // BatchPtrs = gout[kMaxBatch] | coords[kMaxBatch]
    }
    static_assert(sizeof(BatchPtrs) == 2 * kMaxBatch * sizeof(const float*), "BatchPtrs is indexed as one pointer array");
    const float* const* psrc = reinterpret_cast<const float* const*>(&ptrs);
    const bool guarded = geo.gb[l] != 0;
    const int cx_off = geo.ctot * PX + q;

    // zero this CTA's images once; every step leaves them zeroed again
    for (int i = threadIdx.x; i < geo.PX * geo.img_floats / 4; i += blockDim.x) smem4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();

    for (int64_t grp = blockIdx.x; grp < geo.total_groups; grp += gridDim.x) {
        const int b = (int)(grp / geo.groups_per_img);
        const int p0 = (int)(grp - (int64_t)b * geo.groups_per_img) * geo.PX;
        const int p = p0 + q;
        // the first ncu capture of this kernel (profiles/r02_ncu_full_bw3_lookup_bwd_batched_first.json) showed it
        // bound by its own instruction stream (702 M warp instructions, issue slots 77 % busy, DRAM 12 %): 280
        // instructions per window (a branch around every predicated read-modify-write), 120 per written row, a
        // division per staged chunk.  Hence: a predicate-free path for windows inside the level, a branch-free one
        // (out-of-level taps redirected to a scratch word) otherwise, division-free loops everywhere.
        // source of a staged row: (grad_out or coords of the iteration) + element offset of (batch item, channel row, first
        // pixel) -- 32-bit (B * ctot * N < 2^31, checked by the launcher) and fixed per (thread, group)
        const int g_row0 = b * (geo.ctot * N) + p0, c_row0 = b * (2 * N) + p0;
        int my_off[kRowsPerThread];
#pragma unroll
        for (int r = 0; r < kRowsPerThread; ++r) my_off[r] = (my_sel[r] ? c_row0 : g_row0) + my_src[r];
        auto issue = [&](int it, int st) {  // all threads: stage iteration `it` of this group into ring slot `st`
            const uint32_t tile_u32 = stage_u32 + (uint32_t)(st * tile_floats) * 4u;
#pragma unroll
            for (int r = 0; r < kRowsPerThread; ++r) {
                if (my_row[r] < 0) break;
                const float* src = psrc[it + my_sel[r]] + my_off[r];
                const uint32_t dst = tile_u32 + (uint32_t)(my_row[r] * PX) * 4u;
                if (ES == 16 && (PXT == 4 || cpr1)) {
                    cp_async_chunk<ES>(dst, src, true);  // PX == 4 and p0 < N: the one chunk of the row is inside
                } else {
                    for (int k = 0; k < cpr; ++k) {
                        const bool valid = p0 + k * EF < N;  // N is a multiple of EF: a chunk is entirely inside or outside
                        cp_async_chunk<ES>(dst + (uint32_t)(k * EF) * 4u, src + (valid ? k * EF : 0), valid);
                    }
                }
            }
        };
#pragma unroll
        for (int i = 0; i < kStages - 1; ++i) {
            if (i < geo.count) issue(i, i);
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
        int cur_stage = kStages - 1, nxt_stage = kStages - 2;  // ring slots of iteration it / it + kStages - 1 (advanced at the loop top)
        for (int it = 0; it < geo.count; ++it) {
            cur_stage = cur_stage == kStages - 1 ? 0 : cur_stage + 1;
            nxt_stage = nxt_stage == kStages - 1 ? 0 : nxt_stage + 1;
            asm volatile("cp.async.wait_group %0;" ::"n"(kStages - 2) : "memory");
            __syncthreads();  // iteration `it` is staged for everybody; everybody is done with the tile of it - 1
            if (it + kStages - 1 < geo.count) issue(it + kStages - 1, nxt_stage);
            asm volatile("cp.async.commit_group;" ::: "memory");
            if (p >= N) continue;
            const float* tile = stage + cur_stage * tile_floats;
            const float* tg = tile + gbase;
            float g[MAXR + 1];
#pragma unroll
            for (int k = 0; k <= MAXR; ++k) g[k] = tg[k * PX];
            const float cx = tile[cx_off], cy = tile[cx_off + PX];
            const float x = cx * inv, y = cy * inv;
            const float xf = floorf(x), yf = floorf(y);
            const float fx = x - xf, fy = y - yf;
            const int x0 = (int)clamp_coord(xf) - R, y0 = (int)clamp_coord(yf) - R;
            // warp-uniform: windows entirely outside the level contribute nothing (zero padding)
            if (x0 + K < 0 || x0 >= lw || y0 + K < 0 || y0 >= lh) continue;
            const float gx1 = 1.f - fx, gy1 = 1.f - fy;
            float h[MAXR + 1];
#pragma unroll
            for (int k = 0; k <= MAXR; ++k) {
                const float gc = gon[k] ? g[k] : 0.f;
                float gl = __shfl_up_sync(0xffffffffu, gc, 1);  // column c - 1 of the same rows (same lane group for c >= 1)
                gl = c > 0 ? gl : 0.f;
                h[k] = gx1 * gc + fx * gl;
            }
            float* wp = simg + (y0 * lsp + x0 + lat_off);
            if (guarded || (x0 >= 0 && y0 >= 0 && x0 + K < lw && y0 + K < lh)) {
                // the whole (K+1) x (K+1) lattice is inside the level (or its guard band): no bounds predicates
#pragma unroll
                for (int k = 0; k < MAXR; ++k)
                    if (k < r_cnt) wp[k * lsp] += gy1 * h[k + 1] + fy * h[k];
            } else {
                const bool col_ok = (unsigned)(x0 + c) < (unsigned)lw;
#pragma unroll
                for (int k = 0; k < MAXR; ++k)
                    if (k < r_cnt && col_ok && (unsigned)(y0 + r_start + k) < (unsigned)lh) wp[k * lsp] += gy1 * h[k + 1] + fy * h[k];
            }
            __syncwarp();  // the next iteration's window overlaps this one: its lanes own different points of it
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        const int npx = min(PX, N - p0);
        // ---- fold the pooled levels down, coarse to fine (adjoint of F.avg_pool2d(2, stride=2), corr.py:42), on chip:
        // one warp per (pixel, fine row), lanes along the row
        for (int lc = L - 1; lc >= 2; --lc) {
            const int fh = geo.h[lc - 1], fw = geo.w[lc - 1], fsp = geo.sp[lc - 1];
            const int ch = geo.h[lc], cw = geo.w[lc], csp = geo.sp[lc];
            const int fh_in = min(fh, 2 * ch), fw_in = min(fw, 2 * cw);  // rows / columns that have a parent
            for (int qq = 0; qq < npx; ++qq) {
                float* fbase = smem + qq * geo.img_floats + geo.soff[lc - 1];
                const float* cbase = smem + qq * geo.img_floats + geo.soff[lc];
                for (int yy = warp; yy < fh_in; yy += nwarps) {
                    float* frow = fbase + yy * fsp;
                    const float* crow = cbase + (yy >> 1) * csp;
                    for (int xx = lane; xx < fw_in; xx += 32) frow[xx] += 0.25f * crow[xx >> 1];
                }
            }
            __syncthreads();
        }
        // ---- level 0 (+ 1/4 of the folded level 1) leaves as full rows: lane = column (fixed per lane, so the column
        // predicates are formed once), warps walk the rows of the CTA's pixels
        {
            const int h0 = geo.h[0], w0 = geo.w[0], sp0 = geo.sp[0], P0 = geo.P0;
            const int h1 = L > 1 ? geo.h[1] : 1, w1 = L > 1 ? geo.w[1] : 0, sp1 = L > 1 ? geo.sp[1] : 0;
            const int so1 = L > 1 ? geo.soff[1] : geo.soff[0];
            const bool acc = geo.accumulate != 0, rnd = geo.round_out != 0;
            for (int xx = lane; xx < P0; xx += 32) {
                const bool in0 = xx < w0;
                const int x1 = min(xx >> 1, max(w1 - 1, 0));
                const float kx = (xx >> 1) < w1 ? 0.25f : 0.f;  // xx >= w0 implies xx / 2 >= w1
                for (int qq = 0; qq < npx; ++qq) {
                    const float* b0 = smem + qq * geo.img_floats + geo.soff[0] + xx;
                    const float* b1 = smem + qq * geo.img_floats + so1 + x1;
                    float* d = dv0 + ((int64_t)b * N + p0 + qq) * geo.img_stride0 + xx;
                    const float* r0 = b0 + warp * sp0;
                    float* dd = d + warp * P0;
                    const int s0step = nwarps * sp0, dstep = nwarps * P0;
                    if ((nwarps & 1) == 0) {
                        // an even warp count keeps a warp on rows of one parity: the level-1 parent row advances by
                        // nwarps / 2 per step (a running pointer instead of shift + select + multiply), and only the rows
                        // below 2 * h1 have a parent at all (ncu source view of the generic loop: 24 instructions per
                        // 32 floats written, 15 % of the kernel)
                        const int hpar = min(h0, 2 * h1);
                        const float* p1 = b1 + (warp >> 1) * sp1;
                        const int s1step = (nwarps >> 1) * sp1;
                        int yy = warp;
                        for (; yy < hpar; yy += nwarps, r0 += s0step, p1 += s1step, dd += dstep) {
                            float v = fmaf(kx, *p1, in0 ? *r0 : 0.f);
                            if (acc) v += *dd;
                            if (rnd) v = round_tf32_b(v);
                            *dd = v;
                        }
                        for (; yy < h0; yy += nwarps, r0 += s0step, dd += dstep) {  // h0 odd: the last row has no parent
                            float v = in0 ? *r0 : 0.f;
                            if (acc) v += *dd;
                            if (rnd) v = round_tf32_b(v);
                            *dd = v;
                        }
                        continue;
                    }
                    for (int yy = warp; yy < h0; yy += nwarps, r0 += s0step, dd += dstep) {
                        const int y1 = yy >> 1;
                        const bool par = y1 < h1;  // rows without a parent read row 0 with weight 0
                        float v = in0 ? *r0 : 0.f;
                        v = fmaf(par ? kx : 0.f, b1[(par ? y1 : 0) * sp1], v);
                        if (acc) v += *dd;
                        if (rnd) v = round_tf32_b(v);
                        *dd = v;
                    }
                }
            }
        }
        __syncthreads();
        for (int i = threadIdx.x; i < PX * geo.img_floats / 4; i += blockDim.x) smem4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        __syncthreads();
    }
}

// shared-memory pitch of a level image: >= w and == s (mod 32), s chosen so that the lane groups of a window (lattice
// rows r_start(part) + k, K+1 consecutive words each; see the lane mapping in the kernel) land on disjoint banks in
// every step k: r = 4: rows 0 | 4 | 7 -> 0, 12, 21; r = 3: rows 0 | 2 | 4 | 6 -> 0, 8, 16, 24; r = 2: 0, 12, 18, 24, 30
inline int smem_pitch(int w, int K) {
    const int want = K == 9 ? 3 : K == 7 ? 4 : K == 5 ? 6 : 4;
    int sp = w;
    while (((sp - want) & 31) != 0) ++sp;
    return sp;
}

// Shared-memory layout of one source pixel's gradient images; `guards`: levels >= 2 get K guard cells on every side.
void layout_images(BatchGeo& g, const PyramidView& dpv, int K, bool guards) {
    int off = 0;
    for (int l = 0; l < dpv.L; ++l) {
        g.h[l] = dpv.lvl[l].h;
        g.w[l] = dpv.lvl[l].w;
        g.gb[l] = (guards && l >= 2) ? K : 0;
        g.sp[l] = smem_pitch(g.w[l] + 2 * g.gb[l], K);
        g.soff[l] = off + g.gb[l] * (g.sp[l] + 1);
        g.inv_scale[l] = 1.0f / (float)(1 << l);
        off += (g.h[l] + 2 * g.gb[l]) * g.sp[l];
    }
    off += 4;  // slack behind the images
    g.img_floats = (off + 3) & ~3;
}

int pixels_per_cta(const BatchGeo& g, int L, int* per_sm) {
    // two resident CTAs per SM when a power-of-two pixel count allows it (one CTA's write-out overlaps the other's
    // scatter); at most 1024 threads per CTA.  Per pixel: its gradient images + its share of the staging ring.
    const size_t px_bytes = (size_t)g.img_floats * 4 + (size_t)kStages * (g.ctot + 2) * 4;
    const size_t budget2 = 112 * 1024, budget1 = 224 * 1024;  // 2 x (112 + 1 reserved) KB <= 227 KB
    int px = 8;
    while (px > 1 && (px * px_bytes > budget2 || px * L * 32 > 1024)) px >>= 1;
    *per_sm = 2;
    if (px * px_bytes > budget2) {  // a single pixel does not leave room for a second CTA: one CTA per SM
        // at least two pixels (2 L warps) per SM, or the write-out of a 100+ KB image by a handful of warps starves
        px = 2;
        *per_sm = 1;
        if (px * px_bytes > budget1) return 0;
        while (2 * px * px_bytes <= budget1 && 2 * px * L * 32 <= 1024 && px < 8) px <<= 1;
    }
    return px;
}

int make_geo(const PyramidView& dpv, int radius, BatchGeo* out, size_t* smem_bytes, int* threads, int* per_sm) {
    BatchGeo g{};
    const int K = 2 * radius + 1;
    if (dpv.L < 1 || dpv.L > 4) return 1;
    g.L = dpv.L;
    g.B = dpv.B;
    g.N = dpv.N;
    g.ctot = dpv.L * K * K;
    g.P0 = dpv.lvl[0].pitch;
    g.img_stride0 = dpv.lvl[0].img_stride;
    layout_images(g, dpv, K, false);
    int px = pixels_per_cta(g, dpv.L, per_sm);
    if (px == 0) return 1;
    if (const char* e = std::getenv("RAFTCORR_LOOKUP_BWD_PX")) {  // A/B switch: pixels per CTA (one CTA per SM if needed)
        const int want = std::atoi(e);
        const size_t px_bytes = (size_t)g.img_floats * 4 + (size_t)kStages * (g.ctot + 2) * 4;
        if ((want == 1 || want == 2 || want == 4 || want == 8) && want * dpv.L * 32 <= 1024 && want * px_bytes <= 224 * 1024) {
            px = want;
            *per_sm = want * px_bytes <= 112 * 1024 ? 2 : 1;
            g.PX = px;
            g.img_total = px * g.img_floats;
            g.es = (px % 4 == 0 && g.N % 4 == 0) ? 16 : (px % 2 == 0 && g.N % 2 == 0) ? 8 : 4;
            g.groups_per_img = ceil_div(g.N, px);
            g.total_groups = (int64_t)g.groups_per_img * g.B;
            *out = g;
            *smem_bytes = (size_t)px * px_bytes;
            *threads = px * dpv.L * 32;
            return 0;
        }
    }
    if (dpv.L > 2) {
        // the coarse levels are smaller than a window, so every window there hangs over the border: guard bands (a few
        // KB) put all of them on the predicate-free path -- taken when they do not cost pixels per CTA
        BatchGeo gg = g;
        int per_sm_g = 0;
        layout_images(gg, dpv, K, true);
        if (pixels_per_cta(gg, dpv.L, &per_sm_g) == px && per_sm_g == *per_sm) g = gg;
    }
    g.PX = px;
    g.img_total = px * g.img_floats;
    // staging chunks: the largest of 16 / 8 / 4 bytes that divides both a row of PX pixels and the channel stride N
    g.es = (px % 4 == 0 && g.N % 4 == 0) ? 16 : (px % 2 == 0 && g.N % 2 == 0) ? 8 : 4;
    g.groups_per_img = ceil_div(g.N, px);
    g.total_groups = (int64_t)g.groups_per_img * g.B;
    *out = g;
    *smem_bytes = (size_t)px * ((size_t)g.img_floats * 4 + (size_t)kStages * (g.ctot + 2) * 4);
    *threads = px * dpv.L * 32;
    return 0;
}

}  // namespace

bool lookup_bwd_batched_supported(const PyramidView& dpv, int radius) {
    if (radius < 1 || radius > 4) return false;
    if ((int64_t)dpv.B * dpv.L * 81 * dpv.N >= (1LL << 31)) return false;  // 32-bit element offsets into grad_out
    BatchGeo g;
    size_t smem;
    int threads, per_sm;
    return make_geo(dpv, radius, &g, &smem, &threads, &per_sm) == 0;
}

int launch_lookup_bwd_batched(const float* const* gouts, const float* const* coords, int count, const PyramidView& dpv,
                              int radius, bool accumulate, bool round_out, int device, cudaStream_t s) {
    RC_REQUIRE(count >= 1 && count <= kMaxBatch, "lookup_bwd_batched: count %d outside 1..%d", count, kMaxBatch);
    RC_REQUIRE((int64_t)dpv.B * dpv.L * 81 * dpv.N < (1LL << 31), "lookup_bwd_batched: grad_out tensors of 2^31 elements or more");
    BatchGeo g;
    size_t smem;
    int threads, per_sm;
    RC_REQUIRE(make_geo(dpv, radius, &g, &smem, &threads, &per_sm) == 0,
               "lookup_bwd_batched: a source pixel's gradient pyramid does not fit in shared memory (or L > 4)");
    g.count = count;
    g.accumulate = accumulate ? 1 : 0;
    g.round_out = round_out ? 1 : 0;
    BatchPtrs ptrs{};
    for (int i = 0; i < count; ++i) {
        RC_REQUIRE(gouts[i] && coords[i], "lookup_bwd_batched: NULL tensor in the batch");
        ptrs.gout[i] = gouts[i];
        ptrs.coords[i] = coords[i];
    }
    int sms = 0;
    RC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    const int64_t ctas = (int64_t)sms * per_sm;
    const unsigned grid = (unsigned)(g.total_groups < ctas ? g.total_groups : ctas);
    auto launch = [&](auto kern) -> int {
        RC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, threads, smem, s>>>(ptrs, g, dpv.lvl[0].base);
        RC_CUDA(cudaGetLastError());
        return 0;
    };
    auto by_chunk = [&](auto r_tag) -> int {
        constexpr int R = decltype(r_tag)::value;
        if (g.es == 16 && g.PX == 4) return launch(lookup_bwd_batched_kernel<R, 16, 4>);
        if (g.es == 16) return launch(lookup_bwd_batched_kernel<R, 16>);
        if (g.es == 8) return launch(lookup_bwd_batched_kernel<R, 8>);
        return launch(lookup_bwd_batched_kernel<R, 4>);
    };
    switch (radius) {
        case 1: return by_chunk(std::integral_constant<int, 1>{});
        case 2: return by_chunk(std::integral_constant<int, 2>{});
        case 3: return by_chunk(std::integral_constant<int, 3>{});
        case 4: return by_chunk(std::integral_constant<int, 4>{});
        default: return fail("lookup_bwd_batched: radius %d not supported (1..4)", radius);
    }
}

}  // namespace raftcorr

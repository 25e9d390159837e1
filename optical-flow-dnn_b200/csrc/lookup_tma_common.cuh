// Geometry of the TMA lookup (patch boxes per level) and the per-pyramid plan, shared by lookup_tma.cu and
// lookup_conv.cu.
#pragma once
#include "lookup_common.cuh"
#include "tma_util.cuh"

namespace raftcorr {

// IL = row pairs interleaved (raftcorr_pyramid_layout::interleaved): the window is (K+1)/2 + 1 pair rows of
// 2 * (K+1 + 1) floats -- half as many, longer DRAM runs (the box starts at an even column: 16-byte rule again).
template <int R, bool IL>
struct TmaGeo {
    static constexpr int K = 2 * R + 1;
    static constexpr int K1 = K + 1;
    // row-major: box columns cover any 0..3 start shift, 16-byte multiple
    static constexpr int BC = IL ? 2 * ((K1 + 1 + 1) / 2 * 2) : (K1 + 3 + 3) / 4 * 4;  // floats per box row
    static constexpr int BR = IL ? K1 / 2 + 1 : K1;                                     // box rows
    static constexpr int BOX_BYTES = BC * BR * 4;
    static constexpr int SLOT = (BOX_BYTES + 127) / 128 * 128;  // TMA destinations are 128-byte aligned
    static constexpr int KK = K * K;
    // evaluation: lane = (x offset i, row group jg); a lane produces RPL vertically adjacent outputs of column i
    // from RPL+1 horizontal lerps, so every patch value is loaded once per two outputs instead of once per tap
    static constexpr int GROUPS = (32 / K) < K ? (32 / K) : K;
    static constexpr int RPL = (K + GROUPS - 1) / GROUPS;
};

// Row-QUAD layout (raftcorr_pyramid_layout::interleaved == 2): one column of a quad is 16 bytes, so a box starts on
// TMA's 16-byte grid at any x and is EXACTLY K+1 columns wide: GMAX or GMAX-1 quads of 4*(K+1) floats.
template <int R>
struct QuadGeo {
    static constexpr int K = 2 * R + 1;
    static constexpr int K1 = K + 1;
    static constexpr int KK = K * K;
    static constexpr int ROWF = 4 * K1;                     // floats per box row (one quad of the window)
    static constexpr int ROW_BYTES = ROWF * 4;
    static constexpr int GMAX = (K1 + 2) / 4 + 1;           // quads a window can span: ((3 + K1 - 1) >> 2) + 1
    static constexpr int SLOT = (GMAX * ROW_BYTES + 127) / 128 * 128;
    static constexpr int NV = K + 3;                        // quad-relative rows Y = 0 .. K+2 that can be an output row
    static_assert(NV + 1 <= 4 * GMAX, "the horizontal pass reads rows 0 .. K+3");
    // quads the window starting at row-in-quad sy spans
    __host__ __device__ static constexpr int quads(int sy) { return ((sy + K1 - 1) >> 2) + 1; }
};

struct LookupMaps {
    CUtensorMap m[4];
    CUtensorMap ms[4];  // interleaved layouts: one row group shorter, for windows whose start row lets them fit
};

// plans: tensor maps + launch geometry encoded once per pyramid, reused by every refinement iteration
struct LookupPlanData {
    uint32_t magic;
    int radius, groups, lg[2];
    GroupView g[2];
    LookupMaps maps[2];
};
static_assert(sizeof(LookupPlanData) <= RAFTCORR_LOOKUP_PLAN_BYTES, "plan blob too small");
constexpr uint32_t kPlanMagic = 0x52434c50u;  // "RCLP"


}  // namespace raftcorr

// Pieces shared by the lookup kernels (lookup.cu: cp.async forward + adjoint; lookup_tma.cu: TMA forward).
#pragma once
#include "common.cuh"

namespace raftcorr {

constexpr int kWarps = 8;
constexpr int kPix = 32;       // source pixels per CTA = one output line per channel
constexpr int kTilePitch = 33; // +1 float: conflict-free column writes / row reads

struct GroupView {  // <= 4 consecutive levels handled by one launch
    LevelView lvl[4];
    float inv_scale[4];  // 2^-level
    int B, N;
    int ch0, ctot;       // first output channel of this group, total channels
};

__device__ __forceinline__ float clamp_coord(float v) {
    // keeps float->int conversion defined for wild coordinates; everything this far out is
    // outside every level anyway (zero padding).  NaN maps to -2^20 (all taps out of range).
    return fminf(fmaxf(v, -1048576.f), 1048576.f);
}


GroupView make_group(const PyramidView& pv, int l0, int lg, int radius);
int launch_lookup_fwd_tma(const PyramidView& pv, const float* coords, int radius, float* out, cudaStream_t s);
bool use_tma_lookup();
int launch_lookup_fwd_group_cpasync(int radius, int lg, const GroupView& g, const float* coords, float* out,
                                    cudaStream_t s);
int lookup_plan_init(void* plan_blob, const PyramidView& pv, int radius);
int lookup_plan_run(const void* plan_blob, const float* coords, float* out, cudaStream_t s);

}  // namespace raftcorr

// placeholder until the tcgen05 kernel lands (next commit)
#include "common.cuh"
namespace raftcorr {
int launch_build_tc(const float*, const float*, int, float*, const raftcorr_pyramid_layout&, int, cudaStream_t) {
    return fail("build_fwd: TF32 tensor-core mode not built yet");
}
}  // namespace raftcorr

// Standalone probe: does a small 3-D tiled TMA box load with arbitrary signed coordinates work, and
// are out-of-range elements zero-filled?  usage: tma_probe W H PITCH IMGS BC BR X0 Y0 IMG
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 2; } } while (0)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void probe(const __grid_constant__ CUtensorMap map, float* out, int n, int x0, int y0, int img, int bytes) {
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ __align__(8) unsigned long long bar;
    unsigned dst = (unsigned)__cvta_generic_to_shared(sm);
    unsigned b = (unsigned)__cvta_generic_to_shared(&bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     ::"r"(dst), "l"(&map), "r"(b), "r"(x0), "r"(y0), "r"(img) : "memory");
    }
    __syncthreads();
    unsigned ok = 0; long long t0 = clock64();
    while (!ok) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(b) : "memory");
        if (clock64() - t0 > 2000000000LL) { if (threadIdx.x == 0) printf("TIMEOUT waiting for TMA\n"); return; }
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = reinterpret_cast<float*>(sm)[i];
}
int main(int argc, char** argv) {
    if (argc < 10) return 1;
    int W = atoi(argv[1]), H = atoi(argv[2]), P = atoi(argv[3]), IMGS = atoi(argv[4]), BC = atoi(argv[5]), BR = atoi(argv[6]);
    int X0 = atoi(argv[7]), Y0 = atoi(argv[8]), IMG = atoi(argv[9]);
    std::vector<float> h((size_t)IMGS * H * P);
    for (int i = 0; i < IMGS; ++i) for (int y = 0; y < H; ++y) for (int x = 0; x < P; ++x) h[((size_t)i * H + y) * P + x] = 1000.f * i + 10.f * y + 0.01f * x + 1.f;
    float *d, *o; CK(cudaMalloc(&d, h.size() * 4)); CK(cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    int n = BC * BR; CK(cudaMalloc(&o, n * 4)); CK(cudaMemset(o, 0xff, n * 4));
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    CUtensorMap map;
    cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)IMGS};
    cuuint64_t strides[2] = {(cuuint64_t)P * 4, (cuuint64_t)H * P * 4};
    cuuint32_t box[3] = {(cuuint32_t)BC, (cuuint32_t)BR, 1}, es[3] = {1, 1, 1};
    CUresult r = ((EncodeTiledFn)fn)(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 3; }
    probe<<<1, 128, n * 4 + 128>>>(map, o, n, X0, Y0, IMG, n * 4);
    CK(cudaDeviceSynchronize());
    std::vector<float> g(n); CK(cudaMemcpy(g.data(), o, n * 4, cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int r2 = 0; r2 < BR; ++r2) for (int c = 0; c < BC; ++c) {
        int y = Y0 + r2, x = X0 + c;
        float want = (y >= 0 && y < H && x >= 0 && x < W) ? h[((size_t)IMG * H + y) * P + x] : 0.f;
        if (g[r2 * BC + c] != want) { if (bad < 5) printf("  mismatch r%d c%d got %f want %f\n", r2, c, g[r2 * BC + c], want); ++bad; }
    }
    printf("W%d H%d P%d box %dx%d at (%d,%d) img %d: %s (%d bad)\n", W, H, P, BC, BR, X0, Y0, IMG, bad ? "FAIL" : "OK", bad);
    return bad ? 4 : 0;
}

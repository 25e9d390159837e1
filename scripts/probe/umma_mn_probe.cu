// Probe: which shared-memory image + descriptor makes tcgen05.mma kind::tf32 read an MN-major A tile
// (M=128, K=8) correctly?  B is K-major SW128 (known-good path).  usage: umma_mn_probe <variant>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cmath>
#include <algorithm>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); return 2; } } while (0)
constexpr int M = 128, N = 16, K = 8;
__global__ void probe(const float* a_img, const float* b_img, float* out, unsigned long long adesc_hi, unsigned idesc) {
    extern __shared__ __align__(1024) unsigned char sm[];
    __shared__ __align__(8) unsigned long long bar;
    __shared__ unsigned tslot;
    unsigned base = ((unsigned)__cvta_generic_to_shared(sm) + 1023u) & ~1023u;
    float* smf = reinterpret_cast<float*>(sm + (base - (unsigned)__cvta_generic_to_shared(sm)));
    for (int i = threadIdx.x; i < 4096 + 1024; i += blockDim.x) smf[i] = i < 4096 ? a_img[i] : b_img[i - 4096];  // A: 16 KB, B: 4 KB
    unsigned b = (unsigned)__cvta_generic_to_shared(&bar);
    if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b)); asm volatile("fence.mbarrier_init.release.cluster;"); }
    if (threadIdx.x < 32) {
        unsigned ts = (unsigned)__cvta_generic_to_shared(&tslot);
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(ts) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;"); __syncthreads(); asm volatile("tcgen05.fence::after_thread_sync;");
    unsigned tmem = tslot;
    if (threadIdx.x == 0) {
        unsigned long long ad = (unsigned long long)((base >> 4) & 0x3FFF) | adesc_hi;
        unsigned long long bd = (unsigned long long)(((base + 16384) >> 4) & 0x3FFF) | (1ull << 16) | ((unsigned long long)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 0, 1;\n\tsetp.eq.b32 p, 0, 1;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc) : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(b) : "memory");
    }
    unsigned ok = 0;
    while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(b) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;");
    if (threadIdx.x < 128) {
        unsigned r[16];
        unsigned addr = tmem + ((threadIdx.x / 32 * 32) << 16);
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]),"=r"(r[8]),"=r"(r[9]),"=r"(r[10]),"=r"(r[11]),"=r"(r[12]),"=r"(r[13]),"=r"(r[14]),"=r"(r[15]) : "r"(addr) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 16; ++j) out[threadIdx.x * 16 + j] = __uint_as_float(r[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;"); __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tmem) : "memory");
}
static float tf32(float x) { unsigned u; memcpy(&u, &x, 4); u = (u + 0x1000) & ~0x1FFFu; memcpy(&x, &u, 4); return x; }
int main(int argc, char** argv) {
    int variant = argc > 1 ? atoi(argv[1]) : 0;
    std::vector<float> A(M * K), B(N * K);
    srand(1);
    for (auto& v : A) v = tf32((rand() % 2001 - 1000) / 500.f);
    for (auto& v : B) v = tf32((rand() % 2001 - 1000) / 500.f);
    // variants: bit0: chunk stride 1KB(0)/4KB(1); bit1: swap LBO/SBO fields; bit2: image without swizzle; bit3: layout_type NONE(0)
    int chunk = (variant & 1) ? 4096 : 1024, kgrp = (variant & 1) ? 1024 : 4096;
    bool swap = variant & 2, noswz = variant & 4, lt_none = variant & 8;
    std::vector<float> a_img(4096, 0.f), b_img(1024, 0.f);
    for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) {
        int c = m / 32, mm = m % 32;                     // 32-wide MN chunk, element inside
        int byte = c * chunk + k * 128 + mm * 4;         // atom: 8 k-rows of 128 bytes
        if (!noswz) { int chunk16 = (byte >> 4) & 7, row = (byte >> 7) & 7; byte = (byte & ~0x70) | ((chunk16 ^ row) << 4); }
        a_img[byte / 4] = A[m * K + k];
    }
    for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) {  // K-major SW128: row n of 128 bytes, k at k*4
        int byte = n * 128 + k * 4; int chunk16 = (byte >> 4) & 7, row = (byte >> 7) & 7; byte = (byte & ~0x70) | ((chunk16 ^ row) << 4);
        b_img[byte / 4] = B[n * K + k];
    }
    bool kmajor = variant & 16;  // control: K-major A (row m of 128 bytes, k at k*4, 8-row groups 1024 B apart)
    if (kmajor) {
        std::fill(a_img.begin(), a_img.end(), 0.f);
        for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) {
            int byte = m * 128 + k * 4; int chunk16 = (byte >> 4) & 7, row = (byte >> 7) & 7; byte = (byte & ~0x70) | ((chunk16 ^ row) << 4);
            a_img[byte / 4] = A[m * K + k];
        }
    }
    unsigned long long lbo = chunk >> 4, sbo = kgrp >> 4;
    if (kmajor) { lbo = 1; sbo = 1024 >> 4; }
    int ltype = lt_none ? 0 : 2;
    if (variant & 32) {  // SWIZZLE_128B_BASE32B (layout type 1): atoms of 4 k-rows x 128 B, 32-byte chunks swizzled by the row
        std::fill(a_img.begin(), a_img.end(), 0.f);
        for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) {
            int c = m / 32, mm = m % 32;
            int byte = c * chunk + k * 128 + mm * 4;
            if (!noswz) { int c32 = (byte >> 5) & 3, row = (byte >> 7) & 3; byte = (byte & ~0x60) | ((c32 ^ row) << 5); }
            a_img[byte / 4] = A[m * K + k];
        }
        ltype = 1; lbo = chunk >> 4; sbo = 512 >> 4;
        if (swap) { unsigned long long t = lbo; lbo = sbo; sbo = t; }
    }
    if (variant & 64) {  // SWIZZLE_NONE: 8x16-byte core matrices (8 k-rows of 4 mn elements), MN blocks SBO apart, K blocks LBO apart
        std::fill(a_img.begin(), a_img.end(), 0.f);
        for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) a_img[((m / 4) * 128 + k * 16 + (m % 4) * 4) / 4] = A[m * K + k];
        ltype = 0; sbo = 128 >> 4; lbo = 4096 >> 4;
        if (swap) { unsigned long long t = lbo; lbo = sbo; sbo = t; }
    }
    if (swap && !(variant & 96)) { unsigned long long t = lbo; lbo = sbo; sbo = t; }
    unsigned long long hi = (lbo << 16) | (sbo << 32) | (1ull << 46) | ((unsigned long long)ltype << 61);
    unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((kmajor ? 0u : 1u) << 15) | ((unsigned)(N >> 3) << 17) | ((unsigned)(M >> 4) << 24);
    float *da, *db, *dout; CK(cudaMalloc(&da, 16384)); CK(cudaMalloc(&db, 4096)); CK(cudaMalloc(&dout, M * 16 * 4));
    CK(cudaMemcpy(da, a_img.data(), 16384, cudaMemcpyHostToDevice)); CK(cudaMemcpy(db, b_img.data(), 4096, cudaMemcpyHostToDevice));
    CK(cudaMemset(dout, 0xff, M * 16 * 4));
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
    probe<<<1, 128, 16384 + 4096 + 2048>>>(da, db, dout, hi, idesc);
    CK(cudaDeviceSynchronize());
    std::vector<float> o(M * 16); CK(cudaMemcpy(o.data(), dout, M * 16 * 4, cudaMemcpyDeviceToHost));
    int bad = 0, zeros = 0; double maxerr = 0;
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
        double ref = 0; for (int k = 0; k < K; ++k) ref += (double)A[m * K + k] * B[n * K + k];
        double e = fabs(o[m * 16 + n] - ref); if (e > maxerr) maxerr = e; if (e > 1e-3) ++bad; if (o[m * 16 + n] == 0.f) ++zeros;
    }
    printf("variant %d (chunk %d, kgrp %d, swapLS %d, noswz %d, ltNone %d): bad %d/%d zeros %d maxerr %.4f  sample %.4f\n", variant, chunk, kgrp, (int)swap, (int)noswz, (int)lt_none, bad, M * N, zeros, maxerr, o[5 * 16 + 3]);
    return 0;
}

/*
 * Plain-C restatement of the reference's RAFT correlation path -- TEST INFRASTRUCTURE ONLY.
 *
 * Second, independent checker next to oracle/corr_oracle.py (no numpy / torch arithmetic in
 * here: explicit loops, fp32 accumulation in a fixed order).  Never linked into or called by
 * the product library.  Pinned by tests/test_oracle_golden.py against golden vectors produced
 * by the reference itself (tests/golden/make_golden.py).
 *
 * Reference lines restated (relative to /root/reference):
 *   oracle_corr_volume    raft/core/corr.py:93-116   (CorrBlock.corr: matmul + 1/sqrt(D))
 *   oracle_avg_pool2      raft/core/corr.py:41-43    (F.avg_pool2d(corr, 2, stride=2), floor mode)
 *   oracle_lookup         raft/core/corr.py:45-91 + raft/core/utils/utils.py:57-85
 *                         (grid_sample bilinear / zeros / align_corners=True, window index i along x)
 *   oracle_lookup_grad    autograd of the same lines (grid_sampler_2d_backward scatter)
 *   oracle_alt_lookup     raft/core/corr.py:142-167 + raft/alt_cuda_corr/correlation_kernel.cu:18-119
 *
 * Layouts: fmaps [B,D,H,W]; level l [B,N,H_l,W_l] dense (N = H*W); coords [B,2,H,W] (x, y);
 * out [B, L*K*K, H, W], K = 2r+1.
 *
 * Build:  gcc -O3 -march=native -fopenmp -shared -fPIC corr_oracle.c -o libcorr_oracle.so -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* V0[b][m][n] = (sum_k f1[b][k][m] * f2[b][k][n]) * (1/sqrt(D)) */
void oracle_corr_volume(const float* f1, const float* f2, int B, int D, int H, int W, float* v0) {
    const int64_t N = (int64_t)H * W;
    const float scale = 1.0f / sqrtf((float)D);
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b) {
        for (int64_t m = 0; m < N; ++m) {
            float* row = v0 + ((int64_t)b * N + m) * N;
            for (int64_t n = 0; n < N; ++n) row[n] = 0.0f;
            for (int k = 0; k < D; ++k) {
                const float a = f1[((int64_t)b * D + k) * N + m];
                const float* brow = f2 + ((int64_t)b * D + k) * N;
                for (int64_t n = 0; n < N; ++n) row[n] += a * brow[n];
            }
            for (int64_t n = 0; n < N; ++n) row[n] *= scale;
        }
    }
}

/* in: [imgs][Hi][Wi] -> out: [imgs][Hi/2][Wi/2] */
void oracle_avg_pool2(const float* in, int64_t imgs, int Hi, int Wi, float* out) {
    const int Ho = Hi / 2, Wo = Wi / 2;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < imgs; ++i) {
        const float* s = in + i * Hi * Wi;
        float* d = out + i * Ho * Wo;
        for (int y = 0; y < Ho; ++y)
            for (int x = 0; x < Wo; ++x)
                d[y * Wo + x] = (s[(2 * y) * Wi + 2 * x] + s[(2 * y) * Wi + 2 * x + 1] +
                                 s[(2 * y + 1) * Wi + 2 * x] + s[(2 * y + 1) * Wi + 2 * x + 1]) * 0.25f;
    }
}

static inline float tap(const float* img, int Hl, int Wl, int y, int x) {
    return (y >= 0 && y < Hl && x >= 0 && x < Wl) ? img[y * Wl + x] : 0.0f;
}

/* levels[l]: [B][N][H_l][W_l] with (H_l, W_l) = (H >> l, W >> l) by cascaded floor halving */
void oracle_lookup(const float* const* levels, const float* coords, int B, int H, int W, int L, int r,
                   float* out) {
    const int K = 2 * r + 1;
    const int64_t N = (int64_t)H * W;
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b) {
        for (int64_t p = 0; p < N; ++p) {
            const float cx = coords[((int64_t)b * 2 + 0) * N + p];
            const float cy = coords[((int64_t)b * 2 + 1) * N + p];
            int Hl = H, Wl = W;
            for (int l = 0; l < L; ++l) {
                const float* img = levels[l] + ((int64_t)b * N + p) * Hl * Wl;
                const float x = cx / (float)(1 << l), y = cy / (float)(1 << l);
                const float xf = floorf(x), yf = floorf(y);
                const float fx = x - xf, fy = y - yf;
                const int x0 = (int)xf - r, y0 = (int)yf - r;
                for (int i = 0; i < K; ++i) {      /* i: x offset (first window index) */
                    for (int j = 0; j < K; ++j) {  /* j: y offset */
                        const int xx = x0 + i, yy = y0 + j;
                        const float v = (1.0f - fy) * (1.0f - fx) * tap(img, Hl, Wl, yy, xx) +
                                        (1.0f - fy) * fx * tap(img, Hl, Wl, yy, xx + 1) +
                                        fy * (1.0f - fx) * tap(img, Hl, Wl, yy + 1, xx) +
                                        fy * fx * tap(img, Hl, Wl, yy + 1, xx + 1);
                        out[((int64_t)b * L * K * K + (int64_t)l * K * K + i * K + j) * N + p] = v;
                    }
                }
                Hl /= 2;
                Wl /= 2;
            }
        }
    }
}

/* dlevels[l] (zero-initialised by the caller) += adjoint of oracle_lookup applied to grad_out */
void oracle_lookup_grad(const float* coords, const float* grad_out, int B, int H, int W, int L, int r,
                        float* const* dlevels) {
    const int K = 2 * r + 1;
    const int64_t N = (int64_t)H * W;
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b) {
        for (int64_t p = 0; p < N; ++p) { /* each (b,p) owns its own target image: no races */
            const float cx = coords[((int64_t)b * 2 + 0) * N + p];
            const float cy = coords[((int64_t)b * 2 + 1) * N + p];
            int Hl = H, Wl = W;
            for (int l = 0; l < L; ++l) {
                float* img = dlevels[l] + ((int64_t)b * N + p) * Hl * Wl;
                const float x = cx / (float)(1 << l), y = cy / (float)(1 << l);
                const float xf = floorf(x), yf = floorf(y);
                const float fx = x - xf, fy = y - yf;
                const int x0 = (int)xf - r, y0 = (int)yf - r;
                for (int i = 0; i < K; ++i) {
                    for (int j = 0; j < K; ++j) {
                        const float g = grad_out[((int64_t)b * L * K * K + (int64_t)l * K * K + i * K + j) * N + p];
                        const int xx = x0 + i, yy = y0 + j;
                        const float w[4] = {(1.0f - fy) * (1.0f - fx), (1.0f - fy) * fx, fy * (1.0f - fx), fy * fx};
                        const int ys[4] = {yy, yy, yy + 1, yy + 1}, xs[4] = {xx, xx + 1, xx, xx + 1};
                        for (int c = 0; c < 4; ++c)
                            if (ys[c] >= 0 && ys[c] < Hl && xs[c] >= 0 && xs[c] < Wl)
                                img[ys[c] * Wl + xs[c]] += w[c] * g;
                    }
                }
                Hl /= 2;
                Wl /= 2;
            }
        }
    }
}

/* On-the-fly variant.  f2_levels[l]: pooled target features, NCHW [B][D][H_l][W_l]. */
void oracle_alt_lookup(const float* f1, const float* const* f2_levels, const float* coords, int B, int D,
                       int H, int W, int L, int r, float* out) {
    const int K = 2 * r + 1;
    const int64_t N = (int64_t)H * W;
    const float scale = sqrtf((float)D);
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b) {
        for (int64_t p = 0; p < N; ++p) {
            float* s = (float*)malloc(sizeof(float) * (K + 1) * (K + 1));
            const float cx = coords[((int64_t)b * 2 + 0) * N + p];
            const float cy = coords[((int64_t)b * 2 + 1) * N + p];
            int Hl = H, Wl = W;
            for (int l = 0; l < L; ++l) {
                const float* f2 = f2_levels[l] + (int64_t)b * D * Hl * Wl;
                const float x = cx / (float)(1 << l), y = cy / (float)(1 << l);
                const float xf = floorf(x), yf = floorf(y);
                const float dx = x - xf, dy = y - yf;
                for (int iy = 0; iy <= K; ++iy) {
                    for (int ix = 0; ix <= K; ++ix) {
                        const int h2 = (int)yf - r + iy, w2 = (int)xf - r + ix;
                        float acc = 0.0f;
                        if (h2 >= 0 && h2 < Hl && w2 >= 0 && w2 < Wl)
                            for (int k = 0; k < D; ++k)
                                acc += f1[((int64_t)b * D + k) * N + p] * f2[((int64_t)k * Hl + h2) * Wl + w2];
                        s[iy * (K + 1) + ix] = acc;
                    }
                }
                for (int ix = 0; ix < K; ++ix) {
                    for (int iy = 0; iy < K; ++iy) {
                        const float v = s[(iy + 1) * (K + 1) + ix + 1] * dy * dx +          /* nw */
                                        s[(iy + 1) * (K + 1) + ix] * dy * (1.0f - dx) +     /* ne */
                                        s[iy * (K + 1) + ix + 1] * (1.0f - dy) * dx +       /* sw */
                                        s[iy * (K + 1) + ix] * (1.0f - dy) * (1.0f - dx);   /* se */
                        out[((int64_t)b * L * K * K + (int64_t)l * K * K + iy + K * ix) * N + p] = v / scale;
                    }
                }
                Hl /= 2;
                Wl /= 2;
            }
            free(s);
        }
    }
}

// Projection into the estimator's feature space (north_star piece 1; seam S2,
// ref:src/sknnr/_base.py:236-239).  One affine map covers the four float transformers:
//     Z = ((X - center) / scale) @ proj
//   StandardScalerWithDOF   $SP/sklearn/preprocessing/_data.py:1131-1134
//   MahalanobisTransformer  ref:src/sknnr/transformers/_mahalanobis_transformer.py:55
//   CCorATransformer        ref:src/sknnr/transformers/_ccora_transformer.py:70
//   CCATransformer          ref:src/sknnr/transformers/_cca_transformer.py:87
// computed in float64 exactly as written (true division, centre first).  Raw map pixels are
// read from HBM once; the kernel emits both operands the rest of the path needs:
//   z64  [n_q, d_out] float64          - exact re-rank operand
//   qimg [n_qtiles][dpad][256] float32 - search-kernel query tile image, -2 * (Z - mu)
#include "common.cuh"
#include "kernels.h"

namespace sk {

constexpr int PROJ_THREADS = 128;  // one thread per query row; 128 rows per CTA
constexpr int PROJ_KCHUNK = 8;

template <typename TX>
__global__ void __launch_bounds__(PROJ_THREADS)
project_kernel(const TX *__restrict__ X, long long ldx, long long n_q, int d_in, int d_out,
               int dpad, const double *__restrict__ center, const double *__restrict__ scale,
               const double *__restrict__ proj, const double *__restrict__ mu,
               double *__restrict__ z64, float *__restrict__ qimg, float *__restrict__ qimg_tc,
               int tc_mt, const int *__restrict__ n_rows_dev) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int xs_ld = d_in | 1;  // odd row stride (in doubles): conflict-free row reads
    double *xs = reinterpret_cast<double *>(smem_raw);           // [128][xs_ld]
    double *ps = xs + (size_t)PROJ_THREADS * xs_ld;              // [d_in][d_out] (if proj)
    const long long q0 = (long long)blockIdx.x * PROJ_THREADS;
    if (n_rows_dev) {  // compacted launch: only the first *n_rows_dev rows exist
        const long long n_dev = *n_rows_dev;
        if (q0 >= (n_dev + QTILE - 1) / QTILE * QTILE) return;
        n_q = min(n_q, n_dev);
    }
    const int rows = (int)min((long long)PROJ_THREADS, n_q - q0);

    // coalesced load of the tile's rows, centring and scaling on the way in; (row, column) advance
    // incrementally (no integer division per element).  Without a projector the scaled value already
    // is z: it goes to z64 from here, where consecutive threads write consecutive addresses.
    {
        const int dr = PROJ_THREADS / d_in, dc = PROJ_THREADS - dr * d_in;
        int r = threadIdx.x / d_in, c = threadIdx.x - r * d_in;
        const bool z_here = !proj && z64 && d_out == d_in;
        for (int e = threadIdx.x; e < PROJ_THREADS * d_in; e += PROJ_THREADS) {
            double v = 0.0;
            if (r < rows) {
                v = (double)X[(q0 + r) * ldx + c];
                if (center) v -= center[c];
                if (scale) v /= scale[c];
                if (z_here) z64[(q0 + r) * d_out + c] = v;
            }
            xs[r * xs_ld + c] = v;
            r += dr;
            c += dc;
            if (c >= d_in) {
                c -= d_in;
                ++r;
            }
        }
    }
    const bool z_later = z64 && (proj || d_out != d_in);
    if (proj)
        for (int e = threadIdx.x; e < d_in * d_out; e += PROJ_THREADS) ps[e] = proj[e];
    __syncthreads();

    const int r = threadIdx.x;
    const double *xr = xs + r * xs_ld;
    // row (q0 + r) lives in query tile (q0 + r) / QTILE at column (q0 + r) % QTILE
    const long long qrow = q0 + r;
    float *qt = qimg ? qimg + (size_t)(qrow / QTILE) * dpad * QTILE + (qrow % QTILE) : nullptr;
    // tensor-core image: consecutive 128-row operands [chunk][row][4] TF32 (tc_mt of them form
    // one CTA tile), with one extra K block (two chunks) that folds |r|^2 into the contraction:
    // (1,1,1,0 | 0,0,0,0)
    const int kc_tot = dpad / 4 + 2;
    float4 *qc = nullptr;
    if (qimg_tc) {
        qc = reinterpret_cast<float4 *>(qimg_tc) + ((size_t)(qrow / TC_M) * kc_tot) * TC_M + (qrow % TC_M);
    }
    for (int k0 = 0; k0 < dpad; k0 += PROJ_KCHUNK) {
        double z[PROJ_KCHUNK];
#pragma unroll
        for (int kk = 0; kk < PROJ_KCHUNK; ++kk) z[kk] = 0.0;
        if (proj) {
            for (int i = 0; i < d_in; ++i) {
                const double xv = xr[i];
                const double *pr = ps + i * d_out + k0;
#pragma unroll
                for (int kk = 0; kk < PROJ_KCHUNK; ++kk)
                    if (k0 + kk < d_out) z[kk] += xv * pr[kk];
            }
        } else {
#pragma unroll
            for (int kk = 0; kk < PROJ_KCHUNK; ++kk)
                if (k0 + kk < d_out) z[kk] = xr[k0 + kk];
        }
        float svs[PROJ_KCHUNK];
#pragma unroll
        for (int kk = 0; kk < PROJ_KCHUNK; ++kk) {
            const int k = k0 + kk;
            float sv = 0.0f;
            if (k < d_out && r < rows) {
                if (z_later) z64[(q0 + r) * d_out + k] = z[kk];
                sv = (float)(-2.0 * (z[kk] - mu[k]));
            }
            svs[kk] = sv;
            if (qt) qt[(size_t)k * QTILE] = sv;
        }
        if (qc) {
            qc[(size_t)(k0 / 4) * TC_M] =
                make_float4(to_tf32(svs[0]), to_tf32(svs[1]), to_tf32(svs[2]), to_tf32(svs[3]));
            qc[(size_t)(k0 / 4 + 1) * TC_M] =
                make_float4(to_tf32(svs[4]), to_tf32(svs[5]), to_tf32(svs[6]), to_tf32(svs[7]));
        }
    }
    if (qc) {
        const float one = (r < rows) ? 1.0f : 0.0f;
        qc[(size_t)(dpad / 4) * TC_M] = make_float4(one, one, one, 0.0f);
        qc[(size_t)(dpad / 4 + 1) * TC_M] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
}

cudaError_t launch_project(const void *X, int x_is_f32, long long ldx, long long n_q, int d_in,
                           int d_out, int dpad, const double *center, const double *scale,
                           const double *proj, const double *mu, double *z64, float *qimg,
                           float *qimg_tc, int tc_mt, const int *n_rows_dev, cudaStream_t st) {
    (void)tc_mt;
    if (n_q <= 0) return cudaSuccess;
    const int xs_ld = d_in | 1;
    const size_t smem = ((size_t)PROJ_THREADS * xs_ld + (proj ? (size_t)d_in * d_out : 0)) * 8;
    if (smem > 227 * 1024) return cudaErrorInvalidValue;
    // cover whole query tiles so that padding columns of the last tile are zero-filled (every
    // CTA tile size of the search engines divides padded_rows, a multiple of PROJ_THREADS)
    const long long padded = padded_rows(n_q);
    const long long grid = (padded + PROJ_THREADS - 1) / PROJ_THREADS;
    cudaError_t e;
    if (x_is_f32) {
        e = cudaFuncSetAttribute(project_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 227 * 1024);
        if (e != cudaSuccess) return e;
        project_kernel<float><<<(unsigned)grid, PROJ_THREADS, smem, st>>>(
            (const float *)X, ldx, n_q, d_in, d_out, dpad, center, scale, proj, mu, z64, qimg, qimg_tc,
            tc_mt, n_rows_dev);
    } else {
        e = cudaFuncSetAttribute(project_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 227 * 1024);
        if (e != cudaSuccess) return e;
        project_kernel<double><<<(unsigned)grid, PROJ_THREADS, smem, st>>>(
            (const double *)X, ldx, n_q, d_in, d_out, dpad, center, scale, proj, mu, z64, qimg, qimg_tc,
            tc_mt, n_rows_dev);
    }
    return cudaGetLastError();
}

}  // namespace sk

// k5_producers.cu — K5: calibration-map producers, sm_100a.  Streaming float64 reductions over small
// stacks of frames: HBM bound (each input sample is read once, each output written once), no reuse, no tensor cores.
//
//   k5_stack_mean   imgAverage (transform/imgAverage.py:7-22: out = float(i0); out += i_k ...; out /= n, float64, that order)
//                   optionally followed by `img -= bg` (array or number) and, for 3-channel frames, toGray
//                   (transformations.py:126-135: np.average(img, axis=-1, weights=(0.299, 0.587, 0.114)) =
//                   ((r w0 + g w1) + b w2) / ((w0 + w1) + w2), every step rounded) — the first half of
//                   flatFieldFromCloseDistance (camera/flatField/flatFieldFromCloseDistance.py:16-38)
//   k5_scale        img /= mx   (its last line)
//   k5_subsample    img[::sy, ::sx]   (the argument of its median_filter)
//   k5_linear_fit   getLinearityFunction (camera/DarkCurrentMap.py:61-80): per-pixel least-squares line through the exposure
//                   series with samples above mxIntensity masked out, then the reference's clean-up of the ascent.
//                   The regression itself is fancytools.linRegressUsingMasked2dArrays — absent from the reference tree, its
//                   published algorithm (ordinary least squares over the unmasked samples of each pixel, RMSE of the
//                   residuals) is restated here and in oracle/producers.py: PARITY UNPINNED for that ingredient.
#include "imgcorr_kernels.cuh"

namespace imgcorr {

namespace {

template <typename T>
__global__ void __launch_bounds__(256) k5_stack_mean_kernel(const T* __restrict__ frames, int n, size_t elems, const double* __restrict__ minus,
                                                            double minus_scalar, int has_scalar, int gray3, double* __restrict__ out) {
    const size_t n_out = gray3 ? elems / 3 : elems;
    const double w0 = 0.299, w1 = 0.587, w2 = 0.114;
    const double scl = dadd(dadd(w0, w1), w2);
    for (size_t o = blockIdx.x * (size_t)blockDim.x + threadIdx.x; o < n_out; o += (size_t)gridDim.x * blockDim.x) {
        double v[3];
        const int nc = gray3 ? 3 : 1;
        for (int c = 0; c < nc; ++c) {
            const size_t i = gray3 ? o * 3 + c : o;
            double s = (double)__ldg(frames + i);
            for (int k = 1; k < n; ++k) s = dadd(s, (double)__ldg(frames + (size_t)k * elems + i));
            s = ddiv(s, (double)n);
            if (minus) s = dsub(s, __ldg(minus + i));
            else if (has_scalar) s = dsub(s, minus_scalar);
            v[c] = s;
        }
        out[o] = gray3 ? ddiv(dadd(dadd(dmul(v[0], w0), dmul(v[1], w1)), dmul(v[2], w2)), scl) : v[0];
    }
}

__global__ void __launch_bounds__(256) k5_scale_kernel(double* data, size_t elems, double divisor) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < elems; i += (size_t)gridDim.x * blockDim.x)
        data[i] = ddiv(data[i], divisor);
}

__global__ void __launch_bounds__(256) k5_subsample_kernel(const double* __restrict__ src, int H, int W, int sy, int sx, double* __restrict__ dst,
                                                          int h, int w) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x < w && y < h) dst[(size_t)y * w + x] = __ldg(src + (size_t)(y * sy) * W + (size_t)x * sx);
}

// ordinary least squares over the unmasked samples of one pixel:  y = offset + ascent * x
//   ascent = (m Sxy - Sx Sy) / (m Sxx - Sx^2),  offset = (Sy - ascent Sx) / m,  rmse = sqrt(mean(residual^2))   (m = unmasked samples)
template <typename T>
__global__ void __launch_bounds__(256) k5_linear_fit_kernel(const T* __restrict__ frames, int n, size_t px, const double* __restrict__ xs,
                                                            double max_intensity, double min_ascent, double x_mid, double* __restrict__ offset,
                                                            double* __restrict__ ascent, double* __restrict__ rmse) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < px; i += (size_t)gridDim.x * blockDim.x) {
        double sx = 0.0, sy = 0.0, sxy = 0.0, sxx = 0.0;
        int m = 0;
        for (int k = 0; k < n; ++k) {
            const double y = (double)__ldg(frames + (size_t)k * px + i);
            if (y > max_intensity) continue;                       // masked (NaN compares false: kept, as numpy's `imgs > mx`)
            const double x = xs[k];
            sx = dadd(sx, x); sy = dadd(sy, y); sxy = dadd(sxy, dmul(x, y)); sxx = dadd(sxx, dmul(x, x));
            ++m;
        }
        const double dm = (double)m;
        double a = ddiv(dsub(dmul(dm, sxy), dmul(sx, sy)), dsub(dmul(dm, sxx), dmul(sx, sx)));
        double b = ddiv(dsub(sy, dmul(a, sx)), dm);
        double e = 0.0;
        for (int k = 0; k < n; ++k) {
            const double y = (double)__ldg(frames + (size_t)k * px + i);
            if (y > max_intensity) continue;
            const double r = dsub(y, dadd(b, dmul(a, xs[k])));
            e = dadd(e, dmul(r, r));
        }
        e = sqrt(ddiv(e, dm));
        // DarkCurrentMap.py:72-78: ascent[isnan] = 0; where ascent < min_ascent: offset += 0.5 (min t + max t) ascent; ascent = 0
        if (a != a) a = 0.0;
        if (min_ascent > 0.0 && a < min_ascent) {
            b = dadd(b, dmul(x_mid, a));
            a = 0.0;
        }
        offset[i] = b;
        ascent[i] = a;
        if (rmse) rmse[i] = e;
    }
}

int grid_for(size_t n, int sm_count) {
    const size_t want = (n + 255) / 256, cap = (size_t)sm_count * 16;
    return (int)(want < cap ? (want ? want : 1) : cap);
}

}  // namespace

cudaError_t launch_k5_stack_mean(const void* frames, int dtype, int n, size_t elems, const double* minus, double minus_scalar,
                                 int has_scalar, int gray3, double* out, int sm_count, cudaStream_t st, int* launches) {
    if (n <= 0 || elems == 0) return cudaSuccess;
    if (gray3 && elems % 3) return cudaErrorInvalidValue;
    const int grid = grid_for(gray3 ? elems / 3 : elems, sm_count);
    switch (dtype) {
        case DT_U8: k5_stack_mean_kernel<<<grid, 256, 0, st>>>((const uint8_t*)frames, n, elems, minus, minus_scalar, has_scalar, gray3, out); break;
        case DT_U16: k5_stack_mean_kernel<<<grid, 256, 0, st>>>((const uint16_t*)frames, n, elems, minus, minus_scalar, has_scalar, gray3, out); break;
        case DT_F32: k5_stack_mean_kernel<<<grid, 256, 0, st>>>((const float*)frames, n, elems, minus, minus_scalar, has_scalar, gray3, out); break;
        case DT_F64: k5_stack_mean_kernel<<<grid, 256, 0, st>>>((const double*)frames, n, elems, minus, minus_scalar, has_scalar, gray3, out); break;
        default: return cudaErrorInvalidValue;
    }
    if (launches) ++*launches;
    return cudaGetLastError();
}

cudaError_t launch_k5_scale(double* data, size_t elems, double divisor, int sm_count, cudaStream_t st, int* launches) {
    if (elems == 0) return cudaSuccess;
    k5_scale_kernel<<<grid_for(elems, sm_count), 256, 0, st>>>(data, elems, divisor);
    if (launches) ++*launches;
    return cudaGetLastError();
}

cudaError_t launch_k5_subsample(const double* src, int H, int W, int sy, int sx, double* dst, cudaStream_t st, int* launches) {
    if (H <= 0 || W <= 0 || sy <= 0 || sx <= 0) return cudaErrorInvalidValue;
    const int h = (H + sy - 1) / sy, w = (W + sx - 1) / sx;
    k5_subsample_kernel<<<dim3((w + 31) / 32, (h + 7) / 8), 256, 0, st>>>(src, H, W, sy, sx, dst, h, w);
    if (launches) ++*launches;
    return cudaGetLastError();
}

cudaError_t launch_k5_linear_fit(const void* frames, int dtype, int n, size_t px, const double* xs_dev, double max_intensity, double min_ascent,
                                 double x_mid, double* offset, double* ascent, double* rmse, int sm_count, cudaStream_t st, int* launches) {
    if (n <= 0 || px == 0) return cudaSuccess;
    const int grid = grid_for(px, sm_count);
    switch (dtype) {
        case DT_U8: k5_linear_fit_kernel<<<grid, 256, 0, st>>>((const uint8_t*)frames, n, px, xs_dev, max_intensity, min_ascent, x_mid, offset, ascent, rmse); break;
        case DT_U16: k5_linear_fit_kernel<<<grid, 256, 0, st>>>((const uint16_t*)frames, n, px, xs_dev, max_intensity, min_ascent, x_mid, offset, ascent, rmse); break;
        case DT_F32: k5_linear_fit_kernel<<<grid, 256, 0, st>>>((const float*)frames, n, px, xs_dev, max_intensity, min_ascent, x_mid, offset, ascent, rmse); break;
        case DT_F64: k5_linear_fit_kernel<<<grid, 256, 0, st>>>((const double*)frames, n, px, xs_dev, max_intensity, min_ascent, x_mid, offset, ascent, rmse); break;
        default: return cudaErrorInvalidValue;
    }
    if (launches) ++*launches;
    return cudaGetLastError();
}

}  // namespace imgcorr

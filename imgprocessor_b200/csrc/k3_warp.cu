// k3_warp.cu — K3: perspective warp, sm_100a.
//
// Replaces cv2.warpPerspective in PerspectiveCorrection.correct (INTER_LANCZOS4,
// camera/PerspectiveCorrection.py:401-405) and .uncorrect (INTER_CUBIC | WARP_INVERSE_MAP, :374-378), and the
// tilt-factor division in front of it (:394-400).  Arithmetic: imgcorr_warp.cuh (bit-exact with OpenCV).
//
// One thread per output pixel x all frames of the launch: the float64 homography (a division and ~14 DP
// operations) and the 2 x N coefficient fetches are done once, then every frame gathers its N x N window.  N x N
// products + sums per pixel with OpenCV's rounding order cannot be separated into two 1-D passes, so the kernel
// is bound by load/FP32 issue (64 loads + 192 FP32 operations per Lanczos4 pixel), not by HBM; consecutive
// lanes read consecutive source columns, so the gathers are coalesced L1 hits for any sane homography.
#include <algorithm>
#include "imgcorr_kernels.cuh"
#include "imgcorr_warp.cuh"

namespace imgcorr {

namespace {

template <typename T> struct OutCast;
template <> struct OutCast<float> { static __device__ __forceinline__ float cast(float v) { return v; } };
template <> struct OutCast<double> { static __device__ __forceinline__ double cast(double v) { return v; } };
template <> struct OutCast<uint16_t> { static __device__ __forceinline__ uint16_t cast(float v) { return sat_u16(v); } };

template <typename T> __device__ __forceinline__ T ldg(const T* p) { return __ldg(p); }

template <typename T, typename AT, int N>
__device__ __noinline__ AT warp_border_ool(const T* S0, int H, int W, int sx, int sy, const float* wy, const float* wx, AT cv) {
    return warp_sum_border<T, AT, N>(S0, H, W, sx, sy, wy, wx, cv);
}

constexpr int K3_BX = 32, K3_BY = 8;

template <typename T, typename AT, int N>
__global__ void __launch_bounds__(K3_BX* K3_BY) k3_warp_kernel(const K3Args a) {
    __shared__ float tab[N][32];   // transposed: tab[tap][phase] -> lanes with different phases hit different banks
    for (int i = threadIdx.y * K3_BX + threadIdx.x; i < N * 32; i += K3_BX * K3_BY) tab[i % N][i / N] = a.tab[i];
    __syncthreads();
    const int x = blockIdx.x * K3_BX + threadIdx.x, y = blockIdx.y * K3_BY + threadIdx.y;
    if (x >= a.dw || y >= a.dh) return;
    const FixedCoord c = warp_coord(a.wc, x, y);
    constexpr int OFF = N / 2 - 1;
    const int sx = c.ix - OFF, sy = c.iy - OFF;
    float wy[N], wx[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        wy[i] = tab[i][c.fy];
        wx[i] = tab[i][c.fx];
    }
    const int w1 = max(a.W - (N - 1), 0), h1 = max(a.H - (N - 1), 0);
    const bool interior = (unsigned)sx < (unsigned)w1 && (unsigned)sy < (unsigned)h1;
    const bool outside = sx >= a.W || sx + N <= 0 || sy >= a.H || sy + N <= 0;
    const AT cv = (AT)a.border;
    const size_t spx = (size_t)a.H * a.W, dpx = (size_t)a.dh * a.dw;
    const T* S = (const T*)a.src;
    T* D = (T*)a.dst + (size_t)y * a.dw + x;
    const long long o0 = (long long)sy * a.W + sx;
    for (int f = 0; f < a.n_frames; ++f, S += spx, D += dpx) {
        AT v;
        if (interior) {
            const T* P = S + o0;
            v = (AT)0;
#pragma unroll
            for (int r = 0; r < N; ++r) {
                T t[N];
#pragma unroll
                for (int k = 0; k < N; ++k) t[k] = ldg(P + k);
                AT acc = wmul((AT)t[0], (AT)fmul(wy[r], wx[0]));
#pragma unroll
                for (int k = 1; k < N; ++k) acc = wadd(acc, wmul((AT)t[k], (AT)fmul(wy[r], wx[k])));
                v = wadd(v, acc);
                P += a.W;
            }
        } else if (outside) {
            v = cv;
        } else {
            v = warp_border_ool<T, AT, N>(S, a.H, a.W, sx, sy, wy, wx, cv);
        }
        *D = OutCast<T>::cast(v);
    }
}

// image / tiltFactor in float64 (PerspectiveCorrection.py:394-400: np.asfarray(img) / tf)
template <typename T>
__global__ void __launch_bounds__(256) k3_divide_kernel(const T* __restrict__ src, const double* __restrict__ div, double* __restrict__ dst,
                                                         size_t npx, int n_frames) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npx; i += (size_t)gridDim.x * blockDim.x) {
        const double d = div[i];
        for (int f = 0; f < n_frames; ++f) dst[f * npx + i] = ddiv((double)src[f * npx + i], d);
    }
}

template <typename T, typename AT>
cudaError_t launch_typed(const K3Args& a, int interp, cudaStream_t st) {
    dim3 block(K3_BX, K3_BY), grid((a.dw + K3_BX - 1) / K3_BX, (a.dh + K3_BY - 1) / K3_BY);
    if (interp == WARP_LANCZOS4)
        k3_warp_kernel<T, AT, 8><<<grid, block, 0, st>>>(a);
    else
        k3_warp_kernel<T, AT, 4><<<grid, block, 0, st>>>(a);
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_k3(const K3Args& a, int dtype, int interp, cudaStream_t st, int* launches) {
    if (a.n_frames <= 0 || a.dw <= 0 || a.dh <= 0) return cudaSuccess;
    if (interp != WARP_LANCZOS4 && interp != WARP_CUBIC) return cudaErrorInvalidValue;
    cudaError_t e;
    switch (dtype) {
        case DT_U16: e = launch_typed<uint16_t, float>(a, interp, st); break;
        case DT_F32: e = launch_typed<float, float>(a, interp, st); break;
        case DT_F64: e = launch_typed<double, double>(a, interp, st); break;
        default: return cudaErrorInvalidValue;
    }
    if (launches) ++*launches;
    return e;
}

cudaError_t launch_k3_divide(const void* src, int dtype, const double* div, double* dst, size_t npx, int n_frames, int sm_count,
                             cudaStream_t st, int* launches) {
    if (npx == 0 || n_frames <= 0) return cudaSuccess;
    const int grid = (int)std::min<size_t>((npx + 255) / 256, (size_t)sm_count * 8);
    switch (dtype) {
        case DT_U8: k3_divide_kernel<<<grid, 256, 0, st>>>((const uint8_t*)src, div, dst, npx, n_frames); break;
        case DT_U16: k3_divide_kernel<<<grid, 256, 0, st>>>((const uint16_t*)src, div, dst, npx, n_frames); break;
        case DT_F32: k3_divide_kernel<<<grid, 256, 0, st>>>((const float*)src, div, dst, npx, n_frames); break;
        case DT_F64: k3_divide_kernel<<<grid, 256, 0, st>>>((const double*)src, div, dst, npx, n_frames); break;
        default: return cudaErrorInvalidValue;
    }
    if (launches) ++*launches;
    return cudaGetLastError();
}

}  // namespace imgcorr

// k4_ste.cu — K4: single-time-effect-free average of N exposures, sm_100a.
//
// Replaces SingleTimeEffectDetection(images, nStd, noise_level_function).noSTE in the multi-image branch of
// CameraCalibration.correct() (camera/CameraCalibration.py:385-406; features/SingleTimeEffectDetection.py:23-75).
// Arithmetic: imgcorr_ste.cuh (float64, as the reference).
//
// One launch per added image, because the 3x3 "remove single pixels" stencil on the STE mask makes image k+1 depend on
// the neighbours' averages after image k.  A CTA owns a 32x32 tile: every thread computes the STE flags of its four pixels
// (all loads in flight together) and the first 132 threads also compute the flags of the 1-pixel halo (re-reading the neighbours' image /
// average / threshold through L1/L2), the flags meet in shared memory, and each pixel then counts its eight neighbours.
//   FIRST  launch: avg = min(i1, i2), n = 1, thr = nlf(avg) * nStd, then addImage(max(i1, i2))          (:36-48)
//   NEXT   launch: addImage(image)                                                                      (:50-52, 57-75)
// The running average is ping-ponged between two buffers (the halo flags of a neighbouring tile must see the average
// before this launch).  HBM traffic per pixel: FIRST 2 reads of the raw dtype + 20 B written (avg, thr float64, n int32);
// NEXT raw + 20 B read + 8-12 B written — streaming, no reuse beyond the halo.
#include "imgcorr_kernels.cuh"
#include "imgcorr_ste.cuh"

namespace imgcorr {

namespace {

constexpr int K4_TW = 32, K4_TH = 32, K4_THREADS = 256, K4_RPT = K4_TH / (K4_THREADS / K4_TW);   // 4 rows per thread

template <typename T, bool FIRST>
struct SteLoad {
    // value of (image, avg, thr) at one pixel
    static __device__ __forceinline__ void get(const K4Args& a, size_t i, double& img, double& avg, double& thr) {
        if (FIRST) {
            const double p = (double)__ldg((const T*)a.img + i), q = (double)__ldg((const T*)a.img2 + i);
            // np.min / np.max over the pair (:40, :48); a NaN in either propagates as numpy's reduction does
            avg = (p != p || q != q) ? (p + q) : (p < q ? p : q);
            img = (p != p || q != q) ? (p + q) : (p < q ? q : p);
            thr = a.thr_in ? __ldg(a.thr_in + i) : ste_threshold(a.sc, avg);
        } else {
            img = (double)__ldg((const T*)a.img + i);
            avg = __ldg(a.avg_in + i);
            thr = __ldg(a.thr + i);
        }
    }
};

template <typename T, bool FIRST>
__global__ void __launch_bounds__(K4_THREADS) k4_ste_kernel(const K4Args a) {
    __shared__ uint8_t flag[K4_TH + 2][K4_TW + 4];
    const int tx = threadIdx.x % K4_TW, ty = threadIdx.x / K4_TW;
    const int x0 = blockIdx.x * K4_TW, y0 = blockIdx.y * K4_TH;
    const int H = a.H, W = a.W;
    // own pixels first (rows ty, ty + 8, ...): all loads of a thread are in flight together
    const int gx = x0 + tx;
    double img[K4_RPT], avg[K4_RPT], thr[K4_RPT];
    int cnt[K4_RPT];
    bool live[K4_RPT], f[K4_RPT];
#pragma unroll
    for (int j = 0; j < K4_RPT; ++j) {
        const int gy = y0 + ty + j * (K4_THREADS / K4_TW);
        live[j] = gx < W && gy < H;
        img[j] = avg[j] = thr[j] = 0.0;
        cnt[j] = 1;
        if (live[j]) {
            SteLoad<T, FIRST>::get(a, (size_t)gy * W + gx, img[j], avg[j], thr[j]);
            if (!FIRST) cnt[j] = __ldg(a.n + ((size_t)gy * W + gx));
        }
    }
    // halo cells: 2 * (TW + 2) + 2 * TH = 132 cells, one per thread of the first warps
    {
        const int h = threadIdx.x;
        int cx = -2, cy = -2;
        if (h < K4_TW + 2) { cx = h - 1; cy = -1; }
        else if (h < 2 * (K4_TW + 2)) { cx = h - (K4_TW + 2) - 1; cy = K4_TH; }
        else if (h < 2 * (K4_TW + 2) + K4_TH) { cx = -1; cy = h - 2 * (K4_TW + 2); }
        else if (h < 2 * (K4_TW + 2) + 2 * K4_TH) { cx = K4_TW; cy = h - 2 * (K4_TW + 2) - K4_TH; }
        if (cx > -2) {
            const int hx = x0 + cx, hy = y0 + cy;
            bool hf = false;
            if ((unsigned)hx < (unsigned)W && (unsigned)hy < (unsigned)H) {      // outside the image: no neighbour (:19-20)
                double hi, ha, ht;
                SteLoad<T, FIRST>::get(a, (size_t)hy * W + hx, hi, ha, ht);
                hf = ste_flag(hi, ha, ht);
            }
            flag[cy + 1][cx + 1] = hf;
        }
    }
#pragma unroll
    for (int j = 0; j < K4_RPT; ++j) {
        f[j] = live[j] && ste_flag(img[j], avg[j], thr[j]);
        flag[ty + j * (K4_THREADS / K4_TW) + 1][tx + 1] = f[j];
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < K4_RPT; ++j) {
        if (!live[j]) continue;
        const int r = ty + j * (K4_THREADS / K4_TW);
        const size_t i = (size_t)(y0 + r) * W + gx;
        bool s = f[j];
        if (s) {
            // removeSinglePixels: an STE pixel without an STE neighbour is not an STE (:22-33)
            const int nb = flag[r][tx] + flag[r][tx + 1] + flag[r][tx + 2] + flag[r + 1][tx] + flag[r + 1][tx + 2] +
                           flag[r + 2][tx] + flag[r + 2][tx + 1] + flag[r + 2][tx + 2];
            s = nb > 0;
        }
        int n = cnt[j];
        double v = avg[j];
        if (!s) {                                    // clean: MaskedMovingAverage.update (:68-69)
            n += 1;
            v = ste_update_fast<sizeof(T) != 8>(img[j], v, n);
        }
        if (FIRST) a.thr[i] = thr[j];
        a.avg_out[i] = v;                            // ping-pong: neighbouring tiles still read avg_in for their halo flags
        if (FIRST || !s) a.n[i] = n;
        if (a.mask) {
            if (FIRST) a.mask[i] = s;
            else if (s) a.mask[i] = 1;               // mask_STE += ste (:71-72)
        }
    }
}

template <typename T>
cudaError_t launch_typed(const K4Args& a, cudaStream_t st) {
    dim3 grid((a.W + K4_TW - 1) / K4_TW, (a.H + K4_TH - 1) / K4_TH);
    if (a.img2)
        k4_ste_kernel<T, true><<<grid, K4_THREADS, 0, st>>>(a);
    else
        k4_ste_kernel<T, false><<<grid, K4_THREADS, 0, st>>>(a);
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_k4(const K4Args& a, int dtype, cudaStream_t st, int* launches) {
    if (a.H <= 0 || a.W <= 0) return cudaSuccess;
    cudaError_t e;
    switch (dtype) {
        case DT_U8: e = launch_typed<uint8_t>(a, st); break;
        case DT_U16: e = launch_typed<uint16_t>(a, st); break;
        case DT_F32: e = launch_typed<float>(a, st); break;
        case DT_F64: e = launch_typed<double>(a, st); break;
        default: return cudaErrorInvalidValue;
    }
    if (launches) ++*launches;
    return e;
}

}  // namespace imgcorr

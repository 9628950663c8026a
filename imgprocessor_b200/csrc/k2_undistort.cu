// K2 — lens undistortion: per output pixel evaluate the Brown-Conrady model analytically
// (float64, one rounding to float32 — no map arrays are read), convert to OpenCV's 5-bit
// fixed-point coordinates and gather the bilinear blend from the K1 output through L1/L2.
//
// Replaces  LensDistortion.correct()  camera/LensDistortion.py:316-330:
//   cv2.getOptimalNewCameraMatrix (host, stays cv2)  ->  P           :350-353
//   cv2.initUndistortRectifyMap(K, d, None, P, CV_32FC1)              :355-357   (device, analytic)
//   cv2.remap(INTER_LINEAR, BORDER_CONSTANT, borderValue)             :323-326   (device)
//   optional crop to roi                                              :327-329   (output window)
// The same kernel also serves explicit float32 maps (distortImage, parity tests against
// cv2's own maps).  When n_frames > 1 the coordinates and weights of a pixel are computed once
// and applied to every frame of the launch.
//
// A thread owns one output column and four rows of a 32x32 tile (consecutive lanes = consecutive
// columns: coalesced stores, neighbouring gathers).  When P^-1 has no cross terms (what
// getOptimalNewCameraMatrix always produces) x depends on the column only and is hoisted.
#include "imgcorr_kernels.cuh"
#include "imgcorr_tma.cuh"

namespace imgcorr {

#ifndef K2_ROWS_V
#define K2_ROWS_V 2
#endif
#ifndef K2_MINB_V
#define K2_MINB_V 5
#endif
constexpr int K2_BX = 32, K2_BY = 8, K2_ROWS = K2_ROWS_V;     // CTA = 32 x 8 threads, tile = 32 x (8 * K2_ROWS) outputs
constexpr int K2_MINB = K2_MINB_V;

template <typename SrcT> __device__ __forceinline__ SrcT border_cast(double b) { return (SrcT)b; }

// per-pixel interpolation weights, computed once and applied to every frame of the launch
template <typename SrcT> struct Weights {        // float32 / uint16 / float64 images: OpenCV's float32 table entry
    float w00, w01, w10, w11;
    __device__ __forceinline__ void set(const FixedCoord& c) { bilinear_weights(c.fx, c.fy, w00, w01, w10, w11); }
};
template <> struct Weights<uint8_t> {            // uint8 images: int16 fixed-point weights
    int fx, fy;
    __device__ __forceinline__ void set(const FixedCoord& c) { fx = c.fx; fy = c.fy; }
};

template <typename SrcT, typename DstT> struct Blend;
template <> struct Blend<float, float> {
    static __device__ __forceinline__ float run(float a, float b, float c, float d, const Weights<float>& w) {
        return blend_f32(a, b, c, d, w.w00, w.w01, w.w10, w.w11);
    }
};
template <> struct Blend<float, double> {
    static __device__ __forceinline__ double run(float a, float b, float c, float d, const Weights<float>& w) {
        return (double)blend_f32(a, b, c, d, w.w00, w.w01, w.w10, w.w11);
    }
};
template <> struct Blend<double, double> {
    static __device__ __forceinline__ double run(double a, double b, double c, double d, const Weights<double>& w) {
        return blend_f64(a, b, c, d, w.w00, w.w01, w.w10, w.w11);
    }
};
template <> struct Blend<uint16_t, uint16_t> {
    static __device__ __forceinline__ uint16_t run(uint16_t a, uint16_t b, uint16_t c, uint16_t d, const Weights<uint16_t>& w) {
        return sat_u16(blend_f32((float)a, (float)b, (float)c, (float)d, w.w00, w.w01, w.w10, w.w11));
    }
};
template <> struct Blend<uint8_t, uint8_t> {
    static __device__ __forceinline__ uint8_t run(uint8_t a, uint8_t b, uint8_t c, uint8_t d, const Weights<uint8_t>& w) {
        return (uint8_t)blend_u8(a, b, c, d, w.fx, w.fy);
    }
};

// rim of the image: any neighbour outside [0,W)x[0,H) is the border value; a window entirely outside is the border
// value itself (OpenCV does not blend it).  Rare -> kept out of line so the hot path stays small.
template <typename SrcT, typename DstT>
__device__ __noinline__ DstT remap_rim(const SrcT* __restrict__ src, int H, int W, int ix, int iy, Weights<SrcT> w, SrcT bval) {
    if (ix >= W || ix + 1 < 0 || iy >= H || iy + 1 < 0) return (DstT)bval;
    const bool x0 = (unsigned)ix < (unsigned)W, x1 = (unsigned)(ix + 1) < (unsigned)W;
    const bool y0 = (unsigned)iy < (unsigned)H, y1 = (unsigned)(iy + 1) < (unsigned)H;
    const SrcT* p = src + (ptrdiff_t)iy * W + ix;
    const SrcT v00 = (x0 && y0) ? __ldg(p) : bval;
    const SrcT v01 = (x1 && y0) ? __ldg(p + 1) : bval;
    const SrcT v10 = (x0 && y1) ? __ldg(p + W) : bval;
    const SrcT v11 = (x1 && y1) ? __ldg(p + W + 1) : bval;
    return Blend<SrcT, DstT>::run(v00, v01, v10, v11, w);
}

// MODE: 0 explicit maps, 1 analytic general P, 2 analytic separable P^-1 (x = x(u), y = y(v))
//
// The kernel is bound by the latency of the four-neighbour gathers, not by the float64 map evaluation (with explicit
// maps it is no faster, see profiles/).  So a thread first computes the coordinates of all its K2_ROWS pixels, then
// issues all 4*K2_ROWS gathers unconditionally (pixels on the rim gather from offset 0 and are redone out of line),
// then blends and stores: 4*K2_ROWS loads in flight per thread instead of 4.
template <typename SrcT, typename DstT, int MODE>
__global__ void __launch_bounds__(K2_BX * K2_BY, K2_MINB) k2_remap_kernel(K2Args a) {
    const int ox = blockIdx.x * K2_BX + (threadIdx.x % K2_BX);
    const int ty = threadIdx.x / K2_BX;
    if (ox >= a.ow) return;
    const int u = ox + a.x0;
    const SrcT bval = border_cast<SrcT>(a.border);
    const int H = a.H, W = a.W, nf = a.n_frames, ow = a.ow, oh = a.oh;
    const int src_stride = H * W, dst_stride = oh * ow;      // elements; frames are < 2^31 pixels (32767^2)
    // lens constants live in registers for the whole thread.  They are read from a small device buffer, not from
    // the kernel parameters: ptxas re-materialises parameter loads at every use (one LDC per use, ~16 issue
    // slots per pixel); a global load it has to keep.
    LensConst L = a.lens;
    if (MODE == 2) {
        const double2* lp = reinterpret_cast<const double2*>(a.lens_dev);
        double2 q;
        q = __ldg(lp + 0); L.k1 = q.x; L.k2 = q.y;
        q = __ldg(lp + 1); L.k3 = q.x; L.p1 = q.y;
        q = __ldg(lp + 2); L.p2 = q.x; L.p1x2 = q.y;
        q = __ldg(lp + 3); L.p2x2 = q.x; L.fx = q.y;
        q = __ldg(lp + 4); L.fy = q.x; L.cx = q.y;
        q = __ldg(lp + 5); L.cy = q.x; L.ir[0] = q.y;
        q = __ldg(lp + 6); L.ir[2] = q.x; L.ir[4] = q.y;
        q = __ldg(lp + 7); L.ir[5] = q.x;
    }
    double xc = 0.0, xc2 = 0.0;
    if (MODE == 2) {
        xc = fma((double)u, L.ir[0], L.ir[2]);
        xc2 = dmul(xc, xc);
    }
    const int oy0 = blockIdx.y * (K2_BY * K2_ROWS) + ty;

    int off[K2_ROWS];            // element offset of the top-left neighbour; 0 for rim pixels (done in the second pass)
    int cix[K2_ROWS], ciy[K2_ROWS];
    Weights<SrcT> wt[K2_ROWS];
    unsigned rim = 0, live = 0;  // bit j: pixel j lies on the rim / row j exists
#pragma unroll
    for (int j = 0; j < K2_ROWS; ++j) {
        const int oy = oy0 + j * K2_BY;
        const int v = (oy < oh ? oy : oh - 1) + a.y0;
        float mx, my;
        if (MODE == 2) {
            const double y = fma((double)v, L.ir[4], L.ir[5]);
            map_distort(L, xc, y, xc2, dmul(y, y), mx, my);
        } else if (MODE == 1) {
            undistort_map(L, u, v, mx, my);
        } else {
            mx = __ldg(a.mapx + v * W + u);
            my = __ldg(a.mapy + v * W + u);
        }
        const FixedCoord c = fixed_coord(mx, my);
        const bool inner = (unsigned)c.ix < (unsigned)(W - 1) && (unsigned)c.iy < (unsigned)(H - 1);
        off[j] = inner ? c.iy * W + c.ix : 0;
        cix[j] = c.ix; ciy[j] = c.iy;
        wt[j].set(c);
        if (oy < oh) { live |= 1u << j; if (!inner) rim |= 1u << j; }
    }

    const bool tiny = W < 2 || H < 2;                     // no interior 2x2 window exists: offset 0 + 1 + W would overrun
    if (!tiny) {
        const SrcT* src = (const SrcT*)a.src;
        DstT* dst = (DstT*)a.dst + (oy0 * ow + ox);
        const unsigned fast = live & ~rim;
#pragma unroll 1
        for (int f = 0; f < nf; ++f) {
            SrcT v00[K2_ROWS], v01[K2_ROWS], v10[K2_ROWS], v11[K2_ROWS];
#pragma unroll
            for (int j = 0; j < K2_ROWS; ++j) {
                const SrcT* p = src + off[j];
                v00[j] = __ldg(p); v01[j] = __ldg(p + 1); v10[j] = __ldg(p + W); v11[j] = __ldg(p + W + 1);
            }
#pragma unroll
            for (int j = 0; j < K2_ROWS; ++j)
                if (fast & (1u << j)) dst[j * K2_BY * ow] = Blend<SrcT, DstT>::run(v00[j], v01[j], v10[j], v11[j], wt[j]);
            src += src_stride;
            dst += dst_stride;
        }
    } else {
        rim = live;
    }
    // second pass, rare: pixels whose 2x2 window touches or leaves the frame
    if (rim) {
        const SrcT* src = (const SrcT*)a.src;
        DstT* dst = (DstT*)a.dst + (oy0 * ow + ox);
        for (int f = 0; f < nf; ++f) {
#pragma unroll
            for (int j = 0; j < K2_ROWS; ++j)
                if (rim & (1u << j)) dst[j * K2_BY * ow] = remap_rim<SrcT, DstT>(src, H, W, cix[j], ciy[j], wt[j], bval);
            src += src_stride;
            dst += dst_stride;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// Tiled variant (float32 sources — the chain): the source window of a 64x16 output tile is staged in shared memory
// by ONE TMA box per frame, and the four neighbours are gathered from there.  The L1 path above spends ~8 L1
// wavefronts per 32 pixels on the gather (unaligned, two rows) and is bound by them; shared memory serves the same
// gather in ~4-5 conflict-free wavefronts, TMA moves whole lines, and the frames of a launch are double-buffered so
// the box of frame f+1 lands while frame f is blended.
//   * coordinates / weights / shared offsets of the tile's pixels are computed once (4 pixels per thread);
//   * the exact bounding box of the tile's source window comes from a block-wide min/max (REDUX + shared atomics);
//     box origin = (min ix rounded down to 4 columns — TMA's 16-byte rule —, min iy).  If the window does not fit the
//     fixed 80x32 box (very strong distortion) the tile falls back to global gathers: correct for any map;
//   * rim pixels (window touching the frame border) are redone from global memory with per-neighbour border handling.
#ifndef KT_TW_V
#define KT_TW_V 64
#define KT_TH_V 16
#define KT_BW_V 80
#define KT_BH_V 32
#endif
constexpr int KT_TW = KT_TW_V, KT_TH = KT_TH_V, KT_THREADS = 256, KT_PX = KT_TW * KT_TH / KT_THREADS;
constexpr int KT_BW = KT_BW_V, KT_BH = KT_BH_V;
// box width per source type: 16-byte origin granularity costs up to 3 / 7 / 15 extra columns (float32 / uint16 / uint8)
template <typename SrcT> struct KtBox { static constexpr int BW = sizeof(SrcT) == 1 ? KT_BW + 16 : KT_BW, GRAN = 16 / (int)sizeof(SrcT), BYTES = BW * KT_BH * (int)sizeof(SrcT); };
#ifndef KT_NBUF_V
#define KT_NBUF_V 4
#endif
#ifndef KT_MINB_V
#define KT_MINB_V 4
#endif
constexpr int KT_NBUF = KT_NBUF_V;                       // frames of a launch in flight per tile
template <typename SrcT> constexpr int kt_smem() { return KT_NBUF * KtBox<SrcT>::BYTES + 128; }


// predicated global store: a plain `if (flag) *p = v` lets ptxas wrap the whole gather of that pixel in a branch
__device__ __forceinline__ void st_if(float* p, float v, unsigned flag) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q st.global.f32 [%0], %1;\n\t}" ::"l"(p), "f"(v), "r"(flag) : "memory");
}
__device__ __forceinline__ void st_if(double* p, double v, unsigned flag) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q st.global.f64 [%0], %1;\n\t}" ::"l"(p), "d"(v), "r"(flag) : "memory");
}
__device__ __forceinline__ void st_if(uint16_t* p, uint16_t v, unsigned flag) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q st.global.u16 [%0], %1;\n\t}" ::"l"(p), "h"(v), "r"(flag) : "memory");
}
__device__ __forceinline__ void st_if(uint8_t* p, uint8_t v, unsigned flag) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q st.global.u8 [%0], %1;\n\t}" ::"l"(p), "r"((unsigned)v), "r"(flag) : "memory");
}

template <typename SrcT, typename DstT, int MODE>
__global__ void __launch_bounds__(KT_THREADS, KT_MINB_V) k2_tiled_kernel(const __grid_constant__ CUtensorMap tm_src, K2Args a) {
    constexpr int BW = KtBox<SrcT>::BW, KT_BOX_BYTES = KtBox<SrcT>::BYTES;
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* full = (uint64_t*)(smem + KT_NBUF * KT_BOX_BYTES);
    int* red = (int*)(smem + KT_NBUF * KT_BOX_BYTES + 64);     // min ix, max ix, min iy, max iy
    const int tid = threadIdx.x;
    const int c = tid % KT_TW, r0 = tid / KT_TW;               // rows r0 + 4j
    const int ox = blockIdx.x * KT_TW + c;
    const int oy0 = blockIdx.y * KT_TH + r0;
    const int H = a.H, W = a.W, nf = a.n_frames, ow = a.ow, oh = a.oh;
    const int u = (ox < ow ? ox : ow - 1) + a.x0;
    const SrcT bval = border_cast<SrcT>(a.border);

    if (tid == 0) {
        for (int b = 0; b < KT_NBUF; ++b) mbar_init(&full[b], 1);
        red[0] = 0x7fffffff; red[1] = -1; red[2] = 0x7fffffff; red[3] = -1;
        mbar_init_fence();
    }
    LensConst L = a.lens;
    if (MODE == 2) {
        const double2* lp = reinterpret_cast<const double2*>(a.lens_dev);
        double2 q;
        q = __ldg(lp + 0); L.k1 = q.x; L.k2 = q.y;
        q = __ldg(lp + 1); L.k3 = q.x; L.p1 = q.y;
        q = __ldg(lp + 2); L.p2 = q.x; L.p1x2 = q.y;
        q = __ldg(lp + 3); L.p2x2 = q.x; L.fx = q.y;
        q = __ldg(lp + 4); L.fy = q.x; L.cx = q.y;
        q = __ldg(lp + 5); L.cy = q.x; L.ir[0] = q.y;
        q = __ldg(lp + 6); L.ir[2] = q.x; L.ir[4] = q.y;
        q = __ldg(lp + 7); L.ir[5] = q.x;
    }
    double xc = 0.0, xc2 = 0.0;
    if (MODE == 2) {
        xc = fma((double)u, L.ir[0], L.ir[2]);
        xc2 = dmul(xc, xc);
    }
    int cix[KT_PX], ciy[KT_PX];
    Weights<SrcT> wt[KT_PX];
    unsigned rim = 0, fast = 0;
    int mnx = 0x7fffffff, mxx = -1, mny = 0x7fffffff, mxy = -1;
#pragma unroll
    for (int j = 0; j < KT_PX; ++j) {
        const int oy = oy0 + j * (KT_THREADS / KT_TW);
        const int v = (oy < oh ? oy : oh - 1) + a.y0;
        float mx, my;
        if (MODE == 2) {
            const double y = fma((double)v, L.ir[4], L.ir[5]);
            map_distort(L, xc, y, xc2, dmul(y, y), mx, my);
        } else if (MODE == 1) {
            undistort_map(L, u, v, mx, my);
        } else {
            mx = __ldg(a.mapx + v * W + u);
            my = __ldg(a.mapy + v * W + u);
        }
        const FixedCoord fc = fixed_coord(mx, my);
        cix[j] = fc.ix; ciy[j] = fc.iy;
        wt[j].set(fc);
        const bool live = ox < ow && oy < oh;
        const bool inner = (unsigned)fc.ix < (unsigned)(W - 1) && (unsigned)fc.iy < (unsigned)(H - 1);
        if (live && inner) {
            fast |= 1u << j;
            mnx = min(mnx, fc.ix); mxx = max(mxx, fc.ix); mny = min(mny, fc.iy); mxy = max(mxy, fc.iy);
        } else if (live) {
            rim |= 1u << j;
        }
    }
    __syncthreads();                                   // red[] and the barriers are initialised
    mnx = __reduce_min_sync(0xffffffffu, mnx); mxx = __reduce_max_sync(0xffffffffu, mxx);
    mny = __reduce_min_sync(0xffffffffu, mny); mxy = __reduce_max_sync(0xffffffffu, mxy);
    if ((tid & 31) == 0) { atomicMin(&red[0], mnx); atomicMax(&red[1], mxx); atomicMin(&red[2], mny); atomicMax(&red[3], mxy); }
    __syncthreads();
    const int bx = red[0] & ~(KtBox<SrcT>::GRAN - 1), by = red[2];
    const bool any = red[1] >= 0;
    const bool fits = any && (red[1] + 1 - bx) < BW && (red[3] + 1 - by) < KT_BH;

    DstT* dst = (DstT*)a.dst + (oy0 * ow + ox);
    const int dstep = (KT_THREADS / KT_TW) * ow;
    const int src_stride = H * W, dst_stride = oh * ow;
    if (fits) {
        int so[KT_PX];
#pragma unroll
        for (int j = 0; j < KT_PX; ++j) {
            so[j] = (fast & (1u << j)) ? (ciy[j] - by) * BW + (cix[j] - bx) : 0;
            asm volatile("" : "+r"(so[j]));            // keep the offset: ptxas otherwise recomputes it from cix / ciy every frame
        }
        auto issue = [&](int f) {
            uint64_t* bar = &full[f % KT_NBUF];
            mbar_expect_tx(bar, KT_BOX_BYTES);
            tma_load_3d(smem + (f % KT_NBUF) * KT_BOX_BYTES, &tm_src, bar, bx, by, f);
        };
        if (tid == 0) for (int f = 0; f < KT_NBUF && f < nf; ++f) issue(f);
#pragma unroll 1
        for (int f = 0; f < nf; ++f) {
            mbar_wait(&full[f % KT_NBUF], (f / KT_NBUF) & 1);
            const SrcT* box = (const SrcT*)(smem + (f % KT_NBUF) * KT_BOX_BYTES);
#pragma unroll
            for (int j = 0; j < KT_PX; ++j) {
                const SrcT* p = box + so[j];
                const DstT r = Blend<SrcT, DstT>::run(p[0], p[1], p[BW], p[BW + 1], wt[j]);
                st_if(dst + j * dstep, r, fast & (1u << j));            // predicated store, no branch around the gather
            }
            dst += dst_stride;
            __syncthreads();                           // everyone is done with this buffer
            if (tid == 0 && f + KT_NBUF < nf) issue(f + KT_NBUF);
        }
    } else if (any) {
        // the source window of this tile does not fit the staged box: gather from global memory
        const SrcT* src = (const SrcT*)a.src;
        for (int f = 0; f < nf; ++f) {
#pragma unroll
            for (int j = 0; j < KT_PX; ++j)
                if (fast & (1u << j)) {
                    const SrcT* p = src + (ciy[j] * W + cix[j]);
                    dst[j * dstep] = Blend<SrcT, DstT>::run(__ldg(p), __ldg(p + 1), __ldg(p + W), __ldg(p + W + 1), wt[j]);
                }
            src += src_stride;
            dst += dst_stride;
        }
    }
    if (rim) {
        const SrcT* src = (const SrcT*)a.src;
        DstT* d2 = (DstT*)a.dst + (oy0 * ow + ox);
        for (int f = 0; f < nf; ++f) {
#pragma unroll
            for (int j = 0; j < KT_PX; ++j)
                if (rim & (1u << j)) d2[j * dstep] = remap_rim<SrcT, DstT>(src, H, W, cix[j], ciy[j], wt[j], bval);
            src += src_stride;
            d2 += dst_stride;
        }
    }
}

static bool k2_tiled_eligible(const K2Args& a, int src_dtype, int dst_dtype) {
    const bool pair = (src_dtype == DT_F32 && (dst_dtype == DT_F32 || dst_dtype == DT_F64)) || (src_dtype == DT_U16 && dst_dtype == DT_U16) ||
                      (src_dtype == DT_U8 && dst_dtype == DT_U8);
    if (!pair) return false;
    if (a.W < 2 || a.H < 2) return false;
    if (((size_t)a.W * dtype_size(src_dtype)) % 16 || ((uintptr_t)a.src) % 16) return false;
    return tensor_map_encoder() != nullptr;
}

template <typename SrcT, typename DstT>
static cudaError_t launch_tiled_t(const K2Args& a, cudaStream_t st) {
    CUtensorMap tm;
    const CUtensorMapDataType dt = sizeof(SrcT) == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : sizeof(SrcT) == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT16 : CU_TENSOR_MAP_DATA_TYPE_UINT8;
    if (!make_tensor_map(&tm, dt, sizeof(SrcT), a.src, a.W, a.H, a.n_frames, KtBox<SrcT>::BW, KT_BH)) return cudaErrorInvalidValue;
    dim3 grid((a.ow + KT_TW - 1) / KT_TW, (a.oh + KT_TH - 1) / KT_TH);
    constexpr int SMEM = kt_smem<SrcT>();
    if (a.mapx) k2_tiled_kernel<SrcT, DstT, 0><<<grid, KT_THREADS, SMEM, st>>>(tm, a);
    else if (a.lens.affine && a.lens.ir[1] == 0.0 && a.lens.ir[3] == 0.0) k2_tiled_kernel<SrcT, DstT, 2><<<grid, KT_THREADS, SMEM, st>>>(tm, a);
    else k2_tiled_kernel<SrcT, DstT, 1><<<grid, KT_THREADS, SMEM, st>>>(tm, a);
    return cudaGetLastError();
}

__global__ void __launch_bounds__(256) k2_write_maps_kernel(LensConst lens, float* mapx, float* mapy, int H, int W) {
    const int u = blockIdx.x * 32 + (threadIdx.x % 32);
    const int v = blockIdx.y * 8 + (threadIdx.x / 32);
    if (u >= W || v >= H) return;
    float mx, my;
    undistort_map(lens, u, v, mx, my);
    mapx[(size_t)v * W + u] = mx;
    mapy[(size_t)v * W + u] = my;
}

static bool separable(const LensConst& L) { return L.affine && L.ir[1] == 0.0 && L.ir[3] == 0.0; }

template <typename SrcT, typename DstT>
static cudaError_t launch_t(const K2Args& a, cudaStream_t st) {
    dim3 grid((a.ow + K2_BX - 1) / K2_BX, (a.oh + K2_BY * K2_ROWS - 1) / (K2_BY * K2_ROWS));
    if (a.mapx) k2_remap_kernel<SrcT, DstT, 0><<<grid, K2_BX * K2_BY, 0, st>>>(a);
    else if (separable(a.lens)) k2_remap_kernel<SrcT, DstT, 2><<<grid, K2_BX * K2_BY, 0, st>>>(a);
    else k2_remap_kernel<SrcT, DstT, 1><<<grid, K2_BX * K2_BY, 0, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_k2(const K2Args& a, int src_dtype, int dst_dtype, int variant, cudaStream_t st, int* launches) {
    if (a.n_frames <= 0 || a.ow <= 0 || a.oh <= 0) return cudaSuccess;
    if (a.H > 32767 || a.W > 32767) return cudaErrorInvalidValue;      // OpenCV's remap itself is limited to short coordinates
    if (a.x0 < 0 || a.y0 < 0 || a.x0 + a.ow > a.W || a.y0 + a.oh > a.H) return cudaErrorInvalidValue;
    if ((a.mapx == nullptr) != (a.mapy == nullptr)) return cudaErrorInvalidValue;
    if (!a.mapx && !a.lens_dev) return cudaErrorInvalidValue;
    // variant: 0 auto, 1 gathers through L1, 2 shared-memory staged tiles (float32 sources)
    // integer sources: a single frame is faster through L1 (63 vs 69 us at 4096x3000 uint16), batches through the tiles (26 vs 31)
    const bool want_tiles = variant == 2 || (variant == 0 && (src_dtype == DT_F32 || a.n_frames >= 4));
    const bool tiled = want_tiles && k2_tiled_eligible(a, src_dtype, dst_dtype);
    if (variant == 2 && !tiled) return cudaErrorNotSupported;
    if (launches) ++*launches;
    if (tiled) {
        if (src_dtype == DT_U16) return launch_tiled_t<uint16_t, uint16_t>(a, st);
        if (src_dtype == DT_U8) return launch_tiled_t<uint8_t, uint8_t>(a, st);
        return dst_dtype == DT_F32 ? launch_tiled_t<float, float>(a, st) : launch_tiled_t<float, double>(a, st);
    }
    if (src_dtype == DT_F32 && dst_dtype == DT_F32) return launch_t<float, float>(a, st);
    if (src_dtype == DT_F32 && dst_dtype == DT_F64) return launch_t<float, double>(a, st);
    if (src_dtype == DT_F64 && dst_dtype == DT_F64) return launch_t<double, double>(a, st);
    if (src_dtype == DT_U16 && dst_dtype == DT_U16) return launch_t<uint16_t, uint16_t>(a, st);
    if (src_dtype == DT_U8 && dst_dtype == DT_U8) return launch_t<uint8_t, uint8_t>(a, st);
    if (launches) --*launches;
    return cudaErrorInvalidValue;
}

cudaError_t launch_write_maps(const LensConst& lens, float* mapx, float* mapy, int H, int W, cudaStream_t st,
                              int* launches) {
    if (H <= 0 || W <= 0) return cudaSuccess;
    dim3 grid((W + 31) / 32, (H + 7) / 8);
    k2_write_maps_kernel<<<grid, 256, 0, st>>>(lens, mapx, mapy, H, W);
    if (launches) ++*launches;
    return cudaGetLastError();
}

}  // namespace imgcorr

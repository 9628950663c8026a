// K1 — fused dark-current subtraction / flat-field division / nan_to_num / NxN median-threshold.
//
// Replaces, for one frame, the reference's
//   image -= bg                              camera/CameraCalibration.py:502
//   image[flat != 0] /= flat[flat != 0]      camera/CameraCalibration.py:525-526
//   image = np.nan_to_num(image)             camera/CameraCalibration.py:561
//   medianThreshold(image, thr, size)        filters/medianThreshold.py:7-30
// with one pass over HBM: raw (2 or 4 B/px) + dark (4) + flat (4) in, corrected float32 (4) out.
//
// Two staging variants share the same compute phases:
//   generic : coalesced ld.global.nc with reflect indexing (any W, any alignment)
//   tma     : persistent CTAs, cp.async.bulk.tensor tiles (+halo) of raw/dark/flat into a
//             multi-stage shared-memory ring signalled by mbarriers; out-of-bounds halo
//             (TMA zero fill) is replaced by the mirrored in-tile sample (scipy 'reflect').
// Tile = TW x TH outputs; phase A computes the pointwise value of every pixel of the
// tile + halo once into shared memory; phase B walks each thread down a column, sorting one
// horizontal triple per row (shared by three vertically adjacent windows) and combining three
// sorted triples into the median of 9.
#include <cuda.h>
#include "imgcorr_kernels.cuh"

namespace imgcorr {

constexpr int K1_THREADS = 256;
constexpr int K1_TW = 128;          // output columns per tile
constexpr int K1_BOXW = 136;        // staged columns: tx0-4 .. tx0+131 (16-byte multiple for u16 and f32)
constexpr int K1_XOFF = 4;          // column of tx0 inside the staged box

template <typename T> struct RawIO;
template <> struct RawIO<uint8_t>  { using CT = float;  static __device__ __forceinline__ double ld(uint8_t v)  { return (double)(int)v; } };
template <> struct RawIO<uint16_t> { using CT = float;  static __device__ __forceinline__ double ld(uint16_t v) { return (double)(int)v; } };
template <> struct RawIO<float>    { using CT = float;  static __device__ __forceinline__ double ld(float v)    { return (double)v; } };
template <> struct RawIO<double>   { using CT = double; static __device__ __forceinline__ double ld(double v)   { return v; } };

template <typename OutT, typename CT> __device__ __forceinline__ OutT to_out(CT v);
template <> __device__ __forceinline__ float    to_out<float, float>(float v)       { return v; }
template <> __device__ __forceinline__ double   to_out<double, float>(float v)      { return (double)v; }
template <> __device__ __forceinline__ double   to_out<double, double>(double v)    { return v; }
template <> __device__ __forceinline__ uint16_t to_out<uint16_t, float>(float v)    { return sat_u16(v); }
template <> __device__ __forceinline__ uint8_t  to_out<uint8_t, float>(float v)     { return sat_u8(v); }

template <typename T> __device__ __forceinline__ T ldg_stream(const T* p) { return __ldg(p); }

// ------------------------------------------------------------------------------------------
// pointwise only (ksize == 0): pure streaming kernel, 4 pixels per thread
// ------------------------------------------------------------------------------------------
template <typename RawT, typename OutT>
__global__ void __launch_bounds__(256) k1_pointwise_kernel(K1Args a) {
    using CT = typename RawIO<RawT>::CT;
    const size_t npx = (size_t)a.H * a.W;
    const size_t total = npx * a.n_frames;
    const RawT* raw = (const RawT*)a.raw;
    OutT* out = (OutT*)a.out;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        size_t p = i % npx;
        float d = a.dark ? __ldg(a.dark + p) : 0.0f;
        float s = a.ascent ? __ldg(a.ascent + p) : 0.0f;
        float f = a.flat ? __ldg(a.flat + p) : 0.0f;
        CT x = pointwise<CT>(a.pw, RawIO<RawT>::ld(ldg_stream(raw + i)), d, s, f);
        out[i] = to_out<OutT, CT>(x);
    }
}

// ------------------------------------------------------------------------------------------
// phase B: median-threshold of one tile from the shared pointwise tile xs[TH+2h][BOXW]
// ------------------------------------------------------------------------------------------
template <typename CT, typename OutT, int KS, int TH>
__device__ __forceinline__ void median_phase(const CT* __restrict__ xs, const K1Args& a, int frame, int tx0, int ty0) {
    constexpr int HALO = KS / 2;
    constexpr int ROWS_PER_THREAD = TH / (K1_THREADS / K1_TW);
    const int c = threadIdx.x % K1_TW;                 // output column inside the tile
    const int rg = threadIdx.x / K1_TW;                // row group
    const int gx = tx0 + c;
    const int r0 = rg * ROWS_PER_THREAD;               // first output row inside the tile
    if (gx >= a.W) return;
    OutT* out = (OutT*)a.out + ((size_t)frame * a.H) * a.W + gx;
    uint8_t* mask = a.mask ? a.mask + ((size_t)frame * a.H) * a.W + gx : nullptr;
    // column of the window centre inside the staged box
    const CT* col = xs + (K1_XOFF + c);

    if (KS == 3) {
        // rows of xs: tile row r  <->  xs row r + HALO
        const CT* p = col + (size_t)r0 * K1_BOXW;
        Sorted3<CT> s0 = sort3(p[-1], p[0], p[1]);
        p += K1_BOXW;
        CT centre = p[0];
        Sorted3<CT> s1 = sort3(p[-1], centre, p[1]);
#pragma unroll 4
        for (int j = 0; j < ROWS_PER_THREAD; ++j) {
            p += K1_BOXW;
            CT nxt = p[0];
            Sorted3<CT> s2 = sort3(p[-1], nxt, p[1]);
            const int gy = ty0 + r0 + j;
            if (gy < a.H) {
                CT med = median9(s0, s1, s2);
                bool rep = predicate(centre, med, a.pred);
                out[(size_t)gy * a.W] = to_out<OutT, CT>(rep ? med : centre);
                if (mask) mask[(size_t)gy * a.W] = rep ? 1 : 0;
            }
            s0 = s1; s1 = s2; centre = nxt;
        }
    } else {   // KS == 5
        for (int j = 0; j < ROWS_PER_THREAD; ++j) {
            const int gy = ty0 + r0 + j;
            if (gy >= a.H) break;
            const CT* p = col + (size_t)(r0 + j) * K1_BOXW;
            CT w[25];
#pragma unroll
            for (int dy = 0; dy < 5; ++dy)
#pragma unroll
                for (int dx = 0; dx < 5; ++dx) w[dy * 5 + dx] = p[dy * K1_BOXW + dx - 2];
            CT centre = w[12];
            CT med = median25(w);
            bool rep = predicate(centre, med, a.pred);
            out[(size_t)gy * a.W] = to_out<OutT, CT>(rep ? med : centre);
            if (mask) mask[(size_t)gy * a.W] = rep ? 1 : 0;
        }
    }
}

// ------------------------------------------------------------------------------------------
// generic variant: one tile per CTA, reflect indexing on global loads
// ------------------------------------------------------------------------------------------
template <typename RawT, typename OutT, int KS, int TH>
__global__ void __launch_bounds__(K1_THREADS) k1_generic_kernel(K1Args a, int tiles_x, int tiles_y) {
    using CT = typename RawIO<RawT>::CT;
    constexpr int HALO = KS / 2;
    constexpr int LH = TH + 2 * HALO;
    constexpr int LW = K1_TW + 2 * HALO;
    __shared__ CT xs[LH * K1_BOXW];

    int t = blockIdx.x;
    const int txi = t % tiles_x; t /= tiles_x;
    const int tyi = t % tiles_y;
    const int frame = t / tiles_y;
    const int tx0 = txi * K1_TW, ty0 = tyi * TH;
    const RawT* raw = (const RawT*)a.raw + (size_t)frame * a.H * a.W;

    for (int idx = threadIdx.x; idx < LH * LW; idx += K1_THREADS) {
        const int ly = idx / LW, lx = idx - ly * LW;
        const int gy = reflect_index(ty0 - HALO + ly, a.H);
        const int gx = reflect_index(tx0 - HALO + lx, a.W);
        const size_t g = (size_t)gy * a.W + gx;
        float d = a.dark ? __ldg(a.dark + g) : 0.0f;
        float s = a.ascent ? __ldg(a.ascent + g) : 0.0f;
        float f = a.flat ? __ldg(a.flat + g) : 0.0f;
        xs[ly * K1_BOXW + (K1_XOFF - HALO) + lx] = pointwise<CT>(a.pw, RawIO<RawT>::ld(ldg_stream(raw + g)), d, s, f);
    }
    __syncthreads();
    median_phase<CT, OutT, KS, TH>(xs, a, frame, tx0, ty0);
}

// ------------------------------------------------------------------------------------------
// TMA variant
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, uint64_t* bar, int x, int y) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* tm, uint64_t* bar, int x, int y, int z) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z) : "memory");
}

template <typename RawT, int KS, int TH, int NSTAGE>
struct K1TmaSmem {
    static constexpr int HALO = KS / 2;
    static constexpr int LH = TH + 2 * HALO;
    static constexpr int BOX = LH * K1_BOXW;
    static constexpr size_t raw_bytes = (size_t)BOX * sizeof(RawT);
    static constexpr size_t map_bytes = (size_t)BOX * sizeof(float);
    static constexpr size_t align128(size_t v) { return (v + 127) & ~(size_t)127; }
    static constexpr size_t stage_bytes = align128(raw_bytes) + 2 * align128(map_bytes);
    static constexpr size_t xs_off = NSTAGE * stage_bytes;
    static constexpr size_t bar_off = xs_off + align128(map_bytes);
    static constexpr size_t total = bar_off + 128;
};

template <typename RawT, typename OutT, int KS, int TH, int NSTAGE>
__global__ void __launch_bounds__(K1_THREADS)
k1_tma_kernel(const __grid_constant__ CUtensorMap tm_raw, const __grid_constant__ CUtensorMap tm_dark,
              const __grid_constant__ CUtensorMap tm_flat, K1Args a, int tiles_x, int tiles_y, int total_tiles) {
    using CT = float;
    using S = K1TmaSmem<RawT, KS, TH, NSTAGE>;
    constexpr int HALO = S::HALO;
    constexpr int LH = S::LH;
    constexpr int LW = K1_TW + 2 * HALO;
    extern __shared__ __align__(128) uint8_t smem[];
    CT* xs = (CT*)(smem + S::xs_off);
    uint64_t* full = (uint64_t*)(smem + S::bar_off);

    const bool has_dark = a.dark != nullptr, has_flat = a.flat != nullptr;
    const uint32_t tx_bytes = (uint32_t)(S::raw_bytes + (has_dark ? S::map_bytes : 0) + (has_flat ? S::map_bytes : 0));

    auto issue = [&](int stage, int tile) {
        int t = tile;
        const int txi = t % tiles_x; t /= tiles_x;
        const int tyi = t % tiles_y;
        const int frame = t / tiles_y;
        const int x = txi * K1_TW - K1_XOFF, y = tyi * TH - HALO;
        uint8_t* base = smem + (size_t)stage * S::stage_bytes;
        mbar_expect_tx(&full[stage], tx_bytes);
        tma_load_3d(base, &tm_raw, &full[stage], x, y, frame);
        if (has_dark) tma_load_2d(base + S::align128(S::raw_bytes), &tm_dark, &full[stage], x, y);
        if (has_flat) tma_load_2d(base + S::align128(S::raw_bytes) + S::align128(S::map_bytes), &tm_flat, &full[stage], x, y);
    };

    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTAGE; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTAGE; ++s) {
            int tile = blockIdx.x + s * gridDim.x;
            if (tile < total_tiles) issue(s, tile);
        }
    }

    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int stage = it % NSTAGE;
        const uint32_t parity = (it / NSTAGE) & 1;
        int t = tile;
        const int txi = t % tiles_x; t /= tiles_x;
        const int tyi = t % tiles_y;
        const int frame = t / tiles_y;
        const int tx0 = txi * K1_TW, ty0 = tyi * TH;

        mbar_wait(&full[stage], parity);
        const uint8_t* base = smem + (size_t)stage * S::stage_bytes;
        const RawT* sraw = (const RawT*)base;
        const float* sdark = (const float*)(base + S::align128(S::raw_bytes));
        const float* sflat = (const float*)(base + S::align128(S::raw_bytes) + S::align128(S::map_bytes));

        // phase A: pointwise value of every staged pixel the windows need
        for (int idx = threadIdx.x; idx < LH * LW; idx += K1_THREADS) {
            const int ly = idx / LW, lx = idx - ly * LW;
            const int o = ly * K1_BOXW + (K1_XOFF - HALO) + lx;
            float d = has_dark ? sdark[o] : 0.0f;
            float f = has_flat ? sflat[o] : 0.0f;
            xs[o] = pointwise<CT>(a.pw, RawIO<RawT>::ld(sraw[o]), d, 0.0f, f);
        }
        __syncthreads();          // xs complete, stage fully consumed
        if (threadIdx.x == 0) {
            const int nxt = tile + NSTAGE * gridDim.x;
            if (nxt < total_tiles) issue(stage, nxt);
        }
        // scipy 'reflect' at the frame border: TMA zero-filled the out-of-frame part of the halo
        const bool border = (ty0 - HALO < 0) | (ty0 + TH + HALO > a.H) | (tx0 - HALO < 0) | (tx0 + K1_TW + HALO > a.W);
        if (border) {
            for (int idx = threadIdx.x; idx < LH * LW; idx += K1_THREADS) {
                const int ly = idx / LW, lx = idx - ly * LW;
                const int gy = ty0 - HALO + ly, gx = tx0 - HALO + lx;
                if ((unsigned)gy >= (unsigned)a.H || (unsigned)gx >= (unsigned)a.W) {
                    const int sy = reflect_index(gy, a.H) - (ty0 - HALO);
                    const int sx = reflect_index(gx, a.W) - (tx0 - HALO);
                    if ((unsigned)sy < (unsigned)LH && (unsigned)sx < (unsigned)LW)
                        xs[ly * K1_BOXW + (K1_XOFF - HALO) + lx] = xs[sy * K1_BOXW + (K1_XOFF - HALO) + sx];
                }
            }
            __syncthreads();
        }
        median_phase<CT, OutT, KS, TH>(xs, a, frame, tx0, ty0);
        __syncthreads();          // xs is rewritten by the next tile's phase A
    }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)p;
    }
    return fn;
}

static bool make_map(CUtensorMap* tm, CUtensorMapDataType dt, size_t esz, const void* ptr, int W, int H, int N,
                     int boxw, int boxh) {
    PFN_encodeTiled enc = get_encode_fn();
    if (!enc) return false;
    cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)(N > 0 ? N : 1)};
    cuuint64_t strides[2] = {(cuuint64_t)W * esz, (cuuint64_t)W * H * esz};
    cuuint32_t box[3] = {(cuuint32_t)boxw, (cuuint32_t)boxh, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    const int rank = N > 0 ? 3 : 2;
    CUresult r = enc(tm, dt, rank, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

bool k1_tma_eligible(const K1Args& a, int raw_dtype, int out_dtype) {
    if (a.ksize != 3 && a.ksize != 5) return false;
    if (raw_dtype != DT_U16 && raw_dtype != DT_F32) return false;
    if (out_dtype != DT_F32 && !(raw_dtype == DT_U16 && out_dtype == DT_U16)) return false;
    if (a.pw.flags & FLAG_DARK_LINEAR) return false;
    const size_t esz = dtype_size(raw_dtype);
    if (((size_t)a.W * esz) % 16 || ((size_t)a.W * 4) % 16) return false;
    if (((size_t)a.H * a.W * esz) % 16) return false;
    if (((uintptr_t)a.raw) % 16) return false;
    if (a.dark && ((uintptr_t)a.dark) % 16) return false;
    if (a.flat && ((uintptr_t)a.flat) % 16) return false;
    return get_encode_fn() != nullptr;
}

template <typename RawT, typename OutT, int KS, int TH, int NSTAGE>
static cudaError_t launch_tma_t(const K1Args& a, CUtensorMapDataType rdt, int sm_count, cudaStream_t st) {
    using S = K1TmaSmem<RawT, KS, TH, NSTAGE>;
    CUtensorMap tr, td, tf;
    const int boxh = S::LH;
    if (!make_map(&tr, rdt, sizeof(RawT), a.raw, a.W, a.H, a.n_frames, K1_BOXW, boxh)) return cudaErrorInvalidValue;
    // dark / flat maps: encode the raw pointer as a placeholder when absent (never dereferenced)
    if (!make_map(&td, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, a.dark ? (const void*)a.dark : a.raw, a.W, a.H, 0, K1_BOXW, boxh) && a.dark)
        return cudaErrorInvalidValue;
    if (!make_map(&tf, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, a.flat ? (const void*)a.flat : a.raw, a.W, a.H, 0, K1_BOXW, boxh) && a.flat)
        return cudaErrorInvalidValue;
    auto kern = k1_tma_kernel<RawT, OutT, KS, TH, NSTAGE>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::total);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    const int tiles_x = (a.W + K1_TW - 1) / K1_TW, tiles_y = (a.H + TH - 1) / TH;
    const long long total = (long long)tiles_x * tiles_y * a.n_frames;
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, K1_THREADS, S::total);
    if (per_sm < 1) per_sm = 1;
    long long grid = (long long)sm_count * per_sm;
    if (grid > total) grid = total;
    kern<<<(unsigned)grid, K1_THREADS, S::total, st>>>(tr, td, tf, a, tiles_x, tiles_y, (int)total);
    return cudaGetLastError();
}

template <typename RawT, typename OutT, int KS, int TH>
static cudaError_t launch_generic_t(const K1Args& a, cudaStream_t st) {
    const int tiles_x = (a.W + K1_TW - 1) / K1_TW, tiles_y = (a.H + TH - 1) / TH;
    const long long total = (long long)tiles_x * tiles_y * a.n_frames;
    k1_generic_kernel<RawT, OutT, KS, TH><<<(unsigned)total, K1_THREADS, 0, st>>>(a, tiles_x, tiles_y);
    return cudaGetLastError();
}

template <typename RawT, typename OutT>
static cudaError_t launch_pointwise_t(const K1Args& a, int sm_count, cudaStream_t st) {
    const size_t total = (size_t)a.H * a.W * a.n_frames;
    size_t blocks = (total + 255) / 256;
    const size_t cap = (size_t)sm_count * 32;
    if (blocks > cap) blocks = cap;
    k1_pointwise_kernel<RawT, OutT><<<(unsigned)blocks, 256, 0, st>>>(a);
    return cudaGetLastError();
}

template <typename RawT, typename OutT>
static cudaError_t dispatch_ks(const K1Args& a, bool tma, CUtensorMapDataType rdt, int sm_count, cudaStream_t st) {
    if (a.ksize == 0) return launch_pointwise_t<RawT, OutT>(a, sm_count, st);
    if (a.ksize == 3) {
        if constexpr ((sizeof(RawT) == 2 && sizeof(OutT) <= 4) || (sizeof(RawT) == 4 && sizeof(OutT) == 4))
            if (tma) {
                if constexpr (sizeof(RawT) == 2) return launch_tma_t<RawT, OutT, 3, 32, 2>(a, rdt, sm_count, st);
                else return launch_tma_t<RawT, OutT, 3, 16, 2>(a, rdt, sm_count, st);
            }
        return launch_generic_t<RawT, OutT, 3, 32>(a, st);
    }
    if (a.ksize == 5) {
        if constexpr ((sizeof(RawT) == 2 && sizeof(OutT) <= 4) || (sizeof(RawT) == 4 && sizeof(OutT) == 4))
            if (tma) {
                if constexpr (sizeof(RawT) == 2) return launch_tma_t<RawT, OutT, 5, 32, 2>(a, rdt, sm_count, st);
                else return launch_tma_t<RawT, OutT, 5, 16, 2>(a, rdt, sm_count, st);
            }
        return launch_generic_t<RawT, OutT, 5, 32>(a, st);
    }
    return cudaErrorInvalidValue;
}

cudaError_t launch_k1(const K1Args& a, int raw_dtype, int out_dtype, int variant, int sm_count, cudaStream_t st,
                      int* launches) {
    if (a.n_frames <= 0 || a.H <= 0 || a.W <= 0) return cudaSuccess;
    bool tma = variant != 1 && k1_tma_eligible(a, raw_dtype, out_dtype);
    if (variant == 2 && !tma) return cudaErrorNotSupported;
    if (launches) ++*launches;
    switch (raw_dtype) {
        case DT_U8:
            if (out_dtype == DT_F32) return dispatch_ks<uint8_t, float>(a, false, CU_TENSOR_MAP_DATA_TYPE_UINT8, sm_count, st);
            if (out_dtype == DT_U8) return dispatch_ks<uint8_t, uint8_t>(a, false, CU_TENSOR_MAP_DATA_TYPE_UINT8, sm_count, st);
            if (out_dtype == DT_F64) return dispatch_ks<uint8_t, double>(a, false, CU_TENSOR_MAP_DATA_TYPE_UINT8, sm_count, st);
            break;
        case DT_U16:
            if (out_dtype == DT_F32) return dispatch_ks<uint16_t, float>(a, tma, CU_TENSOR_MAP_DATA_TYPE_UINT16, sm_count, st);
            if (out_dtype == DT_U16) return dispatch_ks<uint16_t, uint16_t>(a, tma, CU_TENSOR_MAP_DATA_TYPE_UINT16, sm_count, st);
            if (out_dtype == DT_F64) return dispatch_ks<uint16_t, double>(a, false, CU_TENSOR_MAP_DATA_TYPE_UINT16, sm_count, st);
            break;
        case DT_F32:
            if (out_dtype == DT_F32) return dispatch_ks<float, float>(a, tma, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, sm_count, st);
            if (out_dtype == DT_F64) return dispatch_ks<float, double>(a, false, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, sm_count, st);
            break;
        case DT_F64:
            if (out_dtype == DT_F64) return dispatch_ks<double, double>(a, false, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, sm_count, st);
            break;
    }
    if (launches) --*launches;
    return cudaErrorInvalidValue;
}

}  // namespace imgcorr

// K1, streaming variant (3x3): the fused dark / flat / nan_to_num / median-threshold kernel as a
// warp-specialised TMA pipeline with the stencil window held in registers.
//
//   * work unit = (frame, 120-column strip, segment of rows).  A CTA is 4 consumer warps + 1 producer
//     warp and walks its units top to bottom in chunks of 32 rows.
//   * the producer warp (one elected lane) fills a 2-stage shared-memory ring with
//     cp.async.bulk.tensor boxes of raw / dark / flat rows (16-byte aligned, 16-byte multiple wide,
//     out-of-frame parts zero-filled by TMA), signalled through `full` mbarriers; consumers hand a
//     stage back through `empty` mbarriers.  There is no CTA-wide barrier in the steady state.
//   * each consumer lane owns one image column of its warp's 30-column slice (+1 halo lane on each
//     side; the halo / out-of-frame lanes read the mirrored column, which is scipy's 'reflect').
//     Per row it reads raw, dark, flat from shared memory, computes the pointwise value in float64
//     registers (one rounding to float32), fetches the left / right neighbours with two warp shuffles,
//     sorts the horizontal triple and combines it with the two previous rows' triples (kept in
//     registers) into the median of 9; predicate + select + one coalesced store per row.
//   * vertical 'reflect' needs no halo rows: the first / last row's sorted triple is used twice.
//
// Same arithmetic as the tile kernels (imgcorr_core.cuh) — results are bit-identical.
#include "imgcorr_kernels.cuh"
#include "imgcorr_tma.cuh"

namespace imgcorr {

constexpr int KS_CW = 4;                    // consumer warps per CTA
constexpr int KS_SW = 30;                   // output columns per consumer warp
constexpr int KS_TW = KS_CW * KS_SW;        // 120 output columns per strip
// Pipeline shape: rows per stage x stages x CTAs per SM the register allocation aims at.  Measured on B200 (4096x3000
// uint16, 32 frames per launch): 8 x 4 x 5 (72 registers) 34.5 us/frame; 16 x 3 x 3 32.8; 24 x 3 x 2 30.9; 32 x 2 x 2 (124
// registers) 30.6 — long straight-line chunks with many independent rows in flight per warp beat occupancy, the kernel is
// bound by dependent-issue latency, not by any pipe.  With few frames per launch (one frame: 57 vs 44 us) and for the
// integer-output instantiations (uint16 medianThreshold: 52 vs 41 us/frame) the small shape wins, so both are built.
template <int R_, int NSTAGE_, int MINB_> struct KsShape { static constexpr int R = R_, NSTAGE = NSTAGE_, MINB = MINB_; };
typedef KsShape<8, 4, 5> KsNarrow;
#ifndef KS_WIDE_R
#define KS_WIDE_R 32
#define KS_WIDE_MINB 3
#endif
typedef KsShape<KS_WIDE_R, 2, KS_WIDE_MINB> KsWide;
constexpr int KS_MAPW = 128;                // float32 box: tx0-4 .. tx0+123
constexpr int KS_MAPX = 4;
constexpr int KS_THREADS = (KS_CW + 1) * 32;

template <typename RawT, typename C> struct StreamBox {
    static constexpr int XOFF = 16 / (int)sizeof(RawT);                        // u8 16, u16 8, f32 4
    static constexpr int GRAN = 16 / (int)sizeof(RawT);
    // the box starts at the 16-byte boundary at or left of tx0, minus XOFF: columns tx0-1 .. tx0+120 are inside
    static constexpr int BOXW = ((KS_TW + 2 * XOFF + GRAN - 1) / GRAN) * GRAN;  // u8 160, u16 136, f32 128
    static constexpr size_t raw_bytes = (size_t)C::R * BOXW * sizeof(RawT);     // multiples of 128
    static constexpr size_t map_bytes = (size_t)C::R * KS_MAPW * sizeof(float);
    static constexpr size_t stage_bytes = raw_bytes + 2 * map_bytes;
    static constexpr size_t bar_off = C::NSTAGE * stage_bytes;
    static constexpr size_t total = bar_off + 2 * C::NSTAGE * sizeof(uint64_t) + 64;
};

template <typename T> struct StreamRaw;
template <> struct StreamRaw<uint8_t>  { static __device__ __forceinline__ void ld(uint8_t v, double& d, float& a)  { d = (double)(int)v; a = 0.0f; } };
template <> struct StreamRaw<uint16_t> { static __device__ __forceinline__ void ld(uint16_t v, double& d, float& a) { d = (double)(int)v; a = 0.0f; } };
template <> struct StreamRaw<float>    { static __device__ __forceinline__ void ld(float v, double& d, float& a)    { d = (double)v; a = fabsf(v); } };

template <typename OutT> __device__ __forceinline__ OutT ks_out(float v);
template <> __device__ __forceinline__ float    ks_out<float>(float v)    { return v; }
template <> __device__ __forceinline__ uint16_t ks_out<uint16_t>(float v) { return sat_u16(v); }
template <> __device__ __forceinline__ uint8_t  ks_out<uint8_t>(float v)  { return sat_u8(v); }

__device__ __noinline__ bool ks_exact(float x, float b, double thr, int cond) {
    PredicateConst pc;
    pc.thr = thr; pc.cond = cond; pc.lo = 0.f; pc.hi = 0.f; pc.fast_ok = 0;
    return predicate_exact((double)x, (double)b, pc);
}

struct UnitGeom {
    int frame, tx0, ys, ye, yl0, n_in, nchunk;
};
template <int R>
__device__ __forceinline__ UnitGeom ks_unit(int unit, int strips, int segs, int seg_rows, int H, int n_frames) {
    // frame index fastest: the units that share the dark / flat rows of one (strip, segment) run back to back,
    // so those rows are read from DRAM once per launch and served from L2 for the other frames
    UnitGeom u;
    u.frame = unit % n_frames;
    int t = unit / n_frames;
    const int strip = t % strips;
    const int seg = t / strips;
    (void)segs;
    u.tx0 = strip * KS_TW;
    u.ys = seg * seg_rows;
    u.ye = u.ys + seg_rows < H ? u.ys + seg_rows : H;
    u.yl0 = u.ys > 0 ? u.ys - 1 : 0;
    const int yl1 = u.ye < H ? u.ye : H - 1;
    u.n_in = yl1 - u.yl0 + 1;
    u.nchunk = (u.n_in + R - 1) / R;
    return u;
}

// CFG >= 0 bakes the per-launch switches into the instruction stream (the kernel is issue bound: every
// per-row flag test costs); CFG < 0 reads them from the arguments.
enum : int { KS_DARK = 1, KS_FLAT = 2, KS_N2N = 4, KS_MASK = 8, KS_CHECK = 16, KS_LT = 32, KS_NZ = 64, KS_SWAP = 128, KS_NOMED = 256 };
// KS_NOMED: ksize 0 (threshold <= 0, CameraCalibration.py:556-557 skips the artefact step): the pointwise value itself is
// stored, through the same pipeline (one row late, like the median's centre pixel)
// KS_SWAP: big-endian uint16 samples (reader/RAW.py default) — one PRMT per pixel after the shared-memory load
// KS_NZ: a.flat is the zero-free copy (zeros replaced by 1.0) -> unconditional division
// KS_CHECK: non-finite calibration values or float32 raw samples are possible -> test and fall back per pixel

template <typename RawT, typename OutT, int CFG, typename C>
__global__ void __launch_bounds__(KS_THREADS, C::MINB)
k1_stream_kernel(const __grid_constant__ CUtensorMap tm_raw, const __grid_constant__ CUtensorMap tm_dark,
                 const __grid_constant__ CUtensorMap tm_flat, K1Args a, int strips, int segs, int seg_rows, int total_units) {
    using B = StreamBox<RawT, C>;
    constexpr int KS_R = C::R, KS_NSTAGE = C::NSTAGE;
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* full = (uint64_t*)(smem + B::bar_off);
    uint64_t* empty = full + KS_NSTAGE;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool has_dark = CFG >= 0 ? (CFG & KS_DARK) != 0 : a.dark != nullptr;
    const bool has_flat = CFG >= 0 ? (CFG & KS_FLAT) != 0 : a.flat != nullptr;
    const bool has_mask = CFG >= 0 ? (CFG & KS_MASK) != 0 : a.mask != nullptr;
    const bool check = CFG >= 0 ? (CFG & KS_CHECK) != 0 : true;
    const bool swap = CFG >= 0 ? (CFG & KS_SWAP) != 0 : a.raw_swap != 0;
    const bool nomed = CFG >= 0 ? (CFG & KS_NOMED) != 0 : a.ksize == 0;
    const int flags = CFG >= 0 ? ((CFG & KS_DARK ? FLAG_DARK : 0) | (CFG & KS_FLAT ? FLAG_FLAT : 0) | (CFG & KS_N2N ? FLAG_NAN_TO_NUM : 0))
                               : a.pw.flags;
    const int H = a.H, W = a.W;

    float* sconst = (float*)(smem + B::bar_off + 2 * KS_NSTAGE * sizeof(uint64_t));      // lo, hi, thr (double)
    if (threadIdx.x == 0) {
        for (int s = 0; s < KS_NSTAGE; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], KS_CW); }
        sconst[0] = a.pred.lo; sconst[1] = a.pred.hi; *(double*)(sconst + 2) = a.pred.thr;
        mbar_init_fence();
    }
    __syncthreads();

    if (warp == KS_CW) {
        // ------------------------------------------------------------------ producer
        if (lane != 0) return;
        const uint32_t tx_bytes = (uint32_t)(B::raw_bytes + (has_dark ? B::map_bytes : 0) + (has_flat ? B::map_bytes : 0));
        uint32_t g = 0;
        for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
            const UnitGeom u = ks_unit<KS_R>(unit, strips, segs, seg_rows, H, a.n_frames);
            for (int k = 0; k < u.nchunk; ++k, ++g) {
                const int stage = g % KS_NSTAGE;
                mbar_wait(&empty[stage], ((g / KS_NSTAGE) & 1) ^ 1);
                uint8_t* base = smem + (size_t)stage * B::stage_bytes;
                const int y = u.yl0 + k * KS_R;
                mbar_expect_tx(&full[stage], tx_bytes);
                tma_load_3d(base, &tm_raw, &full[stage], (u.tx0 / B::GRAN) * B::GRAN - B::XOFF, y, u.frame);
                if (has_dark) tma_load_2d(base + B::raw_bytes, &tm_dark, &full[stage], u.tx0 - KS_MAPX, y);
                if (has_flat) tma_load_2d(base + B::raw_bytes + B::map_bytes, &tm_flat, &full[stage], u.tx0 - KS_MAPX, y);
            }
        }
        return;
    }

    // ---------------------------------------------------------------------- consumers
    const PointwiseConst pw = a.pw;
    // per-launch constants come from shared memory: ptxas would re-read kernel parameters at every use
    PredicateConst pred = a.pred;
    pred.lo = sconst[0]; pred.hi = sconst[1]; pred.thr = *(const double*)(sconst + 2);
    if (CFG >= 0) pred.cond = (CFG & KS_LT) ? COND_LT : COND_GT;
    const int lc = warp * KS_SW - 1 + lane;            // strip-local column of this lane: -1 .. 120
    uint32_t g = 0;

    for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
        const UnitGeom u = ks_unit<KS_R>(unit, strips, segs, seg_rows, H, a.n_frames);
        const int gc = u.tx0 + lc;
        const int rc = reflect_index(gc, W);             // scipy 'reflect' in x: halo / outside lanes read the mirrored column
        int mcol = rc - (u.tx0 - KS_MAPX);
        mcol = mcol < 0 ? 0 : (mcol > KS_MAPW - 1 ? KS_MAPW - 1 : mcol);
        int rcol = rc - ((u.tx0 / B::GRAN) * B::GRAN - B::XOFF);
        rcol = rcol < 0 ? 0 : (rcol > B::BOXW - 1 ? B::BOXW - 1 : rcol);
        const bool valid = lane >= 1 && lane <= KS_SW && gc < W;
        const int i_first = u.ys == 0 ? 1 : 2;           // first input row index whose step emits an output row
        const ptrdiff_t o0 = ((ptrdiff_t)u.frame * H + u.yl0 - 1) * W + gc;          // row (yl0 + i - 1) at step i
        OutT* outp = (OutT*)a.out + o0;
        uint8_t* maskp = has_mask ? a.mask + o0 : nullptr;

        Sorted3<float> s0, s1;
        float c1;
        int i = 0;

        auto pixel = [&](const uint8_t* base, int j) -> float {
            RawT rv = ((const RawT*)base)[j * B::BOXW + rcol];
            if (sizeof(RawT) == 2 && swap) rv = (RawT)__byte_perm((unsigned)rv, 0u, 0x0001);
            const float d = has_dark ? ((const float*)(base + B::raw_bytes))[j * KS_MAPW + mcol] : 0.0f;
            const float f = has_flat ? ((const float*)(base + B::raw_bytes + B::map_bytes))[j * KS_MAPW + mcol] : 0.0f;
            double rd; float ra;
            StreamRaw<RawT>::ld(rv, rd, ra);
            bool ok;
            float x = (CFG >= 0 && (CFG & KS_NZ)) ? pointwise_fast_nz(flags, rd, ra, d, f, ok) : pointwise_fast(flags, rd, ra, d, f, ok);
            if (check) { if (!ok) x = pointwise<float>(pw, rd, d, 0.0f, f); }
            return x;
        };
        auto emit = [&](const Sorted3<float>& t2, bool on) {
            if (nomed) {
                if (on) *outp = ks_out<OutT>(c1);
                return;
            }
            const float med = median9(s0, s1, t2);
            bool rep;
            const bool sure = predicate_certain(c1, med, pred, rep);
            if (on) {
                *outp = ks_out<OutT>(rep ? med : c1);
                if (has_mask) *maskp = rep ? 1 : 0;
            }
            // ~1e-5 of the pixels fall inside the guard band of the float32 test: those lanes evaluate the reference's
            // float64 expression out of line and overwrite their result (one warp-uniform branch in the common case)
            if (__any_sync(0xffffffffu, !sure)) {
                if (!sure && on) {
                    const bool r2 = ks_exact(c1, med, pred.thr, pred.cond);
                    *outp = ks_out<OutT>(r2 ? med : c1);
                    if (has_mask) *maskp = r2 ? 1 : 0;
                }
            }
        };
        auto row = [&](const uint8_t* base, int j) {
            const float x = pixel(base, j);
            const float l = __shfl_up_sync(0xffffffffu, x, 1);
            const float r = __shfl_down_sync(0xffffffffu, x, 1);
            const Sorted3<float> t2 = sort3(l, x, r);
            emit(t2, valid && i >= i_first);
            s0 = s1; s1 = t2; c1 = x;
            outp += W;
            if (has_mask) maskp += W;
            ++i;
        };

        for (int k = 0; k < u.nchunk; ++k, ++g) {
            const int stage = g % KS_NSTAGE;
            mbar_wait(&full[stage], (g / KS_NSTAGE) & 1);
            const uint8_t* base = smem + (size_t)stage * B::stage_bytes;
            if (k == 0) {
                // vertical 'reflect' at the top (and a defined s0/s1 elsewhere): the first row's triple is used twice
                const float x = pixel(base, 0);
                const float l = __shfl_up_sync(0xffffffffu, x, 1);
                const float r = __shfl_down_sync(0xffffffffu, x, 1);
                s1 = sort3(l, x, r);
                s0 = s1;
                c1 = x;
            }
            const int rows = u.n_in - k * KS_R;
            if (k > 0 && rows >= KS_R) {
                // steady state: every row emits; the store is unconditional (lanes without an output pixel write to
                // a private scratch slot) and the guard-band test is only accumulated — if any lane of the warp hit the
                // band in this chunk (rare), the chunk is redone below with the exact predicate.
                const Sorted3<float> k0 = s0, k1 = s1;
                const float kc = c1;
                OutT* op = valid ? outp : (OutT*)a.dump + (size_t)blockIdx.x * KS_THREADS + threadIdx.x;
                uint8_t* mp = has_mask ? (valid ? maskp : (uint8_t*)a.dump + (size_t)(gridDim.x + blockIdx.x) * KS_THREADS * sizeof(OutT) + threadIdx.x) : nullptr;
                const int ostride = valid ? W : 0;
                bool unsure = false;
#pragma unroll
                for (int j = 0; j < KS_R; ++j) {
                    const float x = pixel(base, j);
                    if (nomed) {
                        *op = ks_out<OutT>(c1);
                        op += ostride;
                        c1 = x;
                        continue;
                    }
                    const float l = __shfl_up_sync(0xffffffffu, x, 1);
                    const float r = __shfl_down_sync(0xffffffffu, x, 1);
                    const Sorted3<float> t2 = sort3(l, x, r);
                    const float med = median9(s0, s1, t2);
                    bool rep;
                    unsure |= !predicate_certain(c1, med, pred, rep);
                    *op = ks_out<OutT>(rep ? med : c1);
                    if (has_mask) { *mp = rep ? 1 : 0; mp += ostride; }
                    op += ostride;
                    s0 = s1; s1 = t2; c1 = x;
                }
                if (__any_sync(0xffffffffu, unsure)) {
                    s0 = k0; s1 = k1; c1 = kc;
                    for (int j = 0; j < KS_R; ++j) row(base, j);
                } else {
                    outp += (size_t)KS_R * W;
                    if (has_mask) maskp += (size_t)KS_R * W;
                    i += KS_R;
                }
            } else {
                const int n = rows < KS_R ? rows : KS_R;
                for (int j = 0; j < n; ++j) row(base, j);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[stage]);
        }
        if (u.ye == H) {
            // vertical 'reflect' at the bottom: output row H-1 sees (H-2, H-1, H-1)
            const Sorted3<float> t2 = s1;
            emit(t2, valid);
        }
    }
}

// ------------------------------------------------------------------------------------------ host
bool k1_stream_eligible(const K1Args& a, int raw_dtype, int out_dtype) {
    if (a.ksize != 3 && a.ksize != 0) return false;
    if (a.ksize == 0 && a.H < 4) return false;
    if (raw_dtype != DT_U8 && raw_dtype != DT_U16 && raw_dtype != DT_F32) return false;
    if (out_dtype != DT_F32 && !(raw_dtype == DT_U16 && out_dtype == DT_U16) && !(raw_dtype == DT_U8 && out_dtype == DT_U8)) return false;
    if (a.pw.flags & FLAG_DARK_LINEAR) return false;
    if (a.raw_gap) return false;
    const size_t esz = dtype_size(raw_dtype);
    if (((size_t)a.W * esz) % 16 || ((size_t)a.W * 4) % 16) return false;
    if (((size_t)a.H * a.W * esz) % 16) return false;
    if (((uintptr_t)a.raw) % 16) return false;
    if (a.dark && ((uintptr_t)a.dark) % 16) return false;
    if (a.flat && ((uintptr_t)a.flat) % 16) return false;
    return tensor_map_encoder() != nullptr;
}

template <typename RawT, typename OutT, typename C>
static cudaError_t launch_stream_c(const K1Args& a_in, CUtensorMapDataType rdt, int sm_count, int seg_rows, cudaStream_t st) {
    using B = StreamBox<RawT, C>;
    constexpr int KS_R = C::R;
    K1Args a = a_in;
    // pick the instantiation: the hot configurations are fully specialised, the rest read their flags at run time
    const bool check = !a.maps_finite || sizeof(RawT) == 4;
    const int f = a.pw.flags;
    const bool chain = a.ksize == 3 && a.dark && a.flat && a.flat_nz && (f & FLAG_DARK) && (f & FLAG_FLAT) && (f & FLAG_NAN_TO_NUM) && !a.mask &&
                       a.pred.cond == COND_GT;
    const bool plain = a.ksize == 3 && !(f & (FLAG_DARK | FLAG_FLAT | FLAG_NAN_TO_NUM)) && a.mask && a.pred.cond == COND_GT;
    void (*kern)(const CUtensorMap, const CUtensorMap, const CUtensorMap, K1Args, int, int, int, int);
    int slot;
    // threshold <= 0: dark + flat only (nan_to_num is not applied then, CameraCalibration.py:556-561)
    const bool pw_only = a.ksize == 0 && a.dark && a.flat && a.flat_nz && (f & FLAG_DARK) && (f & FLAG_FLAT) && !a.raw_swap && !check;
    if (chain || pw_only) a.flat = a.flat_nz;          // zero-free copy: "divide where flat != 0" becomes an unconditional division
    if (pw_only) {
        kern = (f & FLAG_NAN_TO_NUM) && !a.no_overflow ? k1_stream_kernel<RawT, OutT, KS_DARK | KS_FLAT | KS_N2N | KS_NZ | KS_NOMED, C>
                                                       : k1_stream_kernel<RawT, OutT, KS_DARK | KS_FLAT | KS_NZ | KS_NOMED, C>;
        slot = (f & FLAG_NAN_TO_NUM) && !a.no_overflow ? 7 : 8;
    } else if (a.raw_swap && chain && !check && sizeof(RawT) == 2) {
        kern = a.no_overflow ? k1_stream_kernel<RawT, OutT, KS_DARK | KS_FLAT | KS_NZ | KS_SWAP, C>
                             : k1_stream_kernel<RawT, OutT, KS_DARK | KS_FLAT | KS_N2N | KS_NZ | KS_SWAP, C>;
        slot = a.no_overflow ? 5 : 6;
    } else if (a.raw_swap) { kern = k1_stream_kernel<RawT, OutT, -1, C>; slot = 4; if (chain) a.flat = a_in.flat; }
    else if (chain && !check && a.no_overflow) { kern = k1_stream_kernel<RawT, OutT, KS_DARK | KS_FLAT | KS_NZ, C>; slot = 0; }
    else if (chain && !check) { kern = k1_stream_kernel<RawT, OutT, KS_DARK | KS_FLAT | KS_N2N | KS_NZ, C>; slot = 1; }
    else if (chain) { kern = k1_stream_kernel<RawT, OutT, KS_DARK | KS_FLAT | KS_N2N | KS_NZ | KS_CHECK, C>; slot = 2; }
    else if (plain) { kern = k1_stream_kernel<RawT, OutT, KS_MASK, C>; slot = 3; }
    else { kern = k1_stream_kernel<RawT, OutT, -1, C>; slot = 4; }

    CUtensorMap tr, td, tf;
    if (!make_tensor_map(&tr, rdt, sizeof(RawT), a.raw, a.W, a.H, a.n_frames, B::BOXW, KS_R)) return cudaErrorInvalidValue;
    if (!make_tensor_map(&td, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, a.dark ? (const void*)a.dark : a.raw, a.W, a.H, 0, KS_MAPW, KS_R) && a.dark)
        return cudaErrorInvalidValue;
    if (!make_tensor_map(&tf, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, a.flat ? (const void*)a.flat : a.raw, a.W, a.H, 0, KS_MAPW, KS_R) && a.flat)
        return cudaErrorInvalidValue;
    static int per_sm[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (!per_sm[slot]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)B::total);
        if (e != cudaSuccess) return e;
        int n = 1;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, KS_THREADS, B::total);
        per_sm[slot] = n < 1 ? 1 : n;
    }
    const int strips = (a.W + KS_TW - 1) / KS_TW;
    const long long slots = (long long)sm_count * per_sm[slot];
    if (seg_rows <= 0) {
        // all units cost the same: pick the segment height whose unit count fills whole waves of the resident CTAs best,
        // counting the two re-read halo rows per segment against it.  (Searching beyond 8 waves finds better fills on
        // paper — 231-row segments for 32 frames — but measured slower, 36.1 vs 33.6 us/frame: every unit start refills
        // the pipeline and runs its first chunk row by row.)
        double best = -1.0;
        for (int waves = 1; waves <= 8; ++waves) {
            long long segs_try = slots * waves / ((long long)strips * a.n_frames);
            if (segs_try < 1) continue;
            int rows = (int)((a.H + segs_try - 1) / segs_try);
            if (rows < 2 * KS_R) rows = 2 * KS_R;
            const long long units = (long long)strips * ((a.H + rows - 1) / rows) * a.n_frames;
            const double util = (double)units / (double)(((units + slots - 1) / slots) * slots) * rows / (rows + 2.0);
            if (util > best) { best = util; seg_rows = rows; }
        }
        if (seg_rows <= 0) seg_rows = a.H;
    }
    if (seg_rows < 4) seg_rows = 4;
    const int segs = (a.H + seg_rows - 1) / seg_rows;
    const long long total = (long long)strips * segs * a.n_frames;
    if (total > 0x7fffffffLL) return cudaErrorInvalidValue;
    long long grid = slots;
    if (grid > total) grid = total;
    kern<<<(unsigned)grid, KS_THREADS, B::total, st>>>(tr, td, tf, a, strips, segs, seg_rows, (int)total);
    return cudaGetLastError();
}

template <typename RawT, typename OutT>
static cudaError_t launch_stream_t(const K1Args& a, CUtensorMapDataType rdt, int sm_count, int seg_rows, cudaStream_t st) {
    // the wide shape pays with float32 output and a few frames per launch (see KsShape)
    if (sizeof(OutT) == 4 && a.n_frames >= 4) return launch_stream_c<RawT, OutT, KsWide>(a, rdt, sm_count, seg_rows, st);
    return launch_stream_c<RawT, OutT, KsNarrow>(a, rdt, sm_count, seg_rows, st);
}

cudaError_t launch_k1_stream(const K1Args& a, int raw_dtype, int out_dtype, int sm_count, int seg_rows, cudaStream_t st) {
    switch (raw_dtype) {
        case DT_U8:
            if (out_dtype == DT_F32) return launch_stream_t<uint8_t, float>(a, CU_TENSOR_MAP_DATA_TYPE_UINT8, sm_count, seg_rows, st);
            if (out_dtype == DT_U8) return launch_stream_t<uint8_t, uint8_t>(a, CU_TENSOR_MAP_DATA_TYPE_UINT8, sm_count, seg_rows, st);
            break;
        case DT_U16:
            if (out_dtype == DT_F32) return launch_stream_t<uint16_t, float>(a, CU_TENSOR_MAP_DATA_TYPE_UINT16, sm_count, seg_rows, st);
            if (out_dtype == DT_U16) return launch_stream_t<uint16_t, uint16_t>(a, CU_TENSOR_MAP_DATA_TYPE_UINT16, sm_count, seg_rows, st);
            break;
        case DT_F32:
            if (out_dtype == DT_F32) return launch_stream_t<float, float>(a, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, sm_count, seg_rows, st);
            break;
    }
    return cudaErrorInvalidValue;
}

}  // namespace imgcorr

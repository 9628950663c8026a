// K1, streaming variant 2 (3x3, uint16 / float32 frames): same pipeline idea as k1_stream.cu
// (producer warp -> TMA ring -> independent consumer warps, window in registers, neighbours by warp
// shuffle) restructured around what the profile of variant 1 showed: the kernel is bound by the
// half-rate ALU pipe (FMNMX / FSETP / FSEL), not by DRAM and not by the FP64 pipe.
//
//   * every lane owns TWO adjacent columns: one sorted pair per row serves both horizontal triples
//     (10 min/max per 2 pixels instead of 12), loads are 32/64-bit, shuffles and address updates halve,
//     and 62 of 64 columns of a warp are outputs (30 of 32 before).
//   * the flat-field copy in the context has its zeros replaced by 1.0 at upload, so "divide only where
//     flat != 0" is a plain division (x / 1 == x exactly) — no compare/select per pixel.
//   * when the calibration proves that the quotient cannot overflow float32 (integer frames: |raw - dark| <=
//     65535 + max|dark|, divided by min|flat|), nan_to_num is the identity and its clamp is dropped.
//   * the threshold predicate uses two FFMA-built bounds with absolute margins instead of three compares
//     (see predicate_certain2), and both pixels of a lane share one rarely-taken exact-path branch.
//   * per-launch constants are staged in shared memory once and kept in registers (ptxas otherwise re-reads
//     kernel parameters from the constant bank at every use).
//   * scipy's horizontal 'reflect' is produced by patching the single out-of-frame halo column of an edge
//     strip in shared memory (the warp that owns the column does it for its own rows), vertical 'reflect'
//     by re-using the first / last row's sorted triples.
//
// Strip = 248 output columns (4 consumer warps x 62), staged box = 256 columns starting 8 columns left of
// the strip (16-byte aligned for uint16 and float32 rows; strip k covers output columns 248k-7 .. 248k+240).
// Same arithmetic as every other K1 variant (imgcorr_core.cuh): results are bit-identical.
#include "imgcorr_kernels.cuh"
#include "imgcorr_tma.cuh"

namespace imgcorr {

constexpr int K2S_CW = 4;                      // consumer warps per CTA
constexpr int K2S_WCOLS = 62;                  // output columns per consumer warp
constexpr int K2S_STRIP = K2S_CW * K2S_WCOLS;  // 248
constexpr int K2S_BOXW = 256;
constexpr int K2S_SHIFT = 8;                   // box starts at 248k - 8
constexpr int K2S_THREADS = (K2S_CW + 1) * 32;
#ifndef K2S_MINB_V
#define K2S_MINB_V 6
#endif
#ifndef K2S_UNROLL_V
#define K2S_UNROLL_V 3
#endif
#ifndef K2S_R_V
#define K2S_R_V 6
#endif
#ifndef K2S_NSTAGE_V
#define K2S_NSTAGE_V 2
#endif
constexpr int K2S_UNROLL = K2S_UNROLL_V;
constexpr int K2S_MINB = K2S_MINB_V;           // CTAs per SM the register allocation aims at (6 -> 64 registers / thread)

template <typename RawT, int R, int NSTAGE> struct Stream2Smem {
    static constexpr size_t raw_bytes = (size_t)R * K2S_BOXW * sizeof(RawT);
    static constexpr size_t map_bytes = (size_t)R * K2S_BOXW * sizeof(float);
    static constexpr size_t stage_bytes = raw_bytes + 2 * map_bytes;
    static constexpr size_t bar_off = NSTAGE * stage_bytes;
    static constexpr size_t const_off = bar_off + 2 * NSTAGE * sizeof(uint64_t);
    static constexpr size_t total = const_off + 64;
};

// two adjacent raw samples of a row, as float64
template <typename RawT> __device__ __forceinline__ void s2_load_pair(const RawT* row, int b0, double& v0, double& v1, float& a0, float& a1);
template <> __device__ __forceinline__ void s2_load_pair<uint16_t>(const uint16_t* row, int b0, double& v0, double& v1, float& a0, float& a1) {
    const uint32_t q = *reinterpret_cast<const uint32_t*>(row + b0);
    v0 = (double)(uint16_t)(q & 0xffffu);
    v1 = (double)(uint16_t)(q >> 16);
    a0 = a1 = 0.0f;
}
template <> __device__ __forceinline__ void s2_load_pair<float>(const float* row, int b0, double& v0, double& v1, float& a0, float& a1) {
    const float2 q = *reinterpret_cast<const float2*>(row + b0);
    v0 = (double)q.x; v1 = (double)q.y;
    a0 = fabsf(q.x); a1 = fabsf(q.y);
}

template <typename OutT> __device__ __forceinline__ OutT s2_out(float v);
template <> __device__ __forceinline__ float    s2_out<float>(float v)    { return v; }
template <> __device__ __forceinline__ uint16_t s2_out<uint16_t>(float v) { return sat_u16(v); }

__device__ __noinline__ bool s2_exact(float x, float b, double thr, int cond) {
    PredicateConst pc;
    pc.thr = thr; pc.cond = cond; pc.lo = 0.f; pc.hi = 0.f; pc.fast_ok = 0;
    return predicate_exact((double)x, (double)b, pc);
}

struct S2Unit { int frame, bx0, ys, ye, yl0, n_in, nchunk; };
template <int R>
__device__ __forceinline__ S2Unit s2_unit(int unit, int strips, int segs, int seg_rows, int H) {
    S2Unit u;
    const int strip = unit % strips;
    int t = unit / strips;
    const int seg = t % segs;
    u.frame = t / segs;
    u.bx0 = strip * K2S_STRIP - K2S_SHIFT;          // global column of box column 0
    u.ys = seg * seg_rows;
    u.ye = u.ys + seg_rows < H ? u.ys + seg_rows : H;
    u.yl0 = u.ys > 0 ? u.ys - 1 : 0;
    const int yl1 = u.ye < H ? u.ye : H - 1;
    u.n_in = yl1 - u.yl0 + 1;
    u.nchunk = (u.n_in + R - 1) / R;
    return u;
}

// per-launch constants staged in shared memory
struct S2Const { float lo, hi; double thr; };

// CFG bits as in k1_stream.cu
enum : int { S2_DARK = 1, S2_FLAT = 2, S2_N2N = 4, S2_MASK = 8, S2_CHECK = 16, S2_LT = 32 };

template <typename RawT, typename OutT, int CFG, int R, int NSTAGE>
__global__ void __launch_bounds__(K2S_THREADS, K2S_MINB)
k1_stream2_kernel(const __grid_constant__ CUtensorMap tm_raw, const __grid_constant__ CUtensorMap tm_dark,
                  const __grid_constant__ CUtensorMap tm_flat, K1Args a, int strips, int segs, int seg_rows, int total_units) {
    using S = Stream2Smem<RawT, R, NSTAGE>;
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* full = (uint64_t*)(smem + S::bar_off);
    uint64_t* empty = full + NSTAGE;
    S2Const* sc = (S2Const*)(smem + S::const_off);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr bool has_dark = (CFG & S2_DARK) != 0, has_flat = (CFG & S2_FLAT) != 0, has_mask = (CFG & S2_MASK) != 0;
    constexpr bool check = (CFG & S2_CHECK) != 0;
    constexpr int flags = (has_dark ? FLAG_DARK : 0) | (has_flat ? FLAG_FLAT : 0) | ((CFG & S2_N2N) ? FLAG_NAN_TO_NUM : 0);
    constexpr int cond = (CFG & S2_LT) ? COND_LT : COND_GT;
    const int H = a.H, W = a.W;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], K2S_CW); }
        sc->lo = a.pred.lo; sc->hi = a.pred.hi; sc->thr = a.pred.thr;
        mbar_init_fence();
    }
    __syncthreads();

    if (warp == K2S_CW) {
        // ------------------------------------------------------------------ producer
        if (lane != 0) return;
        constexpr uint32_t tx_bytes = (uint32_t)(S::raw_bytes + (has_dark ? S::map_bytes : 0) + (has_flat ? S::map_bytes : 0));
        uint32_t g = 0;
        for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
            const S2Unit u = s2_unit<R>(unit, strips, segs, seg_rows, H);
            for (int k = 0; k < u.nchunk; ++k, ++g) {
                const int stage = g % NSTAGE;
                mbar_wait(&empty[stage], ((g / NSTAGE) & 1) ^ 1);
                uint8_t* base = smem + (size_t)stage * S::stage_bytes;
                const int y = u.yl0 + k * R;
                mbar_expect_tx(&full[stage], tx_bytes);
                tma_load_3d(base, &tm_raw, &full[stage], u.bx0, y, u.frame);
                if (has_dark) tma_load_2d(base + S::raw_bytes, &tm_dark, &full[stage], u.bx0, y);
                if (has_flat) tma_load_2d(base + S::raw_bytes + S::map_bytes, &tm_flat, &full[stage], u.bx0, y);
            }
        }
        return;
    }

    // ---------------------------------------------------------------------- consumers
    PredicateConst pred;
    pred.lo = sc->lo; pred.hi = sc->hi; pred.thr = sc->thr; pred.cond = cond; pred.fast_ok = 1;
    const PointwiseConst pw = a.pw;
    const int b0 = warp * K2S_WCOLS + 2 * lane;        // box column of this lane's first pixel (second = b0 + 1)
    uint32_t g = 0;

    for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
        const S2Unit u = s2_unit<R>(unit, strips, segs, seg_rows, H);
        const int gc0 = u.bx0 + b0;
        const bool valid0 = lane != 0 && gc0 >= 0 && gc0 < W;
        const bool valid1 = lane != 31 && gc0 + 1 >= 0 && gc0 + 1 < W;
        // horizontal 'reflect': the halo column just outside the frame (box column pl / pr) is a copy of the edge column
        const int pl = -1 - u.bx0;                       // box column of global column -1
        const int pr = W - u.bx0;                        // box column of global column W
        const bool patch_l = pl >= 0 && pl / K2S_WCOLS == warp;          // read as left neighbour of output box column pl + 1
        const bool patch_r = pr >= 2 && pr <= K2S_STRIP + 1 && (pr - 2) / K2S_WCOLS == warp;   // right neighbour of output box column pr - 1
        const int i_first = u.ys == 0 ? 1 : 2;
        const ptrdiff_t o0 = ((ptrdiff_t)u.frame * H + u.yl0 - 1) * W + gc0;      // row (yl0 + i - 1) at step i
        OutT* outp = (OutT*)a.out + o0;
        uint8_t* maskp = has_mask ? a.mask + o0 : nullptr;

        Sorted3<float> a0, a1, c0, c1;          // previous two rows' sorted triples of the first (a*) and second (c*) pixel
        float m0, m1;                            // previous row's centre values
        int i = 0;

        auto load_row = [&](const uint8_t* base, int j, float& x0, float& x1) {
            double v0, v1; float r0, r1;
            s2_load_pair<RawT>((const RawT*)base + j * K2S_BOXW, b0, v0, v1, r0, r1);
            float2 d = make_float2(0.f, 0.f), f = make_float2(1.f, 1.f);
            if (has_dark) d = *reinterpret_cast<const float2*>((const float*)(base + S::raw_bytes) + j * K2S_BOXW + b0);
            if (has_flat) f = *reinterpret_cast<const float2*>((const float*)(base + S::raw_bytes + S::map_bytes) + j * K2S_BOXW + b0);
            bool ok0, ok1;
            x0 = pointwise_fast_nz(flags, v0, r0, d.x, f.x, ok0);
            x1 = pointwise_fast_nz(flags, v1, r1, d.y, f.y, ok1);
            if (check) {
                if (!ok0) x0 = pointwise<float>(pw, v0, d.x, 0.0f, f.x);
                if (!ok1) x1 = pointwise<float>(pw, v1, d.y, 0.0f, f.y);
            }
        };
        auto triples = [&](float x0, float x1, Sorted3<float>& A, Sorted3<float>& C) {
            const float l = __shfl_up_sync(0xffffffffu, x1, 1);      // left neighbour of the first pixel
            const float r = __shfl_down_sync(0xffffffffu, x0, 1);    // right neighbour of the second pixel
            const float lo = fminf(x0, x1), hi = fmaxf(x0, x1);
            float t = fmaxf(l, lo);
            A.lo = fminf(l, lo); A.mid = fminf(t, hi); A.hi = fmaxf(t, hi);
            t = fmaxf(r, lo);
            C.lo = fminf(r, lo); C.mid = fminf(t, hi); C.hi = fmaxf(t, hi);
        };
        auto emit = [&](const Sorted3<float>& A2, const Sorted3<float>& C2, bool on0, bool on1) {
            const float q0 = median9(a0, a1, A2), q1 = median9(c0, c1, C2);
            bool rep0, rep1;
            const bool s0 = predicate_certain2(m0, q0, pred, rep0), s1 = predicate_certain2(m1, q1, pred, rep1);
            // the exact float64 evaluation is needed for ~1e-5 of the pixels: one warp-uniform branch, code out of line
            if (__any_sync(0xffffffffu, !(s0 && s1))) {
                if (!s0) rep0 = s2_exact(m0, q0, pred.thr, cond);
                if (!s1) rep1 = s2_exact(m1, q1, pred.thr, cond);
            }
            if (on0) { outp[0] = s2_out<OutT>(rep0 ? q0 : m0); if (has_mask) maskp[0] = rep0 ? 1 : 0; }
            if (on1) { outp[1] = s2_out<OutT>(rep1 ? q1 : m1); if (has_mask) maskp[1] = rep1 ? 1 : 0; }
        };
        auto patch = [&](uint8_t* base, int rows) {
            // copy edge column -> out-of-frame halo column for raw / dark / flat, rows 0..rows-1 (lanes split the work)
            if (patch_l | patch_r) {
                for (int t = lane; t < rows * 2; t += 32) {
                    const int j = t >> 1, right = t & 1;
                    if (right ? patch_r : patch_l) {
                        const int dstc = right ? pr : pl, srcc = right ? pr - 1 : pl + 1;
                        RawT* rr = (RawT*)base + j * K2S_BOXW;
                        rr[dstc] = rr[srcc];
                        if (has_dark) { float* q = (float*)(base + S::raw_bytes) + j * K2S_BOXW; q[dstc] = q[srcc]; }
                        if (has_flat) { float* q = (float*)(base + S::raw_bytes + S::map_bytes) + j * K2S_BOXW; q[dstc] = q[srcc]; }
                    }
                }
                __syncwarp();
            }
        };

        for (int k = 0; k < u.nchunk; ++k, ++g) {
            const int stage = g % NSTAGE;
            mbar_wait(&full[stage], (g / NSTAGE) & 1);
            uint8_t* base = smem + (size_t)stage * S::stage_bytes;
            const int rows = u.n_in - k * R < R ? u.n_in - k * R : R;
            patch(base, rows);
            if (k == 0) {
                // vertical 'reflect' at the top (and defined state elsewhere): the first row's triples are used twice
                float x0, x1;
                load_row(base, 0, x0, x1);
                triples(x0, x1, a1, c1);
                a0 = a1; c0 = c1; m0 = x0; m1 = x1;
            }
            // rows before i_first (the first one or two input rows of a unit) produce no output
            const bool e0 = valid0 && i >= i_first, e1 = valid1 && i >= i_first;
            if (rows == R && i >= i_first) {
#pragma unroll K2S_UNROLL
                for (int j = 0; j < R; ++j) {
                    float x0, x1;
                    load_row(base, j, x0, x1);
                    Sorted3<float> A2, C2;
                    triples(x0, x1, A2, C2);
                    emit(A2, C2, e0, e1);
                    a0 = a1; a1 = A2; c0 = c1; c1 = C2; m0 = x0; m1 = x1;
                    outp += W; if (has_mask) maskp += W;
                }
                i += R;
            } else {
                for (int j = 0; j < rows; ++j) {
                    float x0, x1;
                    load_row(base, j, x0, x1);
                    Sorted3<float> A2, C2;
                    triples(x0, x1, A2, C2);
                    const bool on = i >= i_first;
                    emit(A2, C2, valid0 && on, valid1 && on);
                    a0 = a1; a1 = A2; c0 = c1; c1 = C2; m0 = x0; m1 = x1;
                    outp += W; if (has_mask) maskp += W;
                    ++i;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[stage]);
        }
        if (u.ye == H) {
            // vertical 'reflect' at the bottom: output row H-1 sees (H-2, H-1, H-1)
            const Sorted3<float> A2 = a1, C2 = c1;
            emit(A2, C2, valid0, valid1);
        }
    }
}

// ------------------------------------------------------------------------------------------ host
// which specialisation serves this call, or -1
static int s2_config(const K1Args& a, int raw_dtype) {
    const int f = a.pw.flags;
    if (a.pred.cond != COND_GT || !a.pred.fast_ok) return -1;
    const bool chain = a.dark && a.flat && (f & FLAG_DARK) && (f & FLAG_FLAT) && (f & FLAG_NAN_TO_NUM) && !a.mask;
    const bool plain = !(f & (FLAG_DARK | FLAG_FLAT | FLAG_NAN_TO_NUM)) && a.mask;
    if (chain) {
        if (!a.flat_nz) return -1;
        const bool check = !a.maps_finite || raw_dtype == DT_F32;
        if (check) return S2_DARK | S2_FLAT | S2_N2N | S2_CHECK;
        return a.no_overflow ? (S2_DARK | S2_FLAT) : (S2_DARK | S2_FLAT | S2_N2N);
    }
    if (plain) return S2_MASK;
    return -1;
}

bool k1_stream2_eligible(const K1Args& a, int raw_dtype, int out_dtype) {
    if (a.ksize != 3) return false;
    if (raw_dtype != DT_U16 && raw_dtype != DT_F32) return false;
    if (out_dtype != DT_F32 && !(raw_dtype == DT_U16 && out_dtype == DT_U16)) return false;
    if (a.pw.flags & FLAG_DARK_LINEAR) return false;
    if (a.raw_swap || a.raw_gap) return false;
    const size_t esz = dtype_size(raw_dtype);
    if (((size_t)a.W * esz) % 16 || ((size_t)a.W * 4) % 16) return false;
    if (((size_t)a.H * a.W * esz) % 16) return false;
    if (((uintptr_t)a.raw) % 16) return false;
    if (a.dark && ((uintptr_t)a.dark) % 16) return false;
    if (a.flat && ((uintptr_t)a.flat) % 16) return false;
    if (s2_config(a, raw_dtype) < 0) return false;
    return tensor_map_encoder() != nullptr;
}

template <typename RawT, typename OutT, int CFG, int R, int NSTAGE>
static cudaError_t launch_s2_cfg(const K1Args& a0, CUtensorMapDataType rdt, int sm_count, int seg_rows, cudaStream_t st) {
    using S = Stream2Smem<RawT, R, NSTAGE>;
    K1Args a = a0;
    if (CFG & S2_FLAT) a.flat = a.flat_nz;          // zero-free copy: division is unconditional
    CUtensorMap tr, td, tf;
    if (!make_tensor_map(&tr, rdt, sizeof(RawT), a.raw, a.W, a.H, a.n_frames, K2S_BOXW, R)) return cudaErrorInvalidValue;
    if (!make_tensor_map(&td, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, a.dark ? (const void*)a.dark : a.raw, a.W, a.H, 0, K2S_BOXW, R) && a.dark)
        return cudaErrorInvalidValue;
    if (!make_tensor_map(&tf, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, a.flat ? (const void*)a.flat : a.raw, a.W, a.H, 0, K2S_BOXW, R) && a.flat)
        return cudaErrorInvalidValue;
    auto kern = k1_stream2_kernel<RawT, OutT, CFG, R, NSTAGE>;
    static int per_sm = 0;
    if (!per_sm) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::total);
        if (e != cudaSuccess) return e;
        int n = 1;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, K2S_THREADS, S::total);
        per_sm = n < 1 ? 1 : n;
    }
    const int strips = (a.W + K2S_SHIFT - 1 + K2S_STRIP - 1) / K2S_STRIP;      // strip k ends at output column 248k + 240
    const long long slots = (long long)sm_count * per_sm;
    if (seg_rows <= 0) {
        // all units cost the same, so pick the segment height whose unit count fills whole waves of the resident CTAs
        // best, counting the two re-read halo rows per segment against it
        double best = -1.0;
        for (int waves = 1; waves <= 8; ++waves) {
            long long segs_try = slots * waves / ((long long)strips * a.n_frames);
            if (segs_try < 1) continue;
            int rows = (int)((a.H + segs_try - 1) / segs_try);
            rows = ((rows + R - 1) / R) * R;
            if (rows < 2 * R) rows = 2 * R;
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
    kern<<<(unsigned)grid, K2S_THREADS, S::total, st>>>(tr, td, tf, a, strips, segs, seg_rows, (int)total);
    return cudaGetLastError();
}

template <typename RawT, typename OutT>
static cudaError_t launch_s2_t(const K1Args& a, int raw_dtype, CUtensorMapDataType rdt, int sm_count, int seg_rows, cudaStream_t st) {
    constexpr int R = K2S_R_V, NS = K2S_NSTAGE_V;   // R a multiple of 3: the three-row register window returns to its registers per chunk
    switch (s2_config(a, raw_dtype)) {
        case S2_DARK | S2_FLAT: return launch_s2_cfg<RawT, OutT, S2_DARK | S2_FLAT, R, NS>(a, rdt, sm_count, seg_rows, st);
        case S2_DARK | S2_FLAT | S2_N2N: return launch_s2_cfg<RawT, OutT, S2_DARK | S2_FLAT | S2_N2N, R, NS>(a, rdt, sm_count, seg_rows, st);
        case S2_DARK | S2_FLAT | S2_N2N | S2_CHECK:
            return launch_s2_cfg<RawT, OutT, S2_DARK | S2_FLAT | S2_N2N | S2_CHECK, R, NS>(a, rdt, sm_count, seg_rows, st);
        case S2_MASK: return launch_s2_cfg<RawT, OutT, S2_MASK, R, NS>(a, rdt, sm_count, seg_rows, st);
    }
    return cudaErrorInvalidValue;
}

cudaError_t launch_k1_stream2(const K1Args& a, int raw_dtype, int out_dtype, int sm_count, int seg_rows, cudaStream_t st) {
    if (raw_dtype == DT_U16 && out_dtype == DT_F32)
        return launch_s2_t<uint16_t, float>(a, raw_dtype, CU_TENSOR_MAP_DATA_TYPE_UINT16, sm_count, seg_rows, st);
    if (raw_dtype == DT_U16 && out_dtype == DT_U16)
        return launch_s2_t<uint16_t, uint16_t>(a, raw_dtype, CU_TENSOR_MAP_DATA_TYPE_UINT16, sm_count, seg_rows, st);
    if (raw_dtype == DT_F32 && out_dtype == DT_F32)
        return launch_s2_t<float, float>(a, raw_dtype, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, sm_count, seg_rows, st);
    return cudaErrorInvalidValue;
}

}  // namespace imgcorr

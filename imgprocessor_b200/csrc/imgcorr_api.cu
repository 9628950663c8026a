// imgcorr_api.cu — the C ABI declared in include/imgcorr.h: context, calibration upload,
// kernel entry points, the device-resident chain and the host-buffer streaming pipeline.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <algorithm>
#include <vector>
#include <thread>

#include "../../include/imgcorr.h"
#include "imgcorr_kernels.cuh"

using namespace imgcorr;

static thread_local std::string g_err;

static int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
static int cuda_fail(cudaError_t e, const char* what) {
    return fail(IMGCORR_ERR_CUDA, "%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
}
#define CK(call)                                             \
    do {                                                     \
        cudaError_t e__ = (call);                            \
        if (e__ != cudaSuccess) return cuda_fail(e__, #call); \
    } while (0)

struct HostSlot {
    void* d_raw = nullptr;
    void* d_out = nullptr;
    void* h_raw = nullptr;
    void* h_out = nullptr;
    cudaEvent_t ev_in = nullptr, ev_k = nullptr, ev_out = nullptr;
};

struct imgcorr_ctx {
    int device = 0, H = 0, W = 0, sm_count = 148;
    float* dark = nullptr;
    float* ascent = nullptr;
    float* flat = nullptr;
    double exposure = 0.0, maxval = 65535.0;
    bool has_lens = false;
    LensConst lens{};
    double lens_in[23] = {0};                 // K, dist, P as last given to imgcorr_set_lens (an identical call is a no-op)
    double* lens_dev = nullptr;
    float4* k2_wtab = nullptr;                // OpenCV's BilinearTab_f [32][32] (K2 tiles)
    struct GeomKey { int x0, y0, ow, oh, esz, g; };
    std::vector<GeomKey> k2_geom;             // staged-box geometry per output window (k2_pick_geometry), reset by set_lens
    // K2 coordinate cache: packed per-pixel source coordinates of the current lens for an output window (4 bytes per pixel),
    // written by the first tiled launch, read by the following ones; dropped by set_lens
    struct K2Cache { int x0, y0, ow, oh, g, esz; unsigned* pack; int2* hdr; cudaEvent_t ev; cudaStream_t st; unsigned long long stamp; };
    std::vector<K2Cache> k2_cache;
    int k2_cache_on = 1, k2_tma_store = 0;
    unsigned long long k2_stamp = 0;
    void* dump = nullptr;                     // scratch for stores of lanes that own no output pixel
    int raw_big_endian = 0;
    long long raw_gap = 0;
    int k1_variant = 0, k2_variant = 0, k3_variant = 0, host_slots = 4, k1_seg_rows = 0, profile = 0, chain_group = 16;
    long long chain_groups_seen = 0;
    double prof_frames[2] = {0.0, 0.0};
    std::vector<cudaEvent_t> prof_ev[2];      // [kernel] start/stop pairs
    size_t mid_frames = 0;
    bool dark_finite = true, flat_finite = true;
    float* flat_nz = nullptr;                 // flat with zeros replaced by 1.0
    double dark_absmax = 0.0, flat_absmin = 1.0;   // over finite entries (flat: non-zero entries)
    long long launches = 0;
    float* mid[2] = {nullptr, nullptr};
    int chain_overlap = 1;                    // K1 of group g+1 on the internal stream while K2 of group g runs: +2-3.5 % on batches
    cudaStream_t s_chain = nullptr;           // high-priority stream of K1 in overlap mode
    cudaEvent_t ev_k1[2] = {nullptr, nullptr}, ev_k2[2] = {nullptr, nullptr}, ev_entry = nullptr;
    double* ste_avg = nullptr;                // K4 scratch: second running-average buffer, thresholds, counts
    double* ste_thr = nullptr;
    int* ste_n = nullptr;
    int16_t* warp_itab[2] = {nullptr, nullptr};   // [bicubic, Lanczos4] fixed-point weights for uint8 images (K3)
    float* warp_tab = nullptr;                // [32][8] Lanczos4 then [32][4] bicubic coefficient tables (K3)
    // host pipeline
    cudaStream_t s_in = nullptr, s_k = nullptr, s_out = nullptr;
    std::vector<HostSlot> slots;
    size_t slot_raw_bytes = 0, slot_out_bytes = 0;
};

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
        if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
#define GUARD(ctx)                                                           \
    if (!(ctx)) return fail(IMGCORR_ERR_INVALID, "null context");            \
    DeviceGuard guard__((ctx)->device);                                      \
    if (!guard__.ok) return cuda_fail(cudaGetLastError(), "cudaSetDevice")

extern "C" IMGCORR_API const char* imgcorr_last_error(void) { return g_err.c_str(); }
extern "C" IMGCORR_API int imgcorr_version(void) { return IMGCORR_VERSION; }

extern "C" IMGCORR_API int imgcorr_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDeviceCount");
    return n;
}

extern "C" IMGCORR_API int imgcorr_ctx_create(int device, int height, int width, imgcorr_ctx** out_ctx) {
    if (!out_ctx) return fail(IMGCORR_ERR_INVALID, "out_ctx is null");
    *out_ctx = nullptr;
    if (height <= 0 || width <= 0) return fail(IMGCORR_ERR_INVALID, "bad frame shape %d x %d", height, width);
    int n = 0;
    CK(cudaGetDeviceCount(&n));
    if (device < 0 || device >= n) return fail(IMGCORR_ERR_INVALID, "device %d out of range (%d devices)", device, n);
    DeviceGuard g(device);
    if (!g.ok) return cuda_fail(cudaGetLastError(), "cudaSetDevice");
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(IMGCORR_ERR_CUDA, "device %d is sm_%d%d; libimgcorr is built for sm_100a only", device, prop.major,
                    prop.minor);
    imgcorr_ctx* c = new (std::nothrow) imgcorr_ctx();
    if (!c) return fail(IMGCORR_ERR_NOMEM, "out of host memory");
    c->device = device;
    c->H = height;
    c->W = width;
    c->sm_count = prop.multiProcessorCount;
    *out_ctx = c;
    return IMGCORR_OK;
}

static void free_k2_cache(imgcorr_ctx* c) {
    for (auto& k : c->k2_cache) {
        cudaFree(k.pack);
        cudaFree(k.hdr);
        if (k.ev) cudaEventDestroy(k.ev);
    }
    c->k2_cache.clear();
}

static void free_slots(imgcorr_ctx* c) {
    for (auto& s : c->slots) {
        if (s.d_raw) cudaFree(s.d_raw);
        if (s.d_out) cudaFree(s.d_out);
        if (s.h_raw) cudaFreeHost(s.h_raw);
        if (s.h_out) cudaFreeHost(s.h_out);
        if (s.ev_in) cudaEventDestroy(s.ev_in);
        if (s.ev_k) cudaEventDestroy(s.ev_k);
        if (s.ev_out) cudaEventDestroy(s.ev_out);
    }
    c->slots.clear();
    c->slot_raw_bytes = c->slot_out_bytes = 0;
}

extern "C" IMGCORR_API int imgcorr_ctx_destroy(imgcorr_ctx* c) {
    if (!c) return IMGCORR_OK;
    DeviceGuard g(c->device);
    cudaDeviceSynchronize();
    free_slots(c);
    if (c->s_in) cudaStreamDestroy(c->s_in);
    if (c->s_k) cudaStreamDestroy(c->s_k);
    if (c->s_out) cudaStreamDestroy(c->s_out);
    if (c->s_chain) cudaStreamDestroy(c->s_chain);
    for (int i = 0; i < 2; ++i) { if (c->ev_k1[i]) cudaEventDestroy(c->ev_k1[i]); if (c->ev_k2[i]) cudaEventDestroy(c->ev_k2[i]); }
    if (c->ev_entry) cudaEventDestroy(c->ev_entry);
    cudaFree(c->dark);
    cudaFree(c->ascent);
    cudaFree(c->flat);
    cudaFree(c->flat_nz);
    cudaFree(c->mid[0]);
    cudaFree(c->mid[1]);
    cudaFree(c->lens_dev);
    cudaFree(c->k2_wtab);
    free_k2_cache(c);
    cudaFree(c->warp_tab);
    cudaFree(c->warp_itab[0]);
    cudaFree(c->warp_itab[1]);
    cudaFree(c->ste_avg);
    cudaFree(c->ste_thr);
    cudaFree(c->ste_n);
    cudaFree(c->dump);
    for (int k = 0; k < 2; ++k) for (auto e : c->prof_ev[k]) cudaEventDestroy(e);
    delete c;
    return IMGCORR_OK;
}

extern "C" IMGCORR_API int imgcorr_set_option(imgcorr_ctx* c, int key, int value) {
    if (!c) return fail(IMGCORR_ERR_INVALID, "null context");
    switch (key) {
        case IMGCORR_OPT_K1_VARIANT:
            if (value < 0 || value > 3) return fail(IMGCORR_ERR_INVALID, "k1 variant %d", value);
            c->k1_variant = value;
            return IMGCORR_OK;
        case IMGCORR_OPT_K2_VARIANT:
            if (value < 0 || value > 2) return fail(IMGCORR_ERR_INVALID, "k2 variant %d", value);
            c->k2_variant = value;
            return IMGCORR_OK;
        case IMGCORR_OPT_CHAIN_OVERLAP:
            c->chain_overlap = value ? 1 : 0;
            return IMGCORR_OK;
        case IMGCORR_OPT_K2_TMA_STORE:
            c->k2_tma_store = value ? 1 : 0;
            return IMGCORR_OK;
        case IMGCORR_OPT_K2_COORD_CACHE:
            c->k2_cache_on = value ? 1 : 0;
            if (!value && !c->k2_cache.empty()) {
                DeviceGuard g(c->device);
                cudaDeviceSynchronize();
                free_k2_cache(c);
            }
            return IMGCORR_OK;
        case IMGCORR_OPT_K3_VARIANT:
            if (value < 0 || value > 2) return fail(IMGCORR_ERR_INVALID, "k3 variant %d", value);
            c->k3_variant = value;
            return IMGCORR_OK;
        case IMGCORR_OPT_PROFILE:
            if (value < 0) return fail(IMGCORR_ERR_INVALID, "profile stride %d", value);
            c->profile = value;
            return IMGCORR_OK;
        case IMGCORR_OPT_CHAIN_GROUP:
            if (value < 1 || value > 64) return fail(IMGCORR_ERR_INVALID, "chain group %d not in [1,64]", value);
            c->chain_group = value;
            return IMGCORR_OK;
        case IMGCORR_OPT_RAW_BIG_ENDIAN:
            c->raw_big_endian = value ? 1 : 0;
            return IMGCORR_OK;
        case IMGCORR_OPT_RAW_FRAME_GAP:
            if (value < 0) return fail(IMGCORR_ERR_INVALID, "frame gap %d", value);
            c->raw_gap = value;
            return IMGCORR_OK;
        case IMGCORR_OPT_K1_SEG_ROWS:
            if (value < 0) return fail(IMGCORR_ERR_INVALID, "seg rows %d", value);
            c->k1_seg_rows = value;
            return IMGCORR_OK;
        case IMGCORR_OPT_HOST_SLOTS:
            if (value < 2 || value > 64) return fail(IMGCORR_ERR_INVALID, "host slots %d not in [2,64]", value);
            if (value != c->host_slots) {
                DeviceGuard g(c->device);
                cudaDeviceSynchronize();
                free_slots(c);
                c->host_slots = value;
            }
            return IMGCORR_OK;
    }
    return fail(IMGCORR_ERR_INVALID, "unknown option %d", key);
}

extern "C" IMGCORR_API long long imgcorr_launch_count(const imgcorr_ctx* c) { return c ? c->launches : 0; }

static void prof_mark(imgcorr_ctx* c, int which, cudaStream_t st) {
    cudaEvent_t e = nullptr;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    cudaEventRecord(e, st);
    c->prof_ev[which].push_back(e);
}

extern "C" IMGCORR_API int imgcorr_profile_read(imgcorr_ctx* c, double out[4]) {
    GUARD(c);
    if (!out) return fail(IMGCORR_ERR_INVALID, "out is null");
    for (int k = 0; k < 2; ++k) {
        double ms = 0.0;
        auto& v = c->prof_ev[k];
        const size_t pairs = v.size() / 2;
        for (size_t i = 0; i < pairs; ++i) {
            float t = 0.f;
            CK(cudaEventSynchronize(v[2 * i + 1]));
            CK(cudaEventElapsedTime(&t, v[2 * i], v[2 * i + 1]));
            ms += t;
        }
        for (auto e : v) cudaEventDestroy(e);
        v.clear();
        out[2 * k] = ms;
        out[2 * k + 1] = c->prof_frames[k];
        c->prof_frames[k] = 0.0;
    }
    return IMGCORR_OK;
}

static bool all_finite(const float* p, size_t n) {
    // |x| <= FLT_MAX fails for NaN and inf; accumulate without branches
    bool ok = true;
    for (size_t i = 0; i < n; ++i) ok &= (fabsf(p[i]) <= 3.402823466e+38f);
    return ok;
}

static int upload_map(imgcorr_ctx* c, float** slot, const float* src, int on_device, bool* finite,
                      std::vector<float>* host_copy = nullptr) {
    const size_t n = (size_t)c->H * c->W, bytes = n * sizeof(float);
    *finite = true;
    if (!src) {
        if (*slot) { CK(cudaDeviceSynchronize()); CK(cudaFree(*slot)); *slot = nullptr; }
        return IMGCORR_OK;
    }
    if (!*slot) CK(cudaMalloc((void**)slot, bytes));
    else CK(cudaDeviceSynchronize());      // launches on other (non-blocking) streams may still read the map being replaced
    CK(cudaMemcpy(*slot, src, bytes, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice));
    std::vector<float> tmp;
    const float* h = src;
    if (on_device || host_copy) {
        std::vector<float>& dst = host_copy ? *host_copy : tmp;
        dst.resize(n);
        if (on_device) CK(cudaMemcpy(dst.data(), *slot, bytes, cudaMemcpyDeviceToHost));
        else memcpy(dst.data(), src, bytes);
        h = dst.data();
    }
    *finite = all_finite(h, n);
    return IMGCORR_OK;
}

extern "C" IMGCORR_API int imgcorr_set_dark(imgcorr_ctx* c, const float* dark, const float* ascent, double exposure_time,
                                int depth_bits, int on_device) {
    GUARD(c);
    if (!dark && ascent) return fail(IMGCORR_ERR_INVALID, "ascent given without offset map");
    if (ascent && (depth_bits < 1 || depth_bits > 62)) return fail(IMGCORR_ERR_INVALID, "depth_bits %d", depth_bits);
    if (ascent && !(exposure_time == exposure_time)) return fail(IMGCORR_ERR_INVALID, "exposure_time is NaN");
    bool fin2 = true;
    std::vector<float> h;
    int r = upload_map(c, &c->dark, dark, on_device, &c->dark_finite, &h);
    if (r) return r;
    c->dark_absmax = 0.0;
    for (float v : h) { const float av = fabsf(v); if (av <= 3.402823466e+38f && av > c->dark_absmax) c->dark_absmax = av; }
    r = upload_map(c, &c->ascent, ascent, on_device, &fin2);
    if (r) return r;
    c->exposure = exposure_time;
    c->maxval = ascent ? ldexp(1.0, depth_bits) - 1.0 : 65535.0;
    return IMGCORR_OK;
}

extern "C" IMGCORR_API int imgcorr_set_flat(imgcorr_ctx* c, const float* flat, int on_device) {
    GUARD(c);
    std::vector<float> h;
    int r = upload_map(c, &c->flat, flat, on_device, &c->flat_finite, &h);
    if (r) return r;
    if (!flat) {
        if (c->flat_nz) { CK(cudaFree(c->flat_nz)); c->flat_nz = nullptr; }
        return IMGCORR_OK;
    }
    // zero-free copy for the streaming kernel + the smallest non-zero magnitude (overflow proof)
    double mn = 3.5e38;
    for (float& v : h) {
        if (v == 0.0f) v = 1.0f;
        const float av = fabsf(v);
        if (av <= 3.402823466e+38f && av < mn) mn = av;
    }
    c->flat_absmin = mn;
    if (!c->flat_nz) CK(cudaMalloc((void**)&c->flat_nz, h.size() * sizeof(float)));
    else CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(c->flat_nz, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice));
    return IMGCORR_OK;
}

extern "C" IMGCORR_API int imgcorr_set_lens(imgcorr_ctx* c, const double K[9], const double dist[5], const double P[9]) {
    if (!c) return fail(IMGCORR_ERR_INVALID, "null context");
    if (!K) { c->has_lens = false; return IMGCORR_OK; }
    if (!dist || !P) return fail(IMGCORR_ERR_INVALID, "dist / P is null");
    {
        // the Python mirror sets the lens on every correct() call (the reference builds a new LensDistortion per call,
        // camera/CameraCalibration.py:565-568): the same lens again must not drop K2's geometry choice and coordinate cache
        double in[23];
        memcpy(in, K, 9 * sizeof(double)); memcpy(in + 9, dist, 5 * sizeof(double)); memcpy(in + 14, P, 9 * sizeof(double));
        if (c->has_lens && c->lens_dev && memcmp(in, c->lens_in, sizeof in) == 0) return IMGCORR_OK;
    }
    LensConst L{};
    if (!invert3x3(P, L.ir)) return fail(IMGCORR_ERR_INVALID, "new camera matrix P is singular");
    L.k1 = dist[0]; L.k2 = dist[1]; L.p1 = dist[2]; L.p2 = dist[3]; L.k3 = dist[4];
    L.p1x2 = L.p1 + L.p1; L.p2x2 = L.p2 + L.p2;
    L.fx = K[0]; L.fy = K[4]; L.cx = K[2]; L.cy = K[5];
    L.affine = (L.ir[6] == 0.0 && L.ir[7] == 0.0 && L.ir[8] == 1.0) ? 1 : 0;
    double pack[LP_COUNT] = {0};
    pack[LP_K1] = L.k1; pack[LP_K2] = L.k2; pack[LP_K3] = L.k3; pack[LP_P1] = L.p1; pack[LP_P2] = L.p2;
    pack[LP_P1X2] = L.p1x2; pack[LP_P2X2] = L.p2x2; pack[LP_FX] = L.fx; pack[LP_FY] = L.fy; pack[LP_CX] = L.cx;
    pack[LP_CY] = L.cy; pack[LP_IR0] = L.ir[0]; pack[LP_IR2] = L.ir[2]; pack[LP_IR4] = L.ir[4]; pack[LP_IR5] = L.ir[5];
    {
        DeviceGuard g(c->device);
        if (!g.ok) return cuda_fail(cudaGetLastError(), "cudaSetDevice");
        if (!c->lens_dev) CK(cudaMalloc((void**)&c->lens_dev, sizeof pack));
        CK(cudaMemcpy(c->lens_dev, pack, sizeof pack, cudaMemcpyHostToDevice));
    }
    c->lens = L;
    c->has_lens = true;
    memcpy(c->lens_in, K, 9 * sizeof(double)); memcpy(c->lens_in + 9, dist, 5 * sizeof(double)); memcpy(c->lens_in + 14, P, 9 * sizeof(double));
    c->k2_geom.clear();
    if (!c->k2_cache.empty()) {
        DeviceGuard g(c->device);
        cudaDeviceSynchronize();                       // launches that still read the old lens' coordinates
        free_k2_cache(c);
    }
    return IMGCORR_OK;
}

static int fill_k1(imgcorr_ctx* c, K1Args& a, const void* raw, int raw_dtype, void* out, uint8_t* mask, int n, double thr,
                   int ksize, int cond, int flags) {
    if (!raw || !out) return fail(IMGCORR_ERR_INVALID, "null image pointer");
    if (n < 0) return fail(IMGCORR_ERR_INVALID, "n_frames %d", n);
    if (!(thr > 0.0)) ksize = 0;                      // also NaN
    if (ksize != 0 && ksize != 3 && ksize != 5) return fail(IMGCORR_ERR_INVALID, "ksize %d (0, 3 or 5)", ksize);
    if (cond != IMGCORR_COND_GT && cond != IMGCORR_COND_LT) return fail(IMGCORR_ERR_INVALID, "cond %d", cond);
    a = K1Args{};
    a.raw = raw; a.out = out; a.mask = ksize ? mask : nullptr;
    a.H = c->H; a.W = c->W; a.n_frames = n; a.ksize = ksize;
    a.raw_swap = (c->raw_big_endian && raw_dtype == DT_U16) ? 1 : 0;
    a.raw_gap = c->raw_gap;
    if (c->raw_big_endian && raw_dtype != DT_U16 && raw_dtype != DT_U8)
        return fail(IMGCORR_ERR_INVALID, "big-endian ingest is implemented for uint16 frames");
    if (c->raw_gap % (long long)dtype_size(raw_dtype)) return fail(IMGCORR_ERR_INVALID, "frame gap must be a multiple of the sample size");
    int f = 0;
    if ((flags & IMGCORR_DO_DARK) && c->dark) {
        f |= FLAG_DARK; a.dark = c->dark;
        if (c->ascent) { f |= FLAG_DARK_LINEAR; a.ascent = c->ascent; }
    }
    if ((flags & IMGCORR_DO_FLAT) && c->flat) { f |= FLAG_FLAT; a.flat = c->flat; }
    if (flags & IMGCORR_DO_NAN_TO_NUM) f |= FLAG_NAN_TO_NUM;
    a.pw.flags = f; a.pw.exposure_time = c->exposure; a.pw.max_value = c->maxval;
    a.maps_finite = ((!a.dark || c->dark_finite) && (!a.flat || c->flat_finite)) ? 1 : 0;
    a.flat_nz = a.flat ? c->flat_nz : nullptr;
    if (!c->dump) CK(cudaMalloc(&c->dump, (size_t)4 << 20));
    a.dump = c->dump;
    // integer frames: |raw - dark| <= raw_max + max|dark|; divided by min|flat| it must stay below FLT_MAX
    a.no_overflow = 0;
    if (a.maps_finite && (raw_dtype == DT_U8 || raw_dtype == DT_U16)) {
        const double num = (raw_dtype == DT_U8 ? 255.0 : 65535.0) + (a.dark ? c->dark_absmax : 0.0);
        const double den = a.flat ? c->flat_absmin : 1.0;
        a.no_overflow = (num / den < 3.0e38) ? 1 : 0;
    }
    a.pred = make_predicate(thr, cond == IMGCORR_COND_LT ? COND_LT : COND_GT);
    return IMGCORR_OK;
}

extern "C" IMGCORR_API int imgcorr_pointwise_median(imgcorr_ctx* c, const void* raw_dev, int raw_dtype, void* out_dev, int out_dtype,
                                        uint8_t* mask_dev, int n_frames, double threshold, int ksize, int cond,
                                        int flags, void* stream) {
    GUARD(c);
    K1Args a;
    int r = fill_k1(c, a, raw_dev, raw_dtype, out_dev, mask_dev, n_frames, threshold, ksize, cond, flags);
    if (r) return r;
    a.out_streaming = 1;
    int l = 0;
    cudaError_t e = launch_k1(a, raw_dtype, out_dtype, c->k1_variant, c->sm_count, c->k1_seg_rows, (cudaStream_t)stream, &l);
    c->launches += l;
    if (e == cudaErrorInvalidValue && l == 0)
        return fail(IMGCORR_ERR_INVALID, "unsupported dtype pair raw=%d out=%d", raw_dtype, out_dtype);
    if (e == cudaErrorNotSupported) return fail(IMGCORR_ERR_INVALID, "requested K1 variant is not eligible for this shape / dtype / alignment");
    if (e != cudaSuccess) return cuda_fail(e, "K1 launch");
    return IMGCORR_OK;
}

static int run_k2(imgcorr_ctx* c, const void* src, int sdt, void* dst, int ddt, int n, const float* mapx,
                  const float* mapy, double border, int x0, int y0, int ow, int oh, cudaStream_t st) {
    if (!src || !dst) return fail(IMGCORR_ERR_INVALID, "null image pointer");
    if (n < 0) return fail(IMGCORR_ERR_INVALID, "n_frames %d", n);
    if (!mapx && !c->has_lens) return fail(IMGCORR_ERR_STATE, "no lens set");
    if (x0 < 0 || y0 < 0 || ow < 0 || oh < 0 || x0 + ow > c->W || y0 + oh > c->H)
        return fail(IMGCORR_ERR_INVALID, "output window (%d,%d,%d,%d) outside %dx%d frame", x0, y0, ow, oh, c->W, c->H);
    K2Args a{};
    a.src = src; a.dst = dst; a.mapx = mapx; a.mapy = mapy;
    a.H = c->H; a.W = c->W; a.n_frames = n;
    a.x0 = x0; a.y0 = y0; a.ow = ow; a.oh = oh;
    a.border = border_for_dtype(sdt == DT_U8, sdt == DT_U16, border);
    a.lens = c->lens;
    a.lens_dev = c->lens_dev;
    a.sm_count = c->sm_count;
    a.tma_store = c->k2_tma_store;
    if (!c->k2_wtab) {
        std::vector<float4> t(32 * 32);
        for (int fy = 0; fy < 32; ++fy)
            for (int fx = 0; fx < 32; ++fx) {
                float4 w;
                bilinear_weights(fx, fy, w.x, w.y, w.z, w.w);
                t[fy * 32 + fx] = w;
            }
        CK(cudaMalloc((void**)&c->k2_wtab, t.size() * sizeof(float4)));
        CK(cudaMemcpy(c->k2_wtab, t.data(), t.size() * sizeof(float4), cudaMemcpyHostToDevice));
    }
    a.wtab = c->k2_wtab;
    a.geometry = 0;
    if (!mapx) {
        bool found = false;
        const int gesz = (int)dtype_size(sdt) > 4 ? 4 : (int)dtype_size(sdt);
        for (const auto& k : c->k2_geom)
            if (k.x0 == x0 && k.y0 == y0 && k.ow == ow && k.oh == oh && k.esz == gesz) { a.geometry = k.g; found = true; break; }
        if (!found) {
            a.geometry = k2_pick_geometry(c->lens, c->H, c->W, x0, y0, ow, oh, gesz);
            if (c->k2_geom.size() > 64) c->k2_geom.clear();
            c->k2_geom.push_back({x0, y0, ow, oh, gesz, a.geometry});
        }
    }
    // coordinate cache (tiled variant, analytic map): the first launch for this lens / window writes it, later ones read it
    imgcorr_ctx::K2Cache* building = nullptr;
    // Only launches of one or two frames use it: a batch amortises the analytic evaluation over its frames and the
    // per-tile kernel then streams faster (19 vs 33 us per 4096x3000 frame) than the cached one (49 vs 65 us for one frame).
    if (!mapx && c->k2_cache_on && n > 0 && n <= 2 && ow > 0 && oh > 0 && k2_will_tile(a, sdt, ddt, c->k2_variant)) {
        const int esz = (int)dtype_size(sdt);
        imgcorr_ctx::K2Cache* hit = nullptr;
        for (auto& k : c->k2_cache)
            if (k.x0 == x0 && k.y0 == y0 && k.ow == ow && k.oh == oh && k.g == a.geometry && k.esz == esz) { hit = &k; break; }
        if (hit) {
            if (hit->st != st) CK(cudaStreamWaitEvent(st, hit->ev, 0));       // written on another stream
            hit->stamp = ++c->k2_stamp;
            a.cpack = hit->pack; a.chdr = hit->hdr;
        } else {
            if (c->k2_cache.size() >= 2) {             // keep at most two windows (keep_size True / False): drop the older one
                size_t old = c->k2_cache[0].stamp < c->k2_cache[1].stamp ? 0 : 1;
                CK(cudaDeviceSynchronize());
                cudaFree(c->k2_cache[old].pack); cudaFree(c->k2_cache[old].hdr);
                if (c->k2_cache[old].ev) cudaEventDestroy(c->k2_cache[old].ev);
                c->k2_cache.erase(c->k2_cache.begin() + old);
            }
            size_t words = 0, tiles = 0;
            k2_cache_size(a.geometry, ow, oh, &words, &tiles);
            imgcorr_ctx::K2Cache k{x0, y0, ow, oh, a.geometry, esz, nullptr, nullptr, nullptr, st, ++c->k2_stamp};
            if (cudaMalloc((void**)&k.pack, words * sizeof(unsigned)) == cudaSuccess &&
                cudaMalloc((void**)&k.hdr, tiles * sizeof(int2)) == cudaSuccess &&
                cudaEventCreateWithFlags(&k.ev, cudaEventDisableTiming) == cudaSuccess) {
                c->k2_cache.push_back(k);
                building = &c->k2_cache.back();
                a.cpack_w = k.pack; a.chdr_w = k.hdr;
            } else {                                   // no memory for the cache: compute every time
                cudaGetLastError();
                cudaFree(k.pack); cudaFree(k.hdr);
                if (k.ev) cudaEventDestroy(k.ev);
            }
        }
    }
    int l = 0;
    cudaError_t e = launch_k2(a, sdt, ddt, c->k2_variant, st, &l);
    if (building) {
        if (e == cudaSuccess) e = cudaEventRecord(building->ev, st);
        if (e != cudaSuccess) {                        // never leave a half-written cache behind
            cudaDeviceSynchronize();
            cudaFree(building->pack); cudaFree(building->hdr); cudaEventDestroy(building->ev);
            c->k2_cache.pop_back();
        }
    }
    c->launches += l;
    if (e == cudaErrorInvalidValue && l == 0)
        return fail(IMGCORR_ERR_INVALID, "unsupported dtype pair src=%d dst=%d", sdt, ddt);
    if (e == cudaErrorNotSupported) return fail(IMGCORR_ERR_INVALID, "requested K2 variant is not eligible for this dtype / alignment");
    if (e != cudaSuccess) return cuda_fail(e, "K2 launch");
    return IMGCORR_OK;
}

extern "C" IMGCORR_API int imgcorr_undistort(imgcorr_ctx* c, const void* src_dev, int src_dtype, void* dst_dev, int dst_dtype,
                                 int n_frames, double border_value, int x0, int y0, int ow, int oh, void* stream) {
    GUARD(c);
    return run_k2(c, src_dev, src_dtype, dst_dev, dst_dtype, n_frames, nullptr, nullptr, border_value, x0, y0, ow, oh,
                  (cudaStream_t)stream);
}

extern "C" IMGCORR_API int imgcorr_remap(imgcorr_ctx* c, const void* src_dev, int src_dtype, void* dst_dev, int dst_dtype,
                             int n_frames, const float* mapx_dev, const float* mapy_dev, double border_value,
                             void* stream) {
    GUARD(c);
    if (!mapx_dev || !mapy_dev) return fail(IMGCORR_ERR_INVALID, "null map pointer");
    return run_k2(c, src_dev, src_dtype, dst_dev, dst_dtype, n_frames, mapx_dev, mapy_dev, border_value, 0, 0, c->W,
                  c->H, (cudaStream_t)stream);
}

extern "C" IMGCORR_API int imgcorr_undistort_maps(imgcorr_ctx* c, float* mapx_dev, float* mapy_dev, void* stream) {
    GUARD(c);
    if (!c->has_lens) return fail(IMGCORR_ERR_STATE, "no lens set");
    if (!mapx_dev || !mapy_dev) return fail(IMGCORR_ERR_INVALID, "null map pointer");
    int l = 0;
    cudaError_t e = launch_write_maps(c->lens, mapx_dev, mapy_dev, c->H, c->W, (cudaStream_t)stream, &l);
    c->launches += l;
    if (e != cudaSuccess) return cuda_fail(e, "map kernel launch");
    return IMGCORR_OK;
}

// ---- K3 -------------------------------------------------------------------------------------
extern "C" IMGCORR_API int imgcorr_warp_perspective(imgcorr_ctx* c, const void* src_dev, int dtype, int src_h, int src_w,
                                        void* dst_dev, int dst_h, int dst_w, int n_frames, const double M[9],
                                        int interpolation, int inverse_map, double border_value, void* stream) {
    GUARD(c);
    if (!src_dev || !dst_dev || !M) return fail(IMGCORR_ERR_INVALID, "null pointer");
    if (dtype < DT_U8 || dtype > DT_F64) return fail(IMGCORR_ERR_INVALID, "bad dtype %d", dtype);
    if (interpolation != IMGCORR_INTER_LANCZOS4 && interpolation != IMGCORR_INTER_CUBIC)
        return fail(IMGCORR_ERR_INVALID, "interpolation must be IMGCORR_INTER_LANCZOS4 or IMGCORR_INTER_CUBIC");
    if (src_h <= 0 || src_w <= 0 || dst_h <= 0 || dst_w <= 0 || src_h > 32767 || src_w > 32767 || dst_h > 32767 || dst_w > 32767)
        return fail(IMGCORR_ERR_INVALID, "bad frame shape (%d x %d -> %d x %d)", src_h, src_w, dst_h, dst_w);
    if (n_frames < 0) return fail(IMGCORR_ERR_INVALID, "n_frames < 0");
    if (!c->warp_tab) {
        float h[32 * 8 + 32 * 4];
        warp_lanczos4_table(h);
        warp_cubic_table(h + 32 * 8);
        CK(cudaMalloc(&c->warp_tab, sizeof(h)));
        CK(cudaMemcpy(c->warp_tab, h, sizeof(h), cudaMemcpyHostToDevice));
    }
    const int lz = interpolation == IMGCORR_INTER_LANCZOS4;
    if (dtype == DT_U8 && !c->warp_itab[lz]) {
        // OpenCV's fixed-point weight table for uint8 images, built on the host once per interpolation
        const int nn = lz ? 8 : 4;
        std::vector<float> t(32 * 8);
        if (lz) warp_lanczos4_table(t.data()); else warp_cubic_table(t.data());
        std::vector<int16_t> it((size_t)32 * 32 * nn * nn);
        warp_fixed_table(t.data(), nn, it.data());
        CK(cudaMalloc(&c->warp_itab[lz], it.size() * sizeof(int16_t)));
        CK(cudaMemcpy(c->warp_itab[lz], it.data(), it.size() * sizeof(int16_t), cudaMemcpyHostToDevice));
    }
    K3Args a;
    a.itab = dtype == DT_U8 ? c->warp_itab[lz] : nullptr;
    a.src = src_dev;
    a.dst = dst_dev;
    a.H = src_h;
    a.W = src_w;
    a.dh = dst_h;
    a.dw = dst_w;
    a.n_frames = n_frames;
    a.border = border_for_dtype(dtype == DT_U8, dtype == DT_U16, border_value);
    a.wc = make_warp_const(M, inverse_map, dst_w, dst_h);
    a.tab = interpolation == IMGCORR_INTER_LANCZOS4 ? c->warp_tab : c->warp_tab + 32 * 8;
    int l = 0;
    cudaError_t e = launch_k3(a, dtype, interpolation, c->k3_variant, (cudaStream_t)stream, &l);
    c->launches += l;
    if (e == cudaErrorNotSupported) return fail(IMGCORR_ERR_INVALID, "the requested K3 variant is not eligible for this call");
    if (e != cudaSuccess) return cuda_fail(e, "K3 launch");
    return IMGCORR_OK;
}

extern "C" IMGCORR_API int imgcorr_divide_f64(imgcorr_ctx* c, const void* src_dev, int src_dtype, const double* divisor_dev,
                                  double* dst_dev, size_t pixels_per_frame, int n_frames, void* stream) {
    GUARD(c);
    if (!src_dev || !divisor_dev || !dst_dev) return fail(IMGCORR_ERR_INVALID, "null pointer");
    if (src_dtype < DT_U8 || src_dtype > DT_F64) return fail(IMGCORR_ERR_INVALID, "bad dtype %d", src_dtype);
    if (n_frames < 0) return fail(IMGCORR_ERR_INVALID, "n_frames < 0");
    int l = 0;
    cudaError_t e = launch_k3_divide(src_dev, src_dtype, divisor_dev, dst_dev, pixels_per_frame, n_frames, c->sm_count,
                                     (cudaStream_t)stream, &l);
    c->launches += l;
    if (e != cudaSuccess) return cuda_fail(e, "divide kernel launch");
    return IMGCORR_OK;
}

// ---- K4 -------------------------------------------------------------------------------------
static int ste_average_impl(imgcorr_ctx* c, const void* frames_dev, int dtype, int n_frames, double* avg_dev,
                            uint8_t* mask_dev, const double nlf[3], double n_std, const double* thr_dev, void* stream);

extern "C" IMGCORR_API int imgcorr_ste_average(imgcorr_ctx* c, const void* frames_dev, int dtype, int n_frames, double* avg_dev,
                                   uint8_t* mask_dev, const double nlf[3], double n_std, void* stream) {
    GUARD(c);
    if (!nlf) return fail(IMGCORR_ERR_INVALID, "null pointer");
    return ste_average_impl(c, frames_dev, dtype, n_frames, avg_dev, mask_dev, nlf, n_std, nullptr, stream);
}

extern "C" IMGCORR_API int imgcorr_ste_average_thr(imgcorr_ctx* c, const void* frames_dev, int dtype, int n_frames, double* avg_dev,
                                       uint8_t* mask_dev, const double* threshold_dev, void* stream) {
    GUARD(c);
    if (!threshold_dev) return fail(IMGCORR_ERR_INVALID, "null pointer");
    const double none[3] = {0.0, 0.0, 0.0};
    return ste_average_impl(c, frames_dev, dtype, n_frames, avg_dev, mask_dev, none, 1.0, threshold_dev, stream);
}

static int ste_average_impl(imgcorr_ctx* c, const void* frames_dev, int dtype, int n_frames, double* avg_dev,
                            uint8_t* mask_dev, const double nlf[3], double n_std, const double* thr_dev, void* stream) {
    if (!frames_dev || !avg_dev || !nlf) return fail(IMGCORR_ERR_INVALID, "null pointer");
    if (dtype < DT_U8 || dtype > DT_F64) return fail(IMGCORR_ERR_INVALID, "bad dtype %d", dtype);
    if (n_frames < 2) return fail(IMGCORR_ERR_INVALID, "single-time-effect removal needs at least 2 images (got %d)", n_frames);
    const size_t npx = (size_t)c->H * c->W;
    if (!c->ste_avg) {
        CK(cudaMalloc(&c->ste_avg, npx * sizeof(double)));
        CK(cudaMalloc(&c->ste_thr, npx * sizeof(double)));
        CK(cudaMalloc(&c->ste_n, npx * sizeof(int)));
    }
    cudaStream_t st = (cudaStream_t)stream;
    // n_frames - 1 launches; ping-pong so that the last one writes avg_dev
    double* buf[2] = {avg_dev, c->ste_avg};
    int cur = (n_frames - 2) % 2;                  // buffer the FIRST launch writes
    K4Args a;
    a.H = c->H;
    a.W = c->W;
    a.thr = c->ste_thr;
    a.thr_in = thr_dev;
    a.n = c->ste_n;
    a.mask = mask_dev;
    a.sc.minY = nlf[0];
    a.sc.ax = nlf[1];
    a.sc.ay = nlf[2];
    a.sc.nstd = n_std;
    const char* f = (const char*)frames_dev;
    const size_t fb = npx * dtype_size(dtype);
    int l = 0;
    cudaError_t e = cudaSuccess;
    for (int k = 1; k < n_frames && e == cudaSuccess; ++k) {
        a.img = k == 1 ? f : f + k * fb;
        a.img2 = k == 1 ? f + fb : nullptr;
        a.avg_in = buf[cur ^ 1];
        a.avg_out = buf[cur];
        e = launch_k4(a, dtype, st, &l);
        cur ^= 1;
    }
    c->launches += l;
    if (e != cudaSuccess) return cuda_fail(e, "K4 launch");
    return IMGCORR_OK;
}

// one frame (or a group of frames) through K1 -> K2 on `st`
static int chain_frames(imgcorr_ctx* c, const void* raw, int raw_dtype, void* out, int out_dtype, int n, double thr,
                        int ksize, int flags, bool lens, double border, int x0, int y0, int ow, int oh, cudaStream_t st) {
    if (out_dtype != DT_F32 && out_dtype != DT_F64) return fail(IMGCORR_ERR_INVALID, "out_dtype must be F32 or F64");
    if (raw_dtype == DT_F64) return fail(IMGCORR_ERR_INVALID, "float64 frames: convert to float32 first (the chain computes in float32)");
    const size_t npx = (size_t)c->H * c->W;
    const size_t raw_stride = npx * dtype_size(raw_dtype) + (size_t)c->raw_gap;
    if (!lens) {
        if (x0 != 0 || y0 != 0 || ow != c->W || oh != c->H)
            return fail(IMGCORR_ERR_INVALID, "output window without a lens");
        K1Args a;
        int r = fill_k1(c, a, raw, raw_dtype, out, nullptr, n, thr, ksize, IMGCORR_COND_GT, flags);
        if (r) return r;
        int l = 0;
        cudaError_t e = launch_k1(a, raw_dtype, out_dtype, c->k1_variant, c->sm_count, c->k1_seg_rows, st, &l);
        c->launches += l;
        if (e != cudaSuccess) return cuda_fail(e, "K1 launch");
        return IMGCORR_OK;
    }
    // K1 -> K2 in groups of `chain_group` frames: the K1 output of a group stays in two alternating scratch
    // buffers (L2-resident for small groups); K2 evaluates each pixel's map once per group.
    const int grp = c->chain_group < n ? c->chain_group : (n > 0 ? n : 1);
    if (c->mid_frames < (size_t)grp) {
        CK(cudaStreamSynchronize(st));
        for (int i = 0; i < 2; ++i) {
            if (c->mid[i]) { CK(cudaFree(c->mid[i])); c->mid[i] = nullptr; }
            CK(cudaMalloc((void**)&c->mid[i], npx * sizeof(float) * grp));
        }
        c->mid_frames = grp;
    }
    const size_t out_stride = (size_t)ow * oh * dtype_size(out_dtype);
    // Overlap mode: K1 of group g+1 runs on an internal high-priority stream while K2 of group g runs on the caller's
    // stream, so that K2's blocks fill the SM slots K1's last, partially filled wave leaves idle (and vice versa).  The
    // two scratch buffers are handed back and forth with events; every K2 stays on the caller's stream, so the call is
    // stream-ordered as before.  A profiled group (IMGCORR_OPT_PROFILE) runs unoverlapped so that its brackets time the
    // kernel alone.
    const bool overlap = c->chain_overlap && n > grp;
    cudaStream_t s1 = st;
    if (overlap) {
        if (!c->s_chain) {
            int lo = 0, hi = 0;
            CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
            CK(cudaStreamCreateWithPriority(&c->s_chain, cudaStreamNonBlocking, hi));
            for (int i = 0; i < 2; ++i) {
                CK(cudaEventCreateWithFlags(&c->ev_k1[i], cudaEventDisableTiming));
                CK(cudaEventCreateWithFlags(&c->ev_k2[i], cudaEventDisableTiming));
            }
            CK(cudaEventCreateWithFlags(&c->ev_entry, cudaEventDisableTiming));
        }
        s1 = c->s_chain;
        CK(cudaEventRecord(c->ev_entry, st));          // K1 must not start before the caller's earlier work on `st`
        CK(cudaStreamWaitEvent(s1, c->ev_entry, 0));
    }
    int gi = 0;
    bool exclusive_next = false;
    for (int f = 0; f < n; f += grp, ++gi) {
        const int nf = n - f < grp ? n - f : grp;
        float* mid = c->mid[gi & 1];
        const bool prof = c->profile > 0 && (c->chain_groups_seen++ % c->profile) == 0;
        if (prof) { c->prof_frames[0] += nf; c->prof_frames[1] += nf; }
        if (overlap) {
            // this buffer was last read by K2 of group gi-2; a profiled group (or the one after it) also waits for gi-1
            if (gi >= 2) CK(cudaStreamWaitEvent(s1, c->ev_k2[gi & 1], 0));
            if ((prof || exclusive_next) && gi >= 1) CK(cudaStreamWaitEvent(s1, c->ev_k2[(gi - 1) & 1], 0));
            exclusive_next = prof;
        }
        K1Args a;
        int r = fill_k1(c, a, (const char*)raw + f * raw_stride, raw_dtype, mid, nullptr, nf, thr, ksize, IMGCORR_COND_GT, flags);
        if (r) return r;
        int l = 0;
        if (prof) prof_mark(c, 0, s1);
        cudaError_t e = launch_k1(a, raw_dtype, DT_F32, c->k1_variant, c->sm_count, c->k1_seg_rows, s1, &l);
        if (prof) prof_mark(c, 0, s1);
        c->launches += l;
        if (e != cudaSuccess) return cuda_fail(e, "K1 launch");
        if (overlap) {
            CK(cudaEventRecord(c->ev_k1[gi & 1], s1));
            CK(cudaStreamWaitEvent(st, c->ev_k1[gi & 1], 0));
        }
        if (prof) prof_mark(c, 1, st);
        r = run_k2(c, mid, DT_F32, (char*)out + f * out_stride, out_dtype, nf, nullptr, nullptr, border, x0, y0, ow, oh, st);
        if (prof) prof_mark(c, 1, st);
        if (r) return r;
        if (overlap) CK(cudaEventRecord(c->ev_k2[gi & 1], st));
    }
    return IMGCORR_OK;
}

extern "C" IMGCORR_API int imgcorr_correct_batch(imgcorr_ctx* c, const void* raw_dev, int raw_dtype, void* out_dev, int out_dtype,
                                     int n_frames, double threshold, int ksize, int flags, int use_lens,
                                     double border_value, int x0, int y0, int ow, int oh, void* stream) {
    GUARD(c);
    if (!raw_dev || !out_dev) return fail(IMGCORR_ERR_INVALID, "null image pointer");
    const bool lens = use_lens && c->has_lens;
    if (!lens) { x0 = 0; y0 = 0; ow = c->W; oh = c->H; }
    return chain_frames(c, raw_dev, raw_dtype, out_dev, out_dtype, n_frames, threshold, ksize, flags, lens, border_value,
                        x0, y0, ow, oh, (cudaStream_t)stream);
}

// ---- host-buffer pipeline ---------------------------------------------------------------------
static bool is_pinned(const void* p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
}

static int ensure_pipeline(imgcorr_ctx* c, size_t raw_bytes, size_t out_bytes) {
    if (!c->s_in) {
        CK(cudaStreamCreateWithFlags(&c->s_in, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&c->s_k, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&c->s_out, cudaStreamNonBlocking));
    }
    if ((int)c->slots.size() == c->host_slots && c->slot_raw_bytes >= raw_bytes && c->slot_out_bytes >= out_bytes)
        return IMGCORR_OK;
    CK(cudaDeviceSynchronize());
    free_slots(c);
    c->slots.resize(c->host_slots);
    for (auto& s : c->slots) {
        CK(cudaMalloc(&s.d_raw, raw_bytes));
        CK(cudaMalloc(&s.d_out, out_bytes));
        CK(cudaEventCreateWithFlags(&s.ev_in, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&s.ev_k, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&s.ev_out, cudaEventDisableTiming));
    }
    c->slot_raw_bytes = raw_bytes;
    c->slot_out_bytes = out_bytes;
    return IMGCORR_OK;
}

extern "C" IMGCORR_API int imgcorr_correct_host(imgcorr_ctx* c, const void* raw_host, int raw_dtype, void* out_host, int out_dtype,
                                    int n_frames, double threshold, int ksize, int flags, int use_lens,
                                    double border_value, int x0, int y0, int ow, int oh) {
    GUARD(c);
    if (!raw_host || !out_host) return fail(IMGCORR_ERR_INVALID, "null image pointer");
    if (n_frames < 0) return fail(IMGCORR_ERR_INVALID, "n_frames %d", n_frames);
    if (raw_dtype < DT_U8 || raw_dtype > DT_F32) return fail(IMGCORR_ERR_INVALID, "raw_dtype %d (U8, U16 or F32)", raw_dtype);
    if (c->raw_gap) return fail(IMGCORR_ERR_INVALID, "IMGCORR_OPT_RAW_FRAME_GAP is not available for host frames (pass the pixel blocks)");
    if (out_dtype != DT_F32 && out_dtype != DT_F64) return fail(IMGCORR_ERR_INVALID, "out_dtype must be F32 or F64");
    const bool lens = use_lens && c->has_lens;
    if (!lens) { x0 = 0; y0 = 0; ow = c->W; oh = c->H; }
    const size_t frame_raw = (size_t)c->H * c->W * dtype_size(raw_dtype);
    const size_t frame_out = (size_t)ow * oh * dtype_size(out_dtype);
    // small frames travel in chunks of G frames per ring slot (up to 32 MB of raw data, at most 16 frames): one copy each
    // way and one K1 / K2 launch per chunk instead of per frame; frames above 16 MB (the 4096x3000 workload) keep one
    // frame per slot
    int G = 1;
    if (frame_raw <= ((size_t)16 << 20)) G = (int)std::min<size_t>(16, ((size_t)32 << 20) / frame_raw);
    if (G > n_frames) G = n_frames > 0 ? n_frames : 1;
    const size_t raw_bytes = frame_raw * G, out_bytes = frame_out * G;
    int r = ensure_pipeline(c, raw_bytes, out_bytes);
    if (r) return r;
    const bool in_pinned = is_pinned(raw_host), out_pinned = is_pinned(out_host);
    for (auto& s : c->slots) {
        if (!in_pinned && !s.h_raw) CK(cudaHostAlloc(&s.h_raw, c->slot_raw_bytes, cudaHostAllocDefault));
        if (!out_pinned && !s.h_out) CK(cudaHostAlloc(&s.h_out, c->slot_out_bytes, cudaHostAllocDefault));
    }
    const int ns = (int)c->slots.size();
    const int chunks = (n_frames + G - 1) / G;
    // whatever happens below, no copy from / into the caller's buffers may still be in flight when this call returns
    struct Drain {
        imgcorr_ctx* c; bool armed = true;
        ~Drain() { if (armed) { cudaStreamSynchronize(c->s_in); cudaStreamSynchronize(c->s_k); cudaStreamSynchronize(c->s_out); } }
    } drain{c};
    auto frames_of = [&](int k) { return n_frames - k * G < G ? n_frames - k * G : G; };
    auto retire = [&](int k) -> int {          // chunk k's D2H has been enqueued on s_out in slot k % ns
        HostSlot& s = c->slots[k % ns];
        CK(cudaEventSynchronize(s.ev_out));
        if (!out_pinned) memcpy((char*)out_host + (size_t)k * out_bytes, s.h_out, frame_out * frames_of(k));
        return IMGCORR_OK;
    };
    for (int k = 0; k < chunks; ++k) {
        HostSlot& s = c->slots[k % ns];
        if (k >= ns) { r = retire(k - ns); if (r) return r; }
        const int nf = frames_of(k);
        const char* src = (const char*)raw_host + (size_t)k * raw_bytes;
        if (!in_pinned) { memcpy(s.h_raw, src, frame_raw * nf); src = (const char*)s.h_raw; }
        CK(cudaMemcpyAsync(s.d_raw, src, frame_raw * nf, cudaMemcpyHostToDevice, c->s_in));
        CK(cudaEventRecord(s.ev_in, c->s_in));
        CK(cudaStreamWaitEvent(c->s_k, s.ev_in, 0));
        r = chain_frames(c, s.d_raw, raw_dtype, s.d_out, out_dtype, nf, threshold, ksize, flags, lens, border_value, x0, y0,
                         ow, oh, c->s_k);
        if (r) return r;
        CK(cudaEventRecord(s.ev_k, c->s_k));
        CK(cudaStreamWaitEvent(c->s_out, s.ev_k, 0));
        void* dst = out_pinned ? (void*)((char*)out_host + (size_t)k * out_bytes) : s.h_out;
        CK(cudaMemcpyAsync(dst, s.d_out, frame_out * nf, cudaMemcpyDeviceToHost, c->s_out));
        CK(cudaEventRecord(s.ev_out, c->s_out));
    }
    for (int k = (chunks > ns ? chunks - ns : 0); k < chunks; ++k) { r = retire(k); if (r) return r; }
    drain.armed = false;                               // every chunk was retired: nothing is in flight
    return IMGCORR_OK;
}

extern "C" IMGCORR_API int imgcorr_selftest_division(imgcorr_ctx* c, int numerators_per_divisor, unsigned long long seed, double out[2]) {
    GUARD(c);
    if (!out || numerators_per_divisor < 1) return fail(IMGCORR_ERR_INVALID, "bad self-test arguments");
    unsigned long long* d = nullptr;
    unsigned long long h[2] = {0, 0};
    CK(cudaMalloc((void**)&d, sizeof h));
    cudaError_t e = cudaMemset(d, 0, sizeof h);
    if (e == cudaSuccess) e = launch_selftest_division(numerators_per_divisor, seed, d, nullptr);
    if (e == cudaSuccess) e = cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return cuda_fail(e, "division self-test");
    c->launches += 1;
    out[0] = (double)h[0];
    double worst;
    memcpy(&worst, &h[1], sizeof worst);
    out[1] = worst;
    return IMGCORR_OK;
}

// ---- K5 -------------------------------------------------------------------------------------
extern "C" IMGCORR_API int imgcorr_stack_mean(imgcorr_ctx* c, const void* frames_dev, int dtype, int n_frames, size_t elems,
                                  const double* minus_dev, double minus_scalar, int use_scalar, int gray3, double* out_dev, void* stream) {
    GUARD(c);
    if (!frames_dev || !out_dev) return fail(IMGCORR_ERR_INVALID, "null pointer");
    if (dtype < DT_U8 || dtype > DT_F64) return fail(IMGCORR_ERR_INVALID, "bad dtype %d", dtype);
    if (n_frames < 1) return fail(IMGCORR_ERR_INVALID, "n_frames %d", n_frames);
    if (gray3 && elems % 3) return fail(IMGCORR_ERR_INVALID, "gray conversion needs 3 interleaved channels");
    int l = 0;
    cudaError_t e = launch_k5_stack_mean(frames_dev, dtype, n_frames, elems, minus_dev, minus_scalar, use_scalar, gray3, out_dev, c->sm_count,
                                         (cudaStream_t)stream, &l);
    c->launches += l;
    if (e != cudaSuccess) return cuda_fail(e, "stack mean launch");
    return IMGCORR_OK;
}

extern "C" IMGCORR_API int imgcorr_scale_f64(imgcorr_ctx* c, double* data_dev, size_t elems, double divisor, void* stream) {
    GUARD(c);
    if (!data_dev) return fail(IMGCORR_ERR_INVALID, "null pointer");
    int l = 0;
    cudaError_t e = launch_k5_scale(data_dev, elems, divisor, c->sm_count, (cudaStream_t)stream, &l);
    c->launches += l;
    if (e != cudaSuccess) return cuda_fail(e, "scale launch");
    return IMGCORR_OK;
}

extern "C" IMGCORR_API int imgcorr_subsample_f64(imgcorr_ctx* c, const double* src_dev, int height, int width, int step_y, int step_x,
                                     double* dst_dev, void* stream) {
    GUARD(c);
    if (!src_dev || !dst_dev) return fail(IMGCORR_ERR_INVALID, "null pointer");
    if (height <= 0 || width <= 0 || step_y <= 0 || step_x <= 0) return fail(IMGCORR_ERR_INVALID, "bad shape / step");
    int l = 0;
    cudaError_t e = launch_k5_subsample(src_dev, height, width, step_y, step_x, dst_dev, (cudaStream_t)stream, &l);
    c->launches += l;
    if (e != cudaSuccess) return cuda_fail(e, "subsample launch");
    return IMGCORR_OK;
}

extern "C" IMGCORR_API int imgcorr_linear_fit(imgcorr_ctx* c, const void* frames_dev, int dtype, int n_frames, size_t pixels, const double* x_host,
                                  double max_intensity, double min_ascent, double* offset_dev, double* ascent_dev, double* rmse_dev,
                                  void* stream) {
    GUARD(c);
    if (!frames_dev || !x_host || !offset_dev || !ascent_dev) return fail(IMGCORR_ERR_INVALID, "null pointer");
    if (dtype < DT_U8 || dtype > DT_F64) return fail(IMGCORR_ERR_INVALID, "bad dtype %d", dtype);
    if (n_frames < 2) return fail(IMGCORR_ERR_INVALID, "a line needs at least 2 exposure times (got %d)", n_frames);
    double mn = x_host[0], mx = x_host[0];
    for (int k = 1; k < n_frames; ++k) { mn = x_host[k] < mn ? x_host[k] : mn; mx = x_host[k] > mx ? x_host[k] : mx; }
    double* xs = nullptr;
    CK(cudaMalloc((void**)&xs, n_frames * sizeof(double)));
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemcpyAsync(xs, x_host, n_frames * sizeof(double), cudaMemcpyHostToDevice, st);
    int l = 0;
    if (e == cudaSuccess)
        e = launch_k5_linear_fit(frames_dev, dtype, n_frames, pixels, xs, max_intensity, min_ascent, 0.5 * (mn + mx), offset_dev, ascent_dev,
                                 rmse_dev, c->sm_count, st, &l);
    c->launches += l;
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);       // x_host / xs may go away when this returns
    cudaFree(xs);
    if (e != cudaSuccess) return cuda_fail(e, "linear fit");
    return IMGCORR_OK;
}

// ---- host-side fingerprint of a calibration array ---------------------------------------------------------
static unsigned long long fp_chunk(const unsigned char* p, size_t n) {
    // four independent multiply-xor lanes over 8-byte words (the loop is load bound), tail bytes folded in at the end
    const unsigned long long K = 0x9E3779B97F4A7C15ull;
    unsigned long long a0 = 1, a1 = 2, a2 = 3, a3 = 4;
    size_t i = 0;
    for (; i + 32 <= n; i += 32) {
        unsigned long long w[4];
        memcpy(w, p + i, 32);
        a0 = (a0 ^ w[0]) * K; a1 = (a1 ^ w[1]) * K; a2 = (a2 ^ w[2]) * K; a3 = (a3 ^ w[3]) * K;
    }
    unsigned long long t = 0;
    for (int s = 0; i < n; ++i, s += 8) t ^= (unsigned long long)p[i] << (s & 63);
    unsigned long long h = (a0 ^ (a1 >> 17) ^ (a1 << 47)) * K;
    h = (h ^ a2 ^ (a3 >> 29) ^ (a3 << 35)) * K;
    h = (h ^ t ^ (unsigned long long)n) * K;
    return h ^ (h >> 32);
}

extern "C" IMGCORR_API int imgcorr_host_fingerprint(const void* host_ptr, size_t bytes, unsigned long long* out) {
    if (!out || (!host_ptr && bytes)) return fail(IMGCORR_ERR_INVALID, "null pointer");
    const unsigned char* p = (const unsigned char*)host_ptr;
    unsigned nt = std::thread::hardware_concurrency();
    if (nt > 16) nt = 16;
    if (nt < 1 || bytes < ((size_t)4 << 20)) nt = 1;
    std::vector<unsigned long long> part(nt, 0);
    const size_t chunk = ((bytes / nt) + 31) & ~(size_t)31;
    auto work = [&](unsigned t) {
        const size_t lo = (size_t)t * chunk < bytes ? (size_t)t * chunk : bytes;
        const size_t hi = lo + chunk < bytes && t + 1 < nt ? lo + chunk : bytes;
        part[t] = fp_chunk(p + lo, hi - lo);
    };
    std::vector<std::thread> th;
    for (unsigned t = 1; t < nt; ++t) th.emplace_back(work, t);
    work(0);
    for (auto& x : th) x.join();
    unsigned long long h = 0x243F6A8885A308D3ull;
    for (unsigned t = 0; t < nt; ++t) h = (h ^ part[t]) * 0x9E3779B97F4A7C15ull + t;
    *out = h;
    return IMGCORR_OK;
}

extern "C" IMGCORR_API int imgcorr_host_alloc(size_t bytes, void** out_ptr) {
    if (!out_ptr) return fail(IMGCORR_ERR_INVALID, "out_ptr is null");
    *out_ptr = nullptr;
    CK(cudaHostAlloc(out_ptr, bytes ? bytes : 1, cudaHostAllocDefault));
    return IMGCORR_OK;
}

extern "C" IMGCORR_API int imgcorr_host_free(void* ptr) {
    if (!ptr) return IMGCORR_OK;
    CK(cudaFreeHost(ptr));
    return IMGCORR_OK;
}

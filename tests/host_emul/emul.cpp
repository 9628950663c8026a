// Host emulation of the K1 / K2 arithmetic: compiles imgprocessor_b200/csrc/imgcorr_core.cuh —
// the very functions the CUDA kernels call per pixel — with g++ and runs them over whole images
// with plain loops.  TEST-ONLY (built into tests/_build/libimgcorr_emul.so by
// tests/test_core_emul.py); it is never loaded by the product and is not a fallback: it exists
// because the build container has no GPU and the per-pixel arithmetic can be pinned against the
// oracle here before any GPU minute is spent.
#include <vector>
#include "../../imgprocessor_b200/csrc/imgcorr_core.cuh"

using namespace imgcorr;

template <typename RawT> static double ld(const void* p, size_t i) { return (double)((const RawT*)p)[i]; }

static double load_raw(const void* raw, int dt, size_t i) {
    switch (dt) {
        case 0: return ld<uint8_t>(raw, i);
        case 1: return ld<uint16_t>(raw, i);
        case 2: return ld<float>(raw, i);
        default: return ld<double>(raw, i);
    }
}

template <typename CT>
static void k1_typed(const void* raw, int raw_dtype, const float* dark, const float* ascent, const float* flat,
                     void* out, int out_dtype, uint8_t* mask, int H, int W, double thr, int ksize, int cond, int flags,
                     double exposure, double maxval) {
    PointwiseConst pw;
    pw.flags = 0;
    if ((flags & 1) && dark) { pw.flags |= FLAG_DARK; if (ascent) pw.flags |= FLAG_DARK_LINEAR; }
    if ((flags & 2) && flat) pw.flags |= FLAG_FLAT;
    if (flags & 4) pw.flags |= FLAG_NAN_TO_NUM;
    pw.exposure_time = exposure;
    pw.max_value = maxval;
    if (!(thr > 0)) ksize = 0;
    PredicateConst pc = make_predicate(thr, cond);
    std::vector<CT> x((size_t)H * W);
    for (size_t i = 0; i < (size_t)H * W; ++i)
        x[i] = pointwise<CT>(pw, load_raw(raw, raw_dtype, i), dark ? dark[i] : 0.f, ascent ? ascent[i] : 0.f,
                             flat ? flat[i] : 0.f);
    auto at = [&](int y, int xx) { return x[(size_t)reflect_index(y, H) * W + reflect_index(xx, W)]; };
    for (int y = 0; y < H; ++y)
        for (int xx = 0; xx < W; ++xx) {
            CT v = x[(size_t)y * W + xx], res = v;
            bool rep = false;
            if (ksize == 3) {
                Sorted3<CT> a = sort3(at(y - 1, xx - 1), at(y - 1, xx), at(y - 1, xx + 1));
                Sorted3<CT> b = sort3(at(y, xx - 1), at(y, xx), at(y, xx + 1));
                Sorted3<CT> c = sort3(at(y + 1, xx - 1), at(y + 1, xx), at(y + 1, xx + 1));
                CT med = median9(a, b, c);
                rep = predicate(v, med, pc);
                if (rep) res = med;
            } else if (ksize == 5) {
                CT w[25];
                for (int dy = 0; dy < 5; ++dy) {
                    for (int dx = 0; dx < 5; ++dx) w[dy * 5 + dx] = at(y + dy - 2, xx + dx - 2);
                    sort5(w + dy * 5);
                }
                CT med = median25_sorted_rows(w);
                rep = predicate(v, med, pc);
                if (rep) res = med;
            }
            size_t i = (size_t)y * W + xx;
            switch (out_dtype) {
                case 0: ((uint8_t*)out)[i] = sat_u8((float)res); break;
                case 1: ((uint16_t*)out)[i] = sat_u16((float)res); break;
                case 2: ((float*)out)[i] = (float)res; break;
                default: ((double*)out)[i] = (double)res; break;
            }
            if (mask) mask[i] = rep;
        }
}

extern "C" __attribute__((visibility("default")))
void emul_k1(const void* raw, int raw_dtype, const float* dark, const float* ascent, const float* flat, void* out,
             int out_dtype, uint8_t* mask, int H, int W, double thr, int ksize, int cond, int flags, double exposure,
             double maxval) {
    if (raw_dtype == 3)
        k1_typed<double>(raw, raw_dtype, dark, ascent, flat, out, out_dtype, mask, H, W, thr, ksize, cond, flags, exposure, maxval);
    else
        k1_typed<float>(raw, raw_dtype, dark, ascent, flat, out, out_dtype, mask, H, W, thr, ksize, cond, flags, exposure, maxval);
}

static bool make_lens(const double* K, const double* dist, const double* P, LensConst& L) {
    if (!invert3x3(P, L.ir)) return false;
    L.k1 = dist[0]; L.k2 = dist[1]; L.p1 = dist[2]; L.p2 = dist[3]; L.k3 = dist[4];
    L.p1x2 = L.p1 + L.p1; L.p2x2 = L.p2 + L.p2;
    L.fx = K[0]; L.fy = K[4]; L.cx = K[2]; L.cy = K[5];
    L.affine = (L.ir[6] == 0.0 && L.ir[7] == 0.0 && L.ir[8] == 1.0) ? 1 : 0;
    return true;
}

extern "C" __attribute__((visibility("default")))
int emul_maps(const double* K, const double* dist, const double* P, int H, int W, float* mapx, float* mapy) {
    LensConst L;
    if (!make_lens(K, dist, P, L)) return -1;
    for (int v = 0; v < H; ++v)
        for (int u = 0; u < W; ++u) undistort_map(L, u, v, mapx[(size_t)v * W + u], mapy[(size_t)v * W + u]);
    return 0;
}

template <typename T> static T fetch(const T* src, int H, int W, int x, int y, T b) {
    return ((unsigned)x < (unsigned)W && (unsigned)y < (unsigned)H) ? src[(size_t)y * W + x] : b;
}

// dtype: 0 u8, 1 u16, 2 f32, 3 f64, 4 f32 -> f64 (widening)
extern "C" __attribute__((visibility("default")))
void emul_remap(const void* src, int dtype, void* dst, int H, int W, const float* mapx, const float* mapy, double border,
                int x0, int y0, int ow, int oh) {
    border = border_for_dtype(dtype == 0, dtype == 1, border);
    for (int oy = 0; oy < oh; ++oy)
        for (int ox = 0; ox < ow; ++ox) {
            const int u = ox + x0, v = oy + y0;
            FixedCoord c = fixed_coord(mapx[(size_t)v * W + u], mapy[(size_t)v * W + u]);
            const bool outside = c.ix >= W || c.ix + 1 < 0 || c.iy >= H || c.iy + 1 < 0;
            float w00, w01, w10, w11;
            bilinear_weights(c.fx, c.fy, w00, w01, w10, w11);
            const size_t o = (size_t)oy * ow + ox;
            if (dtype == 2 || dtype == 4) {
                const float* s = (const float*)src; float b = (float)border;
                float r = outside ? b : blend_f32(fetch(s, H, W, c.ix, c.iy, b), fetch(s, H, W, c.ix + 1, c.iy, b),
                                                  fetch(s, H, W, c.ix, c.iy + 1, b), fetch(s, H, W, c.ix + 1, c.iy + 1, b),
                                                  w00, w01, w10, w11);
                if (dtype == 2) ((float*)dst)[o] = r; else ((double*)dst)[o] = (double)r;
            } else if (dtype == 3) {
                const double* s = (const double*)src; double b = border;
                ((double*)dst)[o] = outside ? b : blend_f64(fetch(s, H, W, c.ix, c.iy, b), fetch(s, H, W, c.ix + 1, c.iy, b),
                                                            fetch(s, H, W, c.ix, c.iy + 1, b), fetch(s, H, W, c.ix + 1, c.iy + 1, b),
                                                            w00, w01, w10, w11);
            } else if (dtype == 1) {
                const uint16_t* s = (const uint16_t*)src; uint16_t b = (uint16_t)border;
                ((uint16_t*)dst)[o] = outside ? b : sat_u16(blend_f32(fetch(s, H, W, c.ix, c.iy, b), fetch(s, H, W, c.ix + 1, c.iy, b),
                                                                      fetch(s, H, W, c.ix, c.iy + 1, b), fetch(s, H, W, c.ix + 1, c.iy + 1, b),
                                                                      w00, w01, w10, w11));
            } else {
                const uint8_t* s = (const uint8_t*)src; uint8_t b = (uint8_t)border;
                ((uint8_t*)dst)[o] = outside ? b : (uint8_t)blend_u8(fetch(s, H, W, c.ix, c.iy, b), fetch(s, H, W, c.ix + 1, c.iy, b),
                                                                     fetch(s, H, W, c.ix, c.iy + 1, b), fetch(s, H, W, c.ix + 1, c.iy + 1, b),
                                                                     c.fx, c.fy);
            }
        }
}

// ---- cv2.warpPerspective (imgcorr_warp.cuh) -------------------------------------------
#include "../../imgprocessor_b200/csrc/imgcorr_warp.cuh"

extern "C" __attribute__((visibility("default")))
void emul_warp_tables(float* lanczos, float* cubic, const double* M, double* Minv) {
    warp_lanczos4_table(lanczos);
    warp_cubic_table(cubic);
    warp_invert3x3(M, Minv);
}

template <typename T, typename AT, int N>
static void warp_typed(const T* src, T* dst, int H, int W, int dh, int dw, const WarpConst& wc, const float* tab, double border) {
    for (int y = 0; y < dh; ++y)
        for (int x = 0; x < dw; ++x) {
            AT v = warp_pixel<T, AT, N>(src, H, W, warp_coord(wc, x, y), tab, (AT)border);
            if (sizeof(T) == 2) dst[(size_t)y * dw + x] = (T)sat_u16((float)v);
            else dst[(size_t)y * dw + x] = (T)v;
        }
}

// dtype: 0 u8, 1 u16, 2 f32, 3 f64;  interp: 2 cubic, 4 lanczos4
extern "C" __attribute__((visibility("default")))
int emul_warp(const void* src, int dtype, void* dst, int H, int W, int dh, int dw, const double* M, int interp, int inverse,
              double border) {
    std::vector<float> tab(32 * 8);
    if (interp == WARP_LANCZOS4) warp_lanczos4_table(tab.data()); else if (interp == WARP_CUBIC) warp_cubic_table(tab.data()); else return -1;
    WarpConst wc = make_warp_const(M, inverse, dw, dh);
    border = border_for_dtype(0, dtype == 1, border);
#define GO(T, AT) { if (interp == WARP_LANCZOS4) warp_typed<T, AT, 8>((const T*)src, (T*)dst, H, W, dh, dw, wc, tab.data(), border); \
                    else warp_typed<T, AT, 4>((const T*)src, (T*)dst, H, W, dh, dw, wc, tab.data(), border); }
    if (dtype == 0) {
        const int n = interp == WARP_LANCZOS4 ? 8 : 4;
        std::vector<int16_t> it((size_t)32 * 32 * n * n);
        warp_fixed_table(tab.data(), n, it.data());
        border = border_for_dtype(1, 0, border);
        for (int y = 0; y < dh; ++y)
            for (int x = 0; x < dw; ++x) {
                const FixedCoord c = warp_coord(wc, x, y);
                const int16_t* w = it.data() + ((size_t)c.fy * 32 + c.fx) * n * n;
                ((uint8_t*)dst)[(size_t)y * dw + x] = n == 8 ? warp_pixel_u8<8>((const uint8_t*)src, H, W, c, w, (int)border)
                                                             : warp_pixel_u8<4>((const uint8_t*)src, H, W, c, w, (int)border);
            }
        return 0;
    }
    if (dtype == 1) GO(uint16_t, float) else if (dtype == 2) GO(float, float) else if (dtype == 3) GO(double, double) else return -1;
#undef GO
    return 0;
}

// ---- single-time-effect-free average (imgcorr_ste.cuh) ---------------------------------
#include "../../imgprocessor_b200/csrc/imgcorr_ste.cuh"

// frames: [n][H][W] float64 (the caller widens); avg out [H][W]; mask out [H][W] or null
extern "C" __attribute__((visibility("default")))
void emul_ste(const double* frames, int n, int H, int W, const double* nlf, double nstd, double* avg, uint8_t* mask) {
    SteConst sc{nlf[0], nlf[1], nlf[2], nstd};
    const size_t npx = (size_t)H * W;
    std::vector<double> thr(npx), img(npx);
    std::vector<int> cnt(npx, 1);
    std::vector<uint8_t> f(npx);
    for (size_t i = 0; i < npx; ++i) {
        const double p = frames[i], q = frames[npx + i];
        avg[i] = (p != p || q != q) ? p + q : (p < q ? p : q);
        img[i] = (p != p || q != q) ? p + q : (p < q ? q : p);
        thr[i] = ste_threshold(sc, avg[i]);
        if (mask) mask[i] = 0;
    }
    for (int k = 1; k < n; ++k) {
        if (k > 1) for (size_t i = 0; i < npx; ++i) img[i] = frames[(size_t)k * npx + i];
        for (size_t i = 0; i < npx; ++i) f[i] = ste_flag(img[i], avg[i], thr[i]);
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                const size_t i = (size_t)y * W + x;
                bool s = f[i];
                if (s) {
                    int nb = 0;
                    for (int dy = -1; dy <= 1; ++dy)
                        for (int dx = -1; dx <= 1; ++dx)
                            if ((dy || dx) && (unsigned)(y + dy) < (unsigned)H && (unsigned)(x + dx) < (unsigned)W) nb += f[(size_t)(y + dy) * W + x + dx];
                    s = nb > 0;
                }
                if (!s) { cnt[i] += 1; avg[i] = ste_update(img[i], avg[i], cnt[i]); }
                else if (mask) mask[i] = 1;
            }
    }
}

// order-preserving integer keys used by the 5x5 streaming kernel (imgcorr_core.cuh: to_key / from_key)
extern "C" __attribute__((visibility("default")))
void emul_keys(const float* x, int n, int* keys, float* back) {
    for (int i = 0; i < n; ++i) { OrdKey k = to_key(x[i]); keys[i] = k.k; back[i] = from_key(k); }
}

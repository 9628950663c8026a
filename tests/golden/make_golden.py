"""Generate tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference) in the build container.  Run from the repo root:

    python tests/golden/make_golden.py

Each .npz holds the inputs and the reference's outputs for one call on the
CameraCalibration.correct() path.  Library versions are stored in
versions.json.  The reference has no golden vectors of its own,
so these files are what pins the oracle (tests/test_oracle_golden.py).
"""
import contextlib
import io
import json
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import ref_shim  # noqa: E402

ref_shim.install()
warnings.simplefilter('ignore')

import cv2  # noqa: E402
import scipy  # noqa: E402
from imgProcessor.camera.CameraCalibration import CameraCalibration  # noqa: E402
from imgProcessor.camera.LensDistortion import LensDistortion  # noqa: E402
from imgProcessor.filters.medianThreshold import medianThreshold  # noqa: E402

from imgprocessor_b200 import synth  # noqa: E402


def quiet(fn, *a, **k):
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        r = fn(*a, **k)
    return r, buf.getvalue()


def make_lens(params, shape):
    l = LensDistortion({})          # fresh dict: the default argument is shared (LensDistortion.py:26)
    l.setCameraParams(*params)
    l._coeffs['shape'] = shape
    return l


def save(name, **arrs):
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **arrs)
    print('wrote', name, {k: (getattr(v, 'shape', None), getattr(v, 'dtype', None)) for k, v in arrs.items()})


def main():
    H, W = 96, 128
    params = synth.lens_moderate(H, W)
    K, d = synth.camera_matrix(params), synth.dist_coeffs(params)

    raw16 = synth.scene(H, W, 1, np.uint16)
    raw32 = synth.scene(H, W, 2, np.float32)
    dark = synth.dark_map(H, W)
    flat = synth.flat_map(H, W, p_zero=2e-3)
    flat[5, 7] = np.float32(1e-42)       # denormal flat -> huge quotient
    flat[6, 9] = np.float32(-0.5)        # negative flat -> negative image values / medians

    # ---- 1/2: full chain, keep_size True / False, u16 ---------------------
    for keep in (True, False):
        cal = CameraCalibration()
        cal.addDarkCurrent(dark)
        cal.addFlatField(flat)
        cal.addLens(make_lens(params, (H, W)))
        out, log = quiet(cal.correct, raw16, threshold=0.1, keep_size=keep)
        save('correct_u16_keep%d' % keep, raw=raw16, dark=dark, flat=flat, K=K, dist=d,
             out=np.ascontiguousarray(out), last_img=cal.last_img, bg=cal.temp['bg'],
             log=np.array(log))

    # ---- 3: f32 frame, threshold 0 (median skipped, no nan_to_num) --------
    cal = CameraCalibration()
    cal.addDarkCurrent(dark)
    cal.addFlatField(flat)
    cal.addLens(make_lens(params, (H, W)))
    out, log = quiet(cal.correct, raw32, threshold=0)
    save('correct_f32_thr0', raw=raw32, dark=dark, flat=flat, K=K, dist=d, out=out, log=np.array(log))

    # ---- 4: only some calibrations present --------------------------------
    cal = CameraCalibration()
    cal.addFlatField(flat)
    out, log = quiet(cal.correct, raw32, threshold=0.25)
    save('correct_f32_flat_only', raw=raw32, flat=flat, out=out, log=np.array(log))

    cal = CameraCalibration()
    out, log = quiet(cal.correct, raw16, threshold=0.1)
    save('correct_u16_nothing', raw=raw16, out=out, log=np.array(log))

    # ---- 5: legacy tuple dark-current entry (offs + ascent*t, clipped) ----
    cal = CameraCalibration()
    offs = dark.astype(np.float64)
    ascent = (synth.dark_map(H, W, seed=5).astype(np.float64) * 40.0)
    ascent[3, 3] = 1e6                       # exercises the 2**depth-1 clip
    cal.coeffs['dark current'].append((cal.currentTime() and __import__('time').localtime(), '', (offs.copy(), ascent.copy()), None))
    cal.coeffs['shape'] = (H, W)
    out, log = quiet(cal.correct, raw16, exposure_time=2.5, threshold=0.1)
    save('correct_u16_legacy_dark', raw=raw16, offs=offs, ascent=ascent, exposure_time=np.float64(2.5),
         out=out, bg=cal.temp['bg'], log=np.array(log))

    # ---- 6: date selection -------------------------------------------------
    cal = CameraCalibration()
    d1 = dark
    d2 = dark + np.float32(50)
    d3 = dark + np.float32(200)
    cal.addDarkCurrent(d1, date='01 Jan 15 - 10:00')
    cal.addDarkCurrent(d3, date='01 Jan 17 - 10:00')
    cal.addDarkCurrent(d2, date='01 Jan 16 - 10:00')
    outs = {}
    for tag, date in (('none', None), ('mid', '01 Jun 16 - 00:00'), ('old', '01 Jan 14 - 00:00'),
                      ('new', '01 Jan 18 - 00:00'), ('bad', 'not a date')):
        o, _ = quiet(cal.correct, raw16, threshold=0, date=date)
        outs['out_' + tag] = o
    save('correct_u16_dates', raw=raw16, d1=d1, d2=d2, d3=d3, **outs)

    # ---- 7: medianThreshold direct ----------------------------------------
    for tag, img in (('u16', raw16), ('f32', raw32)):
        for size in (3, 5):
            for cond in ('>', '<'):
                o, ind = medianThreshold(img, threshold=0.1, size=size, condition=cond, copy=True)
                save('median_%s_s%d_%s' % (tag, size, 'gt' if cond == '>' else 'lt'),
                     img=img, out=o, ind=ind)
    # zeros / negative medians / blur == 0 cases
    z = raw32.copy()
    z[10:40, 10:60] = 0
    z[20, 20] = 5
    z[50:60, 50:60] *= -1
    o, ind = medianThreshold(z, threshold=0.3, size=3)
    save('median_f32_zeros', img=z, out=o, ind=ind, threshold=np.float64(0.3))

    # ---- 8: LensDistortion.correct direct, every dtype the callers use ----
    Hr, Wr = 125, 166                          # realistic coefficients scaled to a quarter-size frame
    pr = list(synth.lens_realistic())
    pr[0] /= 4; pr[1] /= 4; pr[2] /= 4; pr[3] /= 4
    for tag, dt in (('u8', np.uint8), ('u16', np.uint16), ('f32', np.float32), ('f64', np.float64)):
        img = synth.scene(Hr, Wr, 3, dt) if dt != np.float64 else synth.scene(Hr, Wr, 3, np.float32).astype(np.float64) * 1.000000123
        for keep in (False, True):
            l = make_lens(pr, (Hr, Wr))
            o = l.correct(img, keepSize=keep, borderValue=7 if keep else 0)
            extra = dict(mapx=l.mapx, mapy=l.mapy) if (tag == 'f32' and keep) else {}
            save('lens_%s_keep%d' % (tag, keep), img=img, out=np.ascontiguousarray(o), roi=np.array(l.roi),
                 K=l.coeffs['cameraMatrix'], dist=l.coeffs['distortionCoeffs'],
                 border=np.float64(7 if keep else 0), **extra)

    # ---- 9: maps + P + roi for three lenses --------------------------------
    for tag, (hh, ww, pp) in (('moderate', (96, 128, synth.lens_moderate(96, 128))),
                              ('strong', (128, 128, synth.lens_strong(128, 128))),
                              ('realistic', (Hr, Wr, pr))):
        l = make_lens(pp, (hh, ww))
        mx, my = l.getUndistortRectifyMap(ww, hh)
        P, roi = cv2.getOptimalNewCameraMatrix(l.coeffs['cameraMatrix'], l.coeffs['distortionCoeffs'],
                                               (ww, hh), 1, (ww, hh))
        save('maps_' + tag, K=l.coeffs['cameraMatrix'], dist=l.coeffs['distortionCoeffs'], P=P,
             roi=np.array(roi), mapx=mx, mapy=my, shape=np.array([hh, ww]))

    # ---- 10: ingest formats: the reference's own readers on two tiny files -------------
    from imgProcessor.reader.RAW import RAW
    from imgProcessor.reader.elbin import elbin
    rng = np.random.default_rng(77)
    rw, rh = 24, 32                                            # RAW(filename, width, height): array shape (width, height)
    be = rng.integers(0, 65536, (rw, rh)).astype('>u2')
    be.tofile(os.path.join(HERE, 'raw_be_u16_24x32.raw'))
    got = RAW(os.path.join(HERE, 'raw_be_u16_24x32.raw'), rw, rh, '16-bit Unsigned')
    assert got.dtype == np.dtype('>u2') and got.shape == (rw, rh)
    eh, ew, en = 40, 24, 3                                     # elbin header: height, width, frames; frames shaped (width, height)
    with open(os.path.join(HERE, 'stack_3x24x40.elbin'), 'wb') as f:
        f.write(np.array([eh, ew, en], np.uint32).tobytes())
        frames = rng.integers(0, 65536, (en, ew, eh)).astype(np.uint16)
        for i in range(en):
            f.write(np.array([1.5 + i, 30.25 - i], np.float64).tobytes())
            f.write(np.array([3 + i], np.uint32).tobytes())
            f.write(frames[i].tobytes())
    arrs, labels = elbin(os.path.join(HERE, 'stack_3x24x40.elbin'))
    assert np.array_equal(arrs, frames)
    save('readers', raw_be=np.asarray(got, dtype=np.uint16), elbin_frames=arrs,
         elbin_times=np.array([l['exposure time[s]'] for l in labels]),
         elbin_current=np.array([l['current[A]'] for l in labels]), elbin_voltage=np.array([l['voltage[V]'] for l in labels]))

    with open(os.path.join(HERE, 'versions.json'), 'w') as f:
        json.dump({'numpy': np.__version__, 'scipy': scipy.__version__, 'cv2': cv2.__version__,
                   'reference': 'radjkarl/imgProcessor 0.2.5 (/root/reference), unmodified, under ref_shim'}, f, indent=1)


def perspective():
    """PerspectiveCorrection.correct / uncorrect of the unmodified reference
    (camera/PerspectiveCorrection.py:374-406) with a homography and with a quad reference."""
    ref_shim.install_perspective_stubs()
    from imgProcessor.camera.PerspectiveCorrection import PerspectiveCorrection
    H, W = 120, 160
    scene = synth.scene(H, W, seed=21, dtype=np.float64, full_scale=1.0).astype(np.float64)
    quad = np.array([[14.5, 9.0], [150.25, 17.0], [141.0, 110.5], [8.0, 101.0]])
    new_size = (96, 136)                                       # (sizey, sizex)
    out = {}
    for tag, img in (('f64', scene), ('f32', scene.astype(np.float32)), ('u16', (scene * 65535).astype(np.uint16))):
        pc = PerspectiveCorrection(img.shape, new_size=new_size, border=4)
        pc.setReference(quad[[2, 0, 3, 1]])                    # unsorted on purpose: sortCorners orders it
        out['quad_' + tag], log = quiet(pc.correct, img)
        if tag == 'f64':
            out['quad_sorted'] = pc.quad.copy()
            out['quad_homography'] = np.array(pc.homography)
            out['uncorrect_f64'] = pc.uncorrect(out['quad_f64'])
    Hm = np.array([[1.02, 0.03, -6.5], [-0.015, 0.97, 4.25], [2e-5, -4e-5, 1.0]])
    pc = PerspectiveCorrection(scene.shape, new_size=(H, W))
    pc.setReference(Hm)
    out['homography_f64'] = quiet(pc.correct, scene)[0]
    out['homography_f32'] = quiet(pc.correct, scene.astype(np.float32))[0]
    pc = PerspectiveCorrection(scene.shape, new_size=(H + 30, W + 50), cv2_opts={'borderValue': 0.25})
    pc.setReference(Hm)
    out['homography_border_f32'] = quiet(pc.correct, scene.astype(np.float32))[0]
    save('perspective', scene=scene, quad=quad[[2, 0, 3, 1]], new_size=np.array(new_size), border=np.int64(4), Hm=Hm, log=np.array(log), **out)


def ste():
    """the reference's SingleTimeEffectDetection / removeSinglePixels / boundedFunction, unmodified,
    with the restated MaskedMovingAverage injected (ref_shim.install_masked_moving_average)."""
    ref_shim.install_masked_moving_average()
    import importlib
    import imgProcessor.features.SingleTimeEffectDetection as stemod
    importlib.reload(stemod)
    from imgProcessor.filters.removeSinglePixels import removeSinglePixels
    from imgProcessor.camera import NoiseLevelFunction
    rng = np.random.default_rng(31)
    H, W = 72, 100
    base = synth.scene(H, W, seed=5, dtype=np.float64, full_scale=4000.0, hot_dead=False)
    nlf = (6.0, 20.0, 0.9)                                        # minY, ax, ay
    frames = []
    for i in range(5):
        f = base + rng.normal(0, 1, (H, W)) * NoiseLevelFunction.boundedFunction(base, *nlf)
        for k in range(6):                                        # cosmic-ray-like blobs and single pixels
            y, x = rng.integers(2, H - 3), rng.integers(2, W - 3)
            f[y:y + rng.integers(1, 4), x:x + rng.integers(1, 4)] += rng.uniform(300, 3000)
        f[rng.random((H, W)) < 2e-3] += 2000.0
        frames.append(np.clip(np.rint(f), 0, 65535).astype(np.uint16))
    frames = np.stack(frames)
    fn = lambda x: NoiseLevelFunction.boundedFunction(x, *nlf)
    out = {}
    for n in (2, 3, 5):
        det = stemod.SingleTimeEffectDetection(list(frames[:n]), nStd=4, noise_level_function=fn, save_ste_indices=True)
        out['noSTE_%d' % n] = det.noSTE
        out['mask_%d' % n] = det.mask_STE
    det = stemod.SingleTimeEffectDetection(list(frames.astype(np.float32)), nStd=3, noise_level_function=fn)
    out['noSTE_f32_nstd3'] = det.noSTE
    m = rng.random((64, 80)) > 0.93
    m[0, 0] = m[63, 79] = m[0, 79] = True
    m2 = m.copy()
    removeSinglePixels(m2)
    xs = np.concatenate([np.linspace(-50, 5000, 301), [np.nan, np.inf, -np.inf, 20.0]])
    # the multi-image branches of correct() itself (CameraCalibration.py:385-406, 484-498)
    import imgProcessor.camera.CameraCalibration as calmod
    calmod.SingleTimeEffectDetection = stemod.SingleTimeEffectDetection
    dark = synth.dark_map(H, W)
    flat = synth.flat_map(H, W, p_zero=2e-3)
    params = synth.lens_moderate(H, W)
    cal = calmod.CameraCalibration()
    cal.addDarkCurrent(dark)
    cal.addFlatField(flat)
    cal.addLens(make_lens(params, (H, W)))
    cal.addNoise(nlf)
    out['correct_3'], out['correct_3_log'] = quiet(cal.correct, list(frames[:3]), threshold=0.1)
    out['correct_3_log'] = np.array(out['correct_3_log'])
    bgs = np.stack([np.clip(np.rint(dark + rng.normal(0, 3, (H, W))), 0, 65535) for i in range(3)]).astype(np.uint16)
    bgs[1, 10:12, 20:23] += 900
    cal2 = calmod.CameraCalibration()
    cal2.addFlatField(flat)
    cal2.addNoise(nlf)
    cal2.noise_level_function = fn
    out['correct_bg3'], lg = quiet(cal2.correct, frames[0], bgImages=list(bgs), threshold=0.1)
    out['correct_bg3_log'] = np.array(lg)
    out['correct_bg3_bg'] = cal2.temp['bg']
    save('ste', dark=dark, flat=flat, K=synth.camera_matrix(params), dist=synth.dist_coeffs(params), bgs=bgs, frames=frames, nlf=np.array(nlf), rsp_in=m, rsp_out=m2, bf_x=xs,
         bf_y=NoiseLevelFunction.boundedFunction(xs, *nlf), **out)


def ste_nlf():
    """SURVEY §8 a11 / f1 without a 'noise' calibration: the reference then ESTIMATES the noise level function from the images
    (SingleTimeEffectDetection.py:43-45 -> NoiseLevelFunction.oneImageNLF -> calcNLF / _evaluate, unmodified; the restated
    MaskedMovingAverage injected as in ste())."""
    ref_shim.install_masked_moving_average()
    import importlib
    import imgProcessor.features.SingleTimeEffectDetection as stemod
    importlib.reload(stemod)
    from imgProcessor.camera import NoiseLevelFunction
    import imgProcessor.camera.CameraCalibration as calmod
    calmod.SingleTimeEffectDetection = stemod.SingleTimeEffectDetection
    rng = np.random.default_rng(77)
    H, W = 120, 160
    base = synth.scene(H, W, seed=9, dtype=np.float64, full_scale=4000.0, hot_dead=False)
    nlf = (6.0, 20.0, 0.9)
    frames = []
    for i in range(4):
        f = base + rng.normal(0, 1, (H, W)) * NoiseLevelFunction.boundedFunction(base, *nlf)
        for k in range(8):
            y, x = rng.integers(2, H - 3), rng.integers(2, W - 3)
            f[y:y + rng.integers(1, 4), x:x + rng.integers(1, 4)] += rng.uniform(300, 3000)
        frames.append(np.clip(np.rint(f), 0, 65535).astype(np.uint16))
    frames = np.stack(frames)
    out = {}
    avg0 = np.min((frames[0].astype(np.float64), frames[1]), axis=0)
    x, y, w, signal = NoiseLevelFunction.calcNLF(avg0)
    params, fn, valid = NoiseLevelFunction._evaluate(x, y, w)
    out.update(avg0=avg0, nlf_x=x, nlf_y=y, nlf_w=w, nlf_signal=signal, nlf_params=params, nlf_valid=valid,
               nlf_curve_x=np.linspace(0, 4500, 91), nlf_curve=fn(np.linspace(0, 4500, 91)))
    x2, y2, w2, s2 = NoiseLevelFunction.calcNLF(frames[0], frames[1])                 # two-image estimate
    out.update(nlf2_x=x2, nlf2_y=y2, nlf2_w=w2, nlf2_signal=s2)
    for n in (2, 4):
        det = stemod.SingleTimeEffectDetection(list(frames[:n]), nStd=4, save_ste_indices=True)
        out['noSTE_%d' % n] = det.noSTE
        out['mask_%d' % n] = det.mask_STE
    # a noise level function the square-root model cannot describe: the polynomial fallback (smooth)
    xs = np.linspace(100, 3000, 40)
    ys = 30.0 + 1e-5 * (xs - 1500.0) ** 2 - 25.0 * (xs > 2500)
    ws = np.linspace(50, 500, 40)
    sm = NoiseLevelFunction.smooth(xs, ys, ws)
    out.update(smooth_x=xs, smooth_y=ys, smooth_w=ws, smooth_eval_x=np.linspace(-100, 3500, 61), smooth_eval=sm(np.linspace(-100, 3500, 61)))
    dark = synth.dark_map(H, W)
    flat = synth.flat_map(H, W, p_zero=2e-3)
    params_l = synth.lens_moderate(H, W)
    cal = calmod.CameraCalibration()
    cal.addDarkCurrent(dark)
    cal.addFlatField(flat)
    cal.addLens(make_lens(params_l, (H, W)))
    out['correct_3'], lg = quiet(cal.correct, list(frames[:3]), threshold=0.1)        # no addNoise: NLF estimated, then kept
    out['correct_3_log'] = np.array(lg)
    out['correct_3_again'], _ = quiet(cal.correct, list(frames[1:4]), threshold=0.1)  # second call re-uses the kept function
    save('ste_nlf', frames=frames, dark=dark, flat=flat, K=synth.camera_matrix(params_l), dist=synth.dist_coeffs(params_l), **out)


def producers():
    """flatFieldFromCloseDistance, averageSameExpTimes / getDarkCurrentAverages and the map-derived lens
    utilities of the unmodified reference (getLinearityFunction needs the absent fancytools regression: unpinned)."""
    ref_shim.install_masked_moving_average()
    import importlib
    import types
    import imgProcessor.features.SingleTimeEffectDetection as stemod
    importlib.reload(stemod)
    for name in ('imgProcessor.utils.baseClasses',):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                m = types.ModuleType(name)
                m.Iteratives = type('Iteratives', (object,), {'__init__': lambda self, **k: None})
                sys.modules[name] = m
    import imgProcessor.camera.DarkCurrentMap as dcm
    importlib.reload(dcm)
    from imgProcessor.camera.flatField.flatFieldFromCloseDistance import flatFieldFromCloseDistance
    rng = np.random.default_rng(5)
    H, W = 90, 130
    out = {}
    flat = synth.flat_map(H, W, p_zero=0).astype(np.float64)
    imgs = [np.clip(np.rint(flat[..., None] * np.array([900.0, 1400.0, 700.0]) + 40 + rng.normal(0, 6, (H, W, 3))), 0, 65535).astype(np.uint16)
            for _ in range(4)]
    bgs = [np.clip(np.rint(40 + rng.normal(0, 3, (H, W, 3))), 0, 65535).astype(np.uint16) for _ in range(3)]
    out['ff_imgs'], out['ff_bgs'] = np.stack(imgs), np.stack(bgs)
    out['ff_bglist'] = flatFieldFromCloseDistance(list(imgs), list(bgs))
    out['ff_bgnum'] = flatFieldFromCloseDistance(list(imgs), 41.5)
    out['ff_f32'] = flatFieldFromCloseDistance([i.astype(np.float32) for i in imgs], [b.astype(np.float32) for b in bgs])
    # dark-current averages per exposure time (DarkCurrentMap: STE removal with nStd=3, NLF estimated from the images)
    Hd, Wd = 120, 160
    dark = synth.dark_map(Hd, Wd).astype(np.float64)
    times = [1.0, 1.0, 1.0, 4.0, 4.0, 4.0, 4.0, 9.0]
    frames = []
    for t in times:
        f = dark + 2.5 * t + rng.normal(0, 2.0 + 0.2 * t, (Hd, Wd))
        y, x = rng.integers(2, Hd - 4), rng.integers(2, Wd - 4)
        f[y:y + 2, x:x + 3] += 700.0
        frames.append(np.clip(np.rint(f), 0, 65535).astype(np.uint16))
    out['dc_times'], out['dc_frames'] = np.array(times), np.stack(frames)
    out['dc_avg_t4'] = dcm.averageSameExpTimes(list(frames[3:7]))
    xs, av = dcm.getDarkCurrentAverages(times, frames)
    out['dc_x'], out['dc_averages'] = np.array(xs), av
    # map-derived lens utilities (LensDistortion.py:332-340, 382-402)
    Hl, Wl = 96, 128
    lens = make_lens(synth.lens_moderate(Hl, Wl), (Hl, Wl))
    mx, my = lens.getDistortRectifyMap(Wl, Hl)
    out['lens_K'], out['lens_dist'] = lens.coeffs['cameraMatrix'], lens.coeffs['distortionCoeffs']
    out['lens_dmapx'], out['lens_dmapy'] = mx, my
    out['lens_shift'] = lens.getShift(Wl, Hl)
    ux, uy = lens.getDeflection(Wl, Hl)
    out['lens_ux'], out['lens_uy'] = ux, uy
    img8 = synth.scene(Hl, Wl, 4, np.uint8)
    imgf = synth.scene(Hl, Wl, 5, np.float32)
    out['lens_img8'], out['lens_imgf'] = img8, imgf
    out['lens_distort8'] = lens.distortImage(img8)
    out['lens_distortf'] = lens.distortImage(imgf)
    save('producers', **out)


if __name__ == '__main__':
    if sys.argv[1:] == ['producers']:
        producers()
    elif sys.argv[1:] == ['ste_nlf']:
        ste_nlf()
    elif sys.argv[1:] == ['perspective']:
        perspective()
    elif sys.argv[1:] == ['ste']:
        ste()
    else:
        main()
        perspective()
        ste()
        ste_nlf()
        producers()

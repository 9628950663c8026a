"""development: latency of the Python mirror's correct() for one frame (C1 1024x1024 f32, C2 4096x3000 u16)"""
import sys, os, time, io, contextlib, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from imgprocessor_b200 import synth
from imgprocessor_b200.camera import CameraCalibration, LensDistortion
for (H, W, dt) in ((1024, 1024, np.float32), (3000, 4096, np.uint16)):
    raw = synth.scene(H, W, 1, dt)
    cal = CameraCalibration()
    cal.addDarkCurrent(synth.dark_map(H, W))
    cal.addFlatField(synth.flat_map(H, W))
    p = synth.lens_moderate(H, W)
    l = LensDistortion({'cameraMatrix': synth.camera_matrix(p), 'distortionCoeffs': synth.dist_coeffs(p), 'shape': (H, W)})
    cal.addLens(l)
    with contextlib.redirect_stdout(io.StringIO()):
        cal.correct(raw)
        ts = []
        for i in range(5):
            t = time.perf_counter(); out = cal.correct(raw); ts.append(time.perf_counter() - t)
        pr = cProfile.Profile(); pr.enable(); cal.correct(raw); pr.disable()
    print('%dx%d %s: correct() %.1f ms (min of 5)' % (H, W, np.dtype(dt).name, min(ts) * 1e3))
    s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(14); print('\n'.join(s.getvalue().splitlines()[6:24]))

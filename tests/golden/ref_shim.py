"""Import shim that lets the UNMODIFIED reference (/root/reference, imgProcessor
0.2.5) import under numpy 2.x without its absent `fancytools` dependency.
Used only by make_golden.py in the build container; the reference does not
exist on the GPU box.  Nothing here patches reference code: it re-adds two numpy
aliases the reference uses (np.float, np.asfarray; CameraCalibration.py:408,410,
medianThreshold.py:18) and registers empty stand-ins for import-time-only modules
(SingleTimeEffectDetection.py:10, imgSignal.py:8, DarkCurrentMap.py)."""
import sys
import types

import numpy as np

REFERENCE_ROOT = '/root/reference'


def install():
    if not hasattr(np, 'float'):
        np.float = float
    if not hasattr(np, 'asfarray'):
        def asfarray(a, dtype=np.float64):
            dt = np.dtype(dtype)
            if not np.issubdtype(dt, np.inexact):
                dt = np.dtype(np.float64)
            return np.asarray(a, dtype=dt)
        np.asfarray = asfarray

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m

    for pkg in ('fancytools', 'fancytools.math', 'fancytools.os', 'fancytools.fit'):
        mod(pkg)
    mod('fancytools.math.MaskedMovingAverage', MaskedMovingAverage=None)
    mod('fancytools.math.findXAt', findXAt=None)
    mod('fancytools.os.PathStr', PathStr=str)
    mod('fancytools.math.linRegressUsingMasked2dArrays', linRegressUsingMasked2dArrays=None)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def install_perspective_stubs():
    """Extra import-time-only stand-ins for camera/PerspectiveCorrection.py:8-22 (transforms3d and more of
    fancytools are absent).  None of them is reached by PerspectiveCorrection.correct / uncorrect with a
    homography or quad reference and do_correctIntensity=False, which is what the golden files record."""
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m
    mod('transforms3d')
    mod('transforms3d.euler', mat2euler=None, euler2mat=None)
    mod('fancytools.math.Point3D', Point3D=None)
    mod('fancytools.math.vector3d', vectorAngle=None)
    sys.modules['fancytools.math'].line = mod('fancytools.math.line')


def install_masked_moving_average():
    """features/SingleTimeEffectDetection.py:10 imports fancytools.math.MaskedMovingAverage, which is
    absent.  Inject the restatement of its published algorithm (oracle/ste.py — the one ingredient of the STE
    branch whose parity stays unpinned) so that the reference's own STE code can be executed unmodified."""
    from oracle.ste import MaskedMovingAverage
    sys.modules['fancytools.math.MaskedMovingAverage'].MaskedMovingAverage = MaskedMovingAverage

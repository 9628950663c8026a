"""imgprocessor_b200 — the per-frame camera-correction path of radjkarl/imgProcessor
(CameraCalibration.correct(): dark current, flat field, medianThreshold, LensDistortion remap)
as hand-written sm_100a CUDA kernels behind the reference's Python API.

    from imgprocessor_b200.camera import CameraCalibration, LensDistortion
    from imgprocessor_b200.filters import medianThreshold

Submodules are imported lazily so that `import imgprocessor_b200` works on a box without a GPU
(e.g. to build the library); any compute call without a CUDA device raises — there is no CPU fallback.
"""
__version__ = '0.1.0'


def __getattr__(name):
    import importlib
    if name in ('camera', 'filters', 'engine', 'synth', 'sharding', 'build', '_lib', 'imgIO'):
        return importlib.import_module('.' + name, __name__)
    if name in ('CameraCalibration', 'LensDistortion'):
        return getattr(importlib.import_module('.camera', __name__), name)
    if name == 'medianThreshold':
        return importlib.import_module('.filters', __name__).medianThreshold
    raise AttributeError(name)

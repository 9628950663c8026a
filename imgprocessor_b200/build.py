"""Build libimgcorr.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m imgprocessor_b200.build [--force] [-v]

nvcc cross-compiles without a GPU; the resulting .so is git-ignored but travels to the GPU
box with the repository snapshot.
"""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, 'csrc')
ROOT = os.path.dirname(PKG)
OBJ = os.path.join(ROOT, 'build', 'imgcorr')
LIB = os.path.join(PKG, 'libimgcorr.so')
SOURCES = ['k1_pointwise_median.cu', 'k1_stream.cu', 'k1_stream5.cu', 'k2_undistort.cu', 'k3_warp.cu', 'k4_ste.cu', 'selftest.cu', 'k5_producers.cu', 'imgcorr_api.cu']
HEADERS = [os.path.join(CSRC, 'imgcorr_core.cuh'), os.path.join(CSRC, 'imgcorr_kernels.cuh'), os.path.join(CSRC, 'imgcorr_tma.cuh'), os.path.join(CSRC, 'imgcorr_warp.cuh'), os.path.join(CSRC, 'imgcorr_ste.cuh'), os.path.join(CSRC, 'median25_net.inc'), os.path.join(CSRC, 'median25_pair_net.inc'),
           os.path.join(ROOT, 'include', 'imgcorr.h')]
NVCC_FLAGS = ['-O3', '-std=c++17', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo',
              '-Xcompiler', '-fPIC,-ffp-contract=off,-fvisibility=hidden', '--expt-relaxed-constexpr']


def nvcc():
    exe = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(exe):
        raise RuntimeError('nvcc not found: libimgcorr.so cannot be built (there is no CPU fallback)')
    return exe


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, extra_flags=()):
    os.makedirs(OBJ, exist_ok=True)
    exe = nvcc()
    jobs = []
    objs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.replace('.cu', '.o'))
        objs.append(o)
        if force or _stale(o, [s] + HEADERS):
            jobs.append([exe] + NVCC_FLAGS + list(extra_flags) + ['-c', s, '-o', o])

    def run(cmd):
        if verbose:
            print(' '.join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('nvcc failed:\n%s\n%s' % (' '.join(cmd), r.stdout + r.stderr))
        if verbose and (r.stdout or r.stderr):
            print(r.stdout + r.stderr)

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        list(ex.map(run, jobs))
    if force or jobs or _stale(LIB, objs):
        run([exe, '-shared', '-o', LIB] + objs + ['-Xcompiler', '-fvisibility=hidden'])
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))

#!/usr/bin/env python
"""bench.py — corrected Mpx/s of the CameraCalibration.correct() chain on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--frames F] [--impl reference]

Workload (BASELINE.json configs[2]): a batch of 4096x3000 uint16 frames per GPU through the full chain
(dark subtract, flat-field divide, nan_to_num, 3x3 median-threshold, lens undistortion).  One step = one pass
over the whole batch.  Frames are sharded over ranks (weak scaling: F frames on every GPU), calibration maps
are built on rank 0 and broadcast once; there is no collective in the timed region.

One JSON line on rank 0:
  value        device-resident whole-job Mpx/s (inputs in HBM, CUDA-event timed, max over ranks)
  e2e          the same metric through the C-ABI host-buffer call (pinned host frames, H2D + kernels + D2H
               overlapped inside libimgcorr), host wall clock, max over ranks
  roofline     the dominant kernel (K1): algorithmic bytes (14 B/px) / event-timed launch duration
               inside the timed region, against the measured HBM copy peak (MEASURED_PEAKS.json)
  cpu_baseline the oracle's restatement of the reference path (scipy + OpenCV, float64) on one host core
`--impl reference` times that reference path on all host cores instead (rank 0 only).
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W = 3000, 4096
METRIC = 'corrected Mpx/s, CameraCalibration.correct() chain'
K1_BYTES_PER_PX = 14.0      # raw u16 2 + dark 4 + flat 4 + out 4   
K2_BYTES_PER_PX = 8.0       # src 4 + dst 4
FALLBACK_HBM_GBS = 6650.0   # /opt/skills/guides/B200_PROFILING.md


def workload_config(frames, n_gpus):
    return {
        'workload': 'BASELINE configs[2]: batch of %d 4096x3000 uint16 frames per GPU, full correct() chain '
                    '(dark + flat + nan_to_num + 3x3 medianThreshold + LensDistortion remap), moderate 5-coeff lens' % frames,
        'frame_shape': [H, W], 'frames_per_gpu': frames, 'global_frames': frames * n_gpus,
        'threshold': 0.1, 'keep_size': True, 'out_dtype': 'float32',
        'l2': 'inputs (%.1f GB of raw frames per GPU) are far larger than the 126 MB L2; no flush needed' % (frames * H * W * 2 / 1e9),
        'sharding': 'frames split contiguously over ranks, calibration broadcast once, no steady-state collective',
    }


def measured_peak():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    except Exception:
        return FALLBACK_HBM_GBS, 'fallback (B200_PROFILING.md)'


class ClockSampler(object):
    """SM clock + throttle reasons of one GPU, sampled while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {}
        for n in dir(nv):
            if n.startswith('nvmlClocksThrottleReason') or n.startswith('nvmlClocksEventReason'):
                v = getattr(nv, n)
                if isinstance(v, int) and v:
                    names[v] = n.replace('nvmlClocksThrottleReason', '').replace('nvmlClocksEventReason', '')
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    bits = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    bits = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for b, n in names.items():
                    if bits & b and n not in ('None', 'GpuIdle', 'All'):
                        self.reasons.add(n)
            except Exception:
                pass
            self._stop.wait(0.005)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self):
        if not self.samples:
            return {'sm_mhz': None, 'sm_max_mhz': self.max_mhz, 'reasons': ['unavailable']}
        s = sorted(self.samples)
        return {'sm_mhz': s[len(s) // 2], 'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons), 'samples': len(s)}


# ------------------------------------------------------------------------------------------------ reference arm
def _ref_inputs(seed):
    import numpy as np
    from imgprocessor_b200 import synth
    raw = synth.scene(H, W, seed, np.uint16)
    dark, flat = synth.dark_map(H, W), synth.flat_map(H, W)
    p = synth.lens_moderate(H, W)
    return raw, dark, flat, (synth.camera_matrix(p), synth.dist_coeffs(p))


_REF_STATE = {}


def _ref_worker_init():
    import cv2
    cv2.setNumThreads(1)
    raw, dark, flat, lens = _ref_inputs(100 + os.getpid() % 7)
    _REF_STATE.update(raw=raw, dark=dark, flat=flat, lens=lens)


def _ref_worker_step(_):
    from oracle import refpath
    s = _REF_STATE
    out = refpath.correct(s['raw'], s['dark'], s['flat'], s['lens'], 0.1, True)
    return float(out[0, 0])


def run_reference(args):
    """the reference's own CPU implementation of the path (oracle.refpath = its Python chain restated over the
    same scipy / OpenCV / numpy natives), one frame per worker per step, all host cores."""
    import multiprocessing as mp
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    ctx = mp.get_context('fork')
    with ctx.Pool(cores, initializer=_ref_worker_init) as pool:
        for _ in range(max(args.warmup, 1)):
            pool.map(_ref_worker_step, range(cores))
        t0 = time.perf_counter()
        for _ in range(args.steps):
            pool.map(_ref_worker_step, range(cores))
        dt = time.perf_counter() - t0
    px = float(cores) * H * W * args.steps
    val = px / dt / 1e6
    import cv2
    import numpy
    import scipy
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': 'Mpx/s', 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': dt / args.steps * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': workload_config(args.frames, args.gpus),
        'cpu_baseline': {'value': val, 'unit': 'Mpx/s', 'cores': cores, 'kind': 'port',
                         'sample': 'each step = one 4096x3000 uint16 frame per worker through the float64 reference chain '
                                   '(oracle.refpath: scipy %s median_filter, OpenCV %s remap, numpy %s), %d worker processes'
                                   % (scipy.__version__, cv2.__version__, numpy.__version__, cores)},
        'e2e': {'value': val, 'unit': 'Mpx/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


def cpu_baseline_single(frames=3):
    """the oracle's reference-path restatement, one process, one thread, `frames` frames"""
    import cv2
    import numpy
    import scipy
    from oracle import refpath
    cv2.setNumThreads(1)
    raw, dark, flat, lens = _ref_inputs(100)
    refpath.correct(raw[:256, :256], dark[:256, :256], flat[:256, :256], None, 0.1)      # import / warm-up
    best = None
    for _ in range(frames):
        t0 = time.perf_counter()
        refpath.correct(raw, dark, flat, lens, 0.1, True)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return {'value': H * W / best / 1e6, 'unit': 'Mpx/s', 'cores': 1, 'kind': 'port',
            'sample': 'best of %d single 4096x3000 uint16 frames through oracle.refpath.correct (float64; scipy %s '
                      'median_filter single-threaded, OpenCV %s with 1 thread, numpy %s), %.2f s/frame'
                      % (frames, scipy.__version__, cv2.__version__, numpy.__version__, best)}


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from imgprocessor_b200 import _lib, engine, sharding, synth

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py needs a CUDA device; there is no CPU fallback (use --impl reference for the CPU arm)')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    distributed = world > 1
    if distributed:
        dist.init_process_group('nccl', device_id=dev)

    F = args.frames
    eng = engine.Engine(H, W, local)

    # calibration: built on rank 0, broadcast once over NCCL, uploaded into the library's per-device copies
    p = synth.lens_moderate(H, W)
    Kmat, dvec = synth.camera_matrix(p), synth.dist_coeffs(p)
    import cv2
    P, roi = cv2.getOptimalNewCameraMatrix(Kmat, dvec, (W, H), 1, (W, H))
    if distributed:
        maps = {'dark': synth.dark_map(H, W), 'flat': synth.flat_map(H, W)} if rank == 0 else {'dark': None, 'flat': None}
        cal = sharding.broadcast_calibration(maps, src=0, device=dev)
        eng.set_dark(cal['dark'])
        eng.set_flat(cal['flat'])
    else:
        eng.set_dark(synth.dark_map(H, W))
        eng.set_flat(synth.flat_map(H, W))
    eng.set_lens(Kmat, dvec, P)
    eng.set_option(_lib.OPT_CHAIN_GROUP, args.group)
    eng.set_option(_lib.OPT_CHAIN_OVERLAP, int(args.overlap))

    numa = {'bound': False} if os.environ.get('IMGCORR_NO_NUMA_BIND') else sharding.bind_host_to_gpu(local)
    raw = synth.scene_torch(F, H, W, 1000 + rank, dev, 'uint16')
    out = torch.empty((F, H, W), dtype=torch.float32, device=dev)
    px_step = float(F) * H * W

    def barrier():
        if distributed:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize()

    def step():
        eng.correct_batch(raw, threshold=0.1, ksize=3, out=out)

    # ---- device-resident timing ------------------------------------------------------------------
    for _ in range(args.warmup):
        step()
    eng.set_option(_lib.OPT_PROFILE, max(1, F // 32))      # event-bracket ~32 K1 / K2 launches per step
    eng.profile_read()
    launches0 = eng.launch_count
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        ev0.record()
        for _ in range(args.steps):
            step()
        ev1.record()
        torch.cuda.synchronize()
    ms_total = ev0.elapsed_time(ev1)
    barrier()
    launches = eng.launch_count - launches0
    k1_ms, k1_frames, k2_ms, k2_frames = eng.profile_read()
    eng.set_option(_lib.OPT_PROFILE, 0)
    if distributed:
        ms_total = sharding.reduce_max(ms_total, dev)
        launches = sharding.reduce_sum(launches, dev)
    value = px_step * world * args.steps / (ms_total * 1e-3) / 1e6

    # ---- end to end: host frames through the C-ABI host call --------------------------------------
    pool = min(F, args.e2e_pool)
    h_in = engine.pinned_empty((pool, H, W), np.uint16)
    h_out = engine.pinned_empty((pool, H, W), np.float32)
    h_in[...] = raw[:pool].cpu().numpy()
    reps = max(1, F // pool)

    def e2e_step():
        for _ in range(reps):
            eng.correct_host(h_in, out=h_out, threshold=0.1, ksize=3)

    e2e_step()
    barrier()
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    e2e_s = time.perf_counter() - t0
    if distributed:
        e2e_s = sharding.reduce_max(e2e_s, dev)
    e2e_px_step = float(reps * pool) * H * W
    e2e_value = e2e_px_step * world * e2e_steps / e2e_s / 1e6
    checksum = float(h_out[0, H // 2, ::512].astype(np.float64).sum())

    peak, peak_src = measured_peak()
    k1_us = k1_ms * 1e3 / max(k1_frames, 1.0)
    k2_us = k2_ms * 1e3 / max(k2_frames, 1.0)
    achieved = K1_BYTES_PER_PX * H * W / (k1_us * 1e-6) / 1e9 if k1_us > 0 else 0.0
    k2_achieved = K2_BYTES_PER_PX * H * W / (k2_us * 1e-6) / 1e9 if k2_us > 0 else 0.0

    if rank == 0:
        line = {
            'metric': METRIC, 'value': value, 'unit': 'Mpx/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms_total / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic', 'config': workload_config(F, world),
            'clocks': clocks.summary(),
            'e2e': {'value': e2e_value, 'unit': 'Mpx/s', 'h2d_bytes_per_step': int(reps * pool * H * W * 2),
                    'd2h_bytes_per_step': int(reps * pool * H * W * 4), 'frames_per_step_per_gpu': reps * pool,
                    'steps': e2e_steps, 'pinned_pool_frames': pool, 'checksum': checksum,
                    'api': 'imgcorr_correct_host (C ABI) with pinned host buffers; copies inside the timed region',
                    'host_numa_binding': numa},
            'gpu_launches': int(launches),
            'roofline': {'bound': 'hbm', 'kernel': 'K1 k1_stream_kernel (fused dark/flat/nan_to_num/3x3 median-threshold)',
                         'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak, 'traffic': args.k1_traffic,
                         'peak_source': peak_src, 'algorithmic_bytes_per_px': K1_BYTES_PER_PX,
                         'us_per_frame': k1_us, 'frames_timed': k1_frames, 'frames_per_launch': min(args.group, F),
                         'launch_us': k1_us * min(args.group, F),
                         'algorithmic_bytes_per_launch': K1_BYTES_PER_PX * H * W * min(args.group, F),
                         'k2': {'achieved': k2_achieved, 'frac': k2_achieved / peak, 'us_per_frame': k2_us,
                                'algorithmic_bytes_per_px': K2_BYTES_PER_PX},
                         'chain_frac': (K1_BYTES_PER_PX + K2_BYTES_PER_PX) * H * W / ((k1_us + k2_us) * 1e-6) / 1e9 / peak
                                       if k1_us + k2_us > 0 else 0.0,
                         'chain_overlap': bool(args.overlap),
                         'timing_note': 'K1 / K2 durations are event-bracketed launches inside the timed region; with chain_overlap '
                                        'the bracketed frame groups run unoverlapped (kernel alone), the others overlap K1 of group '
                                        'g+1 with K2 of group g, so ms_per_step can be below frames * (K1 + K2)'},
        }
        if world == 1 and not args.no_cpu_baseline:
            line['cpu_baseline'] = cpu_baseline_single()
        print(json.dumps(line), flush=True)
    if distributed:
        dist.barrier(device_ids=[local])
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--frames', type=int, default=256, help='frames per GPU per step')
    ap.add_argument('--group', type=int, default=32, help='frames per K1/K2 launch inside the chain')
    ap.add_argument('--overlap', type=int, default=0, help='1: K1 of the next group overlaps K2 of the current one')
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--e2e-pool', type=int, default=32, help='pinned host frames cycled by the end-to-end leg')
    ap.add_argument('--e2e-steps', type=int, default=3)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    if args.gpus > 1 and 'WORLD_SIZE' not in os.environ and args.impl == 'ours':
        # started as plain `python bench.py --gpus N`: relaunch under torchrun, one rank per GPU (the driver does this itself)
        import socket
        with socket.socket() as sk:
            sk.bind(('127.0.0.1', 0))
            port = sk.getsockname()[1]
        os.execv(sys.executable, [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(args.gpus),
                                  '--master-addr', '127.0.0.1', '--master-port', str(port), os.path.abspath(__file__)] + sys.argv[1:])
    if 'WORLD_SIZE' in os.environ and int(os.environ['WORLD_SIZE']) != args.gpus and args.impl == 'ours':
        raise SystemExit('bench.py: --gpus %d but WORLD_SIZE=%s' % (args.gpus, os.environ['WORLD_SIZE']))
    args.k1_traffic = None
    try:        # dram__bytes_read.sum + dram__bytes_write.sum of one K1 launch from the committed ncu --set full capture
        with open(os.path.join(ROOT, 'profiles', 'roofline_traffic.json')) as f:
            t = json.load(f)
        if int(t['frames_per_launch']) == args.group:          # the capture is of a launch of that many frames
            args.k1_traffic = float(t['k1_dram_bytes_per_launch'])
    except Exception:
        pass
    if args.warmup < 3 and args.impl == 'ours':
        args.warmup = 3
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()

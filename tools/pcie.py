"""Raw page-locked copy bandwidth of the box — the ceiling of every host-buffer number (bench.py e2e, configs[4]).

    python tools/pcie.py [--gpus 1,2,4,8] [--frame-bytes-in N] [--frame-bytes-out N]

For each GPU count N: N processes (one per GPU, started together behind a barrier) run cudaMemcpyAsync of pinned buffers
H2D only, D2H only and both at once (two streams), with the frame sizes of the 4096x3000 uint16 -> float32 chain by
default.  Prints one JSON object {N: {h2d, d2h, both}} with the sum over the N GPUs in GB/s; bench.py measures the same
thing inside its own run (e2e.pcie_ceiling_detail)."""
import argparse
import json
import multiprocessing as mp
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def worker(rank, n, in_bytes, out_bytes, barrier, q):
    import torch
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    h_in = torch.empty(in_bytes, dtype=torch.uint8, pin_memory=True)
    h_out = torch.empty(out_bytes, dtype=torch.uint8, pin_memory=True)
    d_in = torch.empty(in_bytes, dtype=torch.uint8, device=dev)
    d_out = torch.empty(out_bytes, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def run(do_in, do_out, reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            if do_in:
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if do_out:
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        return time.perf_counter() - t0

    res = {}
    for name, di, do in (('h2d', True, False), ('d2h', False, True), ('both', True, True)):
        run(di, do, 2)
        barrier.wait()
        t = run(di, do, 24)
        res[name] = 24 * ((in_bytes if di else 0) + (out_bytes if do else 0)) / t / 1e9
        barrier.wait()
    q.put((rank, res))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', default='1')
    ap.add_argument('--frame-bytes-in', type=int, default=4 * 3000 * 4096 * 2)
    ap.add_argument('--frame-bytes-out', type=int, default=4 * 3000 * 4096 * 4)
    a = ap.parse_args()
    ctx = mp.get_context('spawn')
    table = {}
    for n in [int(v) for v in a.gpus.split(',')]:
        barrier, q = ctx.Barrier(n), ctx.Queue()
        procs = [ctx.Process(target=worker, args=(r, n, a.frame_bytes_in, a.frame_bytes_out, barrier, q)) for r in range(n)]
        for p in procs:
            p.start()
        got = [q.get() for _ in procs]
        for p in procs:
            p.join()
        table[n] = {k: round(sum(r[1][k] for r in got), 1) for k in ('h2d', 'd2h', 'both')}
        table[n]['per_gpu_both'] = round(table[n]['both'] / n, 1)
    print(json.dumps({'pinned_copy_gbs_sum_over_gpus': table, 'cpus': os.cpu_count()}), flush=True)


if __name__ == '__main__':
    main()

"""Generate the selection networks of the 5x5 streaming kernel's row-PAIR scheme (csrc/median25_pair_net.inc).

Two vertically adjacent 5x5 windows share four of their five rows.  With every row already sorted (the kernel sorts the
horizontal quintuple of each input row once):

  net A   4 sorted quintuples (20 values)  ->  the values of rank 7..12 of the 20, in order ("band", 6 values)
  net B   band (6 sorted) + one more sorted quintuple  ->  the value of rank 5 of these 11

The median of 25 = rank 12 of (20 shared values + 5 of the window's own fifth row).  Of the 20, those of rank <= 6 have at
most 6 + 5 = 11 values below them, those of rank >= 13 at least 13: neither can be the median, and exactly 7 values lie
below the band, so the median is the value of rank 12 - 7 = 5 among band + fifth row.  Net A runs once per PAIR of output
rows, net B once per output row.

Comparator (a, b): min -> wire a, max -> wire b.  Correctness by the 0-1 principle restricted to inputs with sorted rows:
6^4 inputs for A, 7 * 6 for B, and the composition is checked on all 6^5 inputs of a full window.  Networks start from
Batcher sorts over several wire orders and are pruned greedily against the required output wires.

    python tools/gen_median25_pair.py
"""
import itertools
import random

import numpy as np


def batcher(n):
    p2 = 1
    while p2 < n:
        p2 *= 2
    net = []
    p = 1
    while p < p2:
        k = p
        while k >= 1:
            for j in range(k % p, p2 - k, 2 * k):
                for i in range(min(k, p2 - j - k)):
                    if (i + j) // (2 * p) == (i + j + k) // (2 * p):
                        a, b = i + j, i + j + k
                        if b < n:
                            net.append((a, b))
            k //= 2
        p *= 2
    return net


def sorted_lists_inputs(lengths):
    """all 0/1 inputs whose consecutive blocks of the given lengths are sorted ascending -> [wires][cases]"""
    blocks = [[np.array([0] * (L - k) + [1] * k, dtype=np.uint8) for k in range(L + 1)] for L in lengths]
    cols = [np.concatenate(c) for c in itertools.product(*blocks)]
    return np.array(cols, dtype=np.uint8).T.copy()


def apply(net, x):
    x = x.copy()
    for a, b in net:
        lo = x[a] & x[b]
        hi = x[a] | x[b]
        x[a], x[b] = lo, hi
    return x


def want_rank(x0, rank):
    n = x0.shape[0]
    ones = x0.sum(axis=0)
    return (ones >= n - rank).astype(np.uint8)          # sorted ascending: position `rank` holds 1 iff ones >= n - rank


def correct(net, x0, outs):
    """outs: list of (wire, rank)"""
    x = apply(net, x0)
    return all(np.array_equal(x[w], want_rank(x0, r)) for w, r in outs)


def prune(net, x0, outs, seed):
    rng = random.Random(seed)
    net = list(net)
    assert correct(net, x0, outs)
    changed = True
    rounds = 0
    while changed and rounds < 8:
        changed = False
        rounds += 1
        idxs = list(range(len(net)))
        if rounds % 2 == 0:
            rng.shuffle(idxs)
        else:
            idxs.reverse()
        for idx in sorted(idxs, reverse=True) if rounds % 2 else idxs:
            if idx >= len(net):
                continue
            trial = net[:idx] + net[idx + 1:]
            if correct(trial, x0, outs):
                net = trial
                changed = True
    return net


def liveness(net, out_wires):
    need = set(out_wires)
    live = []
    for a, b in reversed(net):
        nl, nh = a in need, b in need
        if not (nl or nh):
            continue
        live.append((a, b, nl, nh))
        need.add(a)
        need.add(b)
    live.reverse()
    full = sum(1 for _, _, nl, nh in live if nl and nh)
    half = len(live) - full
    return live, full, half


def search(n, lengths, ranks, orders, seeds=4, presort=None):
    x0 = sorted_lists_inputs(lengths)
    best = None
    for oname, order in orders.items():
        base = [(order[a], order[b]) for a, b in batcher(n)]
        if presort:
            base = presort + base
        outs = [(order[r], r) for r in ranks]
        if not correct(base, x0, outs):
            continue
        for seed in range(seeds):
            p = prune(base, x0, outs, seed)
            live, full, half = liveness(p, [w for w, _ in outs])
            cost = 3 * full + half                      # integer keys: VIMNMX + 2 IMAD per exchange, 1 op for a half exchange
            print('  %-10s seed %d: %2d exchanges + %2d half = %3d min/max ops, cost %3d' % (oname, seed, full, half, 2 * full + half, cost))
            if best is None or cost < best[0]:
                best = (cost, live, outs, full, half, oname)
    return best


def main():
    # ---- net A: wires 5*i + j = j-th smallest of shared row i (4 rows)
    def orders20():
        w = list(range(20))
        return {
            'rowmajor': w,
            'colmajor': sorted(w, key=lambda k: (k % 5, k // 5)),
            'diag': sorted(w, key=lambda k: (k // 5 + k % 5, k // 5)),
            'diag2': sorted(w, key=lambda k: (k // 5 + k % 5, k % 5)),
            'prod': sorted(w, key=lambda k: ((k // 5 + 1) * (k % 5 + 1), k)),
        }
    s4 = [(0, 1), (2, 3), (0, 2), (1, 3), (1, 2)]
    cols = [(5 * a + j, 5 * b + j) for j in range(5) for a, b in s4]
    print('net A (4 sorted quintuples -> ranks 7..12):')
    ba = search(20, [5, 5, 5, 5], range(7, 13), orders20())
    print(' with the rank-columns sorted first:')
    bb = search(20, [5, 5, 5, 5], range(7, 13), orders20(), presort=cols)
    a = ba if ba[0] <= bb[0] else bb
    print('net B (6 sorted + 5 sorted -> rank 5):')
    w11 = list(range(11))
    orders11 = {
        'concat': w11,
        'interleave': [0, 6, 1, 7, 2, 8, 3, 9, 4, 10, 5],
        'interleave2': [6, 0, 7, 1, 8, 2, 9, 3, 10, 4, 5],
        'rev': [6, 7, 8, 9, 10, 0, 1, 2, 3, 4, 5],
    }
    b = search(11, [6, 5], [5], orders11, seeds=6)
    cost_a, live_a, outs_a, fa, ha, na = a
    cost_b, live_b, outs_b, fb, hb, nb = b
    # ---- composition check on a full window: rows 0..3 shared (wires 0..19), row 4 the window's own (wires 20..24)
    x0 = sorted_lists_inputs([5, 5, 5, 5, 5])
    x = apply([(p, q) for p, q, _, _ in live_a], x0)
    band = [w for w, _ in outs_a]                       # wires holding ranks 7..12, ascending
    y = np.concatenate([x[band], x0[20:25]])            # net B input layout: band 0..5, row 6..10
    y = apply([(p, q) for p, q, _, _ in live_b], y)
    med = (x0.sum(axis=0) >= 25 - 12).astype(np.uint8)
    assert np.array_equal(y[outs_b[0][0]], med), 'composition A + B is not the median of 25'
    print('composition verified on %d window inputs' % x0.shape[1])
    per_px = (cost_a / 2.0 + cost_b)
    print('per output pixel: A/2 + B = %.1f integer-key instructions (+ the 9-exchange row sort); was 57 exchanges' % per_px)
    path = 'imgprocessor_b200/csrc/median25_pair_net.inc'
    with open(path, 'w') as f:
        f.write('// generated by tools/gen_median25_pair.py: the row-pair scheme of the 5x5 streaming kernel.\n')
        f.write('// net A (start %s): 4 sorted quintuples, wire 5*i+j = j-th smallest of shared row i -> ranks 7..12 of the 20 on wires\n' % na)
        f.write('//   M25A_BAND0..5; %d exchanges + %d half exchanges.  Verified on all 1296 0/1 inputs with sorted rows.\n' % (fa, ha))
        f.write('// net B (start %s): wires 0..5 = band (ascending), 6..10 = the fifth row (ascending) -> rank 5 of the 11 on wire M25B_RESULT;\n' % nb)
        f.write('//   %d exchanges + %d half exchanges.  Verified on all 42 0/1 inputs; A + B verified on all 7776 window inputs.\n' % (fb, hb))
        f.write('// (a, b): min -> wire a, max -> wire b.\n')
        for k, w in enumerate(band):
            f.write('#define M25A_BAND%d %d\n' % (k, w))
        f.write('#define M25B_RESULT %d\n' % outs_b[0][0])
        f.write('#define M25A_NET \\\n')
        f.write(' \\\n'.join(('    M25_CE(%d, %d)' if nl and nh else '    M25_LO(%d, %d)' if nl else '    M25_HI(%d, %d)') % (p, q) for p, q, nl, nh in live_a) + '\n')
        f.write('#define M25B_NET \\\n')
        f.write(' \\\n'.join(('    M25_CE(%d, %d)' if nl and nh else '    M25_LO(%d, %d)' if nl else '    M25_HI(%d, %d)') % (p, q) for p, q, nl, nh in live_b) + '\n')
    print('->', path)


if __name__ == '__main__':
    main()

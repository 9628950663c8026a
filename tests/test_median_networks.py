"""The generated selection networks of the 5x5 kernels, re-verified from the COMMITTED .inc files (not from the generator):
0-1 principle over every input whose rows are sorted.  csrc/median25_net.inc: median of 25 from five sorted quintuples;
csrc/median25_pair_net.inc: net A (four sorted quintuples -> ranks 7..12) and net B (band + fifth row -> rank 5), and their
composition = the median of 25 (tools/gen_median25.py, tools/gen_median25_pair.py)."""
import itertools
import os
import re

import numpy as np

CSRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'imgprocessor_b200', 'csrc')


def sorted_inputs(lengths):
    blocks = [[np.array([0] * (L - k) + [1] * k, dtype=np.uint8) for k in range(L + 1)] for L in lengths]
    return np.array([np.concatenate(c) for c in itertools.product(*blocks)], dtype=np.uint8).T.copy()


def run(ops, x):
    x = x.copy()
    for kind, a, b in ops:
        lo, hi = x[a] & x[b], x[a] | x[b]
        if kind in ('CE', 'LO'):
            new_a = lo
        if kind in ('CE', 'HI'):
            x[b] = hi
        if kind in ('CE', 'LO'):
            x[a] = new_a
    return x


def parse(text):
    return [(m.group(1), int(m.group(2)), int(m.group(3))) for m in re.finditer(r'M25_(CE|LO|HI)\((\d+),\s*(\d+)\)', text)]


def test_median25_network():
    txt = open(os.path.join(CSRC, 'median25_net.inc')).read()
    wire = int(re.search(r'#define M25_RESULT_WIRE (\d+)', txt).group(1))
    ops = parse(txt)
    x0 = sorted_inputs([5] * 5)
    assert x0.shape == (25, 7776)
    got = run(ops, x0)[wire]
    assert np.array_equal(got, (x0.sum(axis=0) >= 13).astype(np.uint8))


def test_median25_pair_networks():
    txt = open(os.path.join(CSRC, 'median25_pair_net.inc')).read()
    band = [int(re.search(r'#define M25A_BAND%d (\d+)' % k, txt).group(1)) for k in range(6)]
    res = int(re.search(r'#define M25B_RESULT (\d+)', txt).group(1))
    a_txt = txt[txt.index('#define M25A_NET'):txt.index('#define M25B_NET')]
    b_txt = txt[txt.index('#define M25B_NET'):]
    net_a, net_b = parse(a_txt), parse(b_txt)
    assert len(net_a) > 30 and len(net_b) > 8
    # net A alone: ranks 7..12 of the 20 shared values, ascending on the band wires
    xa = sorted_inputs([5] * 4)
    ya = run(net_a, xa)
    ones = xa.sum(axis=0)
    for k, w in enumerate(band):
        assert np.array_equal(ya[w], (ones >= 20 - (7 + k)).astype(np.uint8)), k
    # net B alone: rank 5 of a sorted 6-list and a sorted 5-list
    xb = sorted_inputs([6, 5])
    assert np.array_equal(run(net_b, xb)[res], (xb.sum(axis=0) >= 11 - 5).astype(np.uint8))
    # composition on a whole window: rows 0..3 shared, row 4 the window's own -> the median of 25
    x0 = sorted_inputs([5] * 5)
    y = run(net_a, x0[:20])
    z = run(net_b, np.concatenate([y[band], x0[20:25]]))[res]
    assert np.array_equal(z, (x0.sum(axis=0) >= 13).astype(np.uint8))

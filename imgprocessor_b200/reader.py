"""Readers for the raw formats that feed correct(): mirrors of imgProcessor.reader.RAW.RAW (reader/RAW.py:15-30) and
imgProcessor.reader.elbin.elbin (reader/elbin.py:7-40), plus layout helpers that let the GPU path consume the FILE
BYTES directly (K1 swaps big-endian samples and skips the per-frame headers in its load):

    arr = RAW(path, width, height, '16-bit Unsigned')            # big-endian numpy array, as the reference returns
    out = cal.correct_batch(arr[None])                            # no host byte swap: the bytes go to the device as stored

    lay = elbin_layout(path)                                      # file image + where the frames are
    out = engine.correct_file_bytes(torch.from_numpy(lay['bytes']).cuda(), lay['offset'], lay['frames'], gap=lay['gap'])
"""
from collections import OrderedDict

import numpy as np

STR_TO_DTYPE = OrderedDict((('8-bit', 'u1'), ('16-bit Signed', 'i2'), ('16-bit Unsigned', 'u2'), ('32-bit Signed', 'i4'),
                            ('32-bit Unsigned', 'u4'), ('32-bit Real/floating point', 'f4')))

# exposure times [s] selectable in the RELTRON EL software, indexed by the per-frame header field (reader/elbin.py:17-21)
ELBIN_TIMES = (0.3, 0.4, 0.6, 0.8, 1.2, 1.6, 2.4, 3.2, 4.8, 6.4, 9.6, 12.8, 19.2, 25.6, 38.4, 51.2, 76.8, 102.6, 153.6,
               204.6, 307.2, 409.8, 614.4, 819., 1228.8, 1638.6, 3276.6, 5400., 8100., 12168., 18216., 27324., 41004.,
               61488., 92268.)
ELBIN_FILE_HEADER = 12          # uint32 height, width, frames
ELBIN_FRAME_HEADER = 20         # float64 current, float64 voltage, uint32 exposure-time index


def RAW(filename, width, height, dtype, littleEndian=False):
    """headerless raw image; big-endian unless told otherwise; shaped (width, height) like the reference (which reshapes
    to (s0, s1) = (width, height), RAW.py:22-25) and re-derives the second extent if the file is shorter"""
    dtype = STR_TO_DTYPE.get(dtype, dtype)
    if not littleEndian:
        dtype = '>' + dtype
    arr = np.fromfile(filename, dtype=dtype, count=width * height)
    try:
        return arr.reshape(width, height)
    except ValueError:
        return arr.reshape(width, arr.shape[0] // width)


def elbin(filename):
    """-> (frames uint16 [n][width][height], labels) exactly like the reference reader"""
    lay = elbin_layout(filename)
    b = lay['bytes']
    n, w, h = lay['frames'], lay['shape'][0], lay['shape'][1]
    arrs = np.empty((n, w, h), dtype=np.uint16)
    stride = w * h * 2 + lay['gap']
    for i in range(n):
        o = lay['offset'] + i * stride
        arrs[i] = b[o:o + w * h * 2].view(np.uint16).reshape(w, h)
    return arrs, lay['labels']


def elbin_layout(filename):
    """file image of an .elbin stack and where its frames are: dict(bytes=uint8 array of the whole file, offset=byte
    offset of the first frame's pixels, gap=bytes between frames, frames=n, shape=(rows, cols) of a frame as the
    reference shapes it, labels=[...])"""
    b = np.fromfile(filename, dtype=np.uint8)
    height, width, frames = (int(v) for v in b[:ELBIN_FILE_HEADER].view(np.uint32))
    px = width * height * 2
    labels = []
    for i in range(frames):
        o = ELBIN_FILE_HEADER + i * (ELBIN_FRAME_HEADER + px)
        current, voltage = (float(v) for v in b[o:o + 16].view(np.float64))
        i_time = int(b[o + 16:o + 20].view(np.uint32)[0])
        labels.append({'exposure time[s]': ELBIN_TIMES[i_time], 'current[A]': current, 'voltage[V]': voltage})
    return dict(bytes=b, offset=ELBIN_FILE_HEADER + ELBIN_FRAME_HEADER, gap=ELBIN_FRAME_HEADER, frames=frames,
                shape=(width, height), labels=labels)

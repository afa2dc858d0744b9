"""Loaders for the committed golden fixtures (tests/golden/*.json, made by make_golden.py)."""
import json
import os

import numpy as np

import oracle_lib as ol
from oracle_lib import pyref

HERE = os.path.dirname(os.path.abspath(__file__))


def load(name):
    with open(os.path.join(HERE, "golden", name)) as fh:
        return json.load(fh)


def ints(xs):
    return [int(x, 16) for x in xs]


def g1_from_dlogs(dlogs):
    """bases with the given discrete logs (0 -> infinity) as arkworks affine images, via the oracle."""
    lib = ol.oracle()
    out = np.zeros((len(dlogs), 72), dtype=np.uint8)
    lib.zko_g1_fixed_base(ol._p(ol.fr_np(dlogs)), len(dlogs), out.ctypes.data, 72)
    return out


def g2_from_dlogs(dlogs):
    lib = ol.oracle()
    out = np.zeros((len(dlogs), 136), dtype=np.uint8)
    lib.zko_g2_fixed_base(ol._p(ol.fr_np(dlogs)), len(dlogs), out.ctypes.data, 136)
    return out


def g1_point(js):
    return None if js is None else (int(js[0], 16), int(js[1], 16))


def g2_point(js):
    if js is None:
        return None
    return (pyref.Fq2(int(js[0][0], 16), int(js[0][1], 16)), pyref.Fq2(int(js[1][0], 16), int(js[1][1], 16)))

"""Index arithmetic of the slab-ordered M kernel (thunder_b200/csrc/thb_slab.cuh, the source the CUDA kernel compiles) built
for the host and checked by brute force: for random and degenerate rotations and several slab thicknesses every pixel is a
candidate of the slab that holds its exact cell base, never twice, and the candidate lists stay close to the exact ones
(the kernel's lane efficiency).  CPU only."""
import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np
import pytest

from oracle import portapi
from thunder_b200 import synth

ROOT = Path(__file__).resolve().parent.parent
SRC = ROOT / "tests" / "host_slab" / "slab_host.cpp"
LIB = ROOT / "tests" / "host_slab" / "libslab_host.so"


@pytest.fixture(scope="module")
def lib():
    hdr = ROOT / "thunder_b200" / "csrc" / "thb_slab.cuh"
    if not LIB.exists() or LIB.stat().st_mtime < max(SRC.stat().st_mtime, hdr.stat().st_mtime):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-I", str(ROOT / "thunder_b200" / "csrc"),
                               "-o", os.fspath(LIB), os.fspath(SRC)])
    L = C.CDLL(os.fspath(LIB))
    L.slab_check.restype = C.c_int
    L.slab_check.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    return L


def _check(L, pix, pf, vdim, q, th):
    iCol = np.ascontiguousarray(pix["iCol"], np.int32); iRow = np.ascontiguousarray(pix["iRow"], np.int32)
    q = np.ascontiguousarray(q, np.float64)
    cand = C.c_longlong(0); hits = C.c_longlong(0)
    rc = L.slab_check(pf, len(iCol), iCol.ctypes.data, iRow.ctypes.data, q.ctypes.data, vdim, th, C.byref(cand), C.byref(hits))
    assert rc == 0, f"slab_check rc={rc} for q={q}, th={th}"
    assert hits.value == len(iCol)
    return cand.value / hits.value


@pytest.mark.parametrize("N,r", [(16, 7.0), (64, 31.0), (256, 127.0)])
def test_every_sample_in_exactly_one_slab(lib, N, r):
    pf = 2
    pix = portapi.pixel_list(N, pf, r, 0.0)
    rng = np.random.default_rng(N)
    h = np.sqrt(0.5)
    quats = [[1, 0, 0, 0], [h, h, 0, 0], [h, 0, h, 0], [h, 0, 0, h], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1], [0.5, 0.5, 0.5, 0.5],
             [h, -h, 0, 0], [h, 0, -h, 0], [0, h, h, 0], [0, h, 0, h], [0, 0, h, h]]
    # almost-degenerate: tiny tilts off the axis-aligned slices (slopes of 1e-12 .. 1e-3)
    for eps in (1e-12, 1e-9, 1e-7, 1e-5, 1e-3):
        for base in ([1, 0, 0, 0], [h, 0, h, 0], [h, h, 0, 0]):
            q = np.array(base, float) + eps * rng.normal(size=4)
            quats.append(q / np.linalg.norm(q))
    for q in quats:
        for th in ((1, 3, 7, pf * N) if N < 256 else (11, 22)):
            ratio = _check(lib, pix, pf, pf * N, q, th)
            assert N < 256 or ratio < 4.0, (q, th, ratio)    # even the degenerate slices stay cheap at practical thicknesses
    over = []
    for q in synth.random_quats(12 if N == 256 else 40, rng):
        for th in ((1, 3, 7, pf * N) if N < 256 else (11, 22)):
            over.append(_check(lib, pix, pf, pf * N, q, th))
    # conservative, but not wasteful: at the thickness the kernel uses (box 256: 22 planes) the candidate lists of slices in
    # general position are within 30 % of the exact ones on average
    if N == 256:
        assert np.mean(over) < 1.3, np.mean(over)
        print("candidates / samples at box 256:", np.mean(over), np.max(over))


def test_ring_pixel_list_with_gaps(lib):
    """pixel lists whose rows have holes (rL > 0: the E list) split into several runs per row"""
    pix = portapi.pixel_list(64, 2, 30.0, 6.0)
    rng = np.random.default_rng(5)
    for q in synth.random_quats(20, rng):
        _check(lib, pix, 2, 128, q, 5)

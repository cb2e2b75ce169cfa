"""ctypes wrapper of oracle/_port/libthb_oracle.so - the plain-C restatement (oracle/thb_oracle.c).

TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py cpu_baseline).  Built by
`make oracle` / __graft_entry__.build(); compiled on first use if missing (gcc only).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
SRC = HERE / "thb_oracle.c"
LIB = HERE / "_port" / "libthb_oracle.so"

_p, _i, _f = C.c_void_p, C.c_int, C.c_float
_lib = None


class _Cpx(C.Structure):
    _fields_ = [("re", C.c_float), ("im", C.c_float)]


def build(force=False):
    if force or not LIB.exists() or LIB.stat().st_mtime < SRC.stat().st_mtime:
        LIB.parent.mkdir(exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-std=gnu99", "-o", os.fspath(LIB), os.fspath(SRC), "-lm"])
    return LIB


def _ptr(a):
    return None if a is None else a.ctypes.data_as(_p)


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(os.fspath(LIB))
        L.orc_pixel_list.restype = _i
        L.orc_pixel_list.argtypes = [_i, _i, _f, _f] + [_p] * 6
        L.orc_rotate3D.argtypes = [_p, _p]
        L.orc_translate.argtypes = [_p, _p, _f, _f, _i, _p, _p, _i]
        L.orc_ctf.argtypes = [_p] + [_f] * 8 + [_i, _i, _p, _p, _i]
        L.orc_interp_ft.restype = _Cpx
        L.orc_interp_ft.argtypes = [_p, _i, _f, _f, _f]
        L.orc_project.argtypes = [_p, _p, _i, _i, _p, _p, _p, _i]
        L.orc_logDataVSPrior.restype = _f
        L.orc_logDataVSPrior.argtypes = [_p, _p, _p, _p, _i]
        L.orc_logDataVSPrior_m_n.argtypes = [_p, _p, _p, _p, _i, _i, _p]
        L.orc_expect_local.argtypes = [_p, _i, _i, _i, _p, _p, _i, _p, _p, _p, _i, _i] + [_p] * 9
        L.orc_insertP.argtypes = [_p, _p, _i, _p, _p, _p, _f, _p, _p, _i]
        L.orc_insert_loop.argtypes = [_p, _p, _p, _p, _i, _i, _i, _p, _p, _p, _p, _p, _p, _i, _i, _p, _p, _i]
        L.orc_normalise_TF.argtypes = [_p, _p, C.c_size_t]
        _lib = L
    return _lib


def pixel_list(N, pf, rU, rL):
    cap = (N // 2 + 1) * N
    names = ("iCol", "iRow", "iPxl", "iSig", "iColPad", "iRowPad")
    b = {k: np.empty(cap, np.int32) for k in names}
    n = lib().orc_pixel_list(N, pf, rU, rL, *[_ptr(b[k]) for k in names])
    return {k: v[:n].copy() for k, v in b.items()}


def rotate3D(quat):
    q = np.ascontiguousarray(quat, np.float64)
    m = np.empty(9)
    lib().orc_rotate3D(_ptr(q), _ptr(m))
    return m


def translate(tx, ty, N, iCol, iRow, src=None):
    out = np.empty(len(iCol), np.complex64)
    src = None if src is None else np.ascontiguousarray(src, np.complex64)
    lib().orc_translate(_ptr(out), _ptr(src), tx, ty, N, _ptr(iCol), _ptr(iRow), len(iCol))
    return out


def ctf(pixelSize, voltage, dU, dV, theta, Cs, ac, ps, N, iCol, iRow):
    out = np.empty(len(iCol), np.float32)
    lib().orc_ctf(_ptr(out), pixelSize, voltage, dU, dV, theta, Cs, ac, ps, N, N, _ptr(iCol), _ptr(iRow), len(iCol))
    return out


def project(volFT, pf, mat9, iCol, iRow):
    v = np.ascontiguousarray(volFT, np.complex64)
    m = np.ascontiguousarray(mat9, np.float64)
    out = np.empty(len(iCol), np.complex64)
    lib().orc_project(_ptr(out), _ptr(v), v.shape[0], pf, _ptr(m), _ptr(iCol), _ptr(iRow), len(iCol))
    return out


def logDataVSPrior(dat, pri, ctf_, sigRcp):
    dat = np.ascontiguousarray(dat, np.complex64); pri = np.ascontiguousarray(pri, np.complex64)
    ctf_ = np.ascontiguousarray(ctf_, np.float32); sigRcp = np.ascontiguousarray(sigRcp, np.float32)
    return float(lib().orc_logDataVSPrior(_ptr(dat), _ptr(pri), _ptr(ctf_), _ptr(sigRcp), len(ctf_)))


def logDataVSPrior_m_n(datPM, pri, ctfPM, sigPM, n, m):
    out = np.empty(n, np.float32)
    lib().orc_logDataVSPrior_m_n(_ptr(datPM), _ptr(pri), _ptr(ctfPM), _ptr(sigPM), n, m, _ptr(out))
    return out


def expect_local(volFT, pf, N, iCol, iRow, dat, ctf_, sigRcp, quat, tran, wR, wT):
    """one image; returns dict(uR,uT,uC,base,logL)"""
    v = np.ascontiguousarray(volFT, np.complex64)
    dat = np.ascontiguousarray(dat, np.complex64); ctf_ = np.ascontiguousarray(ctf_, np.float32)
    sigRcp = np.ascontiguousarray(sigRcp, np.float32)
    quat = np.ascontiguousarray(quat, np.float64); tran = np.ascontiguousarray(tran, np.float64)
    wR = np.ascontiguousarray(wR, np.float64); wT = np.ascontiguousarray(wT, np.float64)
    nR, nT = quat.shape[0], tran.shape[0]
    uR = np.empty(nR, np.float32); uT = np.empty(nT, np.float32); uC = np.empty(1, np.float32); base = np.empty(1, np.float32)
    logL = np.empty((nR, nT), np.float32)
    lib().orc_expect_local(_ptr(v), v.shape[0], pf, N, _ptr(iCol), _ptr(iRow), len(iCol), _ptr(dat), _ptr(ctf_), _ptr(sigRcp),
                           nR, nT, _ptr(quat), _ptr(tran), _ptr(wR), _ptr(wT), _ptr(uR), _ptr(uT), _ptr(uC), _ptr(base),
                           _ptr(logL))
    return dict(uR=uR, uT=uT, uC=float(uC[0]), base=float(base[0]), logL=logL)


def insert_loop(n, pf, N, dat, ctf_, w, offS, nr, nt, iCol, iRow, F=None, T=None):
    """returns dict(F, T, O, counter); F/T accumulate in place when given"""
    dat = np.ascontiguousarray(dat, np.complex64); ctf_ = np.ascontiguousarray(ctf_, np.float32)
    nImg, P = dat.shape
    nr = np.ascontiguousarray(nr, np.float64); nt = np.ascontiguousarray(nt, np.float64)
    mReco = nr.shape[1]
    w = np.ascontiguousarray(w, np.float32)
    offS = None if offS is None else np.ascontiguousarray(offS, np.float64)
    shape = (n, n, n // 2 + 1)
    if F is None:
        F = np.zeros(shape, np.complex64)
    if T is None:
        T = np.zeros(shape, np.float32)
    O = np.zeros(3); cnt = np.zeros(1, np.int32)
    lib().orc_insert_loop(_ptr(F), _ptr(T), _ptr(O), _ptr(cnt), n, pf, N, _ptr(dat), _ptr(ctf_), _ptr(w), _ptr(offS), _ptr(nr),
                          _ptr(nt), nImg, mReco, _ptr(iCol), _ptr(iRow), P)
    return dict(F=F, T=T, O=O, counter=int(cnt[0]))


def normalise_TF(F, T):
    lib().orc_normalise_TF(_ptr(F), _ptr(T), T.size)

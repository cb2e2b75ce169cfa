"""ctypes wrapper of oracle/_ref/libthunder_ref.so - the REFERENCE's own CPU classes.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs, never by thunder_b200/.  The library is built by
oracle/build_ref.sh from the sources under /root/reference; every number it returns is
computed by reference code (see oracle/ref_harness.cpp).
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB = HERE / "_ref" / "libthunder_ref.so"

_p, _i, _f, _d = C.c_void_p, C.c_int, C.c_float, C.c_double
_lib = None


def available() -> bool:
    return LIB.exists()


def _ptr(a):
    return None if a is None else a.ctypes.data_as(_p)


def lib():
    global _lib
    if _lib is None:
        if not LIB.exists():
            raise RuntimeError(f"{LIB} missing: run oracle/build_ref.sh where /root/reference exists")
        L = C.CDLL(os.fspath(LIB))
        L.ref_init.argtypes = [_i]
        L.ref_set_seed.argtypes = [C.c_ulong]
        L.ref_alloc_precal_idx.restype = _i
        L.ref_alloc_precal_idx.argtypes = [_i, _i, _f, _f] + [_p] * 6
        L.ref_rotate3D.argtypes = [_p, _p]
        L.ref_translate.argtypes = [_p, _f, _f, _i, _p, _p, _i]
        L.ref_translate_src.argtypes = [_p, _p, _f, _f, _i, _p, _p, _i]
        L.ref_ctf.argtypes = [_p] + [_f] * 8 + [_i, _p, _p, _i]
        L.ref_logDataVSPrior.restype = _f
        L.ref_logDataVSPrior.argtypes = [_p, _p, _p, _p, _i, _i]
        L.ref_logDataVSPrior_m_n.argtypes = [_p, _p, _p, _p, _i, _i, _p, _i]
        L.ref_projector_create.restype = _p
        L.ref_projector_create.argtypes = [_i]
        L.ref_projector_destroy.argtypes = [_p]
        L.ref_projector_set_from_real.argtypes = [_p, _p, _i, _i]
        L.ref_projector_set_padded_ft.argtypes = [_p, _p, _i]
        L.ref_projector_padded_dim.restype = _i
        L.ref_projector_padded_dim.argtypes = [_p]
        L.ref_projector_get_padded_ft.argtypes = [_p, _p]
        L.ref_projector_set_max_radius.argtypes = [_p, _i]
        L.ref_projector_project.argtypes = [_p, _p, _p, _p, _p, _i]
        L.ref_reco_create.restype = _p
        L.ref_reco_create.argtypes = [_i, _i, _i, _i]
        L.ref_reco_destroy.argtypes = [_p]
        L.ref_projector_project_image.argtypes = [_p, _i, _p, _p, _p]
        L.ref_projector2d_create.restype = _p
        L.ref_projector2d_create.argtypes = [_i, _p, _i]
        L.ref_projector2d_project.argtypes = [_p, _p, _p, _p, _p, _i]
        L.ref_reco2d_create.restype = _p
        L.ref_reco2d_create.argtypes = [_i, _i, _i, _i]
        L.ref_reco2d_pad_size.restype = _i
        L.ref_reco2d_pad_size.argtypes = [_p]
        L.ref_reco2d_insert_draw.argtypes = [_p, _p, _p, _i, _p, _p, _i, _p, _p, _p, _f]
        L.ref_reco2d_get.argtypes = [_p, _p, _p, _p, _p]
        L.ref_reco2d_set.argtypes = [_p, _p, _p]
        L.ref_reco2d_reconstruct.restype = _i
        L.ref_reco2d_reconstruct.argtypes = [_p, _p, _i, _i, _p, _i, _i]
        L.ref_projector2d_set_from_real.restype = _i
        L.ref_projector2d_set_from_real.argtypes = [_p, _p, _i, _p]
        L.ref_symmetry_elements.restype = _i
        L.ref_symmetry_elements.argtypes = [C.c_char_p, _p, _i]
        L.ref_reco_symmetrize.argtypes = [_p, C.c_char_p, _i]
        L.ref_reco_set_O.argtypes = [_p, _p, _i]
        L.ref_balance_r_2d.argtypes = [_i, _p, _p]
        L.ref_particle_resample_c.restype = _i
        L.ref_particle_resample_c.argtypes = [_i, _p, _p, _p, _i, _p, _p]
        L.ref_sample_vms.argtypes = [_d, _i, _p]
        L.ref_infer_vms.argtypes = [_i, _p, _p, _p]
        L.ref_pdf_vms.restype = _d
        L.ref_pdf_vms.argtypes = [_p, _p, _d]
        L.ref_norm_residual.argtypes = [_p, _i, _i, _f, _f, _p, _p, _p, _p, _f, _p]
        L.ref_recentre_remask.argtypes = [_p, _p, _i, _d, _d, _f, _i]
        L.ref_sigma_accumulate.argtypes = [_p, _i, _i, _i, _p, _p, _p, _p, _p, _p, _f, _p, _i, _p, _p, _p]
        L.ref_reco_set.argtypes = [_p, _p, _p]
        L.ref_reco_reconstruct.restype = _i
        L.ref_reco_reconstruct.argtypes = [_p, _p, _i, _i, _p, _i, _i]
        L.ref_reco_max_radius.restype = _i
        L.ref_reco_max_radius.argtypes = [_p]
        L.ref_reco_reset.argtypes = [_p, _i]
        L.ref_reco_set_precal.argtypes = [_p, _i, _p, _p, _p, _p]
        L.ref_reco_insertP.argtypes = [_p, _p, _p, _p, _f]
        L.ref_reco_insertDir.argtypes = [_p, _d, _d, _d]
        L.ref_reco_pad_size.restype = _i
        L.ref_reco_pad_size.argtypes = [_p]
        L.ref_reco_get.argtypes = [_p, _p, _p, _p, _p]
        L.ref_reco_prepareTF.argtypes = [_p, _i]
        L.ref_particle_create.restype = _p
        L.ref_particle_create.argtypes = [_i, _i, _i, _i, _d, _d]
        L.ref_particle_destroy.argtypes = [_p]
        L.ref_particle_load.argtypes = [_p, _i, _i, _i, _p, _d, _d, _d, _p, _d, _d, _d, _d, _d]
        L.ref_particle_get_counts.argtypes = [_p, _p]
        L.ref_particle_get.argtypes = [_p] * 13
        L.ref_particle_set.argtypes = [_p] * 9
        L.ref_particle_get_scalars.argtypes = [_p, _p]
        L.ref_particle_set_scalars.argtypes = [_p, _p]
        L.ref_particle_set_u.argtypes = [_p, _i, _p, _i]
        for name in ("perturb",):
            getattr(L, "ref_particle_" + name).argtypes = [_p, _d, _i]
        L.ref_particle_resample.argtypes = [_p, _i, _i]
        L.ref_particle_initD.argtypes = [_p, _i, _d]
        for name in ("calVari", "calRank1st", "keepHalfHeightPeak", "setPeakFactor", "shuffle", "balanceWeight"):
            getattr(L, "ref_particle_" + name).argtypes = [_p, _i]
        for name in ("resetPeakFactor", "normW", "calScore"):
            getattr(L, "ref_particle_" + name).argtypes = [_p]
        for name in ("compressR", "compressT", "variR", "variT", "variD"):
            getattr(L, "ref_particle_" + name).restype = _d
            getattr(L, "ref_particle_" + name).argtypes = [_p]
        L.ref_particle_rand.argtypes = [_p] * 5
        L.ref_particle_rank1st.argtypes = [_p] * 5
        L.ref_expectation_local.argtypes = [_p, _i, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _d, _d, _i, _i, _i, _d, _i,
                                            _i, _i, _p, _p]
        L.ref_insert_loop.argtypes = [_p, _p, _i, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i]
        L.ref_expectation_local_trace.argtypes = [_p, _i, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _d, _d, _i, _i, _i, _d, _i,
                                                  _i, _i, _p, _p, _p, _p, _p, _p, _i, _p, _p, _p, _p, _p, _p, _p]
        L.ref_precal_ctf.argtypes = [_f] * 5 + [_i, _f, _p, _p, _i, _p, _p, _p]
        L.ref_expect_ctf.argtypes = [_p] * 5 + [_f] * 4 + [_p] * 6 + [_d, _p, _p, _i, _i, _i, _i, _i, _i] + [_p] * 6
        L.ref_insert_loop_ctf.argtypes = [_p, _i, _p, _p, _p, _p, _p, _p, _p, _f, _p, _p, _i, _i, _i, _i]
        L.ref_scan.argtypes = [_p, _i, _i, _p, _p, _p, _i, _i, _i, _p, _p, _p, _i, _p, _i, _p, _p, _i, _i, _p, _p, _p, _p]
        L.ref_insert_loop_2d.argtypes = [_p, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i]
        L.ref_particle_from_scan.restype = _i
        L.ref_particle_from_scan.argtypes = [_i, _i, _i, _i, _p, _p, _p, _p, _p, _i, _i, _d, _d, _d, _d, C.c_ulonglong, C.c_ulonglong, C.c_ulonglong,
                                             _p, _p, _p, _p, _p]
        L.ref_rng_replay.argtypes = [_i]
        L.ref_rng_key.argtypes = [C.c_ulonglong] * 3
        L.ref_rng_replay_loop.argtypes = [C.c_ulonglong] * 3
        L.ref_rng_draw.argtypes = [_i, _i, _d, _d, _p]
        L.ref_rng_replay_stride.argtypes = [C.c_ulonglong]
        if hasattr(L, "ref_inferACG"):
            L.ref_inferACG.argtypes = [_p, _i, _p, _p, _p]
            L.ref_pdfACG.restype = _d
            L.ref_pdfACG.argtypes = [_p, _p]
        L.ref_init(1)
        _lib = L
    return _lib


# ---------------------------------------------------------------------------------------------
def pixel_list(N, pf, rU, rL):
    cap = (N // 2 + 1) * N
    names = ("iCol", "iRow", "iPxl", "iSig", "iColPad", "iRowPad")
    b = {k: np.empty(cap, np.int32) for k in names}
    n = lib().ref_alloc_precal_idx(N, pf, rU, rL, *[_ptr(b[k]) for k in names])
    return {k: v[:n].copy() for k, v in b.items()}


def rotate3D(quat):
    q = np.ascontiguousarray(quat, np.float64)
    m = np.empty(9, np.float64)
    lib().ref_rotate3D(_ptr(q), _ptr(m))
    return m  # column-major


def translate(tx, ty, N, iCol, iRow, src=None):
    n = len(iCol)
    out = np.empty(n, np.complex64)
    if src is None:
        lib().ref_translate(_ptr(out), tx, ty, N, _ptr(iCol), _ptr(iRow), n)
    else:
        src = np.ascontiguousarray(src, np.complex64)
        lib().ref_translate_src(_ptr(out), _ptr(src), tx, ty, N, _ptr(iCol), _ptr(iRow), n)
    return out


def ctf(pixelSize, voltage, dU, dV, theta, Cs, ac, ps, N, iCol, iRow):
    n = len(iCol)
    out = np.empty(n, np.float32)
    lib().ref_ctf(_ptr(out), pixelSize, voltage, dU, dV, theta, Cs, ac, ps, N, _ptr(iCol), _ptr(iRow), n)
    return out


def sigma_accumulate(P, imgFT, imgOriFT, quat, tran, offS, ctfAttr, pixelSize, group, nGroup, rSig):
    """per-image part of Optimiser::allReduceSigma with the reference's own functions; P: Projector with max radius rSig.
    Returns sigM, sigN, svd [nGroup][rSig+1] float32 (last column: weight sums)."""
    imgFT = np.ascontiguousarray(imgFT, np.complex64); imgOriFT = np.ascontiguousarray(imgOriFT, np.complex64)
    nImg, N = imgFT.shape[0], imgFT.shape[1]
    quat = np.ascontiguousarray(quat, np.float64); tran = np.ascontiguousarray(tran, np.float64); offS = np.ascontiguousarray(offS, np.float64)
    attr = np.ascontiguousarray(ctfAttr, np.float32); group = np.ascontiguousarray(group, np.int32)
    out = [np.zeros((nGroup, rSig + 1), np.float32) for _ in range(3)]
    lib().ref_projector_set_max_radius(P.h, rSig)
    lib().ref_sigma_accumulate(P.h, nImg, N, rSig, _ptr(imgFT), _ptr(imgOriFT), _ptr(quat), _ptr(tran), _ptr(offS), _ptr(attr), float(pixelSize),
                               _ptr(group), nGroup, _ptr(out[0]), _ptr(out[1]), _ptr(out[2]))
    return out


def sample_vms(k, n):
    """sampleVMS about mu = (1, 0): [n][2]"""
    out = np.zeros((n, 2))
    lib().ref_sample_vms(float(k), int(n), _ptr(out))
    return out


def infer_vms(cs):
    cs = np.ascontiguousarray(cs, np.float64)
    mu = np.zeros(2); k = np.zeros(1)
    lib().ref_infer_vms(cs.shape[0], _ptr(cs), _ptr(mu), _ptr(k))
    return mu, float(k[0])


def pdf_vms(x, mu, k):
    x = np.ascontiguousarray(x, np.float64); mu = np.ascontiguousarray(mu, np.float64)
    return float(lib().ref_pdf_vms(_ptr(x), _ptr(mu), float(k)))


def balance_r_2d(cs):
    cs = np.ascontiguousarray(cs, np.float64)
    w = np.zeros(cs.shape[0])
    lib().ref_balance_r_2d(cs.shape[0], _ptr(cs), _ptr(w))
    return w


def particle_resample_c(c, wC, uC, nOut):
    """Particle::resample(nOut, PAR_C) of a MODE_2D reference particle -> (classes, priors, top class)"""
    c = np.ascontiguousarray(c, np.int32); wC = np.ascontiguousarray(wC, np.float64); uC = np.ascontiguousarray(uC, np.float64)
    cOut = np.zeros(nOut, np.int32); wOut = np.zeros(nOut)
    top = lib().ref_particle_resample_c(len(c), _ptr(c), _ptr(wC), _ptr(uC), int(nOut), _ptr(cOut), _ptr(wOut))
    return cOut, wOut, int(top)


def symmetry_elements(name):
    """R matrices [nElem][9] (column-major dmat33) of the reference's Symmetry(name), e.g. C4, D2, T, O, I"""
    R = np.zeros((128, 9))
    n = lib().ref_symmetry_elements(name.encode(), _ptr(R), 128)
    return R[:n].copy()


def norm_residual(P, imgFT, quat, tran, ctfAttr, pixelSize, rL, rNorm):
    """image loop of Optimiser::normCorrection with the reference's functions; P: Projector whose max radius covers rNorm"""
    imgFT = np.ascontiguousarray(imgFT, np.complex64)
    nImg, N = imgFT.shape[0], imgFT.shape[1]
    quat = np.ascontiguousarray(quat, np.float64); tran = np.ascontiguousarray(tran, np.float64)
    attr = np.ascontiguousarray(ctfAttr, np.float32)
    out = np.zeros(nImg, np.float32)
    lib().ref_projector_set_max_radius(P.h, int(np.ceil(rNorm)) + 1)
    lib().ref_norm_residual(P.h, nImg, N, float(rL), float(rNorm), _ptr(imgFT), _ptr(quat), _ptr(tran), _ptr(attr), float(pixelSize), _ptr(out))
    return out


def recentre_remask(imgOriFT, offset, maskRadiusPx, zeroMask=True):
    """Optimiser::reCentreImg + reMaskImg for one image: half-complex [N][N/2+1] complex64 in and out"""
    src = np.ascontiguousarray(imgOriFT, np.complex64)
    N = src.shape[0]
    out = np.empty_like(src)
    lib().ref_recentre_remask(_ptr(out), _ptr(src), N, float(offset[0]), float(offset[1]), float(maskRadiusPx), int(zeroMask))
    return out


def logDataVSPrior(dat, pri, ctf_, sigRcp, variant=1):
    dat = np.ascontiguousarray(dat, np.complex64); pri = np.ascontiguousarray(pri, np.complex64)
    ctf_ = np.ascontiguousarray(ctf_, np.float32); sigRcp = np.ascontiguousarray(sigRcp, np.float32)
    return float(lib().ref_logDataVSPrior(_ptr(dat), _ptr(pri), _ptr(ctf_), _ptr(sigRcp), len(ctf_), variant))


def logDataVSPrior_m_n(datPM, pri, ctfPM, sigPM, n, m, variant=1):
    """pixel-major arrays [m][n]; returns result[n] (the reference accumulates into a caller-zeroed array)"""
    out = np.zeros(n, np.float32)
    lib().ref_logDataVSPrior_m_n(_ptr(datPM), _ptr(pri), _ptr(ctfPM), _ptr(sigPM), n, m, _ptr(out), variant)
    return out


class Projector:
    def __init__(self, pf=2):
        self.h = lib().ref_projector_create(pf)
        self.pf = pf

    def close(self):
        if self.h:
            lib().ref_projector_destroy(self.h)
            self.h = None

    def set_from_real(self, vol, nThread=1):
        vol = np.ascontiguousarray(vol, np.float32)
        lib().ref_projector_set_from_real(self.h, _ptr(vol), vol.shape[0], nThread)

    def set_padded_ft(self, volFT):
        v = np.ascontiguousarray(volFT, np.complex64)
        lib().ref_projector_set_padded_ft(self.h, _ptr(v), v.shape[0])

    def padded_ft(self):
        n = lib().ref_projector_padded_dim(self.h)
        out = np.empty((n, n, n // 2 + 1), np.complex64)
        lib().ref_projector_get_padded_ft(self.h, _ptr(out))
        return out

    def project(self, mat9, iCol, iRow):
        out = np.empty(len(iCol), np.complex64)
        m = np.ascontiguousarray(mat9, np.float64)
        lib().ref_projector_project(self.h, _ptr(out), _ptr(m), _ptr(iCol), _ptr(iRow), len(iCol))
        return out

    def project_image(self, N, quat, tran, maxRadius):
        """Projector::project(Image&, rot, t): half-complex [N][N/2+1], pixels with |k| < maxRadius"""
        out = np.empty((N, N // 2 + 1), np.complex64)
        q = np.ascontiguousarray(quat, np.float64); t = np.ascontiguousarray(tran, np.float64)
        lib().ref_projector_set_max_radius(self.h, int(maxRadius))
        lib().ref_projector_project_image(self.h, N, _ptr(q), _ptr(t), _ptr(out))
        return out


class Reconstructor:
    def __init__(self, size, N, pf=2, nThread=1):
        self.h = lib().ref_reco_create(size, N, pf, nThread)
        self.nThread = nThread

    def close(self):
        if self.h:
            lib().ref_reco_destroy(self.h)
            self.h = None

    def set_precal(self, iColPad, iRowPad, iPxl, iSig):
        lib().ref_reco_set_precal(self.h, len(iColPad), _ptr(iColPad), _ptr(iRowPad), _ptr(iPxl), _ptr(iSig))

    def insertP(self, src, ctf_, mat9, w):
        src = np.ascontiguousarray(src, np.complex64); ctf_ = np.ascontiguousarray(ctf_, np.float32)
        m = np.ascontiguousarray(mat9, np.float64)
        lib().ref_reco_insertP(self.h, _ptr(src), _ptr(ctf_), _ptr(m), w)

    def insertDir(self, o):
        lib().ref_reco_insertDir(self.h, float(o[0]), float(o[1]), float(o[2]))

    def pad_size(self):
        return lib().ref_reco_pad_size(self.h)

    def get(self):
        m = self.pad_size()
        F = np.empty((m, m, m // 2 + 1), np.complex64); T = np.empty((m, m, m // 2 + 1), np.float32)
        O = np.empty(3); cnt = np.zeros(1, np.int32)
        lib().ref_reco_get(self.h, _ptr(F), _ptr(T), _ptr(O), _ptr(cnt))
        return dict(F=F, T=T, O=O, counter=int(cnt[0]))

    def set(self, F, T):
        F = np.ascontiguousarray(F, np.complex64); T = np.ascontiguousarray(T, np.float32)
        lib().ref_reco_set(self.h, _ptr(F), _ptr(T))

    def max_radius(self):
        return lib().ref_reco_max_radius(self.h)

    def symmetrize(self, name, O=None, counter=0):
        """Reconstructor::symmetrizeT / F / O on the current accumulators"""
        if O is not None:
            O = np.ascontiguousarray(O, np.float64)
            lib().ref_reco_set_O(self.h, _ptr(O), int(counter))
        lib().ref_reco_symmetrize(self.h, name.encode(), self.nThread)

    def reconstruct(self, N, gridCorr=True, joinHalf=False, fsc=None, nThread=8):
        """Reconstructor::reconstruct -> real volume [N][N][N] float32, origin at index 0"""
        out = np.empty((N, N, N), np.float32)
        f = None if fsc is None else np.ascontiguousarray(fsc, np.float32)
        n = lib().ref_reco_reconstruct(self.h, _ptr(out), int(gridCorr), int(joinHalf), _ptr(f), 0 if f is None else len(f), nThread)
        assert n == N, (n, N)
        return out

    def prepareTF(self):
        lib().ref_reco_prepareTF(self.h, self.nThread)

    def insert_loop_ctf(self, dat, w, offS, nr, nt, nd, ctfAttr, pixelSize, iCol, iRow, N, nThread=1):
        """the insert loop with cSearch: per-draw defocus factors nd[nImg][mReco], ctfAttr[nImg][7]"""
        dat = np.ascontiguousarray(dat, np.complex64)
        nImg, P = dat.shape
        nr = np.ascontiguousarray(nr, np.float64); nt = np.ascontiguousarray(nt, np.float64); nd = np.ascontiguousarray(nd, np.float64)
        mReco = nr.shape[1]
        w = np.ascontiguousarray(w, np.float32); ctfAttr = np.ascontiguousarray(ctfAttr, np.float32)
        offS = None if offS is None else np.ascontiguousarray(offS, np.float64)
        lib().ref_insert_loop_ctf(self.h, nImg, _ptr(dat), _ptr(w), _ptr(offS), _ptr(nr), _ptr(nt), _ptr(nd), _ptr(ctfAttr), pixelSize,
                                  _ptr(iCol), _ptr(iRow), P, N, mReco, nThread)

    def insert_loop(self, dat, ctf_, w, offS, nr, nt, iCol, iRow, N, nThread=1, pars=None):
        dat = np.ascontiguousarray(dat, np.complex64); ctf_ = np.ascontiguousarray(ctf_, np.float32)
        nImg, P = dat.shape
        w = None if w is None else np.ascontiguousarray(w, np.float32)
        offS = None if offS is None else np.ascontiguousarray(offS, np.float64)
        if pars is None:
            nr = np.ascontiguousarray(nr, np.float64); nt = np.ascontiguousarray(nt, np.float64)
            mReco = nr.shape[1]
            parr = None
        else:
            mReco = int(nr)
            parr = (_p * nImg)(*[p.h for p in pars])
            nr = nt = None
        lib().ref_insert_loop(self.h, parr, nImg, _ptr(dat), _ptr(ctf_), _ptr(w), _ptr(offS), _ptr(nr), _ptr(nt),
                              _ptr(iCol), _ptr(iRow), P, N, mReco, nThread)


class Particle:
    PAR_C, PAR_R, PAR_T, PAR_D = 0, 1, 2, 3

    def __init__(self, nR, nT, transS=2.0, transQ=0.01):
        self.h = lib().ref_particle_create(1, nR, nT, 1, transS, transQ)

    def close(self):
        if self.h:
            lib().ref_particle_destroy(self.h)
            self.h = None

    def load(self, nR, nT, q, k1, k2, k3, t, s0, s1, d=1.0, s=0.0, score=1.0):
        q = np.ascontiguousarray(q, np.float64); t = np.ascontiguousarray(t, np.float64)
        lib().ref_particle_load(self.h, nR, nT, 1, _ptr(q), k1, k2, k3, _ptr(t), s0, s1, d, s, score)

    def counts(self):
        n = np.zeros(4, np.int32)
        lib().ref_particle_get_counts(self.h, _ptr(n))
        return n

    def get(self):
        nC, nR, nT, nD = self.counts()
        r = np.empty((nR, 4)); t = np.empty((nT, 2)); wR = np.empty(nR); wT = np.empty(nT); uR = np.empty(nR); uT = np.empty(nT)
        lib().ref_particle_get(self.h, None, _ptr(r), _ptr(t), None, None, _ptr(wR), _ptr(wT), None, None, _ptr(uR), _ptr(uT), None)
        return dict(r=r, t=t, wR=wR, wT=wT, uR=uR, uT=uT)

    def get_d(self):
        nC, nR, nT, nD = self.counts()
        d = np.empty(nD); wD = np.empty(nD); uD = np.empty(nD)
        lib().ref_particle_get(self.h, None, None, None, _ptr(d), None, None, None, _ptr(wD), None, None, None, _ptr(uD))
        return dict(d=d, wD=wD, uD=uD)

    def set(self, r=None, t=None, wR=None, wT=None):
        f = lambda a: None if a is None else np.ascontiguousarray(a, np.float64)
        r, t, wR, wT = f(r), f(t), f(wR), f(wT)
        lib().ref_particle_set(self.h, None, _ptr(r), _ptr(t), None, None, _ptr(wR), _ptr(wT), None)

    def scalars(self):
        out = np.empty(19)
        lib().ref_particle_get_scalars(self.h, _ptr(out))
        return out

    def set_scalars(self, s):
        s = np.ascontiguousarray(s, np.float64)
        lib().ref_particle_set_scalars(self.h, _ptr(s))

    def set_u(self, which, u):
        u = np.ascontiguousarray(u, np.float64)
        lib().ref_particle_set_u(self.h, which, _ptr(u), len(u))

    def __getattr__(self, name):
        fn = getattr(lib(), "ref_particle_" + name)
        return lambda *a: fn(self.h, *a)


def expectation_local(pars, proj, datP, ctfP, sigRcpP, iCol, iRow, N, mLR, mLT, pfL=2.0, pfS=0.5, minPhase=3, maxPhase=100,
                      noDecreaseLimit=1, decreaseFactor=0.95, fixedPhases=0, simd=1, nThread=1, want_dvp=False):
    datP = np.ascontiguousarray(datP, np.complex64)
    nImg, P = datP.shape
    ctfP = np.ascontiguousarray(ctfP, np.float32); sigRcpP = np.ascontiguousarray(sigRcpP, np.float32)
    parr = (_p * nImg)(*[p.h for p in pars])
    nPhase = np.zeros(nImg, np.int32)
    dvp = np.zeros((nImg, mLR, mLT), np.float32) if want_dvp else None
    lib().ref_expectation_local(parr, nImg, proj.h, _ptr(datP), _ptr(ctfP), _ptr(sigRcpP), _ptr(iCol), _ptr(iRow), P, N, mLR,
                                mLT, pfL, pfS, minPhase, maxPhase, noDecreaseLimit, decreaseFactor, fixedPhases, simd,
                                nThread, _ptr(nPhase), _ptr(dvp))
    return nPhase, dvp


def expectation_local_trace(pars, proj, datP, ctfP, sigRcpP, iCol, iRow, N, mLR, mLT, fixedPhases, pfL=2.0, pfS=0.5, uRIn=None,
                            uTIn=None, simd=1, nThread=1, rIn=None, tIn=None, want_states=False):
    """the phase loop with its marginal weights per phase returned (own) and, optionally, replaced by uRIn / uTIn
    [phase][nImg][mLR / mLT] before setUR / setUT (ref_expectation_local_trace in ref_harness.cpp)"""
    datP = np.ascontiguousarray(datP, np.complex64)
    nImg, P = datP.shape
    ctfP = np.ascontiguousarray(ctfP, np.float32); sigRcpP = np.ascontiguousarray(sigRcpP, np.float32)
    parr = (_p * nImg)(*[p.h for p in pars])
    nPhase = np.zeros(nImg, np.int32)
    uROwn = np.zeros((fixedPhases, nImg, mLR), np.float32); uTOwn = np.zeros((fixedPhases, nImg, mLT), np.float32)
    f = lambda a, shp: None if a is None else np.ascontiguousarray(a, np.float32).reshape(shp)
    uRIn = f(uRIn, uROwn.shape); uTIn = f(uTIn, uTOwn.shape)
    cond = np.zeros((fixedPhases, nImg))
    g = lambda a, shp: None if a is None else np.ascontiguousarray(a, np.float64).reshape(shp)
    shR, shT = (fixedPhases, nImg, mLR, 4), (fixedPhases, nImg, mLT, 2)
    rIn = g(rIn, shR); tIn = g(tIn, shT)
    rPert = np.zeros(shR) if want_states else None; tPert = np.zeros(shT) if want_states else None
    rRes = np.zeros(shR) if want_states else None; tRes = np.zeros(shT) if want_states else None
    lib().ref_expectation_local_trace(parr, nImg, proj.h, _ptr(datP), _ptr(ctfP), _ptr(sigRcpP), _ptr(iCol), _ptr(iRow), P, N, mLR,
                                      mLT, pfL, pfS, 3, 100, 1, 0.95, fixedPhases, simd, nThread, _ptr(nPhase), None, _ptr(uRIn),
                                      _ptr(uTIn), _ptr(uROwn), _ptr(uTOwn), fixedPhases, _ptr(cond), _ptr(rIn), _ptr(tIn), _ptr(rPert),
                                      _ptr(tPert), _ptr(rRes), _ptr(tRes))
    if want_states:
        return uROwn, uTOwn, cond, dict(rPert=rPert, tPert=tPert, rRes=rRes, tRes=tRes)
    return uROwn, uTOwn, cond


def precal_ctf(voltage, dU, dV, theta, Cs, N, pixelSize, iCol, iRow):
    """allocPreCal for the CTF search, one image: frequency[nPxl], defocusP[nPxl], (K1, K2)"""
    n = len(iCol)
    f = np.zeros(n, np.float32); dp = np.zeros(n, np.float32); k = np.zeros(2, np.float32)
    lib().ref_precal_ctf(voltage, dU, dV, theta, Cs, N, pixelSize, _ptr(iCol), _ptr(iRow), n, _ptr(f), _ptr(dp), _ptr(k))
    return f, dp, k


def expect_ctf(proj, dat, sigRcp, defP, freq, K1, K2, phaseShift, ac, quat, tran, dpar, wR, wT, wD, wC, iCol, iRow, N, simd=1):
    """one phase of the likelihood loop with the defocus dimension, one image (ref_expect_ctf in ref_harness.cpp)"""
    dat = np.ascontiguousarray(dat, np.complex64); sigRcp = np.ascontiguousarray(sigRcp, np.float32)
    defP = np.ascontiguousarray(defP, np.float32); freq = np.ascontiguousarray(freq, np.float32)
    quat = np.ascontiguousarray(quat, np.float64); tran = np.ascontiguousarray(tran, np.float64); dpar = np.ascontiguousarray(dpar, np.float64)
    wR = np.ascontiguousarray(wR, np.float64); wT = np.ascontiguousarray(wT, np.float64); wD = np.ascontiguousarray(wD, np.float64)
    nR, nT, nD, P = len(quat), len(tran), len(dpar), len(iCol)
    oC = np.zeros(1, np.float32); oR = np.zeros(nR, np.float32); oT = np.zeros(nT, np.float32); oD = np.zeros(nD, np.float32)
    base = np.zeros(1, np.float32); logL = np.zeros((nR, nT, nD), np.float32)
    lib().ref_expect_ctf(proj.h, _ptr(dat), _ptr(sigRcp), _ptr(defP), _ptr(freq), K1, K2, phaseShift, ac, _ptr(quat), _ptr(tran), _ptr(dpar),
                         _ptr(wR), _ptr(wT), _ptr(wD), float(wC), _ptr(iCol), _ptr(iRow), P, N, nR, nT, nD, simd, _ptr(oC), _ptr(oR), _ptr(oT),
                         _ptr(oD), _ptr(base), _ptr(logL))
    return dict(uC=oC[0], uR=oR, uT=oT, uD=oD, base=base[0], logL=logL)


def scan(projs, mode2D, datP, ctfP, sigRcpP, iCol, iRow, N, rot, tran, pR, pT, simd=1, nThread=1):
    """the initial phase of the global search / 2D classification (ref_scan in ref_harness.cpp); datP / ctfP / sigRcpP image-major
    [nImg][nPxl] (transposed to the reference's pixel-major layout here)"""
    datP = np.ascontiguousarray(datP, np.complex64)
    nImg, P = datP.shape
    dPM = np.ascontiguousarray(datP.T); cPM = np.ascontiguousarray(np.asarray(ctfP, np.float32).T); sPM = np.ascontiguousarray(np.asarray(sigRcpP, np.float32).T)
    rot = np.ascontiguousarray(rot, np.float64); tran = np.ascontiguousarray(tran, np.float64)
    pR = np.ascontiguousarray(pR, np.float64); pT = np.ascontiguousarray(pT, np.float64)
    nK, nR, nT = len(projs), len(rot), len(tran)
    hs = (_p * nK)(*[p.h for p in projs])
    wC = np.zeros((nImg, nK), np.float32); wR = np.zeros((nK, nImg, nR), np.float32); wT = np.zeros((nK, nImg, nT), np.float32)
    base = np.zeros(nImg, np.float32)
    lib().ref_scan(hs, nK, int(mode2D), _ptr(dPM), _ptr(cPM), _ptr(sPM), nImg, P, N, _ptr(iCol), _ptr(iRow), _ptr(rot), nR, _ptr(tran), nT,
                   _ptr(pR), _ptr(pT), simd, nThread, _ptr(wC), _ptr(wR), _ptr(wT), _ptr(base))
    return dict(wC=wC, wR=wR, wT=wT, base=base)


def insert_loop_2d(recos, dat, ctf_, w, offS, nc, nr, nt, iCol, iRow, N, nThread=1):
    dat = np.ascontiguousarray(dat, np.complex64)
    nImg, P = dat.shape
    ctf_ = np.ascontiguousarray(ctf_, np.float32); w = np.ascontiguousarray(w, np.float32)
    nc = np.ascontiguousarray(nc, np.int32); nr = np.ascontiguousarray(nr, np.float64); nt = np.ascontiguousarray(nt, np.float64)
    offS = None if offS is None else np.ascontiguousarray(offS, np.float64)
    hs = (_p * len(recos))(*[r.h for r in recos])
    lib().ref_insert_loop_2d(hs, nImg, _ptr(dat), _ptr(ctf_), _ptr(w), _ptr(offS), _ptr(nc), _ptr(nr), _ptr(nt), _ptr(iCol), _ptr(iRow), P, N,
                             nc.shape[1], nThread)


def particle_from_scan(mode2D, gridR, gridT, wC, wR, wT, mLR, mLT, kFloor, sFloor, key, transS=2.0, transQ=0.01):
    """post-scan logic of Optimiser::expectation on the reference's Particle (ref_particle_from_scan); wR[nK][nR], wT[nK][nT] of ONE image"""
    gridR = np.ascontiguousarray(gridR, np.float64); gridT = np.ascontiguousarray(gridT, np.float64)
    wC = np.ascontiguousarray(wC, np.float32); wR = np.ascontiguousarray(wR, np.float32); wT = np.ascontiguousarray(wT, np.float32)
    nK, nR, nT = len(wC), len(gridR), len(gridT)
    r = np.zeros((mLR, 4)); t = np.zeros((mLT, 2)); oR = np.zeros(mLR); oT = np.zeros(mLT); sc = np.zeros(19)
    cls = lib().ref_particle_from_scan(int(mode2D), nK, nR, nT, _ptr(gridR), _ptr(gridT), _ptr(wC), _ptr(wR), _ptr(wT), mLR, mLT, kFloor, sFloor,
                                       transS, transQ, key[0], key[1], key[2], _ptr(r), _ptr(t), _ptr(oR), _ptr(oT), _ptr(sc))
    return dict(cls=cls, r=r, t=t, wR=oR, wT=oT, scal=sc)


class replay:
    """context manager: the reference's random engine swapped for the Philox bit generator of the CUDA library (ref_harness.cpp)"""

    def __init__(self, seed=0, stream=0, epoch=0):
        self.key = (seed, stream, epoch)

    def __enter__(self):
        lib().ref_rng_replay(1)
        lib().ref_rng_key(*self.key)
        lib().ref_rng_replay_loop(*self.key)
        lib().ref_rng_replay_stride(1)
        return self

    def __exit__(self, *a):
        lib().ref_rng_replay(0)


def rng_key(seed, stream, epoch):
    lib().ref_rng_key(seed, stream, epoch)


def rng_replay_loop(seed, stream, epoch, stride=1):
    """image l of the driver loops draws from the stream (seed, stream + stride * l, epoch)"""
    lib().ref_rng_replay_loop(seed, stream, epoch)
    lib().ref_rng_replay_stride(stride)


def rng_draw(kind, n, a=0.0, b=0.0):
    out = np.zeros(2 * n if kind == 4 else n)
    lib().ref_rng_draw(kind, n, a, b, _ptr(out))
    return out


# ---------------------------------------------------------------------------------------------- MODE_2D
class Projector2D:
    """reference Projector in MODE_2D around a padded class average [pfN][pfN/2+1]"""

    def __init__(self, pf, imgFT):
        v = np.ascontiguousarray(imgFT, np.complex64)
        self.h = lib().ref_projector2d_create(pf, _ptr(v), v.shape[0])

    def close(self):
        if self.h:
            lib().ref_projector_destroy(self.h)
            self.h = None

    def set_from_real(self, img, pf):
        """Projector::setProjectee(Image): pad, grid correction, FFT; returns the padded FT [pf N][pf N / 2 + 1]"""
        img = np.ascontiguousarray(img, np.float32)
        N = img.shape[0]
        out = np.empty((N * pf, N * pf // 2 + 1), np.complex64)
        n = lib().ref_projector2d_set_from_real(self.h, _ptr(img), N, _ptr(out))
        assert n == N * pf
        return out

    def project(self, cs, iCol, iRow):
        out = np.empty(len(iCol), np.complex64)
        cs = np.ascontiguousarray(cs, np.float64)
        iCol = np.ascontiguousarray(iCol, np.int32); iRow = np.ascontiguousarray(iRow, np.int32)
        lib().ref_projector2d_project(self.h, _ptr(out), _ptr(cs), _ptr(iCol), _ptr(iRow), len(iCol))
        return out


class Reconstructor2D:
    def __init__(self, size, N, pf=2):
        self.h = lib().ref_reco2d_create(size, N, pf, 1)
        self.N = N

    def close(self):
        if self.h:
            lib().ref_reco_destroy(self.h)
            self.h = None

    def pad_size(self):
        return lib().ref_reco2d_pad_size(self.h)

    def set_precal(self, iColPad, iRowPad, iPxl, iSig):
        self._keep = [np.ascontiguousarray(a, np.int32) for a in (iColPad, iRowPad, iPxl, iSig)]
        lib().ref_reco_set_precal(self.h, len(self._keep[0]), *[_ptr(a) for a in self._keep])

    def insert_draw(self, dat, ctf, iCol, iRow, cs, tran, off, w):
        dat = np.ascontiguousarray(dat, np.complex64); ctf = np.ascontiguousarray(ctf, np.float32)
        iCol = np.ascontiguousarray(iCol, np.int32); iRow = np.ascontiguousarray(iRow, np.int32)
        cs = np.ascontiguousarray(cs, np.float64); tran = np.ascontiguousarray(tran, np.float64)
        off = np.ascontiguousarray(off, np.float64) if off is not None else None
        lib().ref_reco2d_insert_draw(self.h, _ptr(dat), _ptr(ctf), self.N, _ptr(iCol), _ptr(iRow), len(iCol), _ptr(cs), _ptr(tran),
                                     _ptr(off) if off is not None else None, float(w))

    def set(self, F, T):
        F = np.ascontiguousarray(F, np.complex64); T = np.ascontiguousarray(T, np.float32)
        lib().ref_reco2d_set(self.h, _ptr(F), _ptr(T))

    def prepareTF(self):
        lib().ref_reco_prepareTF(self.h, 1)

    def reconstruct(self, gridCorr=True, joinHalf=False, fsc=None):
        """Reconstructor::reconstruct in MODE_2D -> the N x N class average (origin at index 0)"""
        out = np.empty((self.N, self.N), np.float32)
        f = None if fsc is None else np.ascontiguousarray(fsc, np.float32)
        n = lib().ref_reco2d_reconstruct(self.h, _ptr(out), int(gridCorr), int(joinHalf), _ptr(f), 0 if f is None else len(f), 1)
        assert n == self.N
        return out

    def get(self):
        m = self.pad_size()
        F = np.empty((m, m // 2 + 1), np.complex64); T = np.empty((m, m // 2 + 1), np.float32)
        O = np.zeros(3); cnt = np.zeros(1, np.int32)
        lib().ref_reco2d_get(self.h, _ptr(F), _ptr(T), _ptr(O), _ptr(cnt))
        return dict(F=F, T=T, O=O, counter=int(cnt[0]))

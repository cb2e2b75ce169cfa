"""oracle/sigma_port.py - numpy restatement of the image loops of Optimiser::allReduceSigma (src/Optimiser.cpp:6428-6600) and
Optimiser::normCorrection (:6201-6350), MODE_3D, OPTIMISER_SIGMA_RANK1ST / OPTIMISER_NORM_MASK, no CTF search.

TEST INFRASTRUCTURE ONLY.  Pinned to the reference's own functions (Projector::project(Image&, rot, t), CTF(Image&, ...),
powerSpectrum, NEG_FT / ADD_FT) through oracle/_ref in tests/test_reco_oracle.py (`-m "not gpu"`).

powerSpectrum (src/Functions/Spectrum.cpp:161-190) walks the WHOLE half plane j in [-N/2, N/2), i in [0, N/2]: both
Hermitian mates of the i = 0 column are counted; ring u = rint(hypot(i, j)), pixels with i^2 + j^2 < r^2 and u < r."""
import numpy as np

from . import portapi as port


def _half_plane(N, r):
    jj = np.fft.fftfreq(N, 1.0 / N).astype(np.int64)[:, None] + np.zeros((1, N // 2 + 1), np.int64)
    ii = np.arange(N // 2 + 1, dtype=np.int64)[None, :] + np.zeros((N, 1), np.int64)
    q = ii * ii + jj * jj
    u = np.rint(np.hypot(ii, jj)).astype(np.int64)
    return ii, jj, q, u


def _model(volFT, pf, N, quat, t, ctfAttr, pixelSize, iCol, iRow):
    """CTF x translated slice on the pixels (iCol, iRow)"""
    p = port.project(volFT, pf, port.rotate3D(quat), iCol.astype(np.int32), iRow.astype(np.int32))
    ph = 2 * np.pi * (iCol * np.float32(t[0]) / N + iRow * np.float32(t[1]) / N)
    ctf = port.ctf(pixelSize, *[float(x) for x in ctfAttr], N, iCol.astype(np.int32), iRow.astype(np.int32))
    return p * np.exp(-1j * ph) * ctf


def sigma_accumulate(volFT, pf, imgFT, imgOriFT, quat, tran, offS, ctfAttr, pixelSize, group, nGroup, rSig):
    """-> sigM, sigN, svd [nGroup][rSig + 1] (last column: number of images of the group)"""
    nImg, N = imgFT.shape[0], imgFT.shape[1]
    ii, jj, q, u = _half_plane(N, rSig)
    sel = (q < rSig * rSig) & (u < rSig)
    iCol, iRow, ring = ii[sel], jj[sel], u[sel]
    cnt = np.bincount(ring, minlength=rSig).astype(np.float64)
    out = [np.zeros((nGroup, rSig + 1)) for _ in range(3)]
    ps = lambda v: np.bincount(ring, weights=np.abs(v) ** 2, minlength=rSig) / cnt
    for l in range(nImg):
        mM = _model(volFT, pf, N, quat[l], tran[l], ctfAttr[l], pixelSize, iCol, iRow)
        mN = _model(volFT, pf, N, quat[l], np.asarray(tran[l]) - np.asarray(offS[l]), ctfAttr[l], pixelSize, iCol, iRow)
        d, dOri = imgFT[l][sel], imgOriFT[l][sel]
        g = int(group[l])
        out[0][g, :rSig] += ps(d - mM) / 2
        out[1][g, :rSig] += ps(dOri - mN) / 2
        out[2][g, :rSig] += np.sqrt(ps(mM) / ps(d))
        for o in out:
            o[g, rSig] += 1
    return out


def norm_residual(volFT, pf, imgFT, quat, tran, ctfAttr, pixelSize, rL, rNorm):
    nImg, N = imgFT.shape[0], imgFT.shape[1]
    ii, jj, q, u = _half_plane(N, rNorm)
    sel = (q >= rL * rL) & (q < rNorm * rNorm)
    iCol, iRow = ii[sel], jj[sel]
    out = np.zeros(nImg)
    for l in range(nImg):
        m = _model(volFT, pf, N, quat[l], tran[l], ctfAttr[l], pixelSize, iCol, iRow)
        out[l] = np.sum(np.abs(imgFT[l][sel] - m) ** 2)
    return out

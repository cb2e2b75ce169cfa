"""oracle/port2d.py - numpy restatement of the MODE_2D (2D classification) arithmetic of the hot path.

TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py's CPU baseline): never imported by the product.
Pinned to the reference's own classes through oracle/_ref (tests/test_mode2d.py, `-m "not gpu"`).

Follows
  rotate2D(dmat22&, dvec2)                             src/Geometry/Euler.cpp:125-131
  Projector::project(Complex*, dmat22, iCol, iRow ..)  src/Projector.cpp:337-354      x = R (pf iCol, pf iRow) in double
  Image::getByInterpolationFT / getFTHalf(w, x0)       src/Image/Image.cpp:345-368, 441-493
  conjHalf                                             include/Image/Image.h:94-104   x < 0 -> negate (x, y), conjugate
  W_BI_INTERP_LINEAR                                   include/Functions/Interpolation.h:109-125  w[j][i] = v0[i] v1[j]
  Reconstructor::insertP (2D)                          src/Reconstructor.cpp:708-780  F += src ctf w, T += ctf^2 w
  Image::addFT / addFTHalf(value, w, x0)               src/Image/Image.cpp:370-401, 495-600
  translate(dst, src, tx, ty, ...)                     src/Image/ImageFunctions.cpp:471-492
  M-step driver (MODE_2D)                              src/Optimiser.cpp:7072-7148    translate by -(t - offset), insertDir
"""
import numpy as np

f32 = np.float32


def rotate2d(cs):
    c, s = float(cs[0]), float(cs[1])
    return np.array([[c, -s], [s, c]])


def _cell(cs, a, b):
    """double matvec -> float32 coordinates -> fold, floor, fractional parts, bilinear weights (all float32)"""
    c, s = float(cs[0]), float(cs[1])
    x = (c * a + (-s) * b).astype(f32)
    y = (s * a + c * b).astype(f32)
    conj = ~(x >= 0)
    x = np.where(conj, -x, x); y = np.where(conj, -y, y)
    fx, fy = np.floor(x), np.floor(y)
    xd, yd = (x - fx).astype(f32), (y - fy).astype(f32)
    x0, y0 = fx.astype(np.int64), fy.astype(np.int64)
    vx = [f32(1) - xd, xd]; vy = [f32(1) - yd, yd]
    w = [[(vx[i] * vy[j]).astype(f32) for i in (0, 1)] for j in (0, 1)]          # w[j][i]
    return conj, x0, y0, w


def project2d(imgFT, pf, cs, iCol, iRow):
    """imgFT: padded class average [n][n/2+1] complex64 -> packed slice [nPxl] complex64"""
    n = imgFT.shape[0]
    a = (np.asarray(iCol, np.int64) * pf).astype(np.float64); b = (np.asarray(iRow, np.int64) * pf).astype(np.float64)
    conj, x0, y0, w = _cell(cs, a, b)
    re = np.zeros(len(a), f32); im = np.zeros(len(a), f32)
    for j in (0, 1):
        for i in (0, 1):
            v = imgFT[(y0 + j) % n, x0 + i]                   # negative rows stored at +n
            re = (re + (v.real.astype(f32) * w[j][i]).astype(f32)).astype(f32)
            im = (im + (v.imag.astype(f32) * w[j][i]).astype(f32)).astype(f32)
    return (re + 1j * np.where(conj, -im, im)).astype(np.complex64)


def translate(dat, tx, ty, N, iCol, iRow):
    rc, rr = f32(tx) / f32(N), f32(ty) / f32(N)
    s = (np.asarray(iCol).astype(f32) * rc + np.asarray(iRow).astype(f32) * rr).astype(f32)
    ph = (6.28318530717959 * s.astype(np.float64)).astype(f32)
    return (dat * (np.cos(-ph.astype(np.float64)) + 1j * np.sin(-ph.astype(np.float64)))).astype(np.complex64)


class Reco2D:
    """F2D / T2D accumulators (float64 sums here: parity with the fp32 atomics of the reference is to ~1e-6 relative)"""

    def __init__(self, n):
        self.n = n
        self.F = np.zeros((n, n // 2 + 1), np.complex128); self.T = np.zeros((n, n // 2 + 1), np.float64)
        self.O = np.zeros(3); self.counter = 0

    def insert_draw(self, dat, ctf, N, iCol, iRow, iColPad, iRowPad, cs, tran, off, w):
        t = np.asarray(tran, np.float64) - (np.asarray(off, np.float64) if off is not None else 0.0)
        src = translate(dat, -t[0], -t[1], N, iCol, iRow)
        val = (src * ctf.astype(f32) * f32(w)).astype(np.complex64)
        tv = ((ctf.astype(f32) * ctf.astype(f32)).astype(f32) * f32(w)).astype(f32)
        conj, x0, y0, wt = _cell(cs, np.asarray(iColPad, np.float64), np.asarray(iRowPad, np.float64))
        val = np.where(conj, np.conj(val), val)
        for j in (0, 1):
            for i in (0, 1):
                np.add.at(self.F, ((y0 + j) % self.n, x0 + i), val * wt[j][i])
                np.add.at(self.T, ((y0 + j) % self.n, x0 + i), tv * wt[j][i])
        d = -rotate2d(cs) @ t
        self.O[:2] += d
        self.counter += 1

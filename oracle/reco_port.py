"""numpy restatement of Reconstructor::reconstruct (MODE_3D, default Config.h switches) and Projector::setProjectee.
TEST INFRASTRUCTURE ONLY (see oracle/thb_oracle.c).  Follows, relative to the THUNDER tree:
  reconstruct        src/Reconstructor.cpp:1129-1831  (the branches in force: RECONSTRUCTOR_WIENER_FILTER_FSC without
                     _FREQ_AVG, RECONSTRUCTOR_CHECK_C_MAX, RECONSTRUCTOR_TRILINEAR_KERNEL, no RECONSTRUCTOR_REMOVE_NEG)
  convoluteC         src/Reconstructor.cpp:2595-2660   checkC :2522-2593
  prepareTF (norm.)  src/Reconstructor.cpp:1056-1091
  kernel table       Reconstructor::init :55-90 (TabFunction over MKB_RL_R2, 1e5 bins on [0,1]), TabFunction.cpp:26-45
  MKB_RL / MKB_RL_R2 src/Functions/Functions.cpp:145-213 (FUNCTIONS_MKB_ORDER_0), TIK_RL :236-239
  setProjectee       src/Projector.cpp:123-148, gridCorrection :573-583, VOL_PAD_RL include/Image/ImageFunctions.h:176-192
  FFT conventions    src/FFT.cpp:346-376 (backward scaled by 1/N)
Pinned against oracle/_ref (the reference classes themselves) by tests/test_reco_oracle.py."""
from __future__ import annotations

import numpy as np
from scipy import special

MIN_N_ITER_BALANCE, MAX_N_ITER_BALANCE = 10, 30
DIFF_C_THRES, DIFF_C_DECREASE_THRES, N_DIFF_C_NO_DECREASE = 1e-2, 0.95, 2
WIENER_FACTOR_MIN_R, FSC_BASE_L, FSC_BASE_H = 5, 1e-3, 1 - 1e-3


def mkb_rl_r2(r2, a, alpha):
    """MKB_RL_R2, order 0, float32 arithmetic where the reference uses RFLOAT"""
    r2 = np.asarray(r2, np.float32)
    u2 = (np.float32((2 * np.pi * a) ** 2) * r2).astype(np.float32)
    al2 = np.float32(alpha * alpha)
    inside = u2 <= al2
    v = np.sqrt(np.where(inside, al2 - u2, u2 - al2).astype(np.float32)).astype(np.float32)
    v = np.maximum(v, np.float32(1e-30))
    w = (np.float32((2 * np.pi) ** 1.5 * a ** 3 / special.i0(alpha)) / np.power(v.astype(np.float64), 1.5)).astype(np.float32)
    bes = np.where(inside, special.iv(1.5, v.astype(np.float64)), special.jv(1.5, v.astype(np.float64)))
    return (w * bes.astype(np.float32)).astype(np.float32)


def kernel_table(a, alpha, n=100000):
    s = np.float32(1.0) / np.float32(n)
    x = (np.arange(n + 1, dtype=np.float32) * s).astype(np.float32)
    return mkb_rl_r2(x, a, alpha), s


def _ft_coords(m):
    """signed (k, j, i) coordinates of a half-complex FT grid [m][m][m/2+1]"""
    g = np.fft.fftfreq(m, 1.0 / m).astype(np.int64)          # 0..m/2-1, -m/2..-1
    return g[:, None, None], g[None, :, None], np.arange(m // 2 + 1, dtype=np.int64)[None, None, :]


def _rl_coords(n):
    g = np.fft.fftfreq(n, 1.0 / n).astype(np.int64)
    return g[:, None, None], g[None, :, None], g[None, None, :]


def tik_rl(r):
    """TIK_RL(r) = j0(pi r)^2, spherical Bessel j0 = sin(x)/x"""
    return (np.sinc(np.asarray(r, np.float64)) ** 2).astype(np.float32)


def reconstruct(F, T, N, pf, size=None, a=1.9, alpha=15.0, grid_corr=True, fsc=None, join_half=False, normalise=True):
    """F complex64 [m][m][m/2+1], T float32 (real part of the reference's T volume), m = pf * size.
    Returns (real volume [N][N][N] float32 with the origin at index 0, number of balancing iterations)."""
    size = N if size is None else size
    m = pf * size
    M = pf * N
    F = np.array(F, np.complex64).reshape(m, m, m // 2 + 1)
    T = np.array(T, np.float32).reshape(m, m, m // 2 + 1)
    max_radius = size // 2 - int(np.ceil(a))
    if normalise:                                            # prepareTF: sf = 1 / Re T[0]
        sf = np.float32(1.0) / T.flat[0]
        T = (T * sf).astype(np.float32)
        F = (F * sf).astype(np.complex64)
    k, j, i = _ft_coords(m)
    r2 = i * i + j * j + k * k
    inside = r2 < (max_radius * pf) ** 2
    if fsc is not None:
        fsc = np.asarray(fsc, np.float32)
        sel = (r2 >= (WIENER_FACTOR_MIN_R * pf) ** 2) & inside
        u = np.rint(np.sqrt(r2.astype(np.float64))).astype(np.int64) // pf
        f = np.where(u >= len(fsc), np.float32(0), fsc[np.minimum(u, len(fsc) - 1)]).astype(np.float32)
        f = np.maximum(np.float32(FSC_BASE_L), np.minimum(np.float32(FSC_BASE_H), f))
        if join_half:
            f = np.sqrt(2 * f / (1 + f)).astype(np.float32)
        T = np.where(sel, T / f, T).astype(np.float32)
    W = inside.astype(np.float32)
    T = np.maximum(T, np.float32(1e-25))
    n_iter = 0
    if grid_corr:
        tab, s = kernel_table(a, alpha)
        nf = mkb_rl_r2(np.float32(0), a, alpha)               # MKB_RL(0, a, alpha)
        kk, jj, ii = _rl_coords(m)
        x = ((ii * ii + jj * jj + kk * kk).astype(np.float64) / np.float32(M * M)).astype(np.float32)
        kern = (tab[np.rint((x / s).astype(np.float64)).astype(np.int64)] / nf).astype(np.float32)
        diff_c, n_no_dec = np.float32(3.4028235e38), 0
        for it in range(MAX_N_ITER_BALANCE):
            C = (T * W).astype(np.float32)
            c_rl = np.fft.irfftn(C.astype(np.complex64), s=(m, m, m), axes=(0, 1, 2)).astype(np.float32)      # unnormalised c2r scaled by 1/m^3
            c_rl = (c_rl * kern).astype(np.float32)
            Cft = np.fft.rfftn(c_rl).astype(np.complex64)
            absC = np.abs(Cft).astype(np.float32)
            W = np.where(inside, W / np.maximum(absC, np.float32(1e-6)), W).astype(np.float32)
            prev, diff_c = diff_c, np.float32(np.abs(absC[inside] - 1).max())
            n_iter = it + 1
            n_no_dec = n_no_dec + 1 if diff_c > prev * np.float32(DIFF_C_DECREASE_THRES) else 0
            if diff_c < DIFF_C_THRES or (it >= MIN_N_ITER_BALANCE and n_no_dec == N_DIFF_C_NO_DECREASE):
                break
    else:
        W = np.where(inside, np.float32(1) / np.maximum(np.abs(T), np.float32(1e-6)), W).astype(np.float32)
    pad = np.zeros((M, M, M // 2 + 1), np.complex64)
    idx = np.nonzero(inside)
    kz = k[idx[0], 0, 0] % M
    jy = j[0, idx[1], 0] % M
    pad[kz, jy, idx[2]] = (F * W)[idx]
    rl = np.fft.irfftn(pad, s=(M, M, M), axes=(0, 1, 2)).astype(np.float32)
    g = np.fft.fftfreq(N, 1.0 / N).astype(np.int64)
    sel = g % M
    out = rl[np.ix_(sel, sel, sel)]
    kk, jj, ii = _rl_coords(N)
    r = np.sqrt((ii * ii + jj * jj + kk * kk).astype(np.float64)) / (pf * N)
    out = (out / tik_rl(r)).astype(np.float32)
    return out, n_iter


def set_projectee(vol, pf):
    """real N^3 (origin at index 0) -> padded, grid-corrected half-complex FT (pf N)^3 as Projector::setProjectee:
    note TIK_RL(|x| / (pf * nColRL)) with nColRL ALREADY the padded size (src/Projector.cpp:573-583)"""
    N = vol.shape[0]
    n = pf * N
    pad = np.zeros((n, n, n), np.float32)
    g = np.fft.fftfreq(N, 1.0 / N).astype(np.int64) % n
    pad[np.ix_(g, g, g)] = vol
    kk, jj, ii = _rl_coords(n)
    r = np.sqrt((ii * ii + jj * jj + kk * kk).astype(np.float64)) / (pf * n)
    pad = (pad / tik_rl(r)).astype(np.float32)
    return np.fft.rfftn(pad).astype(np.complex64)


# ------------------------------------------------------------------------------------------------------------------
# Optimiser::reCentreImg (src/Optimiser.cpp:6065-6091) + reMaskImg (:6093-6151, zeroMask)
#   translate(Image&, const Image&, ...)   src/Image/ImageFunctions.cpp:269-284
#   softMask(Image& mask, r, ew)           src/Functions/Mask.cpp:333-350 ; EDGE_WIDTH_RL = 6 (include/Macro.h:99)
EDGE_WIDTH_RL = 6


def soft_mask(N, r, ew=EDGE_WIDTH_RL):
    g = np.fft.fftfreq(N, 1.0 / N)
    u = np.hypot(g[:, None], g[None, :]).astype(np.float32)
    r = np.float32(r); ew = np.float32(ew)
    edge = (0.5 + 0.5 * np.cos(((u - r) / ew).astype(np.float32).astype(np.float64) * np.pi)).astype(np.float32)
    return np.where(u > r + ew, np.float32(0), np.where(u >= r, edge, np.float32(1))).astype(np.float32)


def recentre_remask(img_ori_ft, offset, mask_radius_px, zero_mask=True):
    """half-complex [N][N/2+1] complex64 -> the masked, re-centred image FT (what _img[l] holds after reCentreImg + reMaskImg)"""
    src = np.asarray(img_ori_ft, np.complex64)
    N = src.shape[0]
    j = np.fft.fftfreq(N, 1.0 / N).astype(np.float32)[:, None]
    i = np.arange(N // 2 + 1, dtype=np.float32)[None, :]
    rCol = np.float32(np.float32(offset[0]) / np.float32(N)); rRow = np.float32(np.float32(offset[1]) / np.float32(N))
    s = (i * rCol + j * rRow).astype(np.float32)
    phase = (6.28318530717959 * s.astype(np.float64)).astype(np.float32)
    polar = (np.cos(-phase.astype(np.float64)) + 1j * np.sin(-phase.astype(np.float64))).astype(np.complex64)
    img = (src * polar).astype(np.complex64)
    if not zero_mask:
        return img
    rl = np.fft.irfft2(img, s=(N, N), axes=(0, 1)).astype(np.float32)
    rl = (rl * soft_mask(N, mask_radius_px)).astype(np.float32)
    return np.fft.rfft2(rl).astype(np.complex64)


def symmetrize(F, T, elems, r):
    """Reconstructor::symmetrizeF / symmetrizeT (src/Reconstructor.cpp:2676-2690) = SYMMETRIZE_FT / VOL_TRANSFORM_MAT_FT
    (include/Geometry/Transformation.h:105-131, 170-194): out(v) = src(v) + sum_e src interpolated at R_e v for every voxel of
    the half volume whose rotated position lies inside radius r.  F complex64 [n][n][n/2+1], T float32 (the real part of the
    reference's complex T volume), elems [nElem][9] column-major dmat33 (Symmetry::get), r = maxRadius * pf + 1."""
    f32 = np.float32
    n = F.shape[0]
    nc = n // 2 + 1
    km, jm, i = np.meshgrid(np.arange(n), np.arange(n), np.arange(nc), indexing="ij")
    a = i.astype(np.float64)
    b = np.where(jm < n // 2, jm, jm - n).astype(np.float64)
    c = np.where(km < n // 2, km, km - n).astype(np.float64)
    outF = F.astype(np.complex64).copy()
    outT = T.astype(f32).copy()
    for e in np.asarray(elems, np.float64).reshape(-1, 9):
        ox = e[0] * a + e[3] * b + e[6] * c
        oy = e[1] * a + e[4] * b + e[7] * c
        oz = e[2] * a + e[5] * b + e[8] * c
        ok = ox * ox + oy * oy + oz * oz < r * r
        x, y, z = ox.astype(f32), oy.astype(f32), oz.astype(f32)
        cj = ~(x >= 0)
        x = np.where(cj, -x, x); y = np.where(cj, -y, y); z = np.where(cj, -z, z)
        fx, fy, fz = np.floor(x), np.floor(y), np.floor(z)
        x0, y0, z0 = fx.astype(np.int64), fy.astype(np.int64), fz.astype(np.int64)
        xd, yd, zd = (x - fx).astype(f32), (y - fy).astype(f32), (z - fz).astype(f32)
        accF = np.zeros(F.shape, np.complex64)
        accT = np.zeros(F.shape, f32)
        for dk in (0, 1):
            for dj in (0, 1):
                for di in (0, 1):
                    w = ((xd if di else f32(1) - xd) * (yd if dj else f32(1) - yd)).astype(f32) * (zd if dk else f32(1) - zd)
                    xi = np.clip(x0 + di, 0, nc - 1)                       # only reached by voxels outside r (masked below)
                    accF = (accF + (F[(z0 + dk) % n, (y0 + dj) % n, xi] * w.astype(f32)).astype(np.complex64)).astype(np.complex64)
                    accT = (accT + T[(z0 + dk) % n, (y0 + dj) % n, xi] * w.astype(f32)).astype(f32)
        accF = np.where(cj, np.conj(accF), accF)
        outF = (outF + np.where(ok, accF, 0)).astype(np.complex64)
        outT = (outT + np.where(ok, accT, 0)).astype(f32)
    return outF, outT

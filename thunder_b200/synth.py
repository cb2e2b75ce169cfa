"""Synthetic cryo-EM stacks for the parity tests and the benchmark.

Pure numpy host code: phantom volume, padded/grid-corrected Fourier volume (what
Projector::setProjectee produces, reference src/Projector.cpp:123-148), per-particle CTF
parameters, noisy packed image arrays.  Slice extraction itself is delegated to a callback
(`project_fn`): the caller decides which projector produces the clean slices.
"""
from __future__ import annotations

import numpy as np


def phantom(N: int, n_blobs: int = 30, seed: int = 1234) -> np.ndarray:
    """sum of 3D Gaussians inside radius 0.35 N, real N^3 float32 (z, y, x), centred at the array
    origin (index 0), as the reference's real-space volumes are."""
    rng = np.random.default_rng(seed)
    g = np.fft.fftfreq(N, 1.0 / N).astype(np.float32)  # 0..N/2-1, -N/2..-1
    z, y, x = np.meshgrid(g, g, g, indexing="ij")
    vol = np.zeros((N, N, N), np.float32)
    for _ in range(n_blobs):
        c = rng.normal(size=3)
        c = c / np.linalg.norm(c) * rng.uniform(0, 0.35 * N / 2)
        s = rng.uniform(2.0, 6.0) * N / 256.0 + 1.0
        a = rng.uniform(0.5, 1.5)
        vol += a * np.exp(-((x - c[0]) ** 2 + (y - c[1]) ** 2 + (z - c[2]) ** 2) / (2 * s * s)).astype(np.float32)
    return vol


def padded_ft(vol: np.ndarray, pf: int = 2) -> np.ndarray:
    """real N^3 -> padded, grid-corrected half-complex FT (pf N)^3, complex64 [z][y][x/2+1].
    Same recipe as Projector::setProjectee: zero-pad in real space about the origin, divide by
    sinc^2(|x| / (pf N)) (TIK_RL), forward FFT (unnormalised)."""
    N = vol.shape[0]
    n = pf * N
    pad = np.zeros((n, n, n), np.float32)
    idx = np.fft.fftfreq(N, 1.0 / N).astype(int)  # signed coordinates of the source voxels
    ii = idx % n
    pad[np.ix_(ii, ii, ii)] = vol
    g = np.fft.fftfreq(n, 1.0 / n).astype(np.float32)
    z, y, x = np.meshgrid(g, g, g, indexing="ij")
    r = np.sqrt(x * x + y * y + z * z) / n
    s = np.sinc(r) ** 2
    pad /= np.where(s > 1e-6, s, 1.0).astype(np.float32)
    return np.fft.rfftn(pad).astype(np.complex64)


def random_hermitian_volume(n: int, seed: int = 7) -> np.ndarray:
    """random complex64 half-volume that is the FT of a real volume (Hermitian-consistent x = 0 plane)"""
    rng = np.random.default_rng(seed)
    return np.fft.rfftn(rng.normal(size=(n, n, n)).astype(np.float32)).astype(np.complex64)


def random_quats(n: int, rng) -> np.ndarray:
    q = rng.normal(size=(n, 4))
    return q / np.linalg.norm(q, axis=1, keepdims=True)


def quat_mul(a, b):
    w = a[..., 0] * b[..., 0] - a[..., 1] * b[..., 1] - a[..., 2] * b[..., 2] - a[..., 3] * b[..., 3]
    x = a[..., 0] * b[..., 1] + a[..., 1] * b[..., 0] + a[..., 2] * b[..., 3] - a[..., 3] * b[..., 2]
    y = a[..., 0] * b[..., 2] - a[..., 1] * b[..., 3] + a[..., 2] * b[..., 0] + a[..., 3] * b[..., 1]
    z = a[..., 0] * b[..., 3] + a[..., 1] * b[..., 2] - a[..., 2] * b[..., 1] + a[..., 3] * b[..., 0]
    return np.stack([w, x, y, z], axis=-1)


def acg_cloud(q, k, n, rng):
    """n quaternions ~ ACG(diag(1,k,k,k)) about q (perturbation applied on the left, as Particle::load)"""
    v = rng.normal(size=(n, 4)) * np.sqrt(np.array([1.0, k, k, k]))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    return quat_mul(v, np.broadcast_to(q, (n, 4)))


def ctf_values(iCol, iRow, N, pixelSize, voltage, dU, dV, theta, Cs, ac, phaseShift=0.0):
    """CTF(RFLOAT*, ...) of the reference (src/CTF.cpp:118-151) in numpy float32/float64 mix"""
    lam = np.float32(12.2643247 / np.sqrt(voltage * (1 + voltage * 0.978466e-6)))
    w1 = np.float32(np.sqrt(1 - ac * ac)); w2 = np.float32(ac)
    K1 = np.float32(np.pi * lam); K2 = np.float32(np.pi / 2 * Cs * lam ** 3)
    u = np.hypot(iCol / (pixelSize * N), iRow / (pixelSize * N)).astype(np.float32)
    ang = (np.arctan2(iRow, iCol) - theta).astype(np.float32)
    defocus = -(dU + dV + (dU - dV) * np.cos(2 * ang)) / 2
    ki = K1 * defocus * u ** 2 + K2 * u ** 4 - phaseShift
    return (-w1 * np.sin(ki) + w2 * np.cos(ki)).astype(np.float32)


def make_particles(nImg, N, pix, project_fn, seed=1234, transS=2.0, snr_scale=1.0, pixelSize=1.32):
    """Packed synthetic stack for one pixel list.

    pix: dict with iCol,iRow (int32).  project_fn(quats[n,4]) -> complex64 [n][nPxl] clean slices.
    Returns dict(dat, ctf, sigRcp, quat, tran, ctfpar).  Image = CTF * slice * shift + noise with
    sigma^2 = sig2 (returned) per complex pixel -> sigRcp = -0.5 / sigma^2 (src/Optimiser.cpp:6708)."""
    rng = np.random.default_rng(seed)
    iCol, iRow = pix["iCol"].astype(np.float64), pix["iRow"].astype(np.float64)
    P = len(iCol)
    quat = random_quats(nImg, rng)
    tran = rng.normal(scale=transS, size=(nImg, 2))
    clean = np.asarray(project_fn(quat), np.complex64)
    dU = rng.uniform(1.0e4, 3.0e4, nImg); dV = dU + rng.uniform(0, 500.0, nImg); th = rng.uniform(0, np.pi, nImg)
    ctf = np.empty((nImg, P), np.float32)
    for l in range(nImg):
        ctf[l] = ctf_values(iCol, iRow, N, pixelSize, 3.0e5, dU[l], dV[l], th[l], 2.7e7, 0.1)
    phase = -2 * np.pi * (iCol[None, :] * tran[:, :1] / N + iRow[None, :] * tran[:, 1:] / N)
    shift = np.exp(1j * phase).astype(np.complex64)
    # Noise level chosen so that the per-pixel SNR is ~0.05 * snr_scale.  The SIGNAL keeps the amplitude
    # of the projector volume (the E-step compares dat with ctf * projection, unscaled); the noise is
    # scaled instead, so that the likelihood is consistent with how the data were made.
    spow = float(np.mean(np.abs(clean * ctf) ** 2)) + 1e-30
    sig2 = spow / (0.05 * snr_scale)
    scale = 1.0
    noise = (rng.normal(size=(nImg, P)) + 1j * rng.normal(size=(nImg, P))) * np.sqrt(sig2 / 2)
    dat = (ctf * clean * shift + noise).astype(np.complex64)
    sigRcp = np.full((nImg, P), -0.5 / sig2, np.float32)
    return dict(dat=dat, ctf=ctf, sigRcp=sigRcp, quat=quat, tran=tran, scale=scale, sig2=sig2,
                ctfpar=np.stack([dU, dV, th], axis=1))


def fsc(a: np.ndarray, b: np.ndarray, rmax: int) -> np.ndarray:
    """Fourier shell correlation between two half-complex volumes [z][y][x/2+1] out to shell rmax
    (definition of the reference's FSC(), src/Functions/Spectrum.cpp:302-337)."""
    n = a.shape[0]
    g = np.fft.fftfreq(n, 1.0 / n)
    z, y = np.meshgrid(g, g, indexing="ij")
    x = np.arange(n // 2 + 1)
    r = np.rint(np.sqrt(z[:, :, None] ** 2 + y[:, :, None] ** 2 + x[None, None, :] ** 2)).astype(int)
    m = r < rmax
    num = np.bincount(r[m], weights=(a[m] * np.conj(b[m])).real, minlength=rmax)
    da = np.bincount(r[m], weights=np.abs(a[m]) ** 2, minlength=rmax)
    db = np.bincount(r[m], weights=np.abs(b[m]) ** 2, minlength=rmax)
    return num / np.sqrt(np.maximum(da * db, 1e-60))

"""Generate tests/golden/mode2d_n16.npz from the REFERENCE's own classes in MODE_2D (oracle/_ref), run in the container
where /root/reference exists:   python tests/golden/make_golden_2d.py
The fixture pins oracle/port2d.py (numpy restatement of the 2D project / insert) where the reference library is absent.
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import refapi  # noqa: E402

N, PF = 16, 2
rng = np.random.default_rng(20260102)
n = N * PF
yy, xx = np.mgrid[-n // 2:n // 2, -n // 2:n // 2]
img = np.zeros((n, n))
for _ in range(5):
    cx, cy = rng.uniform(-0.2 * N, 0.2 * N, 2)
    s = rng.uniform(1.5, 3.0)
    img += rng.uniform(0.5, 1.5) * np.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / (2 * s * s))
refFT = np.fft.rfft2(np.fft.ifftshift(img)).astype(np.complex64)        # padded class average [32][17]
pixE = refapi.pixel_list(N, PF, 7.0, 1.0)
pixM = refapi.pixel_list(N, PF, 7.0, 0.0)

phis = np.array([0.0, np.pi / 2, 0.3, -2.2, 3.0, 1e-9])                 # identity, exact fold, generic, tiny angle
cs = np.stack([np.cos(phis), np.sin(phis)], 1)
P2 = refapi.Projector2D(PF, refFT)
slices = np.stack([P2.project(c, pixE["iCol"], pixE["iRow"]) for c in cs])

PM = len(pixM["iCol"])
nImg, mReco = 3, 4
datM = (rng.normal(size=(nImg, PM)) + 1j * rng.normal(size=(nImg, PM))).astype(np.complex64)
ctfM = rng.uniform(-1, 1, (nImg, PM)).astype(np.float32)
ncs = np.stack([np.cos(a := rng.uniform(-np.pi, np.pi, (nImg, mReco))), np.sin(a)], -1)
nt = rng.normal(size=(nImg, mReco, 2)) * 2
off = rng.normal(size=(nImg, 2))
w = np.full(nImg, 1.0 / mReco, np.float32)
R = refapi.Reconstructor2D(N, N, PF)
R.set_precal(pixM["iColPad"], pixM["iRowPad"], pixM["iPxl"], pixM["iSig"])
for l in range(nImg):
    for m in range(mReco):
        R.insert_draw(datM[l], ctfM[l], pixM["iCol"], pixM["iRow"], ncs[l, m], nt[l, m], off[l], w[l])
acc = R.get()
np.savez_compressed(ROOT / "tests" / "golden" / "mode2d_n16.npz", N=N, pf=PF, refFT=refFT, cs=cs, slices=slices,
                    pixE_iCol=pixE["iCol"], pixE_iRow=pixE["iRow"], pixM_iCol=pixM["iCol"], pixM_iRow=pixM["iRow"],
                    pixM_iColPad=pixM["iColPad"], pixM_iRowPad=pixM["iRowPad"], datM=datM, ctfM=ctfM, ncs=ncs, nt=nt, off=off, w=w,
                    F=acc["F"], T=acc["T"], O=acc["O"], counter=acc["counter"])
print("wrote mode2d_n16.npz:", slices.shape, acc["F"].shape, acc["counter"])

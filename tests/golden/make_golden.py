"""Generate tests/golden/hotpath_n16.npz from the REFERENCE's own classes (oracle/_ref), run in the
container where /root/reference exists:   python tests/golden/make_golden.py
The fixture pins the plain-C oracle (oracle/thb_oracle.c) and, on the GPU box, the CUDA path.
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import refapi  # noqa: E402
from thunder_b200 import synth  # noqa: E402

N, PF = 16, 2
rng = np.random.default_rng(20260101)
refapi.lib().ref_set_seed(99)

vol = synth.phantom(N, n_blobs=8, seed=5)
proj = refapi.Projector(PF)
proj.set_from_real(vol)
volFT = proj.padded_ft()                       # Projector::setProjectee output, (32,32,17)
pixE = refapi.pixel_list(N, PF, 7.0, 1.0)      # E pixel set
pixM = refapi.pixel_list(N, PF, 7.0, 0.0)      # M pixel set (rL = 0)
pix_odd = refapi.pixel_list(N, PF, 5.5, 1.5)   # non-integer radii

nRot = 6
quat = synth.random_quats(nRot, rng)
quat[0] = [1, 0, 0, 0]                         # identity: integer coordinates, zero fractional parts
quat[1] = [np.sqrt(0.5), 0, 0, np.sqrt(0.5)]   # 90 deg about z: exact fold at x = 0
mats = np.stack([refapi.rotate3D(q) for q in quat])
slices = np.stack([proj.project(m, pixE["iCol"], pixE["iRow"]) for m in mats])

tran = np.array([[0.0, 0.0], [1.25, -2.5], [-3.0, 0.75]])
tra = np.stack([refapi.translate(t[0], t[1], N, pixE["iCol"], pixE["iRow"]) for t in tran])
ctf = refapi.ctf(1.32, 3e5, 1.5e4, 1.55e4, 0.3, 2.7e7, 0.1, 0.0, N, pixE["iCol"], pixE["iRow"])
P = len(pixE["iCol"])
dat = (ctf * slices[2] * tra[1] + (rng.normal(size=P) + 1j * rng.normal(size=P)) * 0.7).astype(np.complex64)
sigRcp = np.full(P, -0.5 / 0.49, np.float32)
logL = np.zeros((nRot, len(tran)), np.float32)
logL_simd = np.zeros_like(logL)
for r in range(nRot):
    for t in range(len(tran)):
        pri = (tra[t] * slices[r]).astype(np.complex64)
        logL[r, t] = refapi.logDataVSPrior(dat, pri, ctf, sigRcp, 0)
        logL_simd[r, t] = refapi.logDataVSPrior(dat, pri, ctf, sigRcp, 1)

# insert: 2 images x 3 draws through Reconstructor::insertP / insertDir and the reconstructRef loop
PM = len(pixM["iCol"])
ctfM = refapi.ctf(1.32, 3e5, 1.5e4, 1.55e4, 0.3, 2.7e7, 0.1, 0.0, N, pixM["iCol"], pixM["iRow"])
datM = (rng.normal(size=(2, PM)) + 1j * rng.normal(size=(2, PM))).astype(np.complex64)
ctfM2 = np.stack([ctfM, ctfM[::-1].copy()])
nr = synth.random_quats(6, rng).reshape(2, 3, 4)
nr[0, 0] = [1, 0, 0, 0]
nt = rng.normal(scale=2.0, size=(2, 3, 2))
offS = np.array([[0.5, -0.25], [0.0, 0.0]])
w = np.array([1.0 / 3, 0.7 / 3], np.float32)
reco = refapi.Reconstructor(N, N, PF)
reco.set_precal(pixM["iColPad"], pixM["iRowPad"], pixM["iPxl"], pixM["iSig"])
reco.insert_loop(datM, ctfM2, w, offS, nr, nt, pixM["iCol"], pixM["iRow"], N)
acc = reco.get()
reco.prepareTF()
accN = reco.get()

np.savez_compressed(
    ROOT / "tests" / "golden" / "hotpath_n16.npz",
    N=N, pf=PF, vol=vol, volFT=volFT,
    **{"pixE_" + k: v for k, v in pixE.items()}, **{"pixM_" + k: v for k, v in pixM.items()},
    **{"pixO_" + k: v for k, v in pix_odd.items()},
    quat=quat, mats=mats, slices=slices, tran=tran, tra=tra, ctf=ctf, dat=dat, sigRcp=sigRcp,
    logL=logL, logL_simd=logL_simd,
    datM=datM, ctfM=ctfM2, nr=nr, nt=nt, offS=offS, w=w,
    F=acc["F"], T=acc["T"], O=acc["O"], counter=acc["counter"], Fn=accN["F"], Tn=accN["T"],
)
print("wrote hotpath_n16.npz", P, PM, acc["counter"], acc["O"])

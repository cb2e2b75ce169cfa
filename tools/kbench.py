"""Kernel-level timing of the fused E and M kernels at the box-256 configuration (development aid).
usage: python tools/kbench.py [nImg] [N] [k]"""
import sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from thunder_b200 import capi, synth

nImg = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
N = int(sys.argv[2]) if len(sys.argv) > 2 else 256
kconc = float(sys.argv[3]) if len(sys.argv) > 3 else 7.6e-5
pf, nR, nT, mReco = 2, 125, 9, 100
rng = np.random.default_rng(0)
t0 = time.time()
vol = synth.padded_ft(synth.phantom(N, 30), pf)
print("volume", vol.shape, time.time() - t0, flush=True)
rL = np.floor(N * 1.32 / 200)
pixE = capi.pixel_list(N, pf, N // 2 - 1, rL); pixM = capi.pixel_list(N, pf, N // 2 - 1, 0)
P, PM = len(pixE["iCol"]), len(pixM["iCol"])
ctx = capi.Context(0)
ctx.set_expect_pixels(N, pf, pixE["iCol"], pixE["iRow"]); ctx.set_insert_pixels(N, pf, pixM["iColPad"], pixM["iRowPad"])
ctx.set_volume(0, vol); ctx.set_volume(1, vol)
slot = (np.arange(nImg) % 2).astype(np.int32)
par = synth.make_particles(nImg, N, pixE, lambda q: ctx.project(0, q), seed=1)
ctx.upload_stack(capi.STACK_EXPECT, par["dat"], par["ctf"], par["sigRcp"], slot)
datM = (rng.normal(size=(nImg, PM)) + 1j * rng.normal(size=(nImg, PM))).astype(np.complex64)
ctx.upload_stack(capi.STACK_INSERT, datM, rng.uniform(-1, 1, (nImg, PM)).astype(np.float32), slotOfImg=slot)
quat = np.stack([synth.acg_cloud(par["quat"][l], kconc, nR, rng) for l in range(nImg)])
tran = par["tran"][:, None, :] + rng.normal(scale=0.5, size=(nImg, nT, 2))
wR = np.full((nImg, nR), 1.0 / nR); wT = np.full((nImg, nT), 1.0 / nT)
ctx.enable_timing(True)
ctx.set_option("stats", 1)
# KBENCH_IMPLS = comma-separated E kernels to time, "impl[:rpl]" (default: the context's own)
import os
variants = [v for v in os.environ.get("KBENCH_IMPLS", "").split(",") if v] or [None]
for var in variants:
    if var:
        impl, _, rpl = var.partition(":")
        ctx.set_option("expect_impl", int(impl))
        if rpl:
            ctx.set_option("expect_rpl", int(rpl))
    for it in range(4):
        ctx.kernel_ms(capi.KF_EXPECT, reset=True)
        out = ctx.expect_local(quat, tran, wR, wT, want_logL=False)
        ms, n = ctx.kernel_ms(capi.KF_EXPECT, reset=True)
        bytesE = nImg * (P * 16 + nR * P * 64.0)
        print(f"E[{var or 'default'}] k={kconc:g}: {ms:.2f} ms  {nImg / ms * 1e3:.0f} particle-phases/s  alg {bytesE / ms / 1e6:.0f} GB/s  "
              f"{nImg * nR * P / ms / 1e6:.1f} G pixel-rot/s", flush=True)
        if it == 0 and not var:
            print("   staging:", ctx.expect_stats(), flush=True)
if os.environ.get("KBENCH_NO_INSERT"):
    sys.exit(0)
for s in (0, 1):
    ctx.reco_alloc(s, N * pf)
nr = quat[:, rng.integers(0, nR, mReco)]; nt = tran[:, rng.integers(0, nT, mReco)]
w = np.full(nImg, 1.0 / mReco, np.float32)
for it in range(3):
    ctx.kernel_ms(capi.KF_INSERT, reset=True)
    ctx.insert(w, nr, nt)
    ms, n = ctx.kernel_ms(capi.KF_INSERT, reset=True)
    bytesM = nImg * (PM * 12 + mReco * PM * 8 * 12 * 2.0)
    print(f"M: {ms:.2f} ms  {nImg / ms * 1e3:.0f} particles/s  alg {bytesM / ms / 1e6:.0f} GB/s", flush=True)
print("argmax uR hist", np.bincount(np.argmax(out["uR"], 1), minlength=4)[:4])

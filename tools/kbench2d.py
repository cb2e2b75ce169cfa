"""MODE_2D micro-benchmark (BASELINE config 5 shape: box 200, 20 classes, 100 in-plane rotations x 30 translations, r = 99):
the classification scan (every image x every class) and the class-wise insert, on a resident stack of synthetic images.
Prints one line per kernel family; device time by CUDA events (thb_timer)."""
import argparse
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from thunder_b200 import capi  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--box", type=int, default=200)
    ap.add_argument("--classes", type=int, default=20)
    ap.add_argument("--images", type=int, default=1024)
    ap.add_argument("--nr", type=int, default=100)
    ap.add_argument("--nt", type=int, default=30)
    ap.add_argument("--mreco", type=int, default=100)
    ap.add_argument("--impl", type=int, default=3)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    N, pf, k = a.box, 2, a.classes
    rng = np.random.default_rng(1)
    pixE = capi.pixel_list(N, pf, float(N // 2 - 1), float(int(N * 1.32 / 200)))
    pixM = capi.pixel_list(N, pf, float(N // 2 - 1), 0.0)
    PE, PM = len(pixE["iCol"]), len(pixM["iCol"])
    c = capi.Context(0)
    c.set_mode(capi.MODE_2D)
    c.set_option("expect_impl", a.impl)
    c.set_expect_pixels(N, pf, pixE["iCol"], pixE["iRow"])
    c.set_insert_pixels(N, pf, pixM["iColPad"], pixM["iRowPad"])
    n = N * pf
    for s in range(k):
        c.set_volume(s, (rng.normal(size=(n, n // 2 + 1)) + 1j * rng.normal(size=(n, n // 2 + 1))).astype(np.complex64))
        c.reco_alloc(s, n)
    nImg = a.images
    dat = (rng.normal(size=(nImg, PE)) + 1j * rng.normal(size=(nImg, PE))).astype(np.complex64)
    c.upload_stack(capi.STACK_EXPECT, dat, rng.uniform(-1, 1, (nImg, PE)).astype(np.float32), np.full((nImg, PE), -0.5, np.float32))
    c.upload_stack(capi.STACK_INSERT, dat[:, :PM] if PM <= PE else np.zeros((nImg, PM), np.complex64),
                   rng.uniform(-1, 1, (nImg, PM)).astype(np.float32))
    phi = np.linspace(-np.pi, np.pi, a.nr, endpoint=False)
    cs = np.stack([np.cos(phi), np.sin(phi)], 1); t = rng.normal(size=(a.nt, 2)) * 2
    pR = np.full(a.nr, 1.0 / a.nr); pT = np.full(a.nt, 1.0 / a.nt)
    c.expect_scan(0, cs, t, pR, pT)                       # warm-up
    c.enable_timing(True)
    c.kernel_ms(0, reset=True)
    t0 = time.perf_counter()
    for _ in range(a.reps):
        for s in range(k):
            c.expect_scan(s, cs, t, pR, pT)
    wall = (time.perf_counter() - t0) / a.reps
    ms, nl = c.kernel_ms(0, reset=True)
    ms /= a.reps
    samples = nImg * k * a.nr * PE
    print(f"scan   box {N} P {PE} images {nImg} x classes {k} x nR {a.nr} x nT {a.nt}: kernel {ms:.1f} ms ({nl // a.reps} launches), wall {wall * 1e3:.1f} ms"
          f" -> {nImg / (ms / 1e3):.0f} images/s (all classes), {samples / (ms / 1e3) / 1e9:.1f} G pixel-rot/s, "
          f"{samples * a.nt / (ms / 1e3) / 1e12:.2f} T pixel-rot-trans/s")
    # class-wise insert: mReco draws per image, random classes
    nr = np.stack([np.cos(rng.uniform(-np.pi, np.pi, (nImg, a.mreco))), np.zeros((nImg, a.mreco))], -1)
    nr[..., 1] = np.sqrt(1 - nr[..., 0] ** 2)
    nt = rng.normal(size=(nImg, a.mreco, 2)) * 2
    nc = rng.integers(0, k, (nImg, a.mreco)).astype(np.int32)
    w = np.full(nImg, 1.0 / a.mreco, np.float32)
    c.insert_classes(w, nc, nr, nt)
    c.kernel_ms(1, reset=True)
    for _ in range(a.reps):
        c.insert_classes(w, nc, nr, nt)
    ms, nl = c.kernel_ms(1, reset=True)
    ms /= a.reps
    print(f"insert box {N} P {PM} images {nImg} x mReco {a.mreco} into {k} classes: kernel {ms:.1f} ms -> {nImg / (ms / 1e3):.0f} images/s, "
          f"{nImg * a.mreco * PM / (ms / 1e3) / 1e9:.1f} G pixel-draws/s")
    c.close()


if __name__ == "__main__":
    main()

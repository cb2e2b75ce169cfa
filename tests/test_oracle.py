"""Pin the plain-C oracle (oracle/thb_oracle.c): against the committed golden vectors that were
generated from the reference's own classes, and - where oracle/_ref exists - against the reference
library live on fresh random inputs.  CPU only."""
import numpy as np
import pytest

from thunder_b200 import synth


def test_pixel_list_golden(port, golden):
    N, pf = int(golden["N"]), int(golden["pf"])
    for tag, (rU, rL) in {"pixE_": (7.0, 1.0), "pixM_": (7.0, 0.0), "pixO_": (5.5, 1.5)}.items():
        got = port.pixel_list(N, pf, rU, rL)
        for k, v in got.items():
            assert np.array_equal(v, golden[tag + k]), (tag, k)


@pytest.mark.parametrize("N,rU,rL", [(32, 15, 1), (64, 31, 0), (128, 63, 0), (256, 127, 1), (64, 20.5, 2.5)])
def test_pixel_list_vs_reference(port, ref, N, rU, rL):
    a, b = port.pixel_list(N, 2, rU, rL), ref.pixel_list(N, 2, rU, rL)
    for k in a:
        assert np.array_equal(a[k], b[k]), k
    if (N, rU, rL) == (128, 63, 0):
        assert len(a["iCol"]) == 6141          # nPxl(box 128, r=63)
    if (N, rU, rL) == (256, 127, 1):
        assert len(a["iCol"]) == 25134         # nPxl(box 256, r=127)


def test_rotate_translate_ctf_golden(port, golden):
    for q, m in zip(golden["quat"], golden["mats"]):
        assert np.allclose(port.rotate3D(q), m, rtol=0, atol=1e-15)
    iCol, iRow, N = golden["pixE_iCol"], golden["pixE_iRow"], int(golden["N"])
    for t, ref_t in zip(golden["tran"], golden["tra"]):
        assert np.array_equal(port.translate(t[0], t[1], N, iCol, iRow), ref_t)
    c = port.ctf(1.32, 3e5, 1.5e4, 1.55e4, 0.3, 2.7e7, 0.1, 0.0, N, iCol, iRow)
    assert np.allclose(c, golden["ctf"], rtol=0, atol=2e-7)


def test_project_golden_bit_exact(port, golden):
    iCol, iRow = golden["pixE_iCol"], golden["pixE_iRow"]
    for m, s in zip(golden["mats"], golden["slices"]):
        got = port.project(golden["volFT"], int(golden["pf"]), m, iCol, iRow)
        assert np.array_equal(got, s)


def test_logdatavsprior_golden(port, golden):
    for r in range(golden["logL"].shape[0]):
        for t in range(golden["logL"].shape[1]):
            pri = (golden["tra"][t] * golden["slices"][r]).astype(np.complex64)
            v = port.logDataVSPrior(golden["dat"], pri, golden["ctf"], golden["sigRcp"])
            assert v == golden["logL"][r, t]                       # scalar variant: same operation order
            assert abs(v - golden["logL_simd"][r, t]) <= 1e-5 * abs(v)   # AVX variant: 8-lane partial sums


def test_expect_local_composition_golden(port, golden):
    """the fused per-image body reproduces the reference's pieces composed one by one"""
    nR, nT = golden["logL"].shape
    wR = np.full(nR, 1.0 / nR); wT = np.full(nT, 1.0 / nT)
    out = port.expect_local(golden["volFT"], int(golden["pf"]), int(golden["N"]), golden["pixE_iCol"], golden["pixE_iRow"],
                            golden["dat"], golden["ctf"], golden["sigRcp"], golden["quat"], golden["tran"], wR, wT)
    assert np.allclose(out["logL"], golden["logL"], rtol=1e-6)   # numpy forms tra*pri with a different complex-multiply rounding
    w = np.exp(golden["logL"].astype(np.float64) - golden["logL"].max())
    assert np.allclose(out["uR"], (w * wT).sum(1), rtol=2e-6)
    assert np.allclose(out["uT"], (w * wR[:, None]).sum(0), rtol=2e-6)
    assert abs(out["base"] - golden["logL"].max()) <= 1e-6 * abs(out["base"])


def test_insert_golden(port, golden):
    n = int(golden["N"]) * int(golden["pf"])
    out = port.insert_loop(n, int(golden["pf"]), int(golden["N"]), golden["datM"], golden["ctfM"], golden["w"], golden["offS"],
                           golden["nr"], golden["nt"], golden["pixM_iCol"], golden["pixM_iRow"])
    assert out["counter"] == int(golden["counter"])
    assert np.allclose(out["O"], golden["O"], rtol=1e-13, atol=1e-13)
    # same sequential summation order as the single-threaded reference loop
    assert np.allclose(out["T"], golden["T"], rtol=0, atol=1e-7)
    assert np.allclose(out["F"], golden["F"], rtol=0, atol=2e-7)
    port.normalise_TF(out["F"], out["T"])
    assert np.allclose(out["T"], golden["Tn"], rtol=1e-6, atol=1e-9)
    assert np.allclose(out["F"], golden["Fn"], rtol=1e-6, atol=1e-9)


def test_port_vs_reference_random(port, ref):
    """fresh random problem at a different size: project / translate / likelihood / insert agree"""
    rng = np.random.default_rng(3)
    N, pf = 32, 2
    pix = ref.pixel_list(N, pf, 14.0, 2.0)
    vol = synth.random_hermitian_volume(N * pf, seed=11)
    P = ref.Projector(pf)
    P.set_padded_ft(vol)
    quat = synth.random_quats(5, rng)
    for q in quat:
        m = ref.rotate3D(q)
        assert np.allclose(port.rotate3D(q), m, rtol=0, atol=1e-15)
        a = port.project(vol, pf, m, pix["iCol"], pix["iRow"])
        b = P.project(m, pix["iCol"], pix["iRow"])
        assert np.array_equal(a, b)
    Pn = len(pix["iCol"])
    dat = (rng.normal(size=Pn) + 1j * rng.normal(size=Pn)).astype(np.complex64)
    ctf = rng.uniform(-1, 1, Pn).astype(np.float32)
    sig = -0.5 / rng.uniform(0.5, 2, Pn).astype(np.float32)
    pri = P.project(ref.rotate3D(quat[0]), pix["iCol"], pix["iRow"])
    assert port.logDataVSPrior(dat, pri, ctf, sig) == ref.logDataVSPrior(dat, pri, ctf, sig, 0)
    # pixel-major n-image variant
    n = 7
    datPM = (rng.normal(size=(Pn, n)) + 1j * rng.normal(size=(Pn, n))).astype(np.complex64)
    ctfPM = rng.uniform(-1, 1, (Pn, n)).astype(np.float32); sigPM = np.full((Pn, n), -0.5, np.float32)
    a = port.logDataVSPrior_m_n(datPM, pri, ctfPM, sigPM, n, Pn)
    b = ref.logDataVSPrior_m_n(datPM, pri, ctfPM, sigPM, n, Pn, 0)
    assert np.array_equal(a, b)
    # insert
    pixM = ref.pixel_list(N, pf, 14.0, 0.0)
    PM = len(pixM["iCol"])
    datM = (rng.normal(size=(3, PM)) + 1j * rng.normal(size=(3, PM))).astype(np.complex64)
    ctfM = rng.uniform(-1, 1, (3, PM)).astype(np.float32)
    nr = synth.random_quats(12, rng).reshape(3, 4, 4); nt = rng.normal(scale=2, size=(3, 4, 2))
    w = rng.uniform(0.1, 1, 3).astype(np.float32); offS = rng.normal(size=(3, 2))
    reco = ref.Reconstructor(N, N, pf)
    reco.set_precal(pixM["iColPad"], pixM["iRowPad"], pixM["iPxl"], pixM["iSig"])
    reco.insert_loop(datM, ctfM, w, offS, nr, nt, pixM["iCol"], pixM["iRow"], N)
    g = reco.get()
    o = port.insert_loop(N * pf, pf, N, datM, ctfM, w, offS, nr, nt, pixM["iCol"], pixM["iRow"])
    assert o["counter"] == g["counter"] == 12
    assert np.allclose(o["O"], g["O"], rtol=1e-13)
    assert np.allclose(o["F"], g["F"], rtol=0, atol=3e-7 * np.abs(g["F"]).max())
    assert np.allclose(o["T"], g["T"], rtol=0, atol=3e-7 * np.abs(g["T"]).max())
    reco.close(); P.close()

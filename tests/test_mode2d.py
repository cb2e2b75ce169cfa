"""MODE_2D (2D classification, BASELINE config 5): the in-plane twins of the hot path.

CPU (`-m "not gpu"`): oracle/port2d.py (numpy restatement) pinned to the reference's own Projector / Reconstructor in MODE_2D.
GPU (`-m gpu`): thb_project / thb_expect_local / thb_expect_scan / thb_insert(_classes) in MODE_2D against the reference.
"""
import numpy as np
import pytest

from thunder_b200 import capi, synth


def _ref():
    from oracle import refapi
    if not refapi.available():
        pytest.skip("oracle/_ref not present")
    return refapi


def _class_averages(N, pf, k, seed):
    """k padded class-average FTs [pf N][pf N / 2 + 1]: random blobs, zero outside a disc (like a masked average)"""
    rng = np.random.default_rng(seed)
    n = N * pf
    yy, xx = np.mgrid[-n // 2:n // 2, -n // 2:n // 2]
    out = []
    for _ in range(k):
        img = np.zeros((n, n))
        for _b in range(7):
            cx, cy = rng.uniform(-0.25 * N, 0.25 * N, 2); s = rng.uniform(1.5, 5.0)
            img += rng.uniform(0.5, 1.5) * np.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / (2 * s * s))
        out.append(np.fft.rfft2(np.fft.ifftshift(img)).astype(np.complex64))
    return out


def _unit(phi):
    return np.stack([np.cos(phi), np.sin(phi)], axis=-1)


def _setup(N=64, pf=2, k=3, nImg=12, seed=5, rE=28.0, rM=30.0):
    rng = np.random.default_rng(seed)
    refs = _class_averages(N, pf, k, seed)
    pixE = capi.pixel_list(N, pf, rE, 1.0)
    pixM = capi.pixel_list(N, pf, rM, 0.0)
    PE, PM = len(pixE["iCol"]), len(pixM["iCol"])
    datE = (rng.normal(size=(nImg, PE)) + 1j * rng.normal(size=(nImg, PE))).astype(np.complex64) * np.float32(30)
    ctfE = rng.uniform(-1, 1, (nImg, PE)).astype(np.float32)
    sigE = (-0.5 / rng.uniform(50, 200, (nImg, PE))).astype(np.float32)
    datM = (rng.normal(size=(nImg, PM)) + 1j * rng.normal(size=(nImg, PM))).astype(np.complex64)
    ctfM = rng.uniform(-1, 1, (nImg, PM)).astype(np.float32)
    cls = rng.integers(0, k, nImg).astype(np.int32)
    return dict(N=N, pf=pf, k=k, nImg=nImg, rng=rng, refs=refs, pixE=pixE, pixM=pixM, PE=PE, PM=PM, datE=datE, ctfE=ctfE, sigE=sigE,
                datM=datM, ctfM=ctfM, cls=cls)


# ------------------------------------------------------------------------------------------------ CPU: restatement == reference
def test_port2d_project_matches_reference():
    ref = _ref()
    from oracle import port2d
    s = _setup()
    P = ref.Projector2D(s["pf"], s["refs"][0])
    for phi in (0.0, 0.3, 1.7, 3.0, -2.2, np.pi / 2):
        cs = _unit(phi)
        want = P.project(cs, s["pixE"]["iCol"], s["pixE"]["iRow"])
        got = port2d.project2d(s["refs"][0], s["pf"], cs, s["pixE"]["iCol"], s["pixE"]["iRow"])
        assert np.abs(got - want).max() <= 2e-6 * np.abs(want).max()
    P.close()


def test_port2d_insert_matches_reference():
    ref = _ref()
    from oracle import port2d
    s = _setup(nImg=4)
    N, pf, pixM = s["N"], s["pf"], s["pixM"]
    R = ref.Reconstructor2D(N, N, pf)
    R.set_precal(pixM["iColPad"], pixM["iRowPad"], pixM["iPxl"], pixM["iSig"])
    m = R.pad_size()
    assert m == N * pf
    mine = port2d.Reco2D(m)
    rng = s["rng"]
    for l in range(s["nImg"]):
        for _ in range(3):
            cs = _unit(rng.uniform(-np.pi, np.pi)); t = rng.normal(size=2) * 2; off = rng.normal(size=2)
            R.insert_draw(s["datM"][l], s["ctfM"][l], pixM["iCol"], pixM["iRow"], cs, t, off, 0.25)
            mine.insert_draw(s["datM"][l], s["ctfM"][l], N, pixM["iCol"], pixM["iRow"], pixM["iColPad"], pixM["iRowPad"], cs, t, off, 0.25)
    got = R.get()
    assert got["counter"] == mine.counter == 12
    assert np.allclose(got["O"], mine.O, rtol=1e-12, atol=1e-12)
    assert np.abs(got["F"] - mine.F).max() <= 2e-5 * np.abs(mine.F).max()
    assert np.abs(got["T"] - mine.T).max() <= 2e-5 * np.abs(mine.T).max()
    R.close()


def test_port2d_matches_golden_fixture():
    """the committed MODE_2D fixture (tests/golden/make_golden_2d.py: outputs of the reference's own Projector / Reconstructor
    in MODE_2D) pins oracle/port2d.py also where the reference library is not built"""
    from pathlib import Path
    from oracle import port2d
    g = np.load(Path(__file__).resolve().parent / "golden" / "mode2d_n16.npz")
    N, pf = int(g["N"]), int(g["pf"])
    for c, want in zip(g["cs"], g["slices"]):
        got = port2d.project2d(g["refFT"], pf, c, g["pixE_iCol"], g["pixE_iRow"])
        assert np.abs(got - want).max() <= 2e-6 * np.abs(want).max()
    mine = port2d.Reco2D(N * pf)
    nImg, mReco = g["ncs"].shape[:2]
    for l in range(nImg):
        for m in range(mReco):
            mine.insert_draw(g["datM"][l], g["ctfM"][l], N, g["pixM_iCol"], g["pixM_iRow"], g["pixM_iColPad"], g["pixM_iRowPad"],
                             g["ncs"][l, m], g["nt"][l, m], g["off"][l], g["w"][l])
    assert mine.counter == int(g["counter"])
    assert np.allclose(mine.O, g["O"], atol=1e-12)
    assert np.abs(mine.F - g["F"]).max() <= 2e-5 * np.abs(g["F"]).max()
    assert np.abs(mine.T - g["T"]).max() <= 2e-5 * np.abs(g["T"]).max()


# ------------------------------------------------------------------------------------------------ GPU
@pytest.fixture(params=[3, 1], ids=["cell-kernel", "linear-kernel"])
def ctx2d(request):
    """both E kernels serve MODE_2D: the default one (bilinear cell = one 256-bit load) and the linear-layout one"""
    c = capi.Context(0)
    c.set_mode(capi.MODE_2D)
    c.set_option("expect_impl", request.param)
    yield c
    c.close()


def _load(c, s, slotOfImg=None):
    c.set_expect_pixels(s["N"], s["pf"], s["pixE"]["iCol"], s["pixE"]["iRow"])
    c.set_insert_pixels(s["N"], s["pf"], s["pixM"]["iColPad"], s["pixM"]["iRowPad"])
    for k, r in enumerate(s["refs"]):
        c.set_volume(k, r)
    c.upload_stack(capi.STACK_EXPECT, s["datE"], s["ctfE"], s["sigE"], slotOfImg)
    c.upload_stack(capi.STACK_INSERT, s["datM"], s["ctfM"], slotOfImg=slotOfImg)


@pytest.mark.gpu
def test_project2d_matches_reference(ctx2d):
    ref = _ref()
    s = _setup()
    _load(ctx2d, s)
    phis = np.array([0.0, 0.3, 1.7, 3.0, -2.2, np.pi / 2, 1e-9, -1e-9])
    for k in range(s["k"]):
        assert np.array_equal(ctx2d.get_volume(k), s["refs"][k])
        got = ctx2d.project(k, _unit(phis))
        P = ref.Projector2D(s["pf"], s["refs"][k])
        for i, phi in enumerate(phis):
            want = P.project(_unit(phi), s["pixE"]["iCol"], s["pixE"]["iRow"])
            assert np.abs(got[i] - want).max() <= 2e-6 * np.abs(want).max()
        P.close()


def _ref_logL(ref, s, P, l, cs, t):
    """Optimiser::expectation's inner evaluation with the reference's functions: project, translate, logDataVSPrior"""
    from oracle import port2d
    p = P.project(cs, s["pixE"]["iCol"], s["pixE"]["iRow"])
    p = port2d.translate(p, t[0], t[1], s["N"], s["pixE"]["iCol"], s["pixE"]["iRow"])
    d = s["datE"][l].astype(np.complex128) - s["ctfE"][l].astype(np.float64) * p.astype(np.complex128)
    return float(np.sum((d.real ** 2 + d.imag ** 2) * s["sigE"][l].astype(np.float64)))


@pytest.mark.gpu
def test_expect_local_2d_matches_reference(ctx2d):
    ref = _ref()
    s = _setup()
    _load(ctx2d, s, s["cls"])
    rng = s["rng"]
    nImg, nR, nT = s["nImg"], 9, 5
    cs = _unit(rng.uniform(-np.pi, np.pi, (nImg, nR))); t = rng.normal(size=(nImg, nT, 2)) * 1.5
    wR = rng.uniform(0.5, 1.5, (nImg, nR)); wT = rng.uniform(0.5, 1.5, (nImg, nT))
    out = ctx2d.expect_local(cs, t, wR, wT)
    Ps = [ref.Projector2D(s["pf"], r) for r in s["refs"]]
    for l in range(nImg):
        want = np.array([[_ref_logL(ref, s, Ps[s["cls"][l]], l, cs[l, r], t[l, tt]) for tt in range(nT)] for r in range(nR)])
        assert np.abs(out["logL"][l] - want).max() <= 2e-5 * np.abs(want).max() + 1e-4
        e = np.exp(out["logL"][l].astype(np.float64) - out["base"][l])
        assert np.allclose(out["uR"][l], e @ wT[l], rtol=1e-4)
        assert np.allclose(out["uT"][l], wR[l] @ e, rtol=1e-4)
    for P in Ps:
        P.close()


@pytest.mark.gpu
@pytest.mark.parametrize("nT", [6, 12])            # 12 > 9: the 15-translations-per-pass variant of the default kernel
def test_expect_scan_2d_every_image_against_every_class(ctx2d, nT):
    """the classification scan (src/Optimiser.cpp:756-914): shared in-plane rotations x translations, all images x all classes"""
    ref = _ref()
    s = _setup(nImg=7)
    _load(ctx2d, s, s["cls"])
    rng = s["rng"]
    nR = 20
    cs = _unit(np.linspace(-np.pi, np.pi, nR, endpoint=False)); t = rng.normal(size=(nT, 2)) * 2
    pR = np.full(nR, 1.0 / nR); pT = np.full(nT, 1.0 / nT)
    for k in range(s["k"]):
        out = ctx2d.expect_scan(k, cs, t, pR, pT, want_logL=True)
        P = ref.Projector2D(s["pf"], s["refs"][k])
        for l in (0, 3, 6):
            want = np.array([[_ref_logL(ref, s, P, l, cs[r], t[tt]) for tt in range(nT)] for r in range(nR)])
            assert np.abs(out["logL"][l] - want).max() <= 2e-5 * np.abs(want).max() + 1e-4
            assert out["base"][l] == out["logL"][l].max()
        P.close()


@pytest.mark.gpu
@pytest.mark.parametrize("nT", [6, 12])
def test_expect_scan_all_classes_in_one_launch(ctx2d, nT):
    """thb_expect_scan_classes (ExpectGlobal2D's shape: all classes in one table, one baseline per image across the classes) equals the
    class-by-class scans brought to the common baseline, for the whole stack and for an image sub-range"""
    s = _setup(nImg=7)
    _load(ctx2d, s, s["cls"])
    rng = s["rng"]
    nR, k = 20, s["k"]
    cs = _unit(np.linspace(-np.pi, np.pi, nR, endpoint=False)); t = rng.normal(size=(nT, 2)) * 2
    pR = rng.uniform(0.5, 1.5, nR); pR /= pR.sum()
    pT = rng.uniform(0.5, 1.5, nT); pT /= pT.sum()
    res = [ctx2d.expect_scan(c, cs, t, pR, pT) for c in range(k)]
    bmax = np.max([r["base"] for r in res], axis=0)
    for rng_ in (None, (2, 4)):
        out = ctx2d.expect_scan_classes(k, cs, t, pR, pT, img_range=rng_)
        sl = slice(None) if rng_ is None else slice(rng_[0], rng_[0] + rng_[1])
        assert np.abs(out["base"] - bmax[sl]).max() <= 2e-6 * np.abs(bmax).max() + 1e-4
        # weights are exp(logL - baseline): the log-likelihood tolerance of the 2D tests, 2e-5 |logL| + 1e-4 (the two paths sum the
        # pixels in different orders; this fixture's |logL| is 1e4), is a RELATIVE tolerance of that size on every weight.  Compared
        # where a weight matters for its image, i.e. against the image's largest weight over ALL classes - below that fp32 exp()
        # underflows against the common baseline, here as in the reference
        rtol = 2 * (2e-5 * np.abs(bmax).max() + 1e-4) + 1e-3
        f = [np.exp((res[c]["base"] - bmax).astype(np.float64))[sl] for c in range(k)]
        for key in ("wR", "wT", "wC"):
            want = [(res[c][key][sl] * (f[c][:, None] if key != "wC" else f[c])).astype(np.float64) for c in range(k)]
            got = [out[key][c] if key != "wC" else out["wC"][:, c] for c in range(k)]
            top = np.max([w_.reshape(len(w_), -1).max(1) for w_ in want], axis=0)           # per image, over the classes
            worst = 0.0
            for c in range(k):
                t_ = top[:, None] if key != "wC" else top
                big = want[c] > 1e-4 * t_
                worst = max(worst, float(np.max(np.abs(got[c][big] / want[c][big] - 1.0))) if big.any() else 0.0)
                assert np.allclose(got[c][big], want[c][big], rtol=rtol, atol=0.0), (key, c)
                assert np.all(got[c][~big] <= 2e-4 * np.broadcast_to(t_, want[c].shape)[~big] + 1e-30), (key, c)
            print(f"scan_classes vs class-by-class, {key}: worst relative difference of the weights that matter {worst:.2e} (allowed {rtol:.2e})")
    # ... and THE REFERENCE'S OWN scan loop over all classes (oracle/_ref: ref_scan restates src/Optimiser.cpp:756-914 around the MODE_2D
    # Projector::project + translate + logDataVSPrior_m_n, one baseline per image across the classes): same output layouts
    from oracle import refapi
    if refapi.available():
        projs = [refapi.Projector2D(s["pf"], s["refs"][c]) for c in range(k)]
        want = refapi.scan(projs, True, s["datE"], s["ctfE"], s["sigE"], s["pixE"]["iCol"], s["pixE"]["iRow"], s["N"], cs, t, pR, pT, nThread=4)
        for p_ in projs:
            p_.close()
        out = ctx2d.expect_scan_classes(k, cs, t, pR, pT)
        assert np.abs(out["base"] - want["base"]).max() <= 2e-5 * np.abs(want["base"]).max() + 1e-4
        rtol = 2 * (2e-5 * np.abs(want["base"]).max() + 1e-4) + 1e-3
        for key in ("wR", "wT", "wC"):
            w_, g_ = want[key].astype(np.float64), out[key].astype(np.float64)
            if key == "wC":
                top = w_.max(1, keepdims=True)                        # [nImg][nK]
            else:
                top = w_.max(axis=(0, 2), keepdims=True)              # [nK][nImg][n]: per image over classes and samples
            big = w_ > 1e-4 * top
            assert np.allclose(g_[big], w_[big], rtol=rtol, atol=0.0), key
            assert np.all(g_[~big] <= 2e-4 * np.broadcast_to(top, w_.shape)[~big] + 1e-30), key


@pytest.mark.gpu
@pytest.mark.parametrize("per_draw_classes", [False, True])
def test_insert_2d_matches_reference(ctx2d, per_draw_classes):
    ref = _ref()
    s = _setup(nImg=6)
    N, pf, pixM, k = s["N"], s["pf"], s["pixM"], s["k"]
    _load(ctx2d, s, s["cls"])
    for c in range(k):
        ctx2d.reco_alloc(c, N * pf)
    rng = s["rng"]
    nImg, mReco = s["nImg"], 5
    cs = _unit(rng.uniform(-np.pi, np.pi, (nImg, mReco)))
    cs[:, 3] = cs[:, 1]                                          # duplicate rotations: the merged-draw path
    t = rng.normal(size=(nImg, mReco, 2)) * 2; off = rng.normal(size=(nImg, 2))
    w = np.full(nImg, 1.0 / mReco, np.float32)
    nc = rng.integers(0, k, (nImg, mReco)).astype(np.int32) if per_draw_classes else np.repeat(s["cls"][:, None], mReco, 1)
    if per_draw_classes:
        ctx2d.insert_classes(w, nc, cs, t, offS=off)
    else:
        ctx2d.insert(w, cs, t, offS=off)
    Rs = []
    for c in range(k):
        R = ref.Reconstructor2D(N, N, pf)
        R.set_precal(pixM["iColPad"], pixM["iRowPad"], pixM["iPxl"], pixM["iSig"])
        Rs.append(R)
    for l in range(nImg):
        for m in range(mReco):
            Rs[nc[l, m]].insert_draw(s["datM"][l], s["ctfM"][l], pixM["iCol"], pixM["iRow"], cs[l, m], t[l, m], off[l], w[l])
    for c in range(k):
        want = Rs[c].get(); got = ctx2d.reco_download(c)
        assert got["F"].shape == (N * pf, N * pf // 2 + 1)
        assert got["counter"] == want["counter"]
        assert np.allclose(got["O"], want["O"], rtol=1e-10, atol=1e-10)
        if want["counter"]:
            assert np.abs(got["F"] - want["F"]).max() <= 2e-5 * np.abs(want["F"]).max()
            assert np.abs(got["T"] - want["T"]).max() <= 2e-5 * np.abs(want["T"]).max()
        Rs[c].close()


@pytest.mark.gpu
def test_reconstruct_and_set_projectee_2d(ctx2d):
    """class averages on the device: thb_reconstruct / thb_set_projectee in MODE_2D == Reconstructor::reconstruct (MODE_2D
    branches) / Projector::setProjectee(Image) on the same accumulators"""
    ref = _ref()
    s = _setup(nImg=160, seed=21)
    N, pf, rng = s["N"], s["pf"], s["rng"]
    _load(ctx2d, s)
    ctx2d.reco_alloc(0, N * pf)
    mReco = 6
    cs = _unit(rng.uniform(-np.pi, np.pi, (s["nImg"], mReco))); t = rng.normal(size=(s["nImg"], mReco, 2))
    ctx2d.insert(np.full(s["nImg"], 1.0 / mReco, np.float32), cs, t)
    acc = ctx2d.reco_download(0)
    R = ref.Reconstructor2D(N, N, pf)
    fsc = np.linspace(0.99, 0.2, N // 2 + 1).astype(np.float32)
    for gridCorr, f, joinHalf in ((True, None, False), (False, None, False), (True, fsc, True)):
        ctx2d.reco_upload(0, acc["F"], acc["T"])
        got, nit = ctx2d.reconstruct(0, N, pf, gridCorr=gridCorr, joinHalf=joinHalf, fsc=f)
        assert got.shape == (N, N) and (nit > 0) == gridCorr
        R.set(acc["F"], acc["T"])
        R.prepareTF()
        want = R.reconstruct(gridCorr=gridCorr, joinHalf=joinHalf, fsc=f)
        assert np.linalg.norm(got - want) <= 2e-5 * np.linalg.norm(want), (gridCorr, joinHalf)
    # the projector reference rebuilt from the class average kept on the device, and from an explicit image
    ctx2d.set_projectee(1, None, N, pf)
    P = ref.Projector2D(pf, np.zeros((N * pf, N * pf // 2 + 1), np.complex64))
    want_ft = P.set_from_real(got, pf)
    assert np.linalg.norm(ctx2d.get_volume(1) - want_ft) <= 5e-6 * np.linalg.norm(want_ft)
    ctx2d.set_projectee(2, want, N, pf)
    assert np.linalg.norm(ctx2d.get_volume(2) - P.set_from_real(want, pf)) <= 5e-6 * np.linalg.norm(want_ft)
    sl = ctx2d.project(1, _unit(np.array([0.4])))
    assert np.abs(sl[0] - P.project(_unit(0.4), s["pixE"]["iCol"], s["pixE"]["iRow"])).max() <= 1e-5 * np.abs(sl).max()
    P.close(); R.close()


@pytest.mark.gpu
def test_mode_switch_and_guards(ctx2d):
    s = _setup(nImg=2)
    _load(ctx2d, s)
    with pytest.raises(capi.ThbError):
        ctx2d.pf_load(capi.PFParams(mLR=9, mLT=9, transS=2.0, transQ=0.01, perturbFactorL=2.0, perturbFactorS=0.5, minPhase=3, maxPhase=10,
                                    fixedPhases=0, decreaseFactor=0.95, noDecreaseLimit=1, seed=1),
                      np.tile([1.0, 0, 0, 0], (2, 1)), np.full((2, 3), 1e-4), np.zeros((2, 2)), np.ones((2, 2)))
    ctx2d.set_mode(capi.MODE_3D)                                  # drops the 2D references
    with pytest.raises(capi.ThbError):
        ctx2d.project(0, np.array([[1.0, 0, 0, 0]]))


@pytest.mark.gpu
def test_classification_iteration_on_the_device_scan_handover_phases_insert():
    """a whole 2D classification iteration without the host in the loop of the particles: scan of every image against every class
    (thb_expect_scan) -> choice of the class and support of the local phases from the scan's weights (thb_pf_from_scan: the logic
    of src/Optimiser.cpp:921-1075, pinned to the reference's Particle draw by draw in tests/test_pf_host.py) -> local phases with
    the von Mises operators on the device (thb_expectation) -> class-wise insert (thb_reconstruct_insert).  Synthetic images of
    three distinct class averages at per-pixel SNR 0.5: the class is recovered for nearly all images, the in-plane angle to a
    few degrees, and every draw lands in the accumulator of its image's class."""
    N, pf, k, n = 64, 2, 3, 96
    rng = np.random.default_rng(11)
    refs = _class_averages(N, pf, k, 21)
    pixE = capi.pixel_list(N, pf, 28.0, 1.0); pixM = capi.pixel_list(N, pf, 30.0, 0.0)
    PE, PM = len(pixE["iCol"]), len(pixM["iCol"])
    c = capi.Context(0)
    try:
        c.set_mode(capi.MODE_2D)
        c.set_expect_pixels(N, pf, pixE["iCol"], pixE["iRow"])
        c.set_insert_pixels(N, pf, pixM["iColPad"], pixM["iRowPad"])
        for s_, r_ in enumerate(refs):
            c.set_volume(s_, r_)
            c.reco_alloc(s_, N * pf)
        cls_true = rng.integers(0, k, n); phi = rng.uniform(-np.pi, np.pi, n); tran = rng.normal(scale=1.5, size=(n, 2))

        def simulate(pix):
            c.set_expect_pixels(N, pf, pix["iCol"], pix["iRow"])
            clean = np.empty((n, len(pix["iCol"])), np.complex64)
            for s_ in range(k):
                sel = np.nonzero(cls_true == s_)[0]
                clean[sel] = c.project(s_, _unit(phi[sel]))
            ctf = np.stack([synth.ctf_values(pix["iCol"].astype(float), pix["iRow"].astype(float), N, 1.32, 3e5, 2e4 + 100 * l, 2.02e4 + 100 * l, 0.3,
                                             2.7e7, 0.1) for l in range(n)]).astype(np.float32)
            ph = -2 * np.pi * (pix["iCol"][None] * tran[:, :1] / N + pix["iRow"][None] * tran[:, 1:] / N)
            sig2 = float(np.mean(np.abs(clean * ctf) ** 2)) / 0.5
            noise = (rng.normal(size=clean.shape) + 1j * rng.normal(size=clean.shape)) * np.sqrt(sig2 / 2)
            return (ctf * clean * np.exp(1j * ph) + noise).astype(np.complex64), ctf, sig2
        datM, ctfM, _ = simulate(pixM)
        datE, ctfE, sig2 = simulate(pixE)
        c.set_expect_pixels(N, pf, pixE["iCol"], pixE["iRow"])
        c.upload_stack(capi.STACK_EXPECT, datE, ctfE, np.full((n, PE), -0.5 / sig2, np.float32))
        c.upload_stack(capi.STACK_INSERT, datM, ctfM)
        nR, nT = 100, 30
        ang = np.linspace(-np.pi, np.pi, nR, endpoint=False); cs = _unit(ang)
        trans = rng.normal(scale=1.5, size=(nT, 2)); pR = np.full(nR, 1.0 / nR); pT = np.full(nT, 1.0 / nT)
        res = [c.expect_scan(s_, cs, trans, pR, pT) for s_ in range(k)]
        base = np.max([r_["base"] for r_ in res], axis=0)
        wC = np.stack([r_["wC"] * np.exp(r_["base"] - base) for r_ in res], 1)            # one baseline per image, as ExpectGlobal2D
        wR = np.stack([r_["wR"] for r_ in res]); wT = np.stack([r_["wT"] for r_ in res])
        prm = capi.PFParams(mLR=9, mLT=9, transS=2.0, transQ=0.01, perturbFactorL=0.5, perturbFactorS=0.5, minPhase=3, maxPhase=100,
                            fixedPhases=5, decreaseFactor=0.95, noDecreaseLimit=1, seed=5)
        c.pf_set_image_base(0, 0)
        cls = c.pf_from_scan(prm, cs, trans, wC, wR, wT, kFloor=1.0 / nR / 0.5, sFloor=0.1)
        assert (cls == cls_true).mean() >= 0.9, (cls == cls_true).mean()
        st = c.pf_get()
        assert np.allclose(np.linalg.norm(st["r"][..., :2], axis=2), 1.0, atol=1e-12) and not st["r"][..., 2:].any()
        # every support point is one of the scan's grid points
        d = np.abs(st["r"][:, :, None, :2] - cs[None, None]).sum(-1).min(-1)
        assert d.max() <= 1e-12
        ok = cls == cls_true
        sc0 = st["scal"]
        err0 = np.degrees(np.abs(np.angle(np.exp(1j * (np.arctan2(sc0[:, 7], sc0[:, 6]) - phi)))))
        print(f"\nafter the hand-over: angle error of the scan's best grid point median {np.median(err0[ok]):.2f} deg, k1 median {np.median(sc0[:, 0]):.4f}, "
              f"distinct support points median {np.median([len(np.unique(np.round(r_[:, :2], 9), axis=0)) for r_ in st['r']]):.0f}")
        c.expectation()
        sc = c.pf_get_scal()
        st2 = c.pf_get()
        print(f"after the phases: k1 median {np.median(sc[:, 0]):.5f}, NaN supports {int(np.isnan(st2['r']).any((1, 2)).sum())}")
        err = np.degrees(np.abs(np.angle(np.exp(1j * (np.arctan2(sc[:, 7], sc[:, 6]) - phi)))))
        errT = np.linalg.norm(sc[:, 10:12] - tran, axis=1)
        print(f"\n2D iteration on the device: class recovered for {ok.mean() * 100:.0f} % of {n} images; in-plane angle error median "
              f"{np.median(err[ok]):.2f} deg (grid step {360 / nR:.1f} deg), translation error median {np.median(errT[ok]):.2f} px")
        assert np.median(err[ok]) < 3.0 and np.median(errT[ok]) < 0.8
        mReco = 20
        c.reconstruct_insert(mReco)
        cnt = [c.reco_download(s_)["counter"] for s_ in range(k)]
        assert cnt == [int((cls == s_).sum()) * mReco for s_ in range(k)]
    finally:
        c.close()

"""SURVEY.md section 8(f) row 1.  CPU: the numpy restatement of Reconstructor::reconstruct / Projector::setProjectee
(oracle/reco_port.py) pinned against the reference's own classes (oracle/_ref).  GPU: thb_reconstruct and
thb_set_projectee (cuFFT + elementwise kernels) against both."""
import numpy as np
import pytest

from thunder_b200 import synth


def _accumulators(ref, N, pf, nImg, seed):
    """back-projection of noiseless slices of a phantom at random orientations, by the reference's own classes.
    nImg must cover Fourier space (~N^2/3 images): on under-sampled accumulators the reference's gridding iteration
    itself diverges (max | |C| - 1 | grows to 1e3 within 15 iterations) and amplifies rounding differences between any
    two FFT libraries to percents - observed between FFTW and pocketfft as well as cuFFT."""
    rng = np.random.default_rng(seed)
    vol = synth.phantom(N, 10, seed=3)
    P = ref.Projector(pf)
    P.set_from_real(vol)
    volFT = P.padded_ft()
    pixM = ref.pixel_list(N, pf, float(N // 2 - 2), 0.0)
    PM = len(pixM["iCol"])
    reco = ref.Reconstructor(N, N, pf, 8)
    reco.set_precal(pixM["iColPad"], pixM["iRowPad"], pixM["iPxl"], pixM["iSig"])
    quats = synth.random_quats(nImg, rng)
    dat = np.stack([P.project(ref.rotate3D(q), pixM["iCol"], pixM["iRow"]) for q in quats]).astype(np.complex64)
    ctf = rng.uniform(0.3, 1.0, (nImg, PM)).astype(np.float32)
    reco.insert_loop((dat * ctf).astype(np.complex64), ctf, np.ones(nImg, np.float32), np.zeros((nImg, 2)), quats[:, None, :],
                     np.zeros((nImg, 1, 2)), pixM["iCol"], pixM["iRow"], N)
    acc = reco.get()
    P.close()
    return vol, volFT, reco, acc


CASES = [dict(gridCorr=True, fsc=False, joinHalf=False), dict(gridCorr=False, fsc=False, joinHalf=False),
         dict(gridCorr=True, fsc=True, joinHalf=False), dict(gridCorr=True, fsc=True, joinHalf=True)]


def test_numpy_restatement_matches_reference_classes(ref):
    from oracle import reco_port
    N, pf = 32, 2
    vol, volFT, reco, acc = _accumulators(ref, N, pf, 400, 1)
    mine = reco_port.set_projectee(vol, pf)
    assert np.linalg.norm(mine - volFT) <= 2e-6 * np.linalg.norm(volFT)
    fsc = np.linspace(0.99, 0.2, N // 2 + 1).astype(np.float32)
    for case in CASES:
        reco.set(acc["F"], acc["T"])
        reco.prepareTF()
        want = reco.reconstruct(N, gridCorr=case["gridCorr"], joinHalf=case["joinHalf"], fsc=fsc if case["fsc"] else None)
        got, nit = reco_port.reconstruct(acc["F"], acc["T"], N, pf, grid_corr=case["gridCorr"], fsc=fsc if case["fsc"] else None,
                                         join_half=case["joinHalf"])
        assert np.linalg.norm(got - want) <= 5e-6 * np.linalg.norm(want), case
        if not case["fsc"]:
            assert np.corrcoef(want.ravel(), vol.ravel())[0, 1] > 0.999          # and it IS a reconstruction of the phantom
    reco.close()


@pytest.mark.gpu
@pytest.mark.parametrize("N", [32, 64])
def test_device_reconstruct_and_set_projectee(ctx, N):
    from oracle import reco_port, refapi
    ref = refapi if refapi.available() else None
    pf = 2
    rng = np.random.default_rng(N)
    m = N * pf
    if ref is not None:
        vol, volFT, reco, acc = _accumulators(ref, N, pf, 400 if N <= 32 else 1500, 2)
    else:   # the reference library did not travel: synthetic accumulators with the statistics of a back-projection
        vol = synth.phantom(N, 10, seed=3)
        volFT = reco_port.set_projectee(vol, pf)
        T = (np.abs(rng.normal(size=(m, m, m // 2 + 1))) * 5 + 1).astype(np.float32)
        acc = dict(F=(volFT * T).astype(np.complex64), T=T)
        reco = None
    ctx.reco_alloc(0, m)
    fsc = np.linspace(0.99, 0.2, N // 2 + 1).astype(np.float32)
    for case in CASES:
        f = fsc if case["fsc"] else None
        ctx.reco_upload(0, acc["F"], acc["T"])
        got, nit = ctx.reconstruct(0, N, pf, gridCorr=case["gridCorr"], joinHalf=case["joinHalf"], fsc=f)
        want, nit0 = reco_port.reconstruct(acc["F"], acc["T"], N, pf, grid_corr=case["gridCorr"], fsc=f, join_half=case["joinHalf"])
        assert nit == nit0, case
        assert np.linalg.norm(got - want) <= 2e-5 * np.linalg.norm(want), case
        if reco is not None:
            reco.set(acc["F"], acc["T"])
            reco.prepareTF()
            want2 = reco.reconstruct(N, gridCorr=case["gridCorr"], joinHalf=case["joinHalf"], fsc=f)
            assert np.linalg.norm(got - want2) <= 2e-5 * np.linalg.norm(want2), case
    # setProjectee: explicit volume, and straight from the reconstruction kept on the device
    ctx.set_projectee(1, vol, N, pf)
    got = ctx.get_volume(1)
    assert np.linalg.norm(got - volFT) <= 5e-6 * np.linalg.norm(volFT)
    ctx.reco_upload(0, acc["F"], acc["T"])
    rec, _ = ctx.reconstruct(0, N, pf)
    ctx.set_projectee(1, None, N, pf)
    got2 = ctx.get_volume(1)
    want2 = reco_port.set_projectee(rec, pf)
    assert np.linalg.norm(got2 - want2) <= 5e-6 * np.linalg.norm(want2)
    # and the projector built that way drives the E kernel
    pix = __import__("thunder_b200").capi.pixel_list(N, pf, N // 2 - 3.0, 1.0)
    ctx.set_expect_pixels(N, pf, pix["iCol"], pix["iRow"])
    sl = ctx.project(1, synth.random_quats(2, rng))
    assert np.isfinite(sl).all() and np.abs(sl).max() > 0
    if reco is not None:
        reco.close()


@pytest.mark.gpu
def test_device_reconstruct_box128_against_the_reference_class():
    """thb_reconstruct at box 128 (256^3 transforms, BASELINE config 1's size) against Reconstructor::reconstruct itself: the
    accumulators come from the device insert of 6 000 noiseless slices (enough to cover Fourier space, see _accumulators), both
    sides reconstruct the SAME F / T with and without the FSC weighting"""
    from oracle import refapi
    from thunder_b200 import capi
    if not refapi.available():
        pytest.skip("oracle/_ref not present")
    N, pf, nImg = 128, 2, 6000
    m = N * pf
    rng = np.random.default_rng(128)
    vol = synth.phantom(N, 20, seed=3)
    c = capi.Context(0)
    try:
        pixM = capi.pixel_list(N, pf, float(N // 2 - 2), 0.0)
        PM = len(pixM["iCol"])
        c.set_projectee(0, vol, N, pf)
        c.set_expect_pixels(N, pf, pixM["iCol"], pixM["iRow"])
        c.set_insert_pixels(N, pf, pixM["iColPad"], pixM["iRowPad"])
        c.reco_alloc(0, m)
        for b in range(0, nImg, 1000):
            quats = synth.random_quats(1000, rng)
            ctf = rng.uniform(0.3, 1.0, (1000, PM)).astype(np.float32)
            dat = (c.project(0, quats) * ctf).astype(np.complex64)
            c.upload_stack(capi.STACK_INSERT, dat, ctf)
            c.insert(np.ones(1000, np.float32), quats[:, None, :], np.zeros((1000, 1, 2)))
        acc = c.reco_download(0)
        reco = refapi.Reconstructor(N, N, pf, 16)
        fsc = np.linspace(0.99, 0.2, N // 2 + 1).astype(np.float32)
        for f in (None, fsc):
            c.reco_upload(0, acc["F"], acc["T"])
            got, nit = c.reconstruct(0, N, pf, gridCorr=True, joinHalf=False, fsc=f)
            reco.set(acc["F"], acc["T"])
            reco.prepareTF()
            want = reco.reconstruct(N, gridCorr=True, joinHalf=False, fsc=f, nThread=16)
            assert np.linalg.norm(got - want) <= 2e-5 * np.linalg.norm(want), (f is not None, nit)
            if f is None:
                assert np.corrcoef(want.ravel(), vol.ravel())[0, 1] > 0.99
        reco.close()
    finally:
        c.close()


# ------------------------------------------------------------------------------------------- section 8(f) row 2
def test_numpy_recentre_remask_matches_reference(ref):
    from oracle import reco_port
    rng = np.random.default_rng(0)
    for N in (32, 64):
        ft = np.fft.rfft2(rng.normal(size=(N, N)).astype(np.float32)).astype(np.complex64)
        for zm in (False, True):
            want = ref.recentre_remask(ft, (1.3, -2.7), N * 0.35, zm)
            got = reco_port.recentre_remask(ft, (1.3, -2.7), N * 0.35, zm)
            assert np.linalg.norm(got - want) <= 1e-6 * np.linalg.norm(want)


@pytest.mark.gpu
def test_device_remask_pack(ctx):
    """thb_remask_pack == reCentreImg + reMaskImg + allocPreCal: masked image FTs and the packed E stack"""
    from oracle import reco_port, portapi as port
    from thunder_b200 import capi
    N, pf = 64, 2
    rng = np.random.default_rng(5)
    pixE = port.pixel_list(N, pf, 30.0, 1.0)
    nImg, nGroup, nRing = 7, 2, N // 2 + 1
    ori = np.stack([np.fft.rfft2(rng.normal(size=(N, N)).astype(np.float32)) for _ in range(nImg)]).astype(np.complex64)
    offset = rng.normal(scale=2.0, size=(nImg, 2))
    sigRcpTab = (-0.5 / rng.uniform(0.5, 2.0, (nGroup, nRing))).astype(np.float32)
    group = rng.integers(0, nGroup, nImg).astype(np.int32)
    attr = np.stack([np.full(nImg, 3e5), rng.uniform(1e4, 3e4, nImg), rng.uniform(1e4, 3e4, nImg), rng.uniform(0, np.pi, nImg),
                     np.full(nImg, 2.7e7), np.full(nImg, 0.1), np.zeros(nImg)], axis=1).astype(np.float32)
    ctx.set_expect_pixels(N, pf, pixE["iCol"], pixE["iRow"])
    ctx.stack_reserve(capi.STACK_EXPECT, nImg + 1)
    maskR = 0.35 * N
    for zm in (True, False):
        imgs = ctx.remask_pack(1, ori, offset, maskR, pixE["iPxl"], pixE["iSig"], sigRcpTab, attr, 1.32, zeroMask=zm, groupOfImg=group,
                               want_images=True)
        got = ctx.download_stack(capi.STACK_EXPECT, 1, nImg)
        for l in range(nImg):
            want = reco_port.recentre_remask(ori[l], offset[l], maskR, zm)
            assert np.linalg.norm(imgs[l] - want) <= 2e-6 * np.linalg.norm(want), (zm, l)
            assert np.array_equal(got["dat"][l], imgs[l].ravel()[pixE["iPxl"]])
        assert np.array_equal(got["sigRcp"], sigRcpTab[group][:, pixE["iSig"]])


# ------------------------------------------------------------------------------------------- section 8(f) row 3
@pytest.mark.gpu
def test_device_sigma_accumulate(ctx):
    """thb_sigma_accumulate == the image loop of Optimiser::allReduceSigma driven through the reference's own functions
    (Projector::project(Image&, rot, t), CTF(Image&), powerSpectrum, NEG_FT / ADD_FT)"""
    from oracle import refapi
    from thunder_b200 import capi
    if not refapi.available():
        pytest.skip("oracle/_ref not present")
    N, pf, rSig = 64, 2, 24
    rng = np.random.default_rng(77)
    vol = synth.phantom(N, 10, seed=4)
    P = refapi.Projector(pf)
    P.set_from_real(vol)
    volFT = P.padded_ft()
    pixE = capi.pixel_list(N, pf, float(rSig), 0.0)          # the sigma pixel set: rL = 0, r = rSig
    pixM = capi.pixel_list(N, pf, 29.0, 0.0)                 # the M list reaches further out
    nImg, nGroup = 9, 3
    img = np.stack([np.fft.rfft2(rng.normal(size=(N, N)).astype(np.float32)) for _ in range(nImg)]).astype(np.complex64)
    ori = np.stack([np.fft.rfft2(rng.normal(size=(N, N)).astype(np.float32)) for _ in range(nImg)]).astype(np.complex64)
    img *= np.float32(np.abs(volFT).mean() * 3 / np.abs(img).mean()); ori *= np.float32(np.abs(volFT).mean() * 3 / np.abs(ori).mean())
    quat = synth.random_quats(nImg, rng)
    tran = rng.normal(scale=1.5, size=(nImg, 2)); offS = rng.normal(scale=1.0, size=(nImg, 2))
    group = rng.integers(0, nGroup, nImg).astype(np.int32)
    attr = np.stack([np.full(nImg, 3e5), rng.uniform(1e4, 3e4, nImg), rng.uniform(1e4, 3e4, nImg), rng.uniform(0, np.pi, nImg),
                     np.full(nImg, 2.7e7), np.full(nImg, 0.1), np.zeros(nImg)], axis=1).astype(np.float32)
    want = refapi.sigma_accumulate(P, img, ori, quat, tran, offS, attr, 1.32, group, nGroup, rSig)
    ctx.set_expect_pixels(N, pf, pixE["iCol"], pixE["iRow"])
    ctx.set_insert_pixels(N, pf, pixM["iColPad"], pixM["iRowPad"])
    ctx.set_volume(0, volFT)
    tab = np.full((1, N), -0.5, np.float32)
    ctx.stack_reserve(capi.STACK_EXPECT, nImg); ctx.stack_reserve(capi.STACK_INSERT, nImg)
    ctx.pack_stack(capi.STACK_EXPECT, 0, img, pixE["iPxl"], attr, 1.32, iSig=pixE["iSig"], sigRcpTab=tab)
    ctx.pack_stack(capi.STACK_INSERT, 0, ori, pixM["iPxl"], attr, 1.32)
    got = ctx.sigma_accumulate(quat, tran, offS, group, nGroup, rSig, pixE["iSig"], pixM["iSig"])
    for name, g, w in zip(("sigM", "sigN", "svd"), got, want):
        assert np.array_equal(g[:, -1], w[:, -1]), name                        # weight sums: images per group
        assert np.allclose(g[:, :rSig], w[:, :rSig], rtol=2e-4, atol=0), name
    P.close()


@pytest.mark.gpu
def test_device_norm_correction(ctx):
    """thb_norm_residual == the image loop of Optimiser::normCorrection through the reference's own functions;
    thb_scale_images == the rescaling of _img / _imgOri that follows"""
    from oracle import refapi
    from thunder_b200 import capi
    if not refapi.available():
        pytest.skip("oracle/_ref not present")
    N, pf, r, rL, rNorm = 64, 2, 26.0, 2.0, 21.3
    rng = np.random.default_rng(78)
    P = refapi.Projector(pf)
    P.set_from_real(synth.phantom(N, 10, seed=4))
    volFT = P.padded_ft()
    pixE = capi.pixel_list(N, pf, r, rL)
    pixM = capi.pixel_list(N, pf, 29.0, 0.0)
    nImg = 7
    img = np.stack([np.fft.rfft2(rng.normal(size=(N, N)).astype(np.float32)) for _ in range(nImg)]).astype(np.complex64)
    img *= np.float32(np.abs(volFT).mean() * 3 / np.abs(img).mean())
    quat = synth.random_quats(nImg, rng); tran = rng.normal(scale=1.5, size=(nImg, 2))
    attr = np.stack([np.full(nImg, 3e5), rng.uniform(1e4, 3e4, nImg), rng.uniform(1e4, 3e4, nImg), rng.uniform(0, np.pi, nImg),
                     np.full(nImg, 2.7e7), np.full(nImg, 0.1), np.zeros(nImg)], axis=1).astype(np.float32)
    want = refapi.norm_residual(P, img, quat, tran, attr, 1.32, rL, rNorm)
    ctx.set_expect_pixels(N, pf, pixE["iCol"], pixE["iRow"])
    ctx.set_insert_pixels(N, pf, pixM["iColPad"], pixM["iRowPad"])
    ctx.set_volume(0, volFT)
    ctx.stack_reserve(capi.STACK_EXPECT, nImg); ctx.stack_reserve(capi.STACK_INSERT, nImg)
    ctx.pack_stack(capi.STACK_EXPECT, 0, img, pixE["iPxl"], attr, 1.32, iSig=pixE["iSig"], sigRcpTab=np.full((1, N), -0.5, np.float32))
    ctx.pack_stack(capi.STACK_INSERT, 0, img, pixM["iPxl"], attr, 1.32)
    got = ctx.norm_residual(quat, tran, rL, rNorm)
    assert np.allclose(got, want, rtol=2e-4, atol=0)
    # rescale to the median, as :6371-6391
    scale = np.sqrt(np.median(got) / got).astype(np.float32)
    before = [ctx.download_stack(k, 0, nImg) for k in (capi.STACK_EXPECT, capi.STACK_INSERT)]
    ctx.scale_images(scale)
    after = [ctx.download_stack(k, 0, nImg) for k in (capi.STACK_EXPECT, capi.STACK_INSERT)]
    for b, a in zip(before, after):
        assert np.array_equal(a["dat"], (b["dat"] * scale[:, None]).astype(np.complex64))
        assert np.array_equal(a["ctf"], b["ctf"])
    P.close()


@pytest.mark.gpu
@pytest.mark.parametrize("group", ["C4", "D2", "T"])
def test_device_symmetrize(ctx, group):
    """thb_symmetrize == Reconstructor::symmetrizeT / symmetrizeF / symmetrizeO (SYMMETRIZE_FT) on the same accumulators"""
    from oracle import refapi
    if not refapi.available():
        pytest.skip("oracle/_ref not present")
    N, pf = 32, 2
    rng = np.random.default_rng(90)
    R = refapi.Reconstructor(N, N, pf)
    m = R.pad_size()
    shape = (m, m, m // 2 + 1)
    # a smooth Hermitian-consistent F (FT of a real volume) and a positive T, as accumulators look after the all-reduce
    vol = synth.phantom(N, 8, seed=12)
    F = np.fft.rfftn(np.pad(vol, [(0, m - N)] * 3)).astype(np.complex64)
    T = (np.abs(np.fft.rfftn(np.pad(synth.phantom(N, 5, seed=13), [(0, m - N)] * 3))) + 1.0).astype(np.float32)
    assert F.shape == shape
    O = np.array([1.5, -0.25, 0.75]); counter = 17
    R.set(F, T)
    R.symmetrize(group, O, counter)
    want = R.get()
    elems = refapi.symmetry_elements(group)
    ctx.reco_alloc(0, m)
    ctx.reco_upload(0, F, T)
    ctx.symmetrize(0, elems, R.max_radius() * pf + 1)
    got = ctx.reco_download(0)
    # voxels whose integer |v|^2 equals r^2 exactly sit on the cut `|R v|^2 < r^2`: there the last bit of the rotated
    # coordinates (Eigen's evaluation order vs. ours, 1e-13 on 841) decides whether the element contributes, in the reference
    # as much as here.  They are compared for membership in {with, without the borderline elements}, the rest to 2e-6.
    r = R.max_radius() * pf + 1
    kk, jj, ii = np.meshgrid(np.fft.fftfreq(m, 1 / m), np.fft.fftfreq(m, 1 / m), np.arange(m // 2 + 1), indexing="ij")
    edge = (ii * ii + jj * jj + kk * kk) == r * r
    assert 0 < edge.sum() < 2000
    assert np.abs(got["F"] - want["F"])[~edge].max() <= 2e-6 * np.abs(want["F"]).max()
    assert np.abs(got["T"] - want["T"])[~edge].max() <= 2e-6 * np.abs(want["T"]).max()
    assert np.abs(got["T"] - want["T"])[edge].max() <= len(elems) * np.abs(T).max()
    # O and counter start at zero on the device: the linear map is checked on the reference's numbers
    Rm = elems.reshape(-1, 3, 3).transpose(0, 2, 1)                 # column-major -> matrices
    assert np.allclose(want["O"], O + sum(r @ O for r in Rm), atol=1e-12)
    assert want["counter"] == counter * (1 + len(elems))
    R.close()


# ------------------------------------------------------------------------------------------- CPU pins of the restatements
def _sigma_problem(N=32, rSig=12, nImg=4, seed=5):
    from oracle import refapi
    rng = np.random.default_rng(seed)
    P = refapi.Projector(2)
    P.set_from_real(synth.phantom(N, 6, seed=4))
    volFT = P.padded_ft()
    mk = lambda: np.stack([np.fft.rfft2(rng.normal(size=(N, N)).astype(np.float32)) for _ in range(nImg)]).astype(np.complex64)
    img, ori = mk(), mk()
    sc = np.float32(np.abs(volFT).mean() * 3 / np.abs(img).mean())
    img *= sc; ori *= sc
    quat = synth.random_quats(nImg, rng); tran = rng.normal(scale=1.5, size=(nImg, 2)); offS = rng.normal(size=(nImg, 2))
    attr = np.stack([np.full(nImg, 3e5), rng.uniform(1e4, 3e4, nImg), rng.uniform(1e4, 3e4, nImg), rng.uniform(0, np.pi, nImg),
                     np.full(nImg, 2.7e7), np.full(nImg, 0.1), np.zeros(nImg)], axis=1).astype(np.float32)
    return P, volFT, img, ori, quat, tran, offS, attr


def test_sigma_and_norm_restatements_match_reference():
    """oracle/sigma_port.py == the image loops of allReduceSigma / normCorrection driven through the reference's own functions"""
    from oracle import refapi, sigma_port
    if not refapi.available():
        pytest.skip("oracle/_ref not present")
    N, rSig, nGroup = 32, 12, 2
    P, volFT, img, ori, quat, tran, offS, attr = _sigma_problem(N, rSig)
    group = np.array([0, 1, 1, 0], np.int32)
    want = refapi.sigma_accumulate(P, img, ori, quat, tran, offS, attr, 1.32, group, nGroup, rSig)
    got = sigma_port.sigma_accumulate(volFT, 2, img, ori, quat, tran, offS, attr, 1.32, group, nGroup, rSig)
    for g, w in zip(got, want):
        assert np.array_equal(g[:, -1], w[:, -1])
        assert np.allclose(g[:, :rSig], w[:, :rSig], rtol=2e-4)
    wantN = refapi.norm_residual(P, img, quat, tran, attr, 1.32, 1.0, 10.4)
    gotN = sigma_port.norm_residual(volFT, 2, img, quat, tran, attr, 1.32, 1.0, 10.4)
    assert np.allclose(gotN, wantN, rtol=2e-4)
    P.close()


@pytest.mark.parametrize("group", ["C4", "D2"])
def test_symmetrize_restatement_matches_reference(group):
    from oracle import refapi, reco_port
    if not refapi.available():
        pytest.skip("oracle/_ref not present")
    N, pf = 16, 2
    R = refapi.Reconstructor(N, N, pf)
    m = R.pad_size()
    F = np.fft.rfftn(np.pad(synth.phantom(N, 5, seed=12), [(0, m - N)] * 3)).astype(np.complex64)
    T = (np.abs(np.fft.rfftn(np.pad(synth.phantom(N, 4, seed=13), [(0, m - N)] * 3))) + 1.0).astype(np.float32)
    R.set(F, T)
    R.symmetrize(group)
    want = R.get()
    r = R.max_radius() * pf + 1
    gF, gT = reco_port.symmetrize(F, T, refapi.symmetry_elements(group), r)
    kk, jj, ii = np.meshgrid(np.fft.fftfreq(m, 1 / m), np.fft.fftfreq(m, 1 / m), np.arange(m // 2 + 1), indexing="ij")
    edge = (ii * ii + jj * jj + kk * kk) == r * r          # membership of the cut is decided by the last bit there
    assert np.abs(gF - want["F"])[~edge].max() <= 2e-6 * np.abs(want["F"]).max()
    assert np.abs(gT - want["T"])[~edge].max() <= 2e-6 * np.abs(want["T"]).max()
    R.close()

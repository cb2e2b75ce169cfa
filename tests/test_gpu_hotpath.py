"""GPU parity tests: the CUDA hot path, called through the C ABI, against the oracle.

Oracle = the reference-built library (oracle/_ref) when present, else the pinned plain-C port, plus the
committed golden fixture.  Tolerances (fp32 arithmetic, double coordinates):
  * projected slice values: |d| <= 2e-6 * max|slice|  (FMA contraction vs separate mul/add)
  * log-likelihoods: |d| <= 2e-6 * |logL| + 1e-4  (the reference's own scalar-vs-SIMD tolerance is 1e-5 rel)
  * marginal weights uR/uT/uC: rel 2e-3 after exp() of a difference of large numbers, and
    <= 20 * eps * |logL| in log space
  * F / T volumes: relative L2 <= 1e-6 (only summation order differs), per-shell FSC >= 0.99999
"""
import numpy as np
import pytest

from thunder_b200 import capi, synth

pytestmark = pytest.mark.gpu


def _oracle():
    from oracle import portapi, refapi
    return portapi, (refapi if refapi.available() else None)


def _logL_tol(P, L):
    """|d logL| allowed at full image sizes: the reference accumulates the P terms of a log-likelihood sequentially in
    ONE fp32 register (logDataVSPrior_m_huabin, src/Optimiser.cpp:9187-9213), a random walk of half-ulp errors of the
    running sum: ~ 2^-24 * sqrt(P) * |logL|.  Far inside the 1e-4 * |logL| + 1e-4 of SURVEY.md section 8c."""
    return float(np.finfo(np.float32).eps / 2 * np.sqrt(P) * np.abs(L).max() + 1e-4)


def _rel_l2(a, b):
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-30))


# ------------------------------------------------------------------------------------------- golden, N = 16
def test_project_golden(ctx, golden):
    N, pf = int(golden["N"]), int(golden["pf"])
    ctx.drop_volumes()          # the session context may hold volumes of another size from earlier tests
    ctx.set_expect_pixels(N, pf, golden["pixE_iCol"], golden["pixE_iRow"])
    ctx.set_volume(0, golden["volFT"])
    assert np.array_equal(ctx.get_volume(0), golden["volFT"])
    got = ctx.project(0, golden["quat"])
    ref = golden["slices"]
    assert np.abs(got - ref).max() <= 2e-6 * np.abs(ref).max()
    # identity rotation: integer coordinates, zero fractional weights -> exact voxel values
    assert np.array_equal(got[0], ref[0])


def test_expect_local_golden(ctx, golden):
    N, pf = int(golden["N"]), int(golden["pf"])
    ctx.drop_volumes()          # the session context may hold volumes of another size from earlier tests
    ctx.set_expect_pixels(N, pf, golden["pixE_iCol"], golden["pixE_iRow"])
    ctx.set_volume(0, golden["volFT"])
    ctx.upload_stack(capi.STACK_EXPECT, golden["dat"][None], golden["ctf"][None], golden["sigRcp"][None])
    nR, nT = golden["logL"].shape
    wR = np.full((1, nR), 1.0 / nR); wT = np.full((1, nT), 1.0 / nT)
    out = ctx.expect_local(golden["quat"][None], golden["tran"][None], wR, wT)
    ref = golden["logL"]
    assert np.abs(out["logL"][0] - ref).max() <= 2e-6 * np.abs(ref).max() + 1e-4
    # the baseline is one of the log-likelihoods and carries that sample's tolerance
    assert abs(out["base"][0] - ref.max()) <= 2e-6 * np.abs(ref).max() + 1e-4


def test_insert_golden(ctx, golden):
    N, pf = int(golden["N"]), int(golden["pf"])
    ctx.drop_volumes()
    ctx.set_insert_pixels(N, pf, golden["pixM_iColPad"], golden["pixM_iRowPad"])
    ctx.upload_stack(capi.STACK_INSERT, golden["datM"], golden["ctfM"])
    ctx.reco_alloc(0, N * pf)
    ctx.insert(golden["w"], golden["nr"], golden["nt"], offS=golden["offS"])
    out = ctx.reco_download(0)
    assert out["counter"] == int(golden["counter"])
    assert np.allclose(out["O"], golden["O"], rtol=1e-12, atol=1e-12)
    assert _rel_l2(out["F"], golden["F"]) <= 1e-6
    assert _rel_l2(out["T"], golden["T"]) <= 1e-6
    outn = ctx.reco_download(0, normalise=True)       # prepareTF normalisation, sf = 1 / T[0]
    assert _rel_l2(outn["F"], golden["Fn"]) <= 1e-6
    assert _rel_l2(outn["T"], golden["Tn"]) <= 1e-6
    assert abs(outn["T"].ravel()[0] - 1.0) < 1e-6
    # reset really clears
    ctx.reco_reset(0)
    z = ctx.reco_download(0)
    assert not z["F"].any() and not z["T"].any() and z["counter"] == 0 and not z["O"].any()


# ------------------------------------------------------------------------------------------- random, larger
@pytest.fixture(scope="module")
def problem():
    """N = 64 synthetic problem with two half-set volumes, built with the oracle's projector."""
    port, ref = _oracle()
    N, pf = 64, 2
    rng = np.random.default_rng(77)
    vols = [synth.padded_ft(synth.phantom(N, 12, seed=s), pf) for s in (1, 2)]
    pixE = port.pixel_list(N, pf, 30.0, 1.0)
    pixM = port.pixel_list(N, pf, 31.0, 0.0)
    nImg = 6
    slot = np.array([0, 1, 0, 1, 1, 0], np.int32)

    def project_fn(quats):
        return np.stack([port.project(vols[slot[l]], pf, port.rotate3D(q), pixE["iCol"], pixE["iRow"])
                         for l, q in enumerate(quats)])
    par = synth.make_particles(nImg, N, pixE, project_fn, seed=5, snr_scale=4.0)
    # local-search clouds: nR rotations about the truth (+ a few wild ones), nT translations
    nR, nT = 37, 9
    quat = np.stack([synth.acg_cloud(par["quat"][l], 2e-4, nR, rng) for l in range(nImg)])
    quat[:, -3:] = synth.random_quats(nImg * 3, rng).reshape(nImg, 3, 4)
    quat[:, 0] = par["quat"]
    tran = par["tran"][:, None, :] + rng.normal(scale=0.7, size=(nImg, nT, 2))
    tran[:, 0] = par["tran"]
    wR = rng.uniform(0.5, 1.5, (nImg, nR)); wR /= wR.sum(1, keepdims=True)
    wT = rng.uniform(0.5, 1.5, (nImg, nT)); wT /= wT.sum(1, keepdims=True)
    return dict(N=N, pf=pf, vols=vols, pixE=pixE, pixM=pixM, par=par, slot=slot, quat=quat, tran=tran, wR=wR, wT=wT,
                nImg=nImg, nR=nR, nT=nT, rng=rng)


def _setup_E(ctx, pb):
    ctx.set_expect_pixels(pb["N"], pb["pf"], pb["pixE"]["iCol"], pb["pixE"]["iRow"])
    for s, v in enumerate(pb["vols"]):
        ctx.set_volume(s, v)
    ctx.upload_stack(capi.STACK_EXPECT, pb["par"]["dat"], pb["par"]["ctf"], pb["par"]["sigRcp"], pb["slot"])


def test_project_random(ctx, problem):
    pb = problem
    port, ref = _oracle()
    _setup_E(ctx, pb)
    q = synth.random_quats(16, np.random.default_rng(1))
    got = ctx.project(1, q)
    for i in range(len(q)):
        want = port.project(pb["vols"][1], pb["pf"], port.rotate3D(q[i]), pb["pixE"]["iCol"], pb["pixE"]["iRow"])
        assert np.abs(got[i] - want).max() <= 2e-6 * np.abs(want).max()


def test_expect_local_random(ctx, problem):
    pb = problem
    port, ref = _oracle()
    _setup_E(ctx, pb)
    out = ctx.expect_local(pb["quat"], pb["tran"], pb["wR"], pb["wT"])
    for l in range(pb["nImg"]):
        o = port.expect_local(pb["vols"][pb["slot"][l]], pb["pf"], pb["N"], pb["pixE"]["iCol"], pb["pixE"]["iRow"],
                              pb["par"]["dat"][l], pb["par"]["ctf"][l], pb["par"]["sigRcp"][l], pb["quat"][l], pb["tran"][l],
                              pb["wR"][l], pb["wT"][l])
        L = o["logL"]
        assert np.abs(out["logL"][l] - L).max() <= 2e-6 * np.abs(L).max() + 1e-4
        assert abs(out["base"][l] - o["base"]) <= 2e-6 * abs(o["base"]) + 1e-4
        # marginals: compare against float64 weights of the ORACLE's logL shifted to our baseline
        tol_log = 20 * np.finfo(np.float32).eps * np.abs(L).max()
        for key, want in (("uR", o["uR"]), ("uT", o["uT"])):
            got = out[key][l]
            big = want > 1e-6 * want.max()
            assert np.all(np.abs(np.log(got[big]) - np.log(want[big])) <= tol_log + 2e-3), key
        assert abs(np.log(out["uC"][l]) - np.log(o["uC"])) <= tol_log + 2e-3


def test_expect_local_subset_and_order(ctx, problem):
    """imgIdx selects / permutes images; results follow the selection"""
    pb = problem
    _setup_E(ctx, pb)
    full = ctx.expect_local(pb["quat"], pb["tran"], pb["wR"], pb["wT"])
    idx = np.array([4, 1, 3], np.int32)
    sub = ctx.expect_local(pb["quat"][idx], pb["tran"][idx], pb["wR"][idx], pb["wT"][idx], imgIdx=idx)
    assert np.array_equal(sub["logL"], full["logL"][idx])
    assert np.array_equal(sub["uR"], full["uR"][idx])


def test_expect_local_ragged_shapes(ctx, problem):
    """nT not a multiple of the register chunk (9), nR above one CTA (128), nT = 1, nR = 1"""
    pb = problem
    port, ref = _oracle()
    _setup_E(ctx, pb)
    rng = np.random.default_rng(9)
    for nR, nT in ((1, 1), (3, 10), (130, 2), (5, 19)):
        quat = synth.random_quats(nR, rng)[None]
        tran = rng.normal(scale=1.5, size=(1, nT, 2))
        wR = np.full((1, nR), 1.0 / nR); wT = np.full((1, nT), 1.0 / nT)
        out = ctx.expect_local(quat, tran, wR, wT, imgIdx=np.array([2], np.int32))
        o = port.expect_local(pb["vols"][pb["slot"][2]], pb["pf"], pb["N"], pb["pixE"]["iCol"], pb["pixE"]["iRow"],
                              pb["par"]["dat"][2], pb["par"]["ctf"][2], pb["par"]["sigRcp"][2], quat[0], tran[0], wR[0], wT[0])
        assert np.abs(out["logL"][0] - o["logL"]).max() <= 2e-6 * np.abs(o["logL"]).max() + 1e-4, (nR, nT)


def test_expect_lockstep_radial_order_equals_free_running(ctx, problem):
    """lockstep launch of the several-rotations-per-lane kernel (persistent grid, tile barriers, images of one slot adjacent)
    on the radial pixel order: more images than co-resident CTAs (several waves, the last one partial), both layouts, windows
    0 and 3 - the same log-likelihoods and weights as the default kernel on the default pixel order and as the oracle"""
    pb = problem
    port, ref = _oracle()
    rng = np.random.default_rng(404)
    reps = 130                                    # 780 images: 2 full waves of 296 + a partial one
    nImg, nR, nT = pb["nImg"] * reps, 125, 9
    dat = np.tile(pb["par"]["dat"], (reps, 1)); ctf = np.tile(pb["par"]["ctf"], (reps, 1)); sig = np.tile(pb["par"]["sigRcp"], (reps, 1))
    slot = rng.integers(0, 2, nImg).astype(np.int32)
    q0 = np.tile(pb["par"]["quat"], (reps, 1))
    quat = np.stack([synth.acg_cloud(q0[l], 2e-4, nR, rng) for l in range(nImg)])
    tran = np.tile(pb["par"]["tran"], (reps, 1))[:, None, :] + rng.normal(scale=0.7, size=(nImg, nT, 2))
    wR = rng.uniform(0.5, 1.5, (nImg, nR)); wR /= wR.sum(1, keepdims=True)
    wT = rng.uniform(0.5, 1.5, (nImg, nT)); wT /= wT.sum(1, keepdims=True)

    def run():
        ctx.set_expect_pixels(pb["N"], pb["pf"], pb["pixE"]["iCol"], pb["pixE"]["iRow"])
        for s_, v in enumerate(pb["vols"]):
            ctx.set_volume(s_, v)
        ctx.upload_stack(capi.STACK_EXPECT, dat, ctf, sig, slot)
        return ctx.expect_local(quat, tran, wR, wT)
    try:
        ctx.set_option("expect_order", 0); ctx.set_option("expect_impl", 3); ctx.set_option("expect_lock", 0)
        a = run()                                  # one rotation per lane, 8x8-block pixel order, one CTA per image
        outs = []
        ctx.set_option("expect_order", 1); ctx.set_option("expect_impl", 7); ctx.set_option("expect_lock", 1)
        for oct_, win, tiles in ((1, 2, 1), (0, 0, 1), (0, 3, 2)):
            ctx.set_option("quad_oct", oct_); ctx.set_option("expect_lock_window", win); ctx.set_option("expect_lock_tiles", tiles)
            outs.append(run())
    finally:
        for k, v in (("expect_order", 1), ("expect_impl", 0), ("expect_lock", 1), ("quad_oct", 1), ("expect_lock_window", 1), ("expect_lock_tiles", 1)):
            ctx.set_option(k, v)
        ctx.set_expect_pixels(pb["N"], pb["pf"], pb["pixE"]["iCol"], pb["pixE"]["iRow"])
    tol = 2e-6 * np.abs(a["logL"]).max() + 1e-4
    for b in outs:
        assert np.abs(b["logL"] - a["logL"]).max() <= tol
        assert np.abs(b["base"] - a["base"]).max() <= tol
        assert np.allclose(b["uT"], a["uT"], rtol=2e-3, atol=1e-6 * a["uT"].max())
        assert np.allclose(b["uR"], a["uR"], rtol=5e-3, atol=1e-6 * a["uR"].max())
        assert np.allclose(b["uC"], a["uC"], rtol=5e-3)
    assert np.array_equal(outs[1]["logL"], outs[2]["logL"])      # the barriers change timing, not arithmetic
    l = 5
    want = port.expect_local(pb["vols"][slot[l]], pb["pf"], pb["N"], pb["pixE"]["iCol"], pb["pixE"]["iRow"], dat[l], ctf[l], sig[l], quat[l],
                             tran[l], wR[l], wT[l])
    assert np.abs(outs[0]["logL"][l] - want["logL"]).max() <= tol



def test_expect_scan_matches_local(ctx, problem):
    """global-scan shape: one shared rotation/translation set against every image of a slot"""
    pb = problem
    port, ref = _oracle()
    _setup_E(ctx, pb)
    rng = np.random.default_rng(10)
    nR, nT = 40, 12
    quat = synth.random_quats(nR, rng); tran = rng.normal(scale=2.0, size=(nT, 2))
    pR = np.full(nR, 1.0 / nR); pT = np.full(nT, 1.0 / nT)
    out = ctx.expect_scan(1, quat, tran, pR, pT, want_logL=True)       # shared templates (thb_expect8.cuh), the default
    try:
        ctx.set_option("scan_templates", 0)                             # the fused kernel, every rotation re-gathered per image
        old = ctx.expect_scan(1, quat, tran, pR, pT, want_logL=True)
    finally:
        ctx.set_option("scan_templates", 1)
    assert np.abs(out["logL"] - old["logL"]).max() <= 2e-6 * np.abs(old["logL"]).max() + 1e-4
    assert np.abs(out["base"] - old["base"]).max() <= 2e-6 * np.abs(old["base"]).max() + 1e-4
    assert np.allclose(out["wR"], old["wR"], rtol=5e-3, atol=1e-6 * old["wR"].max()) and np.allclose(out["wC"], old["wC"], rtol=5e-3)
    imgs = np.nonzero(pb["slot"] == 1)[0]
    # pixel-major n-image likelihood of the oracle (logDataVSPrior_m_n) for a few templates
    datPM = np.ascontiguousarray(pb["par"]["dat"][imgs].T); ctfPM = np.ascontiguousarray(pb["par"]["ctf"][imgs].T)
    sigPM = np.ascontiguousarray(pb["par"]["sigRcp"][imgs].T)
    for r in (0, 7, 39):
        pri = port.project(pb["vols"][1], pb["pf"], port.rotate3D(quat[r]), pb["pixE"]["iCol"], pb["pixE"]["iRow"])
        for t in (0, 11):
            tra = port.translate(np.float32(tran[t, 0]), np.float32(tran[t, 1]), pb["N"], pb["pixE"]["iCol"], pb["pixE"]["iRow"])
            want = port.logDataVSPrior_m_n(datPM, (tra * pri).astype(np.complex64), ctfPM, sigPM, len(imgs), datPM.shape[0])
            got = out["logL"][imgs, r, t]
            assert np.abs(got - want).max() <= 2e-6 * np.abs(want).max() + 1e-4
    others = np.nonzero(pb["slot"] != 1)[0]
    assert not out["wR"][others].any() and not out["wC"][others].any()
    for l in imgs:
        w = np.exp(out["logL"][l].astype(np.float64) - out["base"][l])
        assert np.allclose(out["wR"][l], (w * pT).sum(1), rtol=1e-4, atol=1e-30)
        assert np.allclose(out["wT"][l], (w * pR[:, None]).sum(0), rtol=1e-4, atol=1e-30)


def test_expect_scan_large_rotation_set_against_the_reference_loop(ctx, problem):
    """the global-search shape at a size that takes several passes inside the kernel (300 rotations > 128 per pass, 20 translations > 15)
    and an image sub-range, against the reference's own scan loop (oracle/_ref: ref_scan restates src/Optimiser.cpp:756-914 around
    Projector::project + logDataVSPrior_m_n): weights relative to the baseline, marginals with the priors of the other dimension"""
    port, ref = _oracle()
    if ref is None:
        pytest.skip("oracle/_ref not present")
    pb = problem
    _setup_E(ctx, pb)
    rng = np.random.default_rng(31)
    nR, nT = 300, 20
    quat = synth.random_quats(nR, rng); quat[:pb["nImg"]] = pb["par"]["quat"]
    tran = rng.normal(scale=1.5, size=(nT, 2))
    pR = rng.uniform(0.5, 1.5, nR); pR /= pR.sum()
    pT = rng.uniform(0.5, 1.5, nT); pT /= pT.sum()
    imgs = np.nonzero(pb["slot"] == 1)[0]
    P = ref.Projector(pb["pf"])
    P.set_padded_ft(pb["vols"][1])
    want = ref.scan([P], False, pb["par"]["dat"][imgs], pb["par"]["ctf"][imgs], pb["par"]["sigRcp"][imgs], pb["pixE"]["iCol"], pb["pixE"]["iRow"], pb["N"],
                    quat, tran, pR, pT, nThread=4)
    P.close()
    for rng_ in (None, (1, pb["nImg"] - 1)):
        out = ctx.expect_scan(1, quat, tran, pR, pT, img_range=rng_)
        off = 0 if rng_ is None else rng_[0]
        for j, l in enumerate(imgs):
            if l < off:
                continue
            assert abs(out["base"][l - off] - want["base"][j]) <= 2e-6 * abs(want["base"][j]) + 1e-4
            for key, k2 in (("wR", "wR"), ("wT", "wT")):
                a, b = out[key][l - off].astype(np.float64), want[k2][0][j].astype(np.float64)
                big = b > 1e-4 * b.max()
                assert np.allclose(a[big], b[big], rtol=5e-3), (key, l)
            assert np.isclose(out["wC"][l - off], want["wC"][j, 0], rtol=5e-3)


def test_insert_random_two_halves(ctx, problem):
    pb = problem
    port, ref = _oracle()
    N, pf = pb["N"], pb["pf"]
    rng = np.random.default_rng(21)
    PM = len(pb["pixM"]["iCol"])
    nImg, mReco = pb["nImg"], 11
    datM = (rng.normal(size=(nImg, PM)) + 1j * rng.normal(size=(nImg, PM))).astype(np.complex64)
    ctfM = rng.uniform(-1, 1, (nImg, PM)).astype(np.float32)
    nr = np.stack([synth.acg_cloud(pb["par"]["quat"][l], 3e-4, mReco, rng) for l in range(nImg)])
    nt = rng.normal(scale=2.0, size=(nImg, mReco, 2))
    w = (rng.uniform(0.5, 1.0, nImg) / mReco).astype(np.float32)
    offS = rng.normal(scale=0.5, size=(nImg, 2))
    ctx.set_insert_pixels(N, pf, pb["pixM"]["iColPad"], pb["pixM"]["iRowPad"])
    ctx.upload_stack(capi.STACK_INSERT, datM, ctfM, slotOfImg=pb["slot"])
    for s in (0, 1):
        ctx.reco_alloc(s, N * pf)
    ctx.insert(w, nr, nt, offS=offS)
    for s in (0, 1):
        sel = np.nonzero(pb["slot"] == s)[0]
        want = port.insert_loop(N * pf, pf, N, datM[sel], ctfM[sel], w[sel], offS[sel], nr[sel], nt[sel],
                                pb["pixM"]["iCol"], pb["pixM"]["iRow"])
        got = ctx.reco_download(s)
        assert got["counter"] == want["counter"] == len(sel) * mReco
        assert np.allclose(got["O"], want["O"], rtol=1e-11, atol=1e-11)
        assert _rel_l2(got["F"], want["F"]) <= 1e-6
        assert _rel_l2(got["T"], want["T"]) <= 1e-6
        f = synth.fsc(got["F"], want["F"], N * pf // 2 - 2)
        assert f[1:].min() >= 0.99999
    # linearity: inserting the same list again doubles the accumulators (size-independent property)
    before = ctx.reco_download(0)
    ctx.insert(w, nr, nt, offS=offS)
    after = ctx.reco_download(0)
    assert _rel_l2(after["F"], 2 * before["F"]) <= 1e-6
    assert after["counter"] == 2 * before["counter"]


def test_insert_slab_ordering_thin_slabs_and_axis_rotations(ctx, problem):
    """the slab-ordered M kernel (thb_insert2.cuh): every sample must be scattered exactly once whatever the slab thickness -
    1, 3, 7 planes and the automatic one - including the slices on which its interval arithmetic degenerates (identity: z = 0
    everywhere; quarter turns about x / y / z: slices along the axes, x = 0 for a whole image, folds on pixel boundaries) and
    duplicated draws (merged groups); against the oracle and against the image-ordered kernel of round 1"""
    pb = problem
    port, ref = _oracle()
    N, pf = pb["N"], pb["pf"]
    rng = np.random.default_rng(2024)
    PM = len(pb["pixM"]["iCol"])
    nImg, mReco = pb["nImg"], 12
    datM = (rng.normal(size=(nImg, PM)) + 1j * rng.normal(size=(nImg, PM))).astype(np.complex64)
    ctfM = rng.uniform(-1, 1, (nImg, PM)).astype(np.float32)
    nr = np.stack([synth.acg_cloud(pb["par"]["quat"][l], 1e-2, mReco, rng) for l in range(nImg)])
    h = np.sqrt(0.5)
    axis = np.array([[1, 0, 0, 0], [h, h, 0, 0], [h, 0, h, 0], [h, 0, 0, h], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1],
                     [0.5, 0.5, 0.5, 0.5]], float)
    nr[0, :8] = axis
    nr[1, :8] = axis[::-1]
    nr[2, 6:] = nr[2, :6]                  # duplicated rotations: merged groups
    nt = rng.normal(scale=2.0, size=(nImg, mReco, 2))
    w = (rng.uniform(0.5, 1.0, nImg) / mReco).astype(np.float32)
    offS = rng.normal(scale=0.5, size=(nImg, 2))
    ctx.set_insert_pixels(N, pf, pb["pixM"]["iColPad"], pb["pixM"]["iRowPad"])
    ctx.upload_stack(capi.STACK_INSERT, datM, ctfM, slotOfImg=pb["slot"])
    want = {}
    for s in (0, 1):
        ctx.reco_alloc(s, N * pf)
        sel = np.nonzero(pb["slot"] == s)[0]
        want[s] = port.insert_loop(N * pf, pf, N, datM[sel], ctfM[sel], w[sel], offS[sel], nr[sel], nt[sel],
                                   pb["pixM"]["iCol"], pb["pixM"]["iRow"])
    try:
        for impl, planes in ((0, 0), (0, 1), (0, 3), (0, 7), (0, 50), (3, 5), (1, 0)):
            ctx.set_option("insert_impl", impl)
            ctx.set_option("insert_slab_planes", planes)
            for s in (0, 1):
                ctx.reco_reset(s)
            ctx.insert(w, nr, nt, offS=offS)
            for s in (0, 1):
                got = ctx.reco_download(s)
                assert got["counter"] == want[s]["counter"], (impl, planes)
                assert np.allclose(got["O"], want[s]["O"], rtol=1e-11, atol=1e-11)
                # T is a sum of non-negative terms: a sample scattered twice or dropped shows up at full size
                assert _rel_l2(got["T"], want[s]["T"]) <= 1e-6, (impl, planes)
                assert _rel_l2(got["F"], want[s]["F"]) <= 1e-6, (impl, planes)
                assert abs(float(got["T"].sum(dtype=np.float64)) / float(want[s]["T"].sum(dtype=np.float64)) - 1) <= 1e-6
    finally:
        ctx.set_option("insert_impl", 0)
        ctx.set_option("insert_slab_planes", 0)


def test_reference_library_agrees(ctx, problem):
    """where the reference-built oracle travelled to this box: Projector / Reconstructor classes directly"""
    port, ref = _oracle()
    if ref is None:
        pytest.skip("oracle/_ref not present")
    pb = problem
    _setup_E(ctx, pb)
    P = ref.Projector(pb["pf"])
    P.set_padded_ft(pb["vols"][0])
    q = synth.random_quats(4, np.random.default_rng(2))
    got = ctx.project(0, q)
    for i in range(4):
        want = P.project(ref.rotate3D(q[i]), pb["pixE"]["iCol"], pb["pixE"]["iRow"])
        assert np.abs(got[i] - want).max() <= 2e-6 * np.abs(want).max()
    P.close()


def test_error_paths(ctx):
    c2 = capi.Context(0)
    with pytest.raises(capi.ThbError):
        c2.upload_stack(capi.STACK_EXPECT, np.zeros((1, 4), np.complex64), np.zeros((1, 4), np.float32), np.zeros((1, 4), np.float32))
    with pytest.raises(capi.ThbError):
        c2.reco_reset(3)
    with pytest.raises(capi.ThbError):
        c2.set_volume(99, np.zeros((4, 4, 3), np.complex64))
    c2.close()


# ------------------------------------------------------------------------------------------- TMA-staged kernel
@pytest.mark.parametrize("k,nR,nT", [(1e-7, 125, 9), (2e-5, 125, 9), (3e-4, 125, 9), (5e-2, 64, 9), (2e-5, 200, 9), (2e-5, 40, 13),
                                      (2e-5, 131, 20)])
def test_expect_kernels_agree_with_each_other_and_oracle(ctx, problem, k, nR, nT):
    """the three E kernels (quad-layout direct gather = default, TMA-staged box, linear direct gather) against each
    other and the oracle, from clouds that stay inside one staged box (k small) to clouds that spill into its
    L1/L2 path (k large), multi-pass shapes (nR > 128, nT > 9) included"""
    pb = problem
    port, ref = _oracle()
    _setup_E(ctx, pb)
    rng = np.random.default_rng(int(k * 1e9) + nR)
    nImg = pb["nImg"]
    quat = np.stack([synth.acg_cloud(pb["par"]["quat"][l], k, nR, rng) for l in range(nImg)])
    tran = pb["par"]["tran"][:, None, :] + rng.normal(scale=0.7, size=(nImg, nT, 2))
    wR = np.full((nImg, nR), 1.0 / nR); wT = np.full((nImg, nT), 1.0 / nT)
    try:
        ctx.set_option("expect_impl", 3)      # direct gather from the cell layout, one rotation per lane
        ctx.set_option("expect_spread", 0)    # one CTA per image (what a launch of thousands of images runs)
        a = ctx.expect_local(quat, tran, wR, wT)
        ctx.set_option("expect_spread", 1)    # the same image spread over (pixel chunk, rotation group) CTAs, double table
        s1 = ctx.expect_local(quat, tran, wR, wT)
        ctx.set_option("expect_spread", 0)
        ctx.set_option("quad_oct", 0)         # 32-byte quad layout, two occupancy variants
        a3 = ctx.expect_local(quat, tran, wR, wT)
        ctx.set_option("expect_minb", 3)
        a4 = ctx.expect_local(quat, tran, wR, wT)
        ctx.set_option("expect_minb", 2)
        ctx.set_option("quad_oct", 1)
        ctx.set_option("expect_impl", 4)      # two lanes per sample (half the L1 tag cycles), cell and quad layouts
        p4 = ctx.expect_local(quat, tran, wR, wT)
        ctx.set_option("quad_oct", 0)
        p4q = ctx.expect_local(quat, tran, wR, wT)
        ctx.set_option("quad_oct", 1)
        ctx.set_option("expect_impl", 5)      # pixels on the lanes (DRAM page locality), cell and quad layouts
        p5 = ctx.expect_local(quat, tran, wR, wT)
        ctx.set_option("quad_oct", 0)
        p5q = ctx.expect_local(quat, tran, wR, wT)
        ctx.set_option("quad_oct", 1)
        ctx.set_option("expect_impl", 7)      # several rotations per lane (record broadcast amortised), 2 and 4, both layouts
        ctx.set_option("expect_rpl", 2)
        m2 = ctx.expect_local(quat, tran, wR, wT)
        ctx.set_option("quad_oct", 0)
        m2q = ctx.expect_local(quat, tran, wR, wT)
        ctx.set_option("quad_oct", 1)
        ctx.set_option("expect_rpl", 4)
        m4 = ctx.expect_local(quat, tran, wR, wT)
        ctx.set_option("expect_rpl", 2)
        ctx.set_option("expect_impl", 2)      # TMA-staged shared-memory box
        c = ctx.expect_local(quat, tran, wR, wT)
        ctx.set_option("expect_impl", 1)      # direct gather, linear layout, unexpanded likelihood
        b = ctx.expect_local(quat, tran, wR, wT)
    finally:
        ctx.set_option("expect_impl", 0)      # back to the default kernel
        ctx.set_option("quad_oct", 1)
        ctx.set_option("expect_minb", 2)
        ctx.set_option("expect_spread", -1)
        ctx.set_option("expect_rpl", 2)
    # each kernel carries its own fp32 summation error (the linear-layout kernel sums ~3000 terms sequentially)
    tol = 2e-6 * np.abs(b["logL"]).max() + 1e-4
    assert np.array_equal(a["logL"], a3["logL"]) and np.array_equal(a["logL"], a4["logL"])   # layouts / occupancy: same bits
    assert np.abs(a["logL"] - b["logL"]).max() <= 2 * tol
    # spread kernel: same arithmetic per sample, the sum over pixels in another order (partial sums in double)
    assert np.abs(s1["logL"] - a["logL"]).max() <= tol
    assert np.abs(s1["base"] - a["base"]).max() <= tol
    assert np.allclose(s1["uT"], a["uT"], rtol=2e-3, atol=1e-6 * a["uT"].max())
    assert np.allclose(s1["uR"], a["uR"], rtol=5e-3, atol=1e-6 * a["uR"].max())
    # (kernel against kernel: two independent fp32 summation orders, each within tol of the exact sum)
    assert np.array_equal(p4["logL"], p4q["logL"])
    assert np.abs(p4["logL"] - a["logL"]).max() <= 2 * tol
    assert np.allclose(p4["uT"], a["uT"], rtol=2e-3, atol=1e-6 * a["uT"].max()) and np.abs(p4["base"] - a["base"]).max() <= 2 * tol
    assert np.array_equal(p5["logL"], p5q["logL"])
    assert np.abs(p5["logL"] - a["logL"]).max() <= 2 * tol
    assert np.allclose(p5["uT"], a["uT"], rtol=2e-3, atol=1e-6 * a["uT"].max()) and np.abs(p5["base"] - a["base"]).max() <= 2 * tol
    assert np.abs(c["logL"] - b["logL"]).max() <= 2 * tol
    assert np.abs(c["logL"] - a["logL"]).max() <= 2 * tol
    assert np.array_equal(m2["logL"], m2q["logL"])
    for m in (m2, m4):
        assert np.abs(m["logL"] - a["logL"]).max() <= 2 * tol
        assert np.allclose(m["uT"], a["uT"], rtol=2e-3, atol=1e-6 * a["uT"].max()) and np.abs(m["base"] - a["base"]).max() <= tol
        assert np.allclose(m["uR"], a["uR"], rtol=5e-3, atol=1e-6 * a["uR"].max())
    for l in (0, nImg - 1):
        o = port.expect_local(pb["vols"][pb["slot"][l]], pb["pf"], pb["N"], pb["pixE"]["iCol"], pb["pixE"]["iRow"],
                              pb["par"]["dat"][l], pb["par"]["ctf"][l], pb["par"]["sigRcp"][l], quat[l], tran[l], wR[l], wT[l])
        assert np.abs(a["logL"][l] - o["logL"]).max() <= tol
        assert np.abs(c["logL"][l] - o["logL"]).max() <= tol
    big = b["uR"] > 1e-6 * b["uR"].max(axis=1, keepdims=True)
    assert np.all(np.abs(np.log(a["uR"][big]) - np.log(b["uR"][big])) <= 20 * np.finfo(np.float32).eps * np.abs(b["logL"]).max() + 2e-3)


# ------------------------------------------------------------------------------------------- a2: packing on the device
def test_pack_stack_matches_allocPreCal(ctx):
    """thb_pack_stack == Optimiser::allocPreCal (image-major) + CTF(): gather by iPxl, sigma table by iSig, CTF on the fly"""
    port, ref = _oracle()
    N, pf = 64, 2
    rng = np.random.default_rng(31)
    pixE = port.pixel_list(N, pf, 30.0, 1.0)
    pixM = port.pixel_list(N, pf, 31.0, 0.0)
    nImg, nGroup, nRing = 5, 3, N // 2 + 1
    imgFT = (rng.normal(size=(nImg, N, N // 2 + 1)) + 1j * rng.normal(size=(nImg, N, N // 2 + 1))).astype(np.complex64)
    sigRcpTab = (-0.5 / rng.uniform(0.5, 2.0, (nGroup, nRing))).astype(np.float32)
    group = rng.integers(0, nGroup, nImg).astype(np.int32)
    attr = np.stack([np.full(nImg, 3e5), rng.uniform(1e4, 3e4, nImg), rng.uniform(1e4, 3e4, nImg), rng.uniform(0, np.pi, nImg),
                     np.full(nImg, 2.7e7), np.full(nImg, 0.1), rng.uniform(0, 0.3, nImg)], axis=1).astype(np.float32)
    slot = np.array([0, 1, 1, 0, 1], np.int32)
    ctx.set_expect_pixels(N, pf, pixE["iCol"], pixE["iRow"])
    ctx.set_insert_pixels(N, pf, pixM["iColPad"], pixM["iRowPad"])
    flat = imgFT.reshape(nImg, -1)
    for kind, pix in ((capi.STACK_EXPECT, pixE), (capi.STACK_INSERT, pixM)):
        ctx.stack_reserve(kind, nImg + 2)
        E = kind == capi.STACK_EXPECT
        ctx.pack_stack(kind, 1, imgFT, pix["iPxl"], attr, 1.32, iSig=pix["iSig"] if E else None, sigRcpTab=sigRcpTab if E else None,
                       groupOfImg=group if E else None, slotOfImg=slot)
        got = ctx.download_stack(kind, 1, nImg)
        assert np.array_equal(got["dat"], flat[:, pix["iPxl"]])                       # a gather: bit-exact
        if E:
            assert np.array_equal(got["sigRcp"], sigRcpTab[group][:, pix["iSig"]])
        for l in range(nImg):
            want = port.ctf(1.32, *[float(x) for x in attr[l]], N, pix["iCol"], pix["iRow"])
            # the phase reaches O(100) rad, one ulp of it is 3e-5 in the value: the packing kernel takes cos / sin through the
            # double-precision library and rounds once, like glibc's cosf / sinf on the reference's side - all but a handful of
            # pixels agree to the last bit or two of the result
            err = np.abs(got["ctf"][l] - want)
            assert err.max() <= 2e-5 and np.median(err) <= 1e-7 and np.mean(err > 5e-7) <= 1e-2, (err.max(), np.mean(err > 5e-7))
    # the packed E stack drives the kernel like an uploaded one
    ctx.set_volume(0, synth.padded_ft(synth.phantom(N, 6, seed=9), pf)); ctx.set_volume(1, synth.padded_ft(synth.phantom(N, 6, seed=8), pf))
    ctx.stack_reserve(capi.STACK_EXPECT, nImg)
    ctx.pack_stack(capi.STACK_EXPECT, 0, imgFT, pixE["iPxl"], attr, 1.32, iSig=pixE["iSig"], sigRcpTab=sigRcpTab, groupOfImg=group, slotOfImg=slot)
    packed = ctx.download_stack(capi.STACK_EXPECT, 0, nImg)
    q = synth.random_quats(7, rng)[None].repeat(nImg, 0); t = rng.normal(size=(nImg, 3, 2))
    a = ctx.expect_local(q, t, np.full((nImg, 7), 1 / 7), np.full((nImg, 3), 1 / 3))
    ctx.upload_stack(capi.STACK_EXPECT, packed["dat"], packed["ctf"], packed["sigRcp"], slot)
    b = ctx.expect_local(q, t, np.full((nImg, 7), 1 / 7), np.full((nImg, 3), 1 / 3))
    assert np.array_equal(a["logL"], b["logL"])


# ------------------------------------------------------------------------------------------- BASELINE size, box 256
def test_box256_expect_and_insert_against_oracle():
    """the benchmark configuration itself (box 256, pf 2, r = 127: 25 134 pixels, 125 x 9 samples, mReco 100) for a few
    images against the oracle, plus size-independent properties of the full-size insert (linearity, T >= 0, Hermitian
    consistency of the x = 0 plane is NOT enforced by insertP and therefore not asserted)"""
    port, ref = _oracle()
    N, pf = 256, 2
    rng = np.random.default_rng(2560)
    vol = synth.padded_ft(synth.phantom(N, 30), pf)
    pixE = port.pixel_list(N, pf, float(N // 2 - 1), float(np.floor(N * 1.32 / 200.0)))
    pixM = port.pixel_list(N, pf, float(N // 2 - 1), 0.0)
    assert len(pixE["iCol"]) == 25134 and len(pixM["iCol"]) == 25135
    c = capi.Context(0)
    try:
        c.set_expect_pixels(N, pf, pixE["iCol"], pixE["iRow"])
        c.set_volume(0, vol)
        nImg, nR, nT = 2, 125, 9
        par = synth.make_particles(nImg, N, pixE, lambda q: c.project(0, q), seed=11)
        c.upload_stack(capi.STACK_EXPECT, par["dat"], par["ctf"], par["sigRcp"])
        quat = np.stack([synth.acg_cloud(par["quat"][l], 2e-5, nR, rng) for l in range(nImg)])
        tran = par["tran"][:, None, :] + rng.normal(scale=0.7, size=(nImg, nT, 2))
        wR = np.full((nImg, nR), 1.0 / nR); wT = np.full((nImg, nT), 1.0 / nT)
        out = c.expect_local(quat, tran, wR, wT)             # two images: the kernel that spreads an image over the chip
        c.set_option("expect_spread", 0)
        out7 = c.expect_local(quat, tran, wR, wT)            # ... and the default kernel of large launches (two rotations per lane, lockstep)
        c.set_option("expect_spread", -1)
        for l in range(nImg):
            o = port.expect_local(vol, pf, N, pixE["iCol"], pixE["iRow"], par["dat"][l], par["ctf"][l], par["sigRcp"][l], quat[l], tran[l], wR[l], wT[l])
            assert np.abs(out["logL"][l] - o["logL"]).max() <= _logL_tol(len(pixE["iCol"]), o["logL"])
            assert np.abs(out7["logL"][l] - o["logL"]).max() <= _logL_tol(len(pixE["iCol"]), o["logL"])
        # slices at full size: project == oracle
        got = c.project(0, quat[0, :3])
        for i in range(3):
            want = port.project(vol, pf, port.rotate3D(quat[0, i]), pixE["iCol"], pixE["iRow"])
            assert np.abs(got[i] - want).max() <= 2e-6 * np.abs(want).max()
        if ref is not None:
            # ... and against THE REFERENCE'S OWN classes at this size (oracle/_ref travels to the GPU box): Projector::project,
            # translate, logDataVSPrior (the SIMD variant the reference's loops call) for a handful of (rotation, translation) pairs
            Pr = ref.Projector(pf)
            Pr.set_padded_ft(vol)
            for i in range(3):
                want = Pr.project(ref.rotate3D(quat[0, i]), pixE["iCol"], pixE["iRow"])
                assert np.abs(got[i] - want).max() <= 2e-6 * np.abs(want).max()
            for l, r_, t_ in ((0, 0, 0), (0, 57, 4), (1, 124, 8), (1, 3, 2)):
                pri = Pr.project(ref.rotate3D(quat[l, r_]), pixE["iCol"], pixE["iRow"])
                tra = ref.translate(tran[l, t_, 0], tran[l, t_, 1], N, pixE["iCol"], pixE["iRow"])
                want = ref.logDataVSPrior(par["dat"][l], (tra * pri).astype(np.complex64), par["ctf"][l], par["sigRcp"][l])
                assert abs(out["logL"][l, r_, t_] - want) <= _logL_tol(len(pixE["iCol"]), np.array([want]))
            Pr.close()
        del vol
        # M at full size
        PM = len(pixM["iCol"])
        mReco = 100
        datM = (rng.normal(size=(nImg, PM)) + 1j * rng.normal(size=(nImg, PM))).astype(np.complex64)
        ctfM = rng.uniform(-1, 1, (nImg, PM)).astype(np.float32)
        nr = np.stack([synth.acg_cloud(par["quat"][l], 2e-5, mReco, rng) for l in range(nImg)])
        nr[:, 50:] = nr[:, :50]                              # duplicated rotations: the merged-draw path
        nt = rng.normal(scale=2.0, size=(nImg, mReco, 2))
        w = np.full(nImg, 1.0 / mReco, np.float32)
        c.set_insert_pixels(N, pf, pixM["iColPad"], pixM["iRowPad"])
        c.upload_stack(capi.STACK_INSERT, datM, ctfM)
        c.reco_alloc(0, N * pf)
        c.insert(w, nr, nt)
        a = c.reco_download(0)
        want = port.insert_loop(N * pf, pf, N, datM, ctfM, w, np.zeros((nImg, 2)), nr, nt, pixM["iCol"], pixM["iRow"])
        assert a["counter"] == nImg * mReco
        assert _rel_l2(a["F"], want["F"]) <= 1e-6 and _rel_l2(a["T"], want["T"]) <= 1e-6
        assert a["T"].min() >= 0.0
        if ref is not None:                                    # the reference's Reconstructor::insertP / insertDir loop itself
            Rr = ref.Reconstructor(N, N, pf)
            Rr.set_precal(pixM["iColPad"], pixM["iRowPad"], pixM["iPxl"], pixM["iSig"])
            Rr.insert_loop(datM, ctfM, w, np.zeros((nImg, 2)), nr, nt, pixM["iCol"], pixM["iRow"], N, nThread=8)
            wr = Rr.get()
            Rr.close()
            assert _rel_l2(a["F"], wr["F"]) <= 1e-6 and _rel_l2(a["T"], wr["T"]) <= 1e-6 and a["counter"] == wr["counter"]
            del wr
        c.set_option("insert_impl", 2)                         # draw-by-draw insertion gives the same volume
        c.reco_reset(0)
        c.insert(w, nr, nt)
        b = c.reco_download(0)
        c.set_option("insert_impl", 0)
        assert _rel_l2(b["F"], a["F"]) <= 1e-6
        c.insert(w, nr, nt)                                    # linearity: twice the list = twice the volume
        b2 = c.reco_download(0)
        assert _rel_l2(b2["F"], 2 * a["F"]) <= 1e-6 and b2["counter"] == 2 * a["counter"]
    finally:
        c.close()


def test_edge_cases(ctx, problem):
    """single pixel tile remainders, one rotation / one translation, a pixel list with the DC pixel (rL = 0), identity and
    axis-aligned rotations (integer coordinates, samples exactly on the Hermitian fold x = 0)"""
    pb = problem
    port, ref = _oracle()
    N, pf = pb["N"], pb["pf"]
    pix0 = port.pixel_list(N, pf, 9.5, 0.0)                   # includes (0,0); 0 <= i
    ctx.set_expect_pixels(N, pf, pix0["iCol"], pix0["iRow"])
    ctx.set_volume(0, pb["vols"][0]); ctx.set_volume(1, pb["vols"][1])
    rng = np.random.default_rng(4)
    P = len(pix0["iCol"])
    dat = (rng.normal(size=(1, P)) + 1j * rng.normal(size=(1, P))).astype(np.complex64)
    ctf = rng.uniform(-1, 1, (1, P)).astype(np.float32); sig = np.full((1, P), -0.5, np.float32)
    ctx.upload_stack(capi.STACK_EXPECT, dat, ctf, sig)
    s2 = np.sqrt(0.5)
    quat = np.array([[[1, 0, 0, 0], [s2, 0, 0, s2], [s2, s2, 0, 0], [s2, 0, s2, 0], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1], [0.5, 0.5, 0.5, 0.5]]], float)
    tran = np.zeros((1, 1, 2))
    out = ctx.expect_local(quat, tran, np.full((1, 8), 1 / 8), np.ones((1, 1)))
    o = port.expect_local(pb["vols"][0], pf, N, pix0["iCol"], pix0["iRow"], dat[0], ctf[0], sig[0], quat[0], tran[0], np.full(8, 1 / 8), np.ones(1))
    assert np.abs(out["logL"][0] - o["logL"]).max() <= 2e-6 * np.abs(o["logL"]).max() + 1e-4
    got = ctx.project(0, quat[0])
    for i in range(8):
        want = port.project(pb["vols"][0], pf, port.rotate3D(quat[0, i]), pix0["iCol"], pix0["iRow"])
        assert np.abs(got[i] - want).max() <= 2e-6 * np.abs(want).max(), i
    with pytest.raises(capi.ThbError):                         # empty inputs are refused, not silently accepted
        ctx.expect_local(np.zeros((0, 1, 4)), np.zeros((0, 1, 2)), np.zeros((0, 1)), np.zeros((0, 1)))
    with pytest.raises(capi.ThbError):
        ctx.expect_local(quat, tran, np.full((1, 8), 1 / 8), np.ones((1, 1)), imgIdx=np.array([5], np.int32))


# ------------------------------------------------------------------------------------------- other BASELINE shapes
def test_box128_config1_shape_against_oracle():
    """BASELINE config 1 (demo_3D defaults: box 128, r = 63 -> 6 141 pixels, 25 rotations x 9 translations per phase)"""
    port, ref = _oracle()
    N, pf = 128, 2
    rng = np.random.default_rng(128)
    vol = synth.padded_ft(synth.phantom(N, 20), pf)
    pixE = port.pixel_list(N, pf, 63.0, float(np.floor(N * 1.32 / 200.0)))
    assert len(pixE["iCol"]) == 6141
    c = capi.Context(0)
    try:
        c.set_expect_pixels(N, pf, pixE["iCol"], pixE["iRow"])
        c.set_volume(0, vol)
        nImg, nR, nT = 4, 25, 9
        par = synth.make_particles(nImg, N, pixE, lambda q: c.project(0, q), seed=12)
        c.upload_stack(capi.STACK_EXPECT, par["dat"], par["ctf"], par["sigRcp"])
        quat = np.stack([synth.acg_cloud(par["quat"][l], 1e-4, nR, rng) for l in range(nImg)])
        tran = par["tran"][:, None, :] + rng.normal(scale=0.7, size=(nImg, nT, 2))
        wR = rng.uniform(0.5, 1.5, (nImg, nR)); wR /= wR.sum(1, keepdims=True)
        wT = rng.uniform(0.5, 1.5, (nImg, nT)); wT /= wT.sum(1, keepdims=True)
        out = c.expect_local(quat, tran, wR, wT)
        for l in range(nImg):
            o = port.expect_local(vol, pf, N, pixE["iCol"], pixE["iRow"], par["dat"][l], par["ctf"][l], par["sigRcp"][l], quat[l], tran[l], wR[l], wT[l])
            tol = _logL_tol(len(pixE["iCol"]), o["logL"])
            assert np.abs(out["logL"][l] - o["logL"]).max() <= tol
            big = o["uR"] > 1e-6 * o["uR"].max()
            assert np.all(np.abs(np.log(out["uR"][l][big]) - np.log(o["uR"][big])) <= 2 * tol + 2e-3)
    finally:
        c.close()


def test_box512_config4_offsets_beyond_4GB():
    """BASELINE config 4 geometry (box 512, pf 2: 1024^3 padded volume = 4.3 GB linear + 17 GB quad layout, 8.6 GB
    accumulator, 101 726 pixels): every 64-bit offset of the gather and of the scatter, against the oracle for one image.
    The volume is random Fourier data (no 1024^3 FFT on the host); parity does not depend on its meaning."""
    port, ref = _oracle()
    import torch
    free, total = torch.cuda.mem_get_info(0)
    if free < 60 << 30:
        pytest.skip("needs ~45 GB of HBM")
    N, pf = 512, 2
    n = N * pf
    rng = np.random.default_rng(512)
    vol = np.empty((n, n, n // 2 + 1), np.complex64)
    for z in range(0, n, 64):                                 # chunked: keeps the float64 temporaries small
        blk = rng.standard_normal((64, n, n // 2 + 1, 2), dtype=np.float32)
        vol[z:z + 64] = blk[..., 0] + 1j * blk[..., 1]
    pixE = port.pixel_list(N, pf, float(N // 2 - 1), float(np.floor(N * 1.32 / 200.0)))
    pixM = port.pixel_list(N, pf, float(N // 2 - 1), 0.0)
    assert len(pixE["iCol"]) == 101726
    c = capi.Context(0)
    try:
        c.set_expect_pixels(N, pf, pixE["iCol"], pixE["iRow"])
        c.set_volume(0, vol)
        nR, nT = 12, 3
        P = len(pixE["iCol"])
        dat = (rng.normal(size=(1, P)) + 1j * rng.normal(size=(1, P))).astype(np.complex64)
        ctf = rng.uniform(-1, 1, (1, P)).astype(np.float32); sig = np.full((1, P), -0.5e-3, np.float32)
        c.upload_stack(capi.STACK_EXPECT, dat, ctf, sig)
        quat = synth.random_quats(nR, rng)[None]               # random orientations: slices through the whole volume
        tran = rng.normal(scale=2.0, size=(1, nT, 2))
        wR = np.full((1, nR), 1.0 / nR); wT = np.full((1, nT), 1.0 / nT)
        out = c.expect_local(quat, tran, wR, wT)
        o = port.expect_local(vol, pf, N, pixE["iCol"], pixE["iRow"], dat[0], ctf[0], sig[0], quat[0], tran[0], wR[0], wT[0])
        assert np.abs(out["logL"][0] - o["logL"]).max() <= _logL_tol(P, o["logL"])
        # the default kernel of large launches (two rotations per lane; > 64 rotations, one CTA per image) on the 34 GB cell layout
        quat7 = synth.random_quats(70, rng)[None]; wR7 = np.full((1, 70), 1.0 / 70)
        c.set_option("expect_spread", 0)
        out7 = c.expect_local(quat7, tran, wR7, wT)
        c.set_option("expect_spread", -1)
        o7 = port.expect_local(vol, pf, N, pixE["iCol"], pixE["iRow"], dat[0], ctf[0], sig[0], quat7[0], tran[0], wR7[0], wT[0])
        assert np.abs(out7["logL"][0] - o7["logL"]).max() <= _logL_tol(P, o7["logL"])
        got = c.project(0, quat[0, :2])
        for i in range(2):
            want = port.project(vol, pf, port.rotate3D(quat[0, i]), pixE["iCol"], pixE["iRow"])
            assert np.abs(got[i] - want).max() <= 2e-6 * np.abs(want).max()
        del vol
        PM = len(pixM["iCol"])
        mReco = 6
        datM = (rng.normal(size=(1, PM)) + 1j * rng.normal(size=(1, PM))).astype(np.complex64)
        ctfM = rng.uniform(-1, 1, (1, PM)).astype(np.float32)
        nr = synth.random_quats(mReco, rng)[None]; nt = rng.normal(scale=2.0, size=(1, mReco, 2))
        w = np.full(1, 1.0 / mReco, np.float32)
        c.set_insert_pixels(N, pf, pixM["iColPad"], pixM["iRowPad"])
        c.upload_stack(capi.STACK_INSERT, datM, ctfM)
        c.reco_alloc(0, n)
        c.insert(w, nr, nt)
        a = c.reco_download(0)
        want = port.insert_loop(n, pf, N, datM, ctfM, w, np.zeros((1, 2)), nr, nt, pixM["iCol"], pixM["iRow"])
        assert a["counter"] == mReco
        assert _rel_l2(a["F"], want["F"]) <= 1e-6 and _rel_l2(a["T"], want["T"]) <= 1e-6
    finally:
        c.close()


def test_async_upload_matches_blocking_upload(ctx):
    """thb_upload_stack_at_async + thb_upload_wait (second stream, overlaps kernels) fills the resident stack exactly like
    thb_upload_stack_at"""
    import torch
    port, ref = _oracle()
    N, pf = 64, 2
    rng = np.random.default_rng(5)
    pixE = port.pixel_list(N, pf, 30.0, 1.0)
    P, nImg = len(pixE["iCol"]), 24
    ctx.set_expect_pixels(N, pf, pixE["iCol"], pixE["iRow"])
    ctx.set_volume(0, synth.padded_ft(synth.phantom(N, 6, seed=9), pf)); ctx.set_volume(1, synth.padded_ft(synth.phantom(N, 6, seed=8), pf))
    pin = lambda a: torch.from_numpy(a).pin_memory()
    tens = [pin((rng.normal(size=(nImg, P)) + 1j * rng.normal(size=(nImg, P))).astype(np.complex64)),
            pin(rng.uniform(-1, 1, (nImg, P)).astype(np.float32)), pin((-0.5 / rng.uniform(0.5, 2, (nImg, P))).astype(np.float32))]
    dat, ctf, sig = [t.numpy() for t in tens]
    slot = (np.arange(nImg) % 2).astype(np.int32)
    ctx.stack_reserve(capi.STACK_EXPECT, 2 * nImg)
    ctx.upload_stack_at(capi.STACK_EXPECT, 0, dat, ctf, sig, slot)
    q = synth.random_quats(5, rng)[None].repeat(nImg, 0); t = rng.normal(size=(nImg, 3, 2))
    wR, wT = np.full((nImg, 5), 0.2), np.full((nImg, 3), 1 / 3)
    ctx.upload_stack_at_async(capi.STACK_EXPECT, nImg, dat, ctf, sig, slot)       # in flight while the kernel below runs
    a = ctx.expect_local(q, t, wR, wT)
    ctx.upload_wait()
    b = ctx.expect_local(q, t, wR, wT, imgIdx=np.arange(nImg, 2 * nImg, dtype=np.int32))
    assert np.array_equal(a["logL"], b["logL"])
    got = ctx.download_stack(capi.STACK_EXPECT, nImg, nImg)
    assert np.array_equal(got["dat"], dat) and np.array_equal(got["ctf"], ctf) and np.array_equal(got["sigRcp"], sig)
    ctx.upload_wait()                                                             # nothing pending: a no-op


def test_insert_counts_3d_classification():
    """thb_insert_counts (the nC of the reference's InsertFT in 3D classification): only the first nDraw[l] rows of image l are
    inserted == inserting every image with its own truncated list"""
    port, ref = _oracle()
    N, pf = 32, 2
    rng = np.random.default_rng(17)
    pixM = port.pixel_list(N, pf, 15.0, 0.0)
    PM = len(pixM["iCol"])
    nImg, mReco = 6, 5
    datM = (rng.normal(size=(nImg, PM)) + 1j * rng.normal(size=(nImg, PM))).astype(np.complex64)
    ctfM = rng.uniform(-1, 1, (nImg, PM)).astype(np.float32)
    nr = synth.random_quats(nImg * mReco, rng).reshape(nImg, mReco, 4); nt = rng.normal(scale=2.0, size=(nImg, mReco, 2))
    w = np.full(nImg, 1.0 / mReco, np.float32); offS = rng.normal(scale=0.5, size=(nImg, 2))
    nDraw = np.array([5, 0, 3, 1, 7, 2], np.int32)                 # 7 > mReco: clipped to mReco
    ctx = capi.Context(0)                                           # private context: its only accumulator is slot 0
    try:
        ctx.set_insert_pixels(N, pf, pixM["iColPad"], pixM["iRowPad"])
        ctx.upload_stack(capi.STACK_INSERT, datM, ctfM)
        ctx.reco_alloc(0, N * pf)
        ctx.insert_counts(w, nDraw, nr, nt, offS=offS)
        got = ctx.reco_download(0)
        ctx.reco_reset(0)
        for l in range(nImg):
            c = min(int(nDraw[l]), mReco)
            if c:
                ctx.insert(w[l:l + 1], nr[l:l + 1, :c], nt[l:l + 1, :c], offS=offS[l:l + 1], imgIdx=np.array([l], np.int32))
        want = ctx.reco_download(0)
    finally:
        ctx.close()
    assert got["counter"] == want["counter"] == int(np.minimum(nDraw, mReco).sum())
    assert np.allclose(got["O"], want["O"], rtol=1e-12, atol=1e-12)
    assert np.abs(got["F"] - want["F"]).max() <= 2e-6 * np.abs(want["F"]).max()
    assert np.abs(got["T"] - want["T"]).max() <= 2e-6 * np.abs(want["T"]).max()

"""CTF search (SEARCH_TYPE_CTF: the defocus dimension of the E-step and the per-draw CTF of the M-step) on the device against the
reference's own arithmetic (oracle/_ref: ref_precal_ctf / ref_expect_ctf / ref_insert_loop_ctf restate src/Optimiser.cpp:8125-8168,
1236-1402 and 7067-7241 around the reference's Projector, translate, logDataVSPrior, CTF and Reconstructor).
Tolerances: log-likelihoods 5e-6 |logL| + 1e-4 (the on-the-fly CTF goes through sinf / cosf of a phase of ~100 rad on both sides);
marginal weights 5e-3 relative on the entries that carry weight; F / T volumes relative L2 <= 2e-5."""
import numpy as np
import pytest

from thunder_b200 import capi, synth

pytestmark = pytest.mark.gpu


def _ref():
    from oracle import refapi
    if not refapi.available():
        pytest.skip("oracle/_ref not present")
    return refapi


def test_expect_local_with_defocus_dimension():
    ref = _ref()
    from oracle import portapi as port
    N, pf = 32, 2
    rng = np.random.default_rng(404)
    vol = synth.padded_ft(synth.phantom(N, 8, seed=2), pf)
    pix = port.pixel_list(N, pf, 14.0, 1.0)
    iCol, iRow = pix["iCol"], pix["iRow"]
    P = len(iCol)
    nImg, nR, nT, nD = 3, 21, 9, 5
    proj = ref.Projector(pf)
    proj.set_padded_ft(vol)
    par = synth.make_particles(nImg, N, pix, lambda q: np.stack([proj.project(ref.rotate3D(x), iCol, iRow) for x in q]), seed=3, snr_scale=4.0)
    volt, Cs, ac, ps, pixelSize = 3.0e5, 2.7e7, 0.1, 0.3, 1.32
    freq = None
    defP = np.zeros((nImg, P), np.float32); ctfK = np.zeros((nImg, 4), np.float32)
    for l in range(nImg):
        dU, dV, th = par["ctfpar"][l]
        f, dp, k = ref.precal_ctf(volt, dU, dV, th, Cs, N, pixelSize, iCol, iRow)
        freq = f
        defP[l] = dp
        ctfK[l] = (k[0], k[1], ps, ac)
    quat = np.stack([synth.acg_cloud(par["quat"][l], 3e-4, nR, rng) for l in range(nImg)])
    tran = par["tran"][:, None, :] + rng.normal(scale=0.6, size=(nImg, nT, 2))
    dpar = 1.0 + rng.normal(scale=0.02, size=(nImg, nD))
    nrm = lambda a: a / a.sum(-1, keepdims=True)
    wR = nrm(rng.uniform(0.5, 1.5, (nImg, nR))); wT = nrm(rng.uniform(0.5, 1.5, (nImg, nT))); wD = nrm(rng.uniform(0.5, 1.5, (nImg, nD)))
    c = capi.Context(0)
    try:
        c.set_expect_pixels(N, pf, iCol, iRow)
        c.set_volume(0, vol)
        c.upload_stack(capi.STACK_EXPECT, par["dat"], par["ctf"], par["sigRcp"])
        c.set_frequency(freq)
        c.upload_stack_defocus(0, defP)
        out = c.expect_local_ctf(quat, tran, dpar, wR, wT, wD, ctfK)
        # the defocus factor 1 with the reference's OTHER wavelength constant (12.2643274 in allocPreCal, 12.2643247 in CTF()):
        # the search's CTF at d = 1 is the stack's CTF to 3e-6 of its amplitude
        one = c.expect_local_ctf(quat, tran, np.ones((nImg, 1)), wR, wT, np.ones((nImg, 1)), ctfK)
    finally:
        c.close()
    for l in range(nImg):
        o = ref.expect_ctf(proj, par["dat"][l], par["sigRcp"][l], defP[l], freq, ctfK[l, 0], ctfK[l, 1], ps, ac, quat[l], tran[l], dpar[l], wR[l],
                           wT[l], wD[l], 1.0, iCol, iRow, N)
        L = o["logL"]
        assert np.abs(out["logL"][l] - L).max() <= 5e-6 * np.abs(L).max() + 1e-4, np.abs(out["logL"][l] - L).max()
        assert abs(out["base"][l] - L.max()) <= 5e-6 * np.abs(L).max() + 1e-4
        for key in ("uR", "uT", "uD"):
            big = o[key] > 1e-4 * o[key].max()
            assert np.allclose(out[key][l][big], o[key][big], rtol=5e-3), key
        assert np.isclose(out["uC"][l], o["uC"], rtol=5e-3)
        # size-independent property: the marginals of one table sum to the same total
        assert np.isclose((out["uR"][l] * wR[l]).sum(), out["uC"][l], rtol=1e-4)
        assert np.isclose((out["uD"][l] * wD[l]).sum(), out["uC"][l], rtol=1e-4)
    proj.close()


def test_insert_with_per_draw_ctf():
    ref = _ref()
    from oracle import portapi as port
    N, pf = 32, 2
    rng = np.random.default_rng(405)
    pixM = port.pixel_list(N, pf, 15.0, 0.0)
    PM = len(pixM["iCol"])
    nImg, mReco = 4, 9
    datM = (rng.normal(size=(nImg, PM)) + 1j * rng.normal(size=(nImg, PM))).astype(np.complex64)
    nr = synth.random_quats(nImg * mReco, rng).reshape(nImg, mReco, 4)
    nr[:, 5:] = nr[:, :4]                                   # duplicated rotations with different defocus factors: merged groups
    nt = rng.normal(scale=2.0, size=(nImg, mReco, 2))
    nd = 1.0 + rng.normal(scale=0.03, size=(nImg, mReco))
    w = (rng.uniform(0.5, 1.0, nImg) / mReco).astype(np.float32)
    offS = rng.normal(scale=0.5, size=(nImg, 2))
    attr = np.stack([[3.0e5, rng.uniform(1e4, 3e4), 0, rng.uniform(0, np.pi), 2.7e7, 0.1, 0.2] for _ in range(nImg)]).astype(np.float32)
    attr[:, 2] = attr[:, 1] + rng.uniform(0, 500, nImg).astype(np.float32)
    pixelSize = 1.32
    reco = ref.Reconstructor(N, N, pf, 1)
    reco.set_precal(pixM["iColPad"], pixM["iRowPad"], pixM["iPxl"], pixM["iSig"])
    reco.insert_loop_ctf(datM, w, offS, nr, nt, nd, attr, pixelSize, pixM["iCol"], pixM["iRow"], N)
    want = reco.get()
    reco.close()
    c = capi.Context(0)
    try:
        c.set_insert_pixels(N, pf, pixM["iColPad"], pixM["iRowPad"])
        c.upload_stack(capi.STACK_INSERT, datM, np.zeros((nImg, PM), np.float32))      # the resident ctf array is not used by the search
        c.reco_alloc(0, N * pf)
        for planes in (0, 5):
            c.set_option("insert_slab_planes", planes)
            c.reco_reset(0)
            c.insert_ctf(w, nr, nt, nd, attr, pixelSize, offS=offS)
            got = c.reco_download(0)
            rel = lambda a, b: np.linalg.norm((a - b).ravel()) / np.linalg.norm(b.ravel())
            assert got["counter"] == want["counter"] == nImg * mReco
            assert rel(got["F"], want["F"]) <= 2e-5 and rel(got["T"], want["T"]) <= 2e-5, (rel(got["F"], want["F"]), rel(got["T"], want["T"]))
            assert np.allclose(got["O"], want["O"], rtol=1e-10, atol=1e-10)
    finally:
        c.close()


def test_ctf_search_through_the_device_particle_filter():
    """SEARCH_TYPE_CTF end to end on the device: thb_expectation with mLD = 9 (initD, the defocus dimension of the fused kernel, perturb /
    calVari / resample of PAR_D) recovers a 3 % defocus error of the nominal CTF parameters, and thb_reconstruct_insert (one CTF per
    draw, from the draw's own defocus factor) equals thb_insert_ctf fed with the draws it made.  The operators of the defocus dimension
    are pinned to the reference's Particle class draw by draw on the CPU (tests/test_pf_host.py::test_defocus_dimension_replay_exact)."""
    ref = _ref()
    from oracle import portapi as port
    N, pf = 64, 2
    rng = np.random.default_rng(606)
    vol = synth.padded_ft(synth.phantom(N, 14, seed=3), pf)
    pixE = port.pixel_list(N, pf, 28.0, 1.0)
    pixM = port.pixel_list(N, pf, 30.0, 0.0)
    PE, PM = len(pixE["iCol"]), len(pixM["iCol"])
    n, mLR, mLT, mLD, phases, mReco = 64, 125, 9, 9, 6, 20
    volt, Cs, ac, ps, pixelSize = 3.0e5, 2.7e7, 0.1, 0.0, 1.32
    d_true = 1.03
    c = capi.Context(0)
    try:
        c.set_expect_pixels(N, pf, pixE["iCol"], pixE["iRow"])
        c.set_insert_pixels(N, pf, pixM["iColPad"], pixM["iRowPad"])
        c.set_volume(0, vol)
        c.reco_alloc(0, N * pf)
        quat = synth.random_quats(n, rng)
        tran = rng.normal(scale=1.0, size=(n, 2))
        dU = rng.uniform(1.5e4, 2.5e4, n); dV = dU + rng.uniform(0, 300, n); th = rng.uniform(0, np.pi, n)
        attr = np.stack([np.full(n, volt), dU, dV, th, np.full(n, Cs), np.full(n, ac), np.full(n, ps)], axis=1).astype(np.float32)

        def simulate(pix):
            c.set_expect_pixels(N, pf, pix["iCol"], pix["iRow"])
            clean = c.project(0, quat)
            ctf = np.stack([synth.ctf_values(pix["iCol"].astype(float), pix["iRow"].astype(float), N, pixelSize, volt, dU[l] * d_true, dV[l] * d_true,
                                             th[l], Cs, ac, ps) for l in range(n)]).astype(np.float32)
            phs = -2 * np.pi * (pix["iCol"][None] * tran[:, :1] / N + pix["iRow"][None] * tran[:, 1:] / N)
            sig2 = float(np.mean(np.abs(clean * ctf) ** 2)) / 2.0
            noise = (rng.normal(size=clean.shape) + 1j * rng.normal(size=clean.shape)) * np.sqrt(sig2 / 2)
            return (ctf * clean * np.exp(1j * phs) + noise).astype(np.complex64), ctf, sig2
        datM, ctfM, _ = simulate(pixM)
        datE, ctfE, sig2 = simulate(pixE)
        c.set_expect_pixels(N, pf, pixE["iCol"], pixE["iRow"])
        c.upload_stack(capi.STACK_EXPECT, datE, ctfE, np.full((n, PE), -0.5 / sig2, np.float32))
        c.upload_stack(capi.STACK_INSERT, datM, ctfM)
        freq = None
        defP = np.zeros((n, PE), np.float32); ctfK = np.zeros((n, 4), np.float32)
        for l in range(n):
            f, dp, k = ref.precal_ctf(volt, dU[l], dV[l], th[l], Cs, N, pixelSize, pixE["iCol"], pixE["iRow"])     # NOMINAL parameters
            freq = f; defP[l] = dp; ctfK[l] = (k[0], k[1], ps, ac)
        c.set_frequency(freq)
        c.upload_stack_defocus(0, defP)
        prm = capi.PFParams(mLR=mLR, mLT=mLT, transS=2.0, transQ=0.01, perturbFactorL=2.0, perturbFactorS=0.5, minPhase=3, maxPhase=100,
                            fixedPhases=phases, decreaseFactor=0.95, noDecreaseLimit=1, seed=99, mLD=mLD, ctfRefineS=0.02, perturbFactorSCTF=0.5)
        k0 = 1e-4
        q_start = np.stack([synth.acg_cloud(quat[l], k0, 1, rng)[0] for l in range(n)])
        c.pf_set_image_base(0, 0)
        c.pf_load(prm, q_start, np.full((n, 3), k0), tran + rng.normal(scale=0.3, size=(n, 2)), np.full((n, 2), 0.5))
        c.pf_set_ctf(ctfK, attr, pixelSize)
        c.expectation()
        D = c.pf_get_d()
        print(f"\nCTF search: true defocus factor {d_true}; most likely factor after {phases} phases: median {np.median(D['topD']):.4f}, "
              f"quartiles {np.percentile(D['topD'], 25):.4f} .. {np.percentile(D['topD'], 75):.4f}; sigma of the support median {np.median(D['sD']):.4f}")
        assert abs(np.median(D["topD"]) - d_true) < 0.01 and np.median(np.abs(D["topD"] - d_true)) < 0.012
        assert np.allclose(D["wD"].sum(1), 1.0)
        # M: per-draw CTF through the particle filter == explicit lists
        c.reconstruct_insert(mReco)
        a = c.reco_download(0)
        st = c.pf_get()
        dR, dT = c.pf_get_draws(mReco); dD = c.pf_get_draws_d(mReco)
        rows = np.arange(n)[:, None]
        c.reco_reset(0)
        c.insert_ctf(np.full(n, 1.0 / mReco, np.float32), st["r"][rows, dR], st["t"][rows, dT], D["d"][rows, dD], attr, pixelSize)
        b = c.reco_download(0)
        rel = lambda x, y: np.linalg.norm((x - y).ravel()) / np.linalg.norm(y.ravel())
        assert a["counter"] == b["counter"] == n * mReco
        assert rel(a["F"], b["F"]) <= 1e-6 and rel(a["T"], b["T"]) <= 1e-6
        # ... and differs from the insert with the nominal CTF (the search does change the volume)
        c.reco_reset(0)
        c.insert(np.full(n, 1.0 / mReco, np.float32), st["r"][rows, dR], st["t"][rows, dT])
        nominal = c.reco_download(0)
        assert rel(a["F"], nominal["F"]) > 1e-3
    finally:
        c.close()

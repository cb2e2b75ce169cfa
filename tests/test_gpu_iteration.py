"""Iteration-level GPU tests: device particle filter + fused E kernel (thb_expectation) and the
particle-filter-driven insert (thb_reconstruct_insert), against the reference's own loops
(oracle/_ref: Particle + Projector + logDataVSPrior + Reconstructor driven by ref_harness.cpp).

Random streams differ by design (Philox vs the reference's mt19937), so E-step parity is statistical:
orientation / translation accuracy and the FSC between the back-projected volumes.  The plumbing
(particle state -> kernels) is checked exactly.
"""
import numpy as np
import pytest

from thunder_b200 import capi, synth

pytestmark = pytest.mark.gpu


def _ang_deg(q, q0):
    d = np.abs(np.sum(q * q0, axis=-1)).clip(0, 1)
    return np.degrees(2 * np.arccos(d))


def _params(mLR=125, mLT=9, fixed=0, seed=7):
    return capi.PFParams(mLR=mLR, mLT=mLT, transS=2.0, transQ=0.01, perturbFactorL=2.0, perturbFactorS=0.5, minPhase=3,
                         maxPhase=100, fixedPhases=fixed, decreaseFactor=0.95, noDecreaseLimit=1, seed=seed)


@pytest.fixture(scope="module")
def prob():
    from oracle import portapi as port
    N, pf = 64, 2
    rng = np.random.default_rng(123)
    vols = [synth.padded_ft(synth.phantom(N, 14, seed=s), pf) for s in (3, 4)]
    pixE = port.pixel_list(N, pf, 28.0, 1.0)
    pixM = port.pixel_list(N, pf, 30.0, 0.0)
    nImg = 48
    slot = (np.arange(nImg) % 2).astype(np.int32)

    def project_fn(quats, pix=pixE):
        return np.stack([port.project(vols[slot[l]], pf, port.rotate3D(q), pix["iCol"], pix["iRow"]) for l, q in enumerate(quats)])
    par = synth.make_particles(nImg, N, pixE, project_fn, seed=9, snr_scale=40.0)
    # unmasked stack on the M pixel set: same particles, same noise model
    rngM = np.random.default_rng(10)
    cleanM = project_fn(par["quat"], pixM)
    PM = len(pixM["iCol"])
    ctfM = np.stack([synth.ctf_values(pixM["iCol"].astype(float), pixM["iRow"].astype(float), N, 1.32, 3e5, *par["ctfpar"][l], 2.7e7, 0.1)
                     for l in range(nImg)])
    ph = -2 * np.pi * (pixM["iCol"][None] * par["tran"][:, :1] / N + pixM["iRow"][None] * par["tran"][:, 1:] / N)
    datM = (ctfM * cleanM * np.exp(1j * ph) + (rngM.normal(size=(nImg, PM)) + 1j * rngM.normal(size=(nImg, PM))) * np.sqrt(par["sig2"] / 2)).astype(np.complex64)
    # starting guesses: truth disturbed by ~2 degrees per axis and ~0.5 pixel; per-pixel SNR 2 so that the likelihood
    # peak (1 voxel at r = 28 is 2 degrees) is well defined - at SNR 0.3 both filters, ours and the reference's, drift
    k0 = 3e-4
    q_start = np.stack([synth.acg_cloud(par["quat"][l], k0, 1, rng)[0] for l in range(nImg)])
    t_start = par["tran"] + rng.normal(scale=0.5, size=(nImg, 2))
    return dict(N=N, pf=pf, vols=vols, pixE=pixE, pixM=pixM, nImg=nImg, slot=slot, par=par, datM=datM, ctfM=ctfM.astype(np.float32),
                q_start=q_start, t_start=t_start, k0=k0)


def _setup(ctx, pb):
    ctx.set_expect_pixels(pb["N"], pb["pf"], pb["pixE"]["iCol"], pb["pixE"]["iRow"])
    ctx.set_insert_pixels(pb["N"], pb["pf"], pb["pixM"]["iColPad"], pb["pixM"]["iRowPad"])
    for s, v in enumerate(pb["vols"]):
        ctx.set_volume(s, v)
        ctx.reco_alloc(s, pb["N"] * pb["pf"])
    ctx.upload_stack(capi.STACK_EXPECT, pb["par"]["dat"], pb["par"]["ctf"], pb["par"]["sigRcp"], pb["slot"])
    ctx.upload_stack(capi.STACK_INSERT, pb["datM"], pb["ctfM"], slotOfImg=pb["slot"])
    ctx.pf_set_image_base(0, 0)


def _load(ctx, pb, prm):
    n = pb["nImg"]
    ctx.pf_load(prm, pb["q_start"], np.full((n, 3), pb["k0"]), pb["t_start"], np.full((n, 2), 1.0))


def test_pf_state_roundtrip_and_device_ops(ctx, prob):
    """pf_get / pf_set are inverse; deterministic operators on the device equal the host build of the same source"""
    pb = prob
    _setup(ctx, pb)
    _load(ctx, pb, _params())
    st = ctx.pf_get()
    assert np.allclose(np.linalg.norm(st["r"], axis=2), 1.0, atol=1e-12)
    assert np.allclose(st["wR"].sum(1), 1.0) and np.allclose(st["wT"].sum(1), 1.0)
    assert np.median(_ang_deg(st["r"][:, 0], pb["q_start"])) < 6.0
    ctx.pf_set(r=st["r"], t=st["t"], wR=st["wR"], wT=st["wT"], scal=st["scal"])
    st2 = ctx.pf_get()
    for k in st:
        assert np.array_equal(st[k], st2[k]), k
    # balanceWeight on the device == numpy restatement of 1/pdf (bivariate Gaussian, sample sd)
    ctx.pf_op(capi.PF_BALANCE_T)
    t = st["t"]
    m = t.mean(1, keepdims=True); sd = t.std(1, ddof=1, keepdims=True)
    w = 1.0 / (np.exp(-(((t - m) / sd) ** 2).sum(2) / 2) / (2 * np.pi * sd[:, 0, 0:1] * sd[:, 0, 1:2]))
    w /= w.sum(1, keepdims=True)
    assert np.allclose(ctx.pf_get()["wT"], w, rtol=1e-10)
    # keepHalfHeightPeak + rank1st
    rng = np.random.default_rng(0)
    uR = (rng.uniform(0, 1, st["wR"].shape) ** 4).astype(np.float32); uT = rng.uniform(0, 1, st["wT"].shape).astype(np.float32)
    ctx.pf_op(capi.PF_SET_U_KEEP_PEAK, uR=uR, uT=uT)
    ctx.pf_op(capi.PF_RANK1ST)
    sc = ctx.pf_get_scal()
    top = st["r"][np.arange(len(uR)), uR.argmax(1)]
    assert np.array_equal(sc[:, 6:10], top)
    assert np.array_equal(sc[:, 10:12], st["t"][np.arange(len(uT)), uT.argmax(1)])


def test_reconstruct_insert_plumbing_exact(ctx, prob):
    """thb_reconstruct_insert == thb_insert fed with the draws it made == the oracle's insert loop"""
    from oracle import portapi as port
    pb = prob
    _setup(ctx, pb)
    _load(ctx, pb, _params())
    mReco = 7
    offS = np.random.default_rng(1).normal(scale=0.3, size=(pb["nImg"], 2))
    ctx.reconstruct_insert(mReco, parGra=False, offS=offS)
    a = [ctx.reco_download(s) for s in (0, 1)]
    st = ctx.pf_get()
    dR, dT = ctx.pf_get_draws(mReco)
    assert dR.min() >= 0 and dR.max() < 125 and dT.min() >= 0 and dT.max() < 9
    assert len(np.unique(dR)) > 60          # draws spread over the support
    rows = np.arange(pb["nImg"])[:, None]
    nr, nt = st["r"][rows, dR], st["t"][rows, dT]
    w = np.full(pb["nImg"], 1.0 / mReco, np.float32)
    for s in (0, 1):
        ctx.reco_reset(s)
    ctx.insert(w, nr, nt, offS=offS)
    for s in (0, 1):
        b = ctx.reco_download(s)
        assert b["counter"] == a[s]["counter"] == mReco * int((pb["slot"] == s).sum())
        assert np.linalg.norm((a[s]["F"] - b["F"]).ravel()) <= 1e-6 * np.linalg.norm(b["F"].ravel())
        sel = np.nonzero(pb["slot"] == s)[0]
        want = port.insert_loop(pb["N"] * pb["pf"], pb["pf"], pb["N"], pb["datM"][sel], pb["ctfM"][sel], w[sel], offS[sel], nr[sel],
                                nt[sel], pb["pixM"]["iCol"], pb["pixM"]["iRow"])
        assert np.linalg.norm((a[s]["F"] - want["F"]).ravel()) <= 1e-6 * np.linalg.norm(want["F"].ravel())
        assert np.linalg.norm((a[s]["T"] - want["T"]).ravel()) <= 1e-6 * np.linalg.norm(want["T"].ravel())
        assert np.allclose(a[s]["O"], want["O"], rtol=1e-10, atol=1e-10)
    # particle grading: w = compressR / mReco
    for s in (0, 1):
        ctx.reco_reset(s)
    ctx.reconstruct_insert(mReco, parGra=True, offS=offS)
    g = ctx.reco_download(0)
    sc = ctx.pf_get_scal()
    cr = (sc[:, 0] * sc[:, 1] * sc[:, 2]) ** (-1.0 / 6)
    sel = pb["slot"] == 0
    assert abs(g["T"].sum() / a[0]["T"].sum() / cr[sel].mean() - 1) < 0.2


def test_expectation_phase_equals_explicit_kernel_call(ctx, prob):
    """one phase of thb_expectation == thb_expect_local on the state it saw (read back between the two)"""
    pb = prob
    _setup(ctx, pb)
    _load(ctx, pb, _params(fixed=1))
    # reproduce phase 0 by hand: perturb (device op), read the state, explicit kernel call
    ctx.pf_op(capi.PF_PERTURB_R, 2.0)
    ctx.pf_op(capi.PF_PERTURB_T, 2.0)
    st = ctx.pf_get()
    out = ctx.expect_local(st["r"], st["t"], st["wR"], st["wT"])
    ctx.pf_op(capi.PF_SET_U_KEEP_PEAK, uR=out["uR"], uT=out["uT"])
    ctx.pf_op(capi.PF_RANK1ST)
    sc = ctx.pf_get_scal()
    rows = np.arange(pb["nImg"])
    assert np.array_equal(sc[:, 6:10], st["r"][rows, out["uR"].argmax(1)])
    # the most likely support point is closer to the truth than the typical one
    err_top = _ang_deg(sc[:, 6:10], pb["par"]["quat"])
    err_all = _ang_deg(st["r"], pb["par"]["quat"][:, None, :])
    assert np.median(err_top) < np.median(err_all)


def test_expectation_statistical_parity_and_fsc(ctx, prob):
    from oracle import refapi as ref
    if not ref.available():
        pytest.skip("oracle/_ref not present")
    pb = prob
    n, N, pf = pb["nImg"], pb["N"], pb["pf"]
    phases, mReco = 8, 50
    _setup(ctx, pb)
    _load(ctx, pb, _params(fixed=phases))
    nph = ctx.expectation(want_phases=True)
    assert np.all(nph == phases)
    sc = ctx.pf_get_scal()
    err0 = _ang_deg(pb["q_start"], pb["par"]["quat"])
    errR = _ang_deg(sc[:, 6:10], pb["par"]["quat"])
    errT = np.linalg.norm(sc[:, 10:12] - pb["par"]["tran"], axis=1)
    ctx.reconstruct_insert(mReco)
    ours = [ctx.reco_download(s) for s in (0, 1)]

    # ---- the reference's own loop on the same inputs, twice: two seeds of ITS random stream give the yardstick for
    # how far two statistically equivalent runs of the particle filter differ
    def run_reference(seed):
        ref.lib().ref_set_seed(seed)
        pars = []
        for l in range(n):
            p = ref.Particle(125, 9, 2.0, 0.01)
            p.load(125, 9, pb["q_start"][l], pb["k0"], pb["k0"], pb["k0"], pb["t_start"][l], 1.0, 1.0)
            pars.append(p)
        refR = np.zeros((n, 4)); refT = np.zeros((n, 2))
        refacc = []
        for s in (0, 1):
            sel = np.nonzero(pb["slot"] == s)[0]
            P = ref.Projector(pf)
            P.set_padded_ft(pb["vols"][s])
            sub = [pars[l] for l in sel]
            ref.expectation_local(sub, P, pb["par"]["dat"][sel], pb["par"]["ctf"][sel], pb["par"]["sigRcp"][sel], pb["pixE"]["iCol"],
                                  pb["pixE"]["iRow"], N, 125, 9, fixedPhases=phases, nThread=8)
            for l in sel:
                q = np.zeros(4); t = np.zeros(2); c = np.zeros(1, np.int32); d = np.zeros(1)
                ref.lib().ref_particle_rank1st(pars[l].h, c.ctypes.data, q.ctypes.data, t.ctypes.data, d.ctypes.data)
                refR[l], refT[l] = q, t
            reco = ref.Reconstructor(N, N, pf, 8)
            reco.set_precal(pb["pixM"]["iColPad"], pb["pixM"]["iRowPad"], pb["pixM"]["iPxl"], pb["pixM"]["iSig"])
            reco.insert_loop(pb["datM"][sel], pb["ctfM"][sel], None, None, mReco, None, pb["pixM"]["iCol"], pb["pixM"]["iRow"], N,
                             nThread=8, pars=sub)
            refacc.append(reco.get())
            reco.close(); P.close()
        for p in pars:
            p.close()
        return refR, refT, refacc

    refR, refT, refacc = run_reference(4242)
    refR2, refT2, refacc2 = run_reference(777)
    refErrR = _ang_deg(refR, pb["par"]["quat"]); refErrT = np.linalg.norm(refT - pb["par"]["tran"], axis=1)
    refErrR2 = _ang_deg(refR2, pb["par"]["quat"])
    print(f"\nstart {np.median(err0):.3f} deg | ours {np.median(errR):.3f} deg, {np.median(errT):.3f} px | "
          f"reference {np.median(refErrR):.3f} deg, {np.median(refErrT):.3f} px | reference, other seed {np.median(refErrR2):.3f} deg")
    assert np.median(errR) < np.median(err0)                       # the filter converges towards the truth
    assert np.median(errR) <= 1.5 * max(np.median(refErrR), np.median(refErrR2)) + 0.25      # as accurate as the reference's
    assert np.median(errT) <= 1.5 * np.median(refErrT) + 0.15
    # volumes: FSC between ours and the reference's back-projections of the same images, against the FSC between two
    # runs of the reference itself (different random streams -> different support points -> the volumes of two
    # equivalent runs differ at high resolution; bit-level volume parity for IDENTICAL orientation lists is
    # test_reconstruct_insert_plumbing_exact / test_insert_random_two_halves, FSC >= 0.99999)
    rmax = 30
    for s in (0, 1):
        assert ours[s]["counter"] == refacc[s]["counter"]
        f = synth.fsc(ours[s]["F"], refacc[s]["F"], rmax)
        f0 = synth.fsc(refacc2[s]["F"], refacc[s]["F"], rmax)
        ft = synth.fsc(ours[s]["T"].astype(np.complex64), refacc[s]["T"].astype(np.complex64), rmax)
        ft0 = synth.fsc(refacc2[s]["T"].astype(np.complex64), refacc[s]["T"].astype(np.complex64), rmax)
        print(f"slot {s}: FSC(F) ours-vs-ref min {f[1:].min():.5f} mean {f[1:].mean():.5f} | ref-vs-ref min {f0[1:].min():.5f} mean "
              f"{f0[1:].mean():.5f}; FSC(T) min {ft[1:].min():.5f} | {ft0[1:].min():.5f}")
        # 24 particles per half map: the seed-to-seed spread of these curves is several percent
        # (a sanity band, not a parity gate: the deterministic volume parity is in the tests named above)
        assert f[1:6].min() >= f0[1:6].min() - 0.1
        assert f[1:].mean() >= f0[1:].mean() - 0.15
        assert ft[1:].mean() >= ft0[1:].mean() - 0.15


def test_adaptive_stop_rule_runs(ctx, prob):
    """data-dependent phase count (MIN 3, MAX 100, 5% variance-decrease rule) is evaluated on the device"""
    pb = prob
    _setup(ctx, pb)
    prm = _params(fixed=0)
    prm.maxPhase = 20
    _load(ctx, pb, prm)
    nph = ctx.expectation(want_phases=True)
    assert nph.min() >= 4 and nph.max() <= 20
    assert len(np.unique(nph)) > 1


def test_closed_loop_iterations_on_device():
    """The whole loop without leaving the device: E-step (device particle filter + fused kernel) -> insert -> (all-reduce)
    -> reconstruct -> setProjectee -> next E-step, two half sets, two iterations, starting from a blurred reference.
    Checks that the pieces fit: the reconstructed half maps correlate with the phantom and with each other."""
    from oracle import reco_port
    N, pf = 64, 2
    rng = np.random.default_rng(2024)
    truth = synth.phantom(N, 14, seed=6)
    truthFT = reco_port.set_projectee(truth, pf)
    # the E-step frequency limit follows the resolution at which the half maps still agree (Optimiser's _r): shell 14 of 32
    pixE = capi.pixel_list(N, pf, 14.0, 1.0)
    pixM = capi.pixel_list(N, pf, 30.0, 0.0)
    PE, PM = len(pixE["iCol"]), len(pixM["iCol"])
    nImg = 600
    slot = (np.arange(nImg) % 2).astype(np.int32)
    c = capi.Context(0)
    try:
        c.set_volume(0, truthFT)
        quat = synth.random_quats(nImg, rng)
        tran = rng.normal(scale=1.5, size=(nImg, 2))
        dU = rng.uniform(1.0e4, 3.0e4, nImg); dV = dU + rng.uniform(0, 500.0, nImg); th = rng.uniform(0, np.pi, nImg)

        def simulate(pix):
            c.set_expect_pixels(N, pf, pix["iCol"], pix["iRow"])
            clean = c.project(0, quat)
            ctf = np.stack([synth.ctf_values(pix["iCol"].astype(float), pix["iRow"].astype(float), N, 1.32, 3e5, dU[l], dV[l], th[l], 2.7e7, 0.1)
                            for l in range(nImg)]).astype(np.float32)
            ph = -2 * np.pi * (pix["iCol"][None] * tran[:, :1] / N + pix["iRow"][None] * tran[:, 1:] / N)
            sig2 = float(np.mean(np.abs(clean * ctf) ** 2)) / 1.0                     # per-pixel SNR 1
            noise = (rng.normal(size=clean.shape) + 1j * rng.normal(size=clean.shape)) * np.sqrt(sig2 / 2)
            return (ctf * clean * np.exp(1j * ph) + noise).astype(np.complex64), ctf, sig2

        datM, ctfM, _ = simulate(pixM)
        datE, ctfE, sig2 = simulate(pixE)
        c.set_expect_pixels(N, pf, pixE["iCol"], pixE["iRow"])
        c.set_insert_pixels(N, pf, pixM["iColPad"], pixM["iRowPad"])
        c.upload_stack(capi.STACK_EXPECT, datE, ctfE, np.full((nImg, PE), -0.5 / sig2, np.float32), slot)
        c.upload_stack(capi.STACK_INSERT, datM, ctfM, slotOfImg=slot)
        # initial reference: the phantom blurred (Gaussian, half amplitude at ~shell 11 of 32)
        g = np.fft.fftfreq(N)[:, None, None] ** 2 + np.fft.fftfreq(N)[None, :, None] ** 2 + np.fft.rfftfreq(N)[None, None, :] ** 2
        blurred = np.fft.irfftn(np.fft.rfftn(truth) * np.exp(-g / (2 * 0.15 ** 2)), s=(N, N, N), axes=(0, 1, 2)).astype(np.float32)
        truthF = np.fft.rfftn(truth).astype(np.complex64)
        for s in (0, 1):
            c.set_projectee(s, blurred, N, pf)
            c.reco_alloc(s, N * pf)
        k0 = 1e-4
        q_start = np.stack([synth.acg_cloud(quat[l], k0, 1, rng)[0] for l in range(nImg)])
        t_start = tran + rng.normal(scale=0.5, size=(nImg, 2))
        k123 = np.full((nImg, 3), k0); s01 = np.full((nImg, 2), 1.0)
        c.pf_set_image_base(0, 0)
        fsc_truth, fsc_half, err = [], [], []
        for it in range(3):
            prm = _params(fixed=6, seed=100 + it)
            c.pf_load(prm, q_start, k123, t_start, s01)
            c.expectation()
            sc = c.pf_get_scal()
            q_start, t_start = sc[:, 6:10].copy(), sc[:, 10:12].copy()       # next iteration starts from this one's best
            k123 = np.maximum(sc[:, 0:3], 1e-6); s01 = np.maximum(sc[:, 3:5], 0.1)
            for s in (0, 1):
                c.reco_reset(s)
            c.reconstruct_insert(20)
            c.allreduce()
            vols = []
            for s in (0, 1):
                v, nit = c.reconstruct(s, N, pf)
                assert np.isfinite(v).all()
                c.set_projectee(s, None, N, pf)                               # the next E-step projects from the new map
                vols.append(v)
            vF = [np.fft.rfftn(v).astype(np.complex64) for v in vols]
            ft = [synth.fsc(x, truthF, N // 2 - 4) for x in vF]
            f = synth.fsc(vF[0], vF[1], N // 2 - 4)
            fsc_truth.append(ft); fsc_half.append(f); err.append(float(np.median(_ang_deg(q_start, quat))))
            print(f"\niteration {it}: FSC with the phantom at shells 2/8/16 " + " ".join(f"{x[2]:.3f}/{x[8]:.3f}/{x[16]:.3f}" for x in ft)
                  + f", half-map FSC {f[2]:.3f}/{f[8]:.3f}/{f[16]:.3f}, median orientation error {err[-1]:.2f} deg")
        # the filter's own spread on this low-resolution phantom is several degrees (the reference's is the same:
        # test_expectation_statistical_parity_and_fsc); what is checked here is that the closed loop does not drift
        assert err[-1] < 10.0 and err[-1] <= err[0] + 2.0
        assert min(x[1:9].min() for x in fsc_truth[-1]) > 0.9                 # ... and the maps stay the phantom
        assert fsc_half[-1][1:9].min() > 0.9
    finally:
        c.close()

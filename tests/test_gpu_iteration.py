"""Iteration-level GPU tests: device particle filter + fused E kernel (thb_expectation) and the
particle-filter-driven insert (thb_reconstruct_insert), against the reference's own loops
(oracle/_ref: Particle + Projector + logDataVSPrior + Reconstructor driven by ref_harness.cpp).

The library's random numbers are GSL's distributions over a Philox bit stream; with the same bit generator plugged into the
reference (oracle/ref_harness.cpp: ref_rng_replay) both particle filters see the same random numbers, so the E -> M chain is
compared deterministically (test_expectation_replay_*) and, on 2 000 particles, by the north star's FSC >= 0.999 gate.  The
older statistical comparison against two seeds of the reference's own mt19937 stream stays as a cross-check.
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


def test_device_pf_operators_equal_host_build_with_the_same_stream(ctx, prob):
    """every particle-filter operator on the device (one warp per particle, lane-parallel ACG inference, CUDA's libm) against the
    host build of the same source (tests/pf_host, the build tests/test_pf_host.py pins to the reference's Particle class draw by
    draw) with the same random stream: key (seed, particle, epoch << 20)"""
    import ctypes as C
    import os
    import subprocess
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    src, lib = root / "tests" / "pf_host" / "pf_host.cpp", root / "tests" / "pf_host" / "libpf_host.so"
    if not lib.exists():
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-I", str(root / "thunder_b200" / "csrc"), "-o", os.fspath(lib), os.fspath(src)])
    L = C.CDLL(os.fspath(lib))
    _p, _i, _d, _u = C.c_void_p, C.c_int, C.c_double, C.c_ulonglong
    L.pfh_run_s.restype = _i
    L.pfh_run_s.argtypes = [_i, _d, _i, _i] + [_p] * 9 + [_d, _d, _u, _u, _u]
    pb = prob
    n, mLR, mLT = pb["nImg"], 125, 9
    prm = _params(mLR, mLT, seed=4711)
    _setup(ctx, pb)
    E = 900
    ctx.pf_set_epoch(E)
    _load(ctx, pb, prm)
    E += 1
    rng = np.random.default_rng(8)

    def host_apply(st, op, arg, epoch, uR=None, uT=None):
        out = {k: v.copy() for k, v in st.items()}
        for p in range(n):
            r = np.ascontiguousarray(st["r"][p].T); t = np.ascontiguousarray(st["t"][p].T)
            wR = st["wR"][p].copy(); wT = st["wT"][p].copy(); scal = st["scal"][p].copy()
            hu = np.zeros(mLR); ht = np.zeros(mLT)
            uRf = None if uR is None else np.ascontiguousarray(uR[p], np.float32)
            uTf = None if uT is None else np.ascontiguousarray(uT[p], np.float32)
            pt = lambda a: None if a is None else a.ctypes.data_as(_p)
            if "uR" in st:
                hu[:] = st["uR"][p]; ht[:] = st["uT"][p]
            L.pfh_run_s(op, arg, mLR, mLT, pt(r), pt(t), pt(wR), pt(wT), pt(hu), pt(ht), pt(uRf), pt(uTf), pt(scal), 2.0, 0.01, prm.seed, p, epoch << 20)
            out["r"][p] = r.T; out["t"][p] = t.T; out["wR"][p] = wR; out["wT"][p] = wT; out["scal"][p] = scal
            out.setdefault("uR", np.zeros((n, mLR)))[p] = hu
            out.setdefault("uT", np.zeros((n, mLT)))[p] = ht
        return out

    def same(a, b, what, tol=1e-8):
        dr = np.abs(a["r"] - b["r"]).max(); dt = np.abs(a["t"] - b["t"]).max()
        dk = np.abs(a["scal"][:, 0:5] / b["scal"][:, 0:5] - 1).max()
        assert dr <= tol and dt <= tol and dk <= 1e-6, (what, dr, dt, dk)

    st = ctx.pf_get()
    host = dict(st)
    for rep in range(3):
        for op, arg, name in ((capi.PF_PERTURB_R, 2.0 if rep == 0 else 0.5, "perturb R"), (capi.PF_PERTURB_T, 0.5, "perturb T")):
            ctx.pf_op(op, arg); E += 1
            host = host_apply(host, op, arg, E)
            st = ctx.pf_get()
            same(st, host, name)
        # weights within a factor 3: most of the support survives the resampling (a cloud resampled down to a dozen distinct
        # points makes the ACG inference rank-deficient and its result a matter of the last bits - see tests/test_pf_host.py)
        uR = np.exp(-rng.uniform(0, 1, (n, mLR)) ** 2).astype(np.float32); uT = rng.uniform(0.1, 1, (n, mLT)).astype(np.float32)
        ctx.pf_op(capi.PF_SET_U_KEEP_PEAK, uR=uR, uT=uT); E += 1
        host = host_apply(host, capi.PF_SET_U_KEEP_PEAK, 0.0, E, uR, uT)
        ctx.pf_op(capi.PF_CALVARI); E += 1
        host = host_apply(host, capi.PF_CALVARI, 0.0, E)
        st = ctx.pf_get()
        same(st, host, "calVari")
        ctx.pf_op(capi.PF_RESAMPLE); E += 1
        host = host_apply(host, capi.PF_RESAMPLE, 0.0, E)
        st = ctx.pf_get()
        same(st, host, "resample")
        assert np.allclose(st["wR"], host["wR"], rtol=1e-9) and np.allclose(st["wT"], host["wT"], rtol=1e-9)


def _ref_particles(ref, pb, idx, prm, epoch_key, mLR, mLT):
    """the reference's Particle objects loaded with the random numbers thb_pf_load gave particle p: key (seed, p, epoch_key)"""
    pars = []
    for l in idx:
        p = ref.Particle(mLR, mLT, 2.0, 0.01)
        ref.rng_key(prm.seed, int(l), epoch_key)
        p.load(mLR, mLT, pb["q_start"][l], pb["k0"], pb["k0"], pb["k0"], pb["t_start"][l], 1.0, 1.0)
        pars.append(p)
    return pars


def _replay_reference(ref, pars, vols, slot, pf, N, pixE, dat, ctf, sigRcp, mLR, mLT, phases, prm, E0, uRt, uTt, tr, nThread):
    """the reference's phase loop with the device's random numbers (replay), per half set.  After ITS OWN perturbation of every
    phase (returned) the reference's support is set to the device's, it computes ITS OWN marginal weights from there (returned),
    then resamples from the device's weights; its support after the resampling is returned as well."""
    n = len(pars)
    own = dict(uR=np.zeros((phases, n, mLR), np.float32), uT=np.zeros((phases, n, mLT), np.float32), cond=np.zeros((phases, n)),
               rPert=np.zeros((phases, n, mLR, 4)), tPert=np.zeros((phases, n, mLT, 2)), rRes=np.zeros((phases, n, mLR, 4)),
               tRes=np.zeros((phases, n, mLT, 2)))
    for s in (0, 1):
        sel = np.nonzero(slot == s)[0]
        assert np.array_equal(sel, 2 * np.arange(len(sel)) + s)          # image j of the sub-list is particle 2 j + s: stream s, stride 2
        P = ref.Projector(pf)
        P.set_padded_ft(vols[s])
        ref.rng_replay_loop(prm.seed, s, (E0 + 2) << 20, stride=2)
        oR, oT, cnd, st = ref.expectation_local_trace([pars[l] for l in sel], P, dat[sel], ctf[sel], sigRcp[sel], pixE["iCol"], pixE["iRow"], N,
                                                      mLR, mLT, phases, uRIn=uRt[:, sel], uTIn=uTt[:, sel], rIn=tr["rPert"][:, sel],
                                                      tIn=tr["tPert"][:, sel], want_states=True, nThread=nThread)
        own["uR"][:, sel] = oR; own["uT"][:, sel] = oT; own["cond"][:, sel] = cnd
        for k in ("rPert", "tPert", "rRes", "tRes"):
            own[k][:, sel] = st[k]
        P.close()
    return own


def _check_replay(own, uRt, uTt, tr, phases, label, base, P):
    """the assertions of the deterministic E chain; returns the summary line.
    Tolerance of the marginal weights: they are exp(logL - max logL), and the reference sums the P terms of a log-likelihood in one
    fp32 register (logDataVSPrior_m_huabin, src/Optimiser.cpp:9187-9213): a random walk of half-ulp errors of the running sum,
    sigma = 2^-24 sqrt(P) |logL| (the same bound tests/test_gpu_hotpath.py uses for the log-likelihoods themselves); a weight may be
    off by the difference of two such errors, taken at 6 sigma over the 10^3 - 10^6 entries compared."""
    rel = lambda a, b: np.abs(a.astype(np.float64) - b.astype(np.float64)).max(-1) / np.maximum(a.astype(np.float64).max(-1), 1e-300)
    wdiff = np.maximum(rel(own["uR"], uRt), rel(own["uT"], uTt))                       # [phase][image]
    dPert = np.abs(own["rPert"] - tr["rPert"]).max((2, 3)); dPertT = np.abs(own["tPert"] - tr["tPert"]).max((2, 3))
    dRes = np.abs(own["rRes"] - tr["rRes"]).max((2, 3)); dResT = np.abs(own["tRes"] - tr["tRes"]).max((2, 3))
    # the perturbation of phase p >= 1 is taken about the ACG mean of the cloud resampled in phase p - 1
    well = np.ones_like(dPert, bool)
    well[1:] = own["cond"][:phases - 1] < 1e6
    tolW_ = np.expm1(6 * np.sqrt(2) * (np.finfo(np.float32).eps / 2 * np.sqrt(P) * np.abs(base).astype(np.float64)) + 2e-4) + 1e-3
    line = (f"{label}: load + first perturbation: support {dPert[0].max():.1e} / {dPertT[0].max():.1e} (rotations / translations); marginal weights "
            f"of the fused kernel vs the reference's from the same support, all {wdiff.size} (image, phase) pairs: median {np.median(wdiff):.1e}, "
            f"max {wdiff.max():.1e} of the maximum (bound from the reference's fp32 summation: median {np.median(tolW_):.1e}, max {tolW_.max():.1e}); support after calVari + resample: max {dRes.max():.1e} / {dResT.max():.1e}; later "
            f"perturbations about a well-defined mean (condition < 1e6): {int(well[1:].sum())} of {well[1:].size}, support max "
            f"{dPert[1:][well[1:]].max() if well[1:].any() else 0:.1e}; about the mean of a collapsed cloud: {int((~well).sum())}, of which "
            f"{int((dPert[~well] > 1e-6).sum())} differ by more than 1e-6; translations max {dPertT.max():.1e}")
    print("\n" + line)
    try:      # arrays of the comparison, for offline reading of a failure
        import os
        os.makedirs("gpurun_out", exist_ok=True)
        np.savez_compressed(f"gpurun_out/replay_dump_{len(wdiff[0])}.npz", wdiff=wdiff, tolW=tolW_, cond=own["cond"], dPert=dPert, dRes=dRes, base=base,
                            uR_own=own["uR"][:, :64], uR_dev=uRt[:, :64])
    except Exception:
        pass
    assert dPert[0].max() <= 1e-9 and dPertT[0].max() <= 1e-11                         # Particle::load + perturb(L)
    # E kernel inside the loop, every image, every phase: 99.9 % of the (image, phase) pairs inside the fp32 bound, none far outside
    assert (wdiff <= tolW_).mean() >= 0.999 and (wdiff / tolW_).max() <= 10.0, ((wdiff <= tolW_).mean(), float((wdiff / tolW_).max()))
    assert np.median(wdiff) <= 2e-3
    # calRank1st / calVari / resample from identical inputs: identical support, but for a rare resampling decision that sits within the
    # 1e-9 by which the two inferences of the balancing weights (1 / pdfACG) differ
    assert (dRes > 1e-9).mean() <= 0.005 and dResT.max() <= 1e-11, (dRes > 1e-9).mean()
    assert dPertT.max() <= 1e-9                                                         # translations: no ill-posed step anywhere
    # perturb(S) wherever the reference's own mean is well-posed (and the resampling before it took the same decisions)
    ok = well[1:] & (dRes[:-1] <= 1e-9)
    assert ok.mean() >= 0.3
    assert np.median(dPert[1:][ok]) <= 1e-9 and (dPert[1:][ok] > 1e-6).mean() <= 0.02, (np.median(dPert[1:][ok]), (dPert[1:][ok] > 1e-6).mean())
    return line


def test_expectation_replay_deterministic_chain(ctx, prob):
    """E -> M against the reference's own loops with the SAME random numbers (replay of the bit generator, GSL's distributions on
    both sides), phase by phase and operator by operator:
      * thb_pf_load + the first perturbation == Particle::load + perturb(L)                                     (1e-9)
      * in every phase the reference computes the marginal weights with ITS Projector::project + translate + logDataVSPrior
        from the same support: they equal the fused kernel's                                               (5e-3 of the maximum)
      * from identical support and identical weights, setU / keepHalfHeightPeak / calRank1st / calVari / resample give the same
        support on both sides                                                                                   (1e-9)
      * the following perturbation gives the same support wherever the reference's ACG mean is well-posed       (1e-6).
        It is not where a sharply peaked likelihood resampled the cloud down to a handful of distinct points (typically phase 0):
        the reference infers the mean by a fixed-point iteration on a then rank-deficient 4x4 matrix, whose result depends on
        the last bits of the arithmetic - the host build of this library's own source differs from itself there when compiled
        with FMA contraction (tests/test_pf_host.py).  The reference's support is therefore set to the device's after each
        perturbation, which keeps every later comparison exact.
      * the draws of thb_reconstruct_insert are Particle::rand's, and the back-projected half maps agree (relative L2 1e-5)."""
    from oracle import refapi as ref
    if not ref.available():
        pytest.skip("oracle/_ref not present")
    pb = prob
    n, N, pf = pb["nImg"], pb["N"], pb["pf"]
    mLR, mLT, phases, mReco = 125, 9, 6, 40
    prm = _params(mLR, mLT, fixed=phases, seed=31337)
    _setup(ctx, pb)
    E0 = 5000
    ctx.pf_set_epoch(E0)
    _load(ctx, pb, prm)                                  # key ((E0 + 1) << 20)
    st0 = ctx.pf_get()
    ctx.pf_trace(phases)
    try:
        ctx.expectation()                                # keys ((E0 + 2) << 20) + {0, phase + 1}
    finally:
        ctx.pf_trace(0)
    uRt, uTt = ctx.pf_get_trace(phases)
    tr = ctx.pf_get_trace_states(phases)
    st1 = ctx.pf_get()
    assert np.array_equal(st1["r"], tr["rRes"][-1])
    ctx.reconstruct_insert(mReco)                        # key ((E0 + 3) << 20)
    ours = [ctx.reco_download(s) for s in (0, 1)]

    with ref.replay(prm.seed, 0, 0):
        pars = _ref_particles(ref, pb, range(n), prm, (E0 + 1) << 20, mLR, mLT)
        for l in range(n):
            g = pars[l].get()
            assert np.abs(g["r"] - st0["r"][l]).max() <= 1e-9 and np.abs(g["t"] - st0["t"][l]).max() <= 1e-11, l
        own = _replay_reference(ref, pars, pb["vols"], pb["slot"], pf, N, pb["pixE"], pb["par"]["dat"], pb["par"]["ctf"], pb["par"]["sigRcp"],
                                mLR, mLT, phases, prm, E0, uRt, uTt, tr, 8)
        print("\n" + _check_replay(own, uRt, uTt, tr, phases, "replay, 48 images", ctx.trace_base, len(pb["pixE"]["iCol"])))
        # M: the reference's insert loop draws with Particle::rand from the same stream, from the support it ended with
        for s in (0, 1):
            sel = np.nonzero(pb["slot"] == s)[0]
            reco = ref.Reconstructor(N, N, pf, 1)
            reco.set_precal(pb["pixM"]["iColPad"], pb["pixM"]["iRowPad"], pb["pixM"]["iPxl"], pb["pixM"]["iSig"])
            ref.rng_replay_loop(prm.seed, s, (E0 + 3) << 20, stride=2)
            reco.insert_loop(pb["datM"][sel], pb["ctfM"][sel], None, None, mReco, None, pb["pixM"]["iCol"], pb["pixM"]["iRow"], N,
                             nThread=8, pars=[pars[l] for l in sel])
            want = reco.get()
            reco.close()
            assert ours[s]["counter"] == want["counter"] == len(sel) * mReco
            relF = np.linalg.norm(ours[s]["F"] - want["F"]) / np.linalg.norm(want["F"])
            relT = np.linalg.norm(ours[s]["T"] - want["T"]) / np.linalg.norm(want["T"])
            f = synth.fsc(ours[s]["F"], want["F"], 30)
            print(f"slot {s}: half-map accumulators vs the reference's, relative L2: F {relF:.2e}, T {relT:.2e}; FSC min {f[1:].min():.7f}")
            assert relF <= 1e-5 and relT <= 1e-5 and f[1:].min() >= 0.999
            assert np.allclose(ours[s]["O"], want["O"], rtol=1e-6, atol=1e-6)
    for p in pars:
        p.close()


def test_iteration_fsc_gate_2000_particles():
    """the north star's acceptance gate at iteration level: 2 000 particles (1 000 per half set), box 64, the whole E step
    (8 phases x 125 x 9) and M step (mReco 50) on the device against the reference's own loops on the same stacks, with the same
    random numbers (replay bit generator) and the reference resampling from the device's marginal weights - which it checks, phase
    by phase, against the weights it computes itself from its own particles.  Per-shell FSC(ours, reference) >= 0.999 to the ring
    exercised, for F and for T, and at least FSC(reference, reference with another random stream) - 0.002.
    (Free-running - each side resampling from its own fp32 likelihoods - the two filters take different resampling decisions
    within a phase or two, as two builds of the reference would, and the half maps then differ like those of two random streams:
    FSC 0.92 against 0.82 reference-vs-reference at the last ring in this configuration.)"""
    from oracle import refapi as ref, portapi as port
    if not ref.available():
        pytest.skip("oracle/_ref not present")
    import os
    N, pf = 64, 2
    n, mLR, mLT, phases, mReco = 2000, 125, 9, 8, 50
    nThread = os.cpu_count() or 8
    rng = np.random.default_rng(2000)
    vols = [synth.padded_ft(synth.phantom(N, 14, seed=s), pf) for s in (3, 4)]
    pixE = port.pixel_list(N, pf, 28.0, 1.0)
    pixM = port.pixel_list(N, pf, 30.0, 0.0)
    slot = (np.arange(n) % 2).astype(np.int32)
    c = capi.Context(0)
    try:
        c.set_expect_pixels(N, pf, pixE["iCol"], pixE["iRow"])
        c.set_insert_pixels(N, pf, pixM["iColPad"], pixM["iRowPad"])
        for s in (0, 1):
            c.set_volume(s, vols[s])
            c.reco_alloc(s, N * pf)

        def project_fn(quats, pix=pixE):
            out = np.empty((n, len(pix["iCol"])), np.complex64)
            c.set_expect_pixels(N, pf, pix["iCol"], pix["iRow"])
            for s in (0, 1):
                sel = np.nonzero(slot == s)[0]
                out[sel] = c.project(s, quats[sel])
            c.set_expect_pixels(N, pf, pixE["iCol"], pixE["iRow"])
            return out
        par = synth.make_particles(n, N, pixE, project_fn, seed=19, snr_scale=40.0)
        quat = par["quat"]
        PM = len(pixM["iCol"])
        cleanM = project_fn(quat, pixM)
        ctfM = np.stack([synth.ctf_values(pixM["iCol"].astype(float), pixM["iRow"].astype(float), N, 1.32, 3e5, *par["ctfpar"][l], 2.7e7, 0.1)
                         for l in range(n)]).astype(np.float32)
        ph = -2 * np.pi * (pixM["iCol"][None] * par["tran"][:, :1] / N + pixM["iRow"][None] * par["tran"][:, 1:] / N)
        datM = (ctfM * cleanM * np.exp(1j * ph) + (rng.normal(size=(n, PM)) + 1j * rng.normal(size=(n, PM))) * np.sqrt(par["sig2"] / 2)).astype(np.complex64)
        k0 = 3e-4
        q_start = np.stack([synth.acg_cloud(quat[l], k0, 1, rng)[0] for l in range(n)])
        t_start = par["tran"] + rng.normal(scale=0.5, size=(n, 2))
        c.upload_stack(capi.STACK_EXPECT, par["dat"], par["ctf"], par["sigRcp"], slot)
        c.upload_stack(capi.STACK_INSERT, datM, ctfM, slotOfImg=slot)
        prm = _params(mLR, mLT, fixed=phases, seed=777001)
        E0 = 100
        c.pf_set_image_base(0, 0)
        c.pf_set_epoch(E0)
        c.pf_load(prm, q_start, np.full((n, 3), k0), t_start, np.full((n, 2), 1.0))
        c.pf_trace(phases)
        c.expectation()
        uRt, uTt = c.pf_get_trace(phases)
        trace_base = c.trace_base
        tr = c.pf_get_trace_states(phases)
        st1 = c.pf_get()
        sc = st1["scal"]
        c.reconstruct_insert(mReco)
        ours = [c.reco_download(s) for s in (0, 1)]
    finally:
        c.close()
    pb = dict(q_start=q_start, t_start=t_start, k0=k0)

    def run_reference(replay):
        pars = []
        own = None
        if replay:
            ref.lib().ref_rng_replay(1)
            pars = _ref_particles(ref, pb, range(n), prm, (E0 + 1) << 20, mLR, mLT)
            own = _replay_reference(ref, pars, vols, slot, pf, N, pixE, par["dat"], par["ctf"], par["sigRcp"], mLR, mLT, phases, prm, E0,
                                    uRt, uTt, tr, nThread)
        else:
            ref.lib().ref_set_seed(4242)
            for l in range(n):
                p = ref.Particle(mLR, mLT, 2.0, 0.01)
                p.load(mLR, mLT, q_start[l], k0, k0, k0, t_start[l], 1.0, 1.0)
                pars.append(p)
        acc = []
        top = np.zeros((n, 4))
        for s in (0, 1):
            sel = np.nonzero(slot == s)[0]
            sub = [pars[l] for l in sel]
            if not replay:
                P = ref.Projector(pf)
                P.set_padded_ft(vols[s])
                ref.expectation_local(sub, P, par["dat"][sel], par["ctf"][sel], par["sigRcp"][sel], pixE["iCol"], pixE["iRow"], N, mLR, mLT,
                                      fixedPhases=phases, nThread=nThread)
                P.close()
            for l in sel:
                q = np.zeros(4); t = np.zeros(2); cc = np.zeros(1, np.int32); d = np.zeros(1)
                ref.lib().ref_particle_rank1st(pars[l].h, cc.ctypes.data, q.ctypes.data, t.ctypes.data, d.ctypes.data)
                top[l] = q
            reco = ref.Reconstructor(N, N, pf, nThread)
            reco.set_precal(pixM["iColPad"], pixM["iRowPad"], pixM["iPxl"], pixM["iSig"])
            if replay:
                ref.rng_replay_loop(prm.seed, s, (E0 + 3) << 20, stride=2)
            reco.insert_loop(datM[sel], ctfM[sel], None, None, mReco, None, pixM["iCol"], pixM["iRow"], N, nThread=nThread, pars=sub)
            acc.append(reco.get())
            reco.close()
        for p in pars:
            p.close()
        ref.lib().ref_rng_replay(0)
        return acc, top, own

    refA, topA, own = run_reference(True)
    refB, topB, _ = run_reference(False)
    print("\n" + _check_replay(own, uRt, uTt, tr, phases, "replay, 2000 particles", trace_base, len(pixE["iCol"])))
    print(f"median orientation error: ours {np.median(_ang_deg(sc[:, 6:10], quat)):.3f} deg, reference (same random numbers) "
          f"{np.median(_ang_deg(topA, quat)):.3f} deg, reference (another random stream, free-running) {np.median(_ang_deg(topB, quat)):.3f} deg")
    rmax = 30            # the M pixel list reaches ring 30 of 32
    for s in (0, 1):
        assert ours[s]["counter"] == refA[s]["counter"]
        f = synth.fsc(ours[s]["F"], refA[s]["F"], rmax)
        f0 = synth.fsc(refB[s]["F"], refA[s]["F"], rmax)
        ft = synth.fsc(ours[s]["T"].astype(np.complex64), refA[s]["T"].astype(np.complex64), rmax)
        ft0 = synth.fsc(refB[s]["T"].astype(np.complex64), refA[s]["T"].astype(np.complex64), rmax)
        relF = np.linalg.norm(ours[s]["F"] - refA[s]["F"]) / np.linalg.norm(refA[s]["F"])
        print(f"slot {s}: FSC(F) ours-vs-reference min {f[1:].min():.7f} (relative L2 {relF:.1e}) | reference-vs-reference(other stream) min "
              f"{f0[1:].min():.5f}; FSC(T) {ft[1:].min():.7f} | {ft0[1:].min():.5f}")
        assert f[1:].min() >= 0.999 and ft[1:].min() >= 0.999
        assert np.all(f[1:] >= f0[1:] - 0.002) and np.all(ft[1:] >= ft0[1:] - 0.002)


def test_global_search_iteration_scan_handover_phases(ctx, prob):
    """a global-search E-step in MODE_3D on the device: scan over a shared grid of 3 000 random rotations x 30 translations
    (thb_expect_scan, one launch per image chunk) -> support of the local phases from the scan's weights (thb_pf_from_scan: class
    choice trivial with k = 1, setPeakFactor / keepHalfHeightPeak / resample down to 125 x 9, variances with the scan's floors) ->
    phases (thb_expectation).  The scan's best grid point is ~10 degrees off (grid spacing), the phases bring it down."""
    pb = prob
    n, N = pb["nImg"], pb["N"]
    _setup(ctx, pb)
    rng = np.random.default_rng(31)
    nR, nT = 3000, 30
    grid = synth.random_quats(nR, rng)
    trans = rng.normal(scale=2.0, size=(nT, 2))
    pR = np.full(nR, 1.0 / nR); pT = np.full(nT, 1.0 / nT)
    res = [ctx.expect_scan(s, grid, trans, pR, pT) for s in (0, 1)]
    # k = 1: every image has ONE reference, its half-set's (the scan of the other slot leaves its rows zero)
    wC = (res[0]["wC"] + res[1]["wC"])[:, None]
    wR = (res[0]["wR"] + res[1]["wR"])[None]; wT = (res[0]["wT"] + res[1]["wT"])[None]
    scanMinStdR = nR ** (-1.0 / 3); pfS = 0.5
    prm = _params(125, 9, fixed=8, seed=77)
    prm.perturbFactorL = pfS                       # global search: no large first perturbation (phases start at 1 in the reference)
    slot_before = pb["slot"].copy()
    ctx.pf_set_image_base(0, 0)
    cls = ctx.pf_from_scan(prm, grid, trans, wC, wR, wT, kFloor=(scanMinStdR / pfS) ** 2, sFloor=0.3)
    assert not cls.any()
    # with k = 1 the slots (the two half sets) are left alone: the phases below run every image against its own half-set reference
    sc0 = ctx.pf_get_scal()
    err0 = _ang_deg(sc0[:, 6:10], pb["par"]["quat"])
    ctx.expectation()
    sc = ctx.pf_get_scal()
    err = _ang_deg(sc[:, 6:10], pb["par"]["quat"])
    print(f"\nglobal search: best grid point of the scan {np.median(err0):.2f} deg off (median), after 8 phases {np.median(err):.2f} deg; "
          f"translation {np.median(np.linalg.norm(sc[:, 10:12] - pb['par']['tran'], axis=1)):.2f} px")
    assert np.median(err0) < 15.0
    assert np.median(err) < 0.6 * np.median(err0)


def test_adaptive_stop_rule_runs(ctx, prob):
    """data-dependent phase count (MIN 3, MAX 100, 5% variance-decrease rule) is evaluated on the device"""
    pb = prob
    _setup(ctx, pb)
    prm = _params(fixed=0)
    prm.maxPhase = 20
    ctx.pf_set_epoch(1000)                      # a known position of the random streams: the run is repeated below
    _load(ctx, pb, prm)
    nph = ctx.expectation(want_phases=True)
    assert nph.min() >= 4 and nph.max() <= 20
    assert len(np.unique(nph)) > 1
    # the E launches of the later phases take the compacted list of unfinished particles (a tail of a few particles then runs spread
    # over the chip).  With the per-image kernel forced for both (the spread kernel sums the pixels in another order, and a last-bit
    # difference of a weight can flip a resampling decision), launches over the compacted list and launches over all particles with
    # the finished ones skipped give the same phase counts and the same final state, bit for bit
    runs = []
    try:
        ctx.set_option("expect_spread", 0)
        for compact in (1, 0):
            ctx.set_option("pf_compact", compact)
            ctx.pf_set_epoch(1000)
            _load(ctx, pb, prm)
            runs.append((ctx.expectation(want_phases=True), ctx.pf_get()))
    finally:
        ctx.set_option("pf_compact", 1)
        ctx.set_option("expect_spread", -1)
    assert np.array_equal(runs[0][0], runs[1][0])
    for key in ("r", "t", "wR", "wT", "scal"):
        assert np.array_equal(runs[0][1][key], runs[1][1][key], equal_nan=True), key


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
            c.reconstruct_insert(50)
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
        assert err[-1] < 10.0 and err[-1] <= err[0] + 3.0
        assert min(x[1:9].min() for x in fsc_truth[-1]) > 0.9                 # ... and the maps stay the phantom
        assert fsc_half[-1][1:9].min() > 0.9
    finally:
        c.close()

"""Particle-filter operators (thunder_b200/csrc/thb_pf.cuh, the source the CUDA kernels compile)
built for the host and compared with the reference's Particle / DirectionalStat classes.
Deterministic operators: exact (1e-9) parity.  Stochastic ones (different RNG by design): structural
and statistical properties.  CPU only."""
import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np
import pytest

from thunder_b200 import synth

ROOT = Path(__file__).resolve().parent.parent
SRC = ROOT / "tests" / "pf_host" / "pf_host.cpp"
LIB = ROOT / "tests" / "pf_host" / "libpf_host.so"
_p, _i, _d = C.c_void_p, C.c_int, C.c_double


@pytest.fixture(scope="module")
def pfh():
    hdr = ROOT / "thunder_b200" / "csrc" / "thb_pf.cuh"
    hdr2 = ROOT / "thunder_b200" / "csrc" / "thb_pf2d.cuh"
    if not LIB.exists() or LIB.stat().st_mtime < max(SRC.stat().st_mtime, hdr.stat().st_mtime, hdr2.stat().st_mtime):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-I", str(ROOT / "thunder_b200" / "csrc"),
                               "-o", os.fspath(LIB), os.fspath(SRC)])
    L = C.CDLL(os.fspath(LIB))
    L.pfh_run.restype = _i
    L.pfh_run.argtypes = [_i, _d, _i, _i] + [_p] * 9 + [_d, _d, C.c_ulonglong, C.c_ulonglong]
    L.pfh_infer_acg.argtypes = [_i, _p, _p, _p]
    L.pfh_rng.argtypes = [C.c_ulonglong, C.c_ulonglong, C.c_ulonglong, _i, _p, _p]
    L.pfh_rng_draw.argtypes = [C.c_ulonglong, C.c_ulonglong, C.c_ulonglong, _i, _i, _d, _d, _p]
    L.pfh_run_d.restype = _i
    L.pfh_from_scan.restype = _i
    L.pfh_from_scan.argtypes = [_i, _i, _i, _i] + [_p] * 5 + [_i, _i, _d, _d] + [_p] * 5 + [C.c_ulonglong] * 3
    L.pfh_run_d.argtypes = [_i, _d, _i, _p, _p, _p, _p, _p, C.c_ulonglong, C.c_ulonglong, C.c_ulonglong]
    L.pfh_sample_vms.argtypes = [C.c_ulonglong, _d, _i, _p]
    L.pfh_infer_vms.argtypes = [_i, _p, _p, _p]
    L.pfh_pdf_vms.restype = _d
    L.pfh_pdf_vms.argtypes = [_p, _p, _d]
    L.pfh_perturb_r_2d.argtypes = [_i, _p, _d, _d, C.c_ulonglong]
    L.pfh_balance_r_2d.argtypes = [_i, _p, _p]
    L.pfh_resample_c.restype = _i
    L.pfh_resample_c.argtypes = [_i, _p, _p, _p, _i, _p, _p, C.c_ulonglong]
    return L


class HostParticle:
    """one particle in the component-major layout of the device state"""

    def __init__(self, L, mLR, mLT, transS=2.0, transQ=0.01, seed=1):
        self.L, self.mLR, self.mLT = L, mLR, mLT
        self.r = np.zeros((4, mLR)); self.t = np.zeros((2, mLT)); self.wR = np.full(mLR, 1.0 / mLR); self.wT = np.full(mLT, 1.0 / mLT)
        self.uR = np.zeros(mLR); self.uT = np.zeros(mLT); self.scal = np.zeros(20)
        self.scal[16] = 1e-3
        self.transS, self.transQ, self.seed, self.epoch = transS, transQ, seed, 0

    def run(self, op, arg=0.0, uRf=None, uTf=None):
        self.epoch += 1
        uRf = None if uRf is None else np.ascontiguousarray(uRf, np.float32)
        uTf = None if uTf is None else np.ascontiguousarray(uTf, np.float32)
        pt = lambda a: None if a is None else a.ctypes.data_as(_p)
        return self.L.pfh_run(op, arg, self.mLR, self.mLT, pt(self.r), pt(self.t), pt(self.wR), pt(self.wT), pt(self.uR),
                              pt(self.uT), pt(uRf), pt(uTf), pt(self.scal), self.transS, self.transQ, self.seed, self.epoch)


def _cloud(rng, n, k=(3e-4, 1e-4, 6e-4)):
    q0 = synth.random_quats(1, rng)[0]
    v = rng.normal(size=(n, 4)) * np.sqrt(np.array([1.0, *k]))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    return synth.quat_mul(v, np.broadcast_to(q0, (n, 4))), q0


def test_rng_uniform_normal(pfh):
    n = 200000
    u = np.empty(n); g = np.empty(n)
    pfh.pfh_rng(7, 3, 9, n, u.ctypes.data_as(_p), g.ctypes.data_as(_p))
    assert 0 < u.min() and u.max() < 1
    assert abs(u.mean() - 0.5) < 4e-3 and abs(u.var() - 1 / 12) < 2e-3
    assert abs(g.mean()) < 1e-2 and abs(g.var() - 1) < 2e-2 and abs((g ** 4).mean() - 3) < 0.1
    u2 = np.empty(n); g2 = np.empty(n)
    pfh.pfh_rng(7, 4, 9, n, u2.ctypes.data_as(_p), g2.ctypes.data_as(_p))     # another stream: decorrelated
    assert abs(np.corrcoef(u, u2)[0, 1]) < 1e-2


def test_infer_acg_matches_reference(pfh, ref):
    rng = np.random.default_rng(0)
    for n in (125, 37, 9):
        r, q0 = _cloud(rng, n)
        A_ref = np.empty(16); k_ref = np.empty(3); m_ref = np.empty(4)
        rr = np.ascontiguousarray(r)
        ref.lib().ref_inferACG(rr.ctypes.data_as(_p), n, A_ref.ctypes.data_as(_p), k_ref.ctypes.data_as(_p), m_ref.ctypes.data_as(_p))
        rc = np.ascontiguousarray(r.T)
        A = np.empty(16); m = np.empty(4)
        pfh.pfh_infer_acg(n, rc.ctypes.data_as(_p), A.ctypes.data_as(_p), m.ctypes.data_as(_p))
        assert np.abs(A - A_ref).max() <= 1e-8 * np.abs(A_ref).max()   # A is near-singular for a tight cloud
        assert np.allclose(np.array([A[5], A[10], A[15]]) / A[0], k_ref, rtol=1e-6)
        assert min(np.abs(m - m_ref).max(), np.abs(m + m_ref).max()) < 1e-9        # eigenvector up to sign
        assert abs(abs(np.dot(m, q0)) - 1) < 1e-2


def _ref_particle(ref, r, t, wR=None, wT=None, k=(1e-4, 1e-4, 1e-4), s=(1.0, 1.0)):
    p = ref.Particle(len(r), len(t))
    p.load(len(r), len(t), r[0], k[0], k[1], k[2], t[0], s[0], s[1])
    p.set(r=r, t=t, wR=wR, wT=wT)
    return p


def test_balance_calvari_keep_peak_match_reference(pfh, ref):
    rng = np.random.default_rng(1)
    mLR, mLT = 125, 9
    r, q0 = _cloud(rng, mLR)
    t = rng.normal(scale=1.3, size=(mLT, 2)) + [0.5, -1.0]
    P = _ref_particle(ref, r, t)
    H = HostParticle(pfh, mLR, mLT)
    H.r[:] = r.T; H.t[:] = t.T
    # balanceWeight
    P.balanceWeight(P.PAR_R); P.balanceWeight(P.PAR_T)
    H.run(7); H.run(8)
    g = P.get()
    assert np.allclose(H.wR, g["wR"], rtol=1e-9)
    assert np.allclose(H.wT, g["wT"], rtol=1e-9)
    # calVari (R rotates the cloud by conj(mean) and back: compare the cloud too)
    P.calVari(P.PAR_R); P.calVari(P.PAR_T)
    H.run(5)
    sc = P.scalars()
    assert np.allclose(H.scal[0:3], sc[0:3], rtol=1e-9)          # k1 k2 k3
    assert np.allclose(H.scal[3:5], sc[3:5], rtol=1e-12)         # s0 s1
    assert np.allclose(H.r.T, P.get()["r"], atol=1e-12)
    assert abs(P.variR() - (H.scal[0] * H.scal[1] * H.scal[2]) ** (1 / 6)) < 1e-12
    assert abs(P.variT() - H.scal[3] * H.scal[4]) < 1e-6          # the reference evaluates this in float (mat22)
    # setU + keepHalfHeightPeak(R) + calRank1st
    uR = rng.uniform(0, 1, mLR).astype(np.float32) ** 6
    uT = rng.uniform(0, 1, mLT).astype(np.float32)
    sc[16] = 0.3                                                   # peakFactorR
    P.set_scalars(sc); H.scal[16] = 0.3
    P.set_u(1, uR.astype(np.float64)); P.set_u(2, uT.astype(np.float64))
    P.keepHalfHeightPeak(P.PAR_R)
    P.calRank1st(P.PAR_R); P.calRank1st(P.PAR_T)
    H.run(3, uRf=uR, uTf=uT); H.run(4)
    g = P.get()
    assert np.array_equal(H.uR, g["uR"]) and np.array_equal(H.uT, g["uT"])
    sc = P.scalars()
    assert np.allclose(H.scal[6:10], sc[8:12], atol=0) and np.allclose(H.scal[10:12], sc[12:14], atol=0)
    P.close()


def test_resample_is_systematic(pfh):
    """each support point i is copied floor(n w_i) or ceil(n w_i) times (systematic resampling with one u0);
    the new prior is 1/u (PARTICLE_PRIOR_ONE), normalised; top = argmax u"""
    rng = np.random.default_rng(2)
    mLR, mLT = 125, 9
    for trial in range(20):
        H = HostParticle(pfh, mLR, mLT, seed=trial)
        r, _ = _cloud(rng, mLR); t = rng.normal(size=(mLT, 2))
        H.r[:] = r.T; H.t[:] = t.T
        H.wR[:] = rng.uniform(0.5, 1.5, mLR); H.wR /= H.wR.sum()
        H.wT[:] = rng.uniform(0.5, 1.5, mLT); H.wT /= H.wT.sum()
        H.uR[:] = rng.uniform(0, 1, mLR) ** 8; H.uT[:] = rng.uniform(0, 1, mLT) ** 2
        H.uR[rng.integers(0, mLR, 30)] = 0.0
        wpost = H.wR * H.uR; wpost /= wpost.sum()
        key = {tuple(np.round(q, 12)): (w, u) for q, w, u in zip(r, wpost, H.uR)}
        top = r[np.argmax(H.uR)]
        tpost = H.wT * H.uT; tpost /= tpost.sum()
        tkey = {tuple(np.round(x, 12)): w for x, w in zip(t, tpost)}
        H.run(6)
        out = H.r.T
        counts = {}
        for q in out:
            counts[tuple(np.round(q, 12))] = counts.get(tuple(np.round(q, 12)), 0) + 1
        for k_, c in counts.items():
            assert k_ in key
            e = mLR * key[k_][0]
            assert np.floor(e) - 1e-9 <= c <= np.ceil(e) + 1e-9
        for k_, (w, u) in key.items():
            if mLR * w >= 1.0:
                assert k_ in counts
        assert np.allclose(H.scal[6:10], top)
        prior = np.array([1.0 / key[tuple(np.round(q, 12))][1] for q in out]); prior /= prior.sum()
        assert np.allclose(H.wR, prior, rtol=1e-12)
        tc = {}
        for x in H.t.T:
            tc[tuple(np.round(x, 12))] = tc.get(tuple(np.round(x, 12)), 0) + 1
        for k_, c in tc.items():
            e = mLT * tkey[k_]
            assert np.floor(e) - 1e-9 <= c <= np.ceil(e) + 1e-9
        assert abs(H.wT.sum() - 1) < 1e-12


def test_perturb_statistics_match_reference(pfh, ref):
    """perturb(pf, R/T): same distribution as the reference (different random numbers): compare the spread of the
    perturbed cloud about its mean over many repetitions"""
    rng = np.random.default_rng(3)
    mLR, mLT = 125, 9
    k = (4e-4, 4e-4, 4e-4)
    r, q0 = _cloud(rng, mLR, k)
    t = rng.normal(scale=1.0, size=(mLT, 2))
    ref.lib().ref_set_seed(5)

    def spread(rr):
        d = synth.quat_mul(rr, np.broadcast_to(q0 * [1, -1, -1, -1], rr.shape))
        d *= np.sign(d[:, :1])
        return (d[:, 1:] ** 2).sum(1).mean()

    s_ref, s_our, st_ref, st_our = [], [], [], []
    for rep in range(400):
        P = _ref_particle(ref, r, t, k=k)
        sc = P.scalars(); sc[0:3] = k; sc[3:5] = (0.8, 1.2); sc[8:12] = q0
        P.set_scalars(sc); P.set(r=r, t=t)
        P.perturb(2.0, P.PAR_R); P.perturb(2.0, P.PAR_T)
        g = P.get()
        s_ref.append(spread(g["r"])); st_ref.append(((g["t"] - t) ** 2).mean(0))
        P.close()
        H = HostParticle(pfh, mLR, mLT, seed=100 + rep)
        H.r[:] = r.T; H.t[:] = t.T; H.scal[0:3] = k; H.scal[3:5] = (0.8, 1.2)
        H.run(1, 2.0); H.run(2, 2.0)
        s_our.append(spread(H.r.T)); st_our.append(((H.t.T - t) ** 2).mean(0))
        assert abs(np.linalg.norm(H.r, axis=0) - 1).max() < 1e-12
        assert abs(H.wR.sum() - 1) < 1e-12 and abs(H.wT.sum() - 1) < 1e-12
    assert abs(np.mean(s_our) / np.mean(s_ref) - 1) < 0.05
    assert np.allclose(np.mean(st_our, 0), np.mean(st_ref, 0), rtol=0.12)
    assert np.allclose(np.mean(st_our, 0), [(0.8 * 2) ** 2, (1.2 * 2) ** 2], rtol=0.1)


def test_load_and_stop_rule(pfh, ref):
    rng = np.random.default_rng(4)
    mLR, mLT = 125, 9
    q0 = synth.random_quats(1, rng)[0]
    H = HostParticle(pfh, mLR, mLT, seed=11)
    H.scal[0:3] = (2e-4, 3e-4, 4e-4); H.scal[3:5] = (1.5, 0.7); H.scal[6:10] = q0; H.scal[10:12] = (1.0, -2.0)
    H.run(100)
    assert abs(np.linalg.norm(H.r, axis=0) - 1).max() < 1e-12
    d = np.abs(H.r.T @ q0)
    assert np.median(d) > 0.999 and np.quantile(d, 0.1) > 0.9    # cloud about +-q0 (the ACG has heavy tails)
    assert (H.r.T @ q0 > 0).sum() not in (0, mLR)            # both hemisphere signs occur, as in Particle::load
    assert np.allclose(H.t.mean(1), (1.0, -2.0), atol=1.5)
    # calVari ran: k of the order of the input concentration
    assert np.all(H.scal[0:3] > 2e-5) and np.all(H.scal[0:3] < 4e-3)
    # stop rule: first check (phase == minPhase) always resets the counter; a non-decreasing variance then stops
    assert H.run(101, 2.0) == 0
    assert H.run(101, 3.0) == 0
    assert H.run(101, 4.0) == 1


# ------------------------------------------------------------------------------------------- MODE_2D rotation operators
def test_vms_operators_match_reference(pfh):
    """thb_pf2d.cuh against the reference's von Mises-like family (src/Geometry/DirectionalStat.cpp:252-384): inferVMS and pdfVMS
    are deterministic (1e-12); sampleVMS draws from a different random stream by design, so the sample's mean resultant length
    (the statistic inferVMS uses) and its circular symmetry are compared within sampling error"""
    from oracle import refapi as ref
    if not ref.available():
        pytest.skip("oracle/_ref not present")
    rng = np.random.default_rng(4)
    pt = lambda a: a.ctypes.data_as(_p)
    # infer: component-major layout r[c * mLR + i]
    for n in (9, 100):
        phi = rng.normal(scale=0.3, size=n) + 1.0
        cs = np.stack([np.cos(phi), np.sin(phi)], 1)
        r = np.zeros((4, n)); r[0], r[1] = cs[:, 0], cs[:, 1]
        mu = np.zeros(2); k = np.zeros(1)
        pfh.pfh_infer_vms(n, pt(r), pt(mu), pt(k))
        mu0, k0 = ref.infer_vms(cs)
        assert np.allclose(mu, mu0, atol=1e-12) and abs(k[0] - k0) <= 1e-12
    # pdf: both branches (kappa < 5: Bessel form, else the Gaussian approximation)
    for k in (0.9, 0.5, 0.2, 0.05, 0.01):
        for ang in (0.0, 0.4, 2.5):
            x = np.array([np.cos(ang), np.sin(ang)]); mu = np.array([np.cos(0.1), np.sin(0.1)])
            want = ref.pdf_vms(x, mu, k)
            assert abs(pfh.pfh_pdf_vms(pt(x), pt(mu), k) - want) <= 1e-10 * max(want, 1e-300) + 1e-300
    # sampling: uniform branch (kappa < 0.1), rejection branch at several concentrations
    n = 40000
    for k in (0.97, 0.6, 0.2, 0.02, 1e-3):
        mine = np.zeros((n, 2))
        pfh.pfh_sample_vms(12345, k, n, pt(mine))
        theirs = ref.sample_vms(k, n)
        assert np.allclose(np.hypot(mine[:, 0], mine[:, 1]), 1.0, atol=1e-12)
        R1, R0 = mine[:, 0].mean(), theirs[:, 0].mean()                      # mean resultant length along mu = (1, 0)
        se = max(mine[:, 0].std(), theirs[:, 0].std()) * np.sqrt(2.0 / n)
        assert abs(R1 - R0) <= 5 * se + 1e-9, (k, R1, R0)
        assert abs(mine[:, 1].mean()) <= 5 * mine[:, 1].std() / np.sqrt(n) + 1e-9   # symmetric about the mode
        assert abs(mine[:, 1].std() - theirs[:, 1].std()) <= 0.03 * theirs[:, 1].std() + 1e-9
    # perturb: every support point turned by its own draw; k1 re-inferred from the cloud grows accordingly
    m = 2000
    phi0 = 0.7
    r = np.zeros((4, m)); r[0], r[1] = np.cos(phi0), np.sin(phi0)
    pfh.pfh_perturb_r_2d(m, pt(r), 0.01, 2.0, 99)
    assert np.allclose(np.hypot(r[0], r[1]), 1.0, atol=1e-12) and not r[2:].any()
    d = np.arctan2(r[1], r[0]) - phi0
    ref_d = ref.sample_vms(min(1.0, 0.01 * 2.0), m)
    assert abs(np.std(d) - np.std(np.arctan2(ref_d[:, 1], ref_d[:, 0]))) <= 0.1 * np.std(d)


def test_class_resampling_and_2d_balance_match_reference(pfh):
    """resample(n, PAR_C) and balanceWeight(PAR_R) of a MODE_2D particle: the balance is deterministic (1e-10 against inferVMS /
    pdfVMS composed as the reference does); the class resampling draws its shuffle and u0 from another stream, so ours and the
    reference's (run on the reference's own Particle class) are both checked for the invariants of systematic resampling:
    count_i in {floor, ceil}(n w_i u_i / sum), prior 1 / u of the source class, top class = argmax u"""
    from oracle import refapi as ref
    if not ref.available():
        pytest.skip("oracle/_ref not present")
    rng = np.random.default_rng(8)
    pt = lambda a: a.ctypes.data_as(_p)
    n = 50
    phi = rng.normal(scale=0.4, size=n) - 0.5
    cs = np.stack([np.cos(phi), np.sin(phi)], 1)
    r = np.zeros((4, n)); r[0], r[1] = cs[:, 0], cs[:, 1]
    w = np.zeros(n)
    pfh.pfh_balance_r_2d(n, pt(r), pt(w))
    assert np.allclose(w, ref.balance_r_2d(cs), rtol=1e-10)

    nIn, nOut = 20, 12
    c0 = np.arange(nIn, dtype=np.int32); wC0 = rng.uniform(0.5, 1.5, nIn); uC0 = rng.uniform(0.0, 1.0, nIn) ** 3
    expect = nOut * wC0 * uC0 / np.sum(wC0 * uC0)

    def check(cOut, wOut, top):
        cnt = np.bincount(cOut, minlength=nIn)
        assert cnt.sum() == nOut
        assert np.all(cnt >= np.floor(expect - 1e-9)) and np.all(cnt <= np.ceil(expect + 1e-9))
        want = 1.0 / uC0[cOut]                                   # the reference normalises all weights at the end of resample()
        assert np.allclose(wOut / wOut.sum(), want / want.sum(), rtol=1e-12)
        assert top == int(np.argmax(uC0))

    for seed in range(5):
        c, wC, uC = c0.copy(), wC0.copy(), uC0.copy()
        cOut = np.zeros(nOut, np.int32); wOut = np.zeros(nOut)
        top = pfh.pfh_resample_c(nIn, pt(c), pt(wC), pt(uC), nOut, pt(cOut), pt(wOut), 1000 + seed)
        check(cOut, wOut, top)
        check(*ref.particle_resample_c(c0, wC0, uC0, nOut))


# ------------------------------------------------------------------------------------------- replay: exact parity of the
# stochastic operators.  The reference's engine is swapped for the library's bit generator (oracle/ref_harness.cpp:
# ref_rng_replay), GSL's own distribution code runs on top of it, and the reference's Particle consumes it in its own order.
def test_gsl_entry_points_bit_exact(pfh, ref):
    """pf::Rng's restatement of gsl_rng_uniform / gsl_ran_gaussian / gsl_rng_uniform_int / gsl_ran_flat /
    gsl_ran_bivariate_gaussian against GSL 2.4 itself (the reference's vendored copy) on the same bit stream"""
    n = 5000
    for kind, a, b in ((0, 0, 0), (1, 1.0, 0), (1, 0.37, 0), (2, 125, 0), (2, 9, 0), (2, 1, 0), (2, 3000000000, 0), (3, -1.0, 1.0),
                       (3, 0.0, 1.0 / 125), (4, 1.3, 0.7)):
        with ref.replay(11, 5, 77):
            want = ref.rng_draw(kind, n, a, b)
        got = np.zeros_like(want)
        pfh.pfh_rng_draw(11, 5, 77, kind, n, a, b, got.ctypes.data_as(_p))
        assert np.array_equal(got, want), (kind, a, b, np.abs(got - want).max())


_worst = [0.0, 0, 0]


def _same_state(H, P, tol=1e-9, weights=True, sync=True):
    """tol: the rotations are products with the ACG mean, the top eigenvector of a matrix whose condition number is 1 / k
    (1e4 .. 1e7 for the clouds of a converging filter) - Eigen's LLT / inverse / eigen-solver and the library's cofactor
    inverse / Jacobi sweep agree to 1e-16 x that, divided by the eigenvalue gap for the eigenvector: observed 3e-8 for clouds of
    1e-2 rad and 1e-6 for clouds that cover a good part of the sphere (where the mean itself is ill-defined)."""
    g = P.get()
    if tol < 1.0:
        _worst[0] = max(_worst[0], np.abs(H.r.T - g["r"]).max())
        _worst[1] += 1
    else:
        _worst[2] += 1
    assert np.abs(H.r.T - g["r"]).max() <= tol, np.abs(H.r.T - g["r"]).max()
    assert np.abs(H.t.T - g["t"]).max() <= 1e-11
    if weights:
        assert np.allclose(H.wR / H.wR.sum(), g["wR"] / g["wR"].sum(), rtol=1e-4, atol=0)   # 1 / pdfACG: x' A^-1 x of a near-singular A, squared
        assert np.allclose(H.wT / H.wT.sum(), g["wT"] / g["wT"].sum(), rtol=1e-9, atol=0)
    if sync:
        # carry on from the reference's state: the comparison of the NEXT operator starts from bit-identical inputs (the
        # reference's inferACG stops on `diff > 1e-3` and returns the last-but-one iterate: its result is discontinuous in its
        # input, so differences of 1e-8 would otherwise be amplified to 1e-4 within a few phases - in the reference itself, too)
        H.r[:] = g["r"].T; H.t[:] = g["t"].T; H.wR[:] = g["wR"]; H.wT[:] = g["wT"]
        sc = P.scalars()
        H.scal[0:5] = sc[0:5]



def _acg_tol(ref, r):
    """how well the ACG mean of the cloud r [n][4] is determined at all: 1e-11 x the condition number of the inferred A.
    The reference infers A by a fixed-point iteration that stops on `diff > 1e-3` and returns the last-but-one iterate
    (DirectionalStat.cpp:93-145); for a cloud collapsed onto a handful of distinct points A is rank-deficient (condition 1e9
    and more), the iteration count depends on the last bits, and two correct implementations differ by 1e-4 in the mean."""
    rr = np.ascontiguousarray(r, np.float64)
    A = np.empty(16); k = np.empty(3); m = np.empty(4)
    ref.lib().ref_inferACG(rr.ctypes.data_as(_p), len(rr), A.ctypes.data_as(_p), k.ctypes.data_as(_p), m.ctypes.data_as(_p))
    e = np.linalg.eigvalsh(A.reshape(4, 4))
    cond = e[-1] / max(e[0], 1e-300)
    return 4.0 if cond > 1e9 else float(np.clip(1e-11 * cond, 1e-9, 1e-2))     # rank-deficient: nothing to compare


def test_load_replay_exact(pfh, ref):
    rng = np.random.default_rng(21)
    for trial in range(5):
        mLR, mLT = (125, 9) if trial else (37, 5)
        q0 = synth.random_quats(1, rng)[0]; t0 = rng.normal(size=2)
        k = rng.uniform(1e-5, 1e-3, 3); s01 = rng.uniform(0.5, 2.0, 2)
        H = HostParticle(pfh, mLR, mLT, seed=99)
        H.scal[0:3] = k; H.scal[3:5] = s01; H.scal[6:10] = q0; H.scal[10:12] = t0
        H.run(100)                                   # epoch 1
        P = ref.Particle(mLR, mLT)
        with ref.replay(99, 0, 1):
            P.load(mLR, mLT, q0, k[0], k[1], k[2], t0, s01[0], s01[1])
        _same_state(H, P, tol=1e-12)
        sc = P.scalars()
        assert np.allclose(H.scal[0:3], sc[0:3], rtol=1e-8) and np.allclose(H.scal[3:5], sc[3:5], rtol=1e-12)
        P.close()


def test_phase_sequence_replay_exact(pfh, ref):
    """perturb (L) -> [set weights, keepHalfHeightPeak, calRank1st, calVari, resample, perturb (S)] x 6 with the same random
    numbers on both sides: support points, weights, variances and the drawn (rotation, translation) indices agree to 1e-9
    after EVERY operator - a sign error in a perturbation or a biased resampler cannot hide in statistics"""
    rng = np.random.default_rng(33)
    for trial in range(6):
        mLR, mLT = (125, 9) if trial % 2 == 0 else (25, 9)
        q0 = synth.random_quats(1, rng)[0]; t0 = rng.normal(size=2)
        H = HostParticle(pfh, mLR, mLT, seed=1234 + trial)
        H.scal[0:3] = 3e-4; H.scal[3:5] = 1.0; H.scal[6:10] = q0; H.scal[10:12] = t0
        H.run(100)
        P = ref.Particle(mLR, mLT)
        seed = 1234 + trial
        with ref.replay(seed, 0, H.epoch):
            P.load(mLR, mLT, q0, 3e-4, 3e-4, 3e-4, t0, 1.0, 1.0)
            _same_state(H, P, 1e-12)
            for phase in range(6):
                pfac = 2.0 if phase == 0 else 0.5
                tol = _acg_tol(ref, H.r.T)
                H.run(1, pfac); ref.rng_key(seed, 0, H.epoch); P.perturb(pfac, P.PAR_R)
                _same_state(H, P, tol, weights=tol < 1e-6)
                H.run(2, pfac); ref.rng_key(seed, 0, H.epoch); P.perturb(pfac, P.PAR_T)
                _same_state(H, P)
                # likelihood weights of this phase: peaked around a random support point, as the E kernel would return them
                c = rng.integers(mLR)
                d2 = 1 - np.abs(H.r.T @ H.r.T[c]) ** 2
                uR = np.exp(-d2 / (2 * np.quantile(d2, 0.5) * rng.uniform(0.5, 2.0))).astype(np.float32)   # about half of the support survives
                uT = np.exp(-0.5 * ((H.t.T - H.t.T[rng.integers(mLT)]) ** 2).sum(1) / rng.uniform(0.2, 2.0)).astype(np.float32)
                P.set_u(1, uR.astype(np.float64)); P.keepHalfHeightPeak(P.PAR_R); P.set_u(2, uT.astype(np.float64))
                P.calRank1st(P.PAR_R); P.calRank1st(P.PAR_T)
                H.run(3, uRf=uR, uTf=uT); H.run(4)
                g = P.get()
                assert np.array_equal(H.uR, g["uR"]) and np.array_equal(H.uT, g["uT"])
                tol = _acg_tol(ref, H.r.T)
                H.run(5); ref.rng_key(seed, 0, H.epoch); P.calVari(P.PAR_R); P.calVari(P.PAR_T)
                sc = P.scalars()
                assert np.allclose(H.scal[0:3], sc[0:3], rtol=max(1e-5, 10 * tol)) and np.allclose(H.scal[3:5], sc[3:5], rtol=1e-10)
                _same_state(H, P, tol, weights=False)
                H.run(6); ref.rng_key(seed, 0, H.epoch); P.resample(mLR, P.PAR_R); P.resample(mLT, P.PAR_T)
                _same_state(H, P)
                sc = P.scalars()
                assert np.allclose(H.scal[6:10], sc[8:12], atol=1e-5) and np.allclose(H.scal[10:12], sc[12:14], atol=1e-11)   # topR, topT
            # Particle::rand x min(100, mLR)
            nDraw = min(100, mLR)
            H.run(102, float(nDraw)); ref.rng_key(seed, 0, H.epoch)
            g = P.get()
            for m in range(nDraw):
                cls = np.zeros(1, np.int32); q = np.zeros(4); t = np.zeros(2); d = np.zeros(1)
                P.rand(cls.ctypes.data_as(_p), q.ctypes.data_as(_p), t.ctypes.data_as(_p), d.ctypes.data_as(_p))
                iR, iT = divmod(int(H.uR[m]), 1000)
                assert np.array_equal(q, g["r"][iR]) and np.array_equal(t, g["t"][iT])
        P.close()
    print("largest support-point difference over the sequences:", _worst[0], "in", _worst[1], "comparisons;", _worst[2], "rank-deficient clouds skipped")
    assert _worst[2] <= _worst[1] // 20


def test_defocus_dimension_replay_exact(pfh, ref):
    """the CTF search's defocus dimension (initD, perturb / balanceWeight / calVari / calRank1st / resample of PAR_D) against the
    reference's Particle with the same random numbers: exact to 1e-12 (one-dimensional Gaussian statistics, nothing ill-posed)"""
    rng = np.random.default_rng(91)
    for trial in range(4):
        mLD = (9, 5, 2, 9)[trial]
        seed = 700 + trial
        P = ref.Particle(25, 9)
        q0 = synth.random_quats(1, rng)[0]
        d = np.zeros(mLD + 1); wD = np.zeros(mLD); uD = np.zeros(mLD); scal = np.zeros(20)
        pt = lambda a: None if a is None else a.ctypes.data_as(_p)
        epoch = 0
        with ref.replay(seed, 0, 0):
            P.load(25, 9, q0, 3e-4, 3e-4, 3e-4, np.zeros(2), 1.0, 1.0)

            def both(op, arg, fn, uDf=None):
                nonlocal epoch
                epoch += 1
                assert pfh.pfh_run_d(op, arg, mLD, pt(d), pt(wD), pt(uD), pt(uDf), pt(scal), seed, 3, epoch) == 0
                ref.rng_key(seed, 3, epoch)
                fn()
                g = P.get_d()
                assert np.abs(g["d"] - d[:mLD]).max() <= 1e-12, (op, np.abs(g["d"] - d[:mLD]).max())
                assert np.allclose(g["wD"] / g["wD"].sum(), wD / wD.sum(), rtol=1e-10), op

            both(200, 0.05, lambda: P.initD(mLD, 0.05))
            for phase in range(5):
                uDf = np.exp(-0.5 * ((d[:mLD] - 1.01) / 0.03) ** 2).astype(np.float32) + np.float32(1e-3)
                both(205, 0.0, lambda: (P.set_u(3, uDf.astype(np.float64)), P.calRank1st(P.PAR_D)), uDf)
                both(202, 0.0, lambda: P.calVari(P.PAR_D))
                assert abs(scal[19] - P.variD()) <= 1e-14
                both(203, 0.0, lambda: P.resample(mLD, P.PAR_D))
                assert abs(d[mLD] - P.scalars()[14]) <= 1e-14                 # _topD
                both(201, 0.5, lambda: P.perturb(0.5, P.PAR_D))
        P.close()


@pytest.mark.parametrize("mode2D", [0, 1])
def test_scan_hand_over_replay_exact(pfh, ref, mode2D):
    """from the global scan to the support of the local phases (src/Optimiser.cpp:921-1075): class choice, peak factors, resampling of
    the scan grid down to (mLR, mLT), variances and floors - against the reference's Particle driven through the same calls with the
    same random numbers.  Chosen class and support points identical; variances to 1e-9 (3D: where the ACG of the resampled support is
    well-posed - it holds many duplicates; 2D: von Mises, always)."""
    rng = np.random.default_rng(77 + mode2D)
    pt = lambda a: a.ctypes.data_as(_p)
    skipped = 0
    for trial in range(12):
        nK = (1, 3, 20)[trial % 3]
        nR, nT = (200, 30) if trial % 2 == 0 else (1000, 45)
        mLR, mLT = (9, 9) if mode2D else (125, 9)
        if mode2D:
            ang = rng.uniform(-np.pi, np.pi, nR); gridR = np.stack([np.cos(ang), np.sin(ang)], 1)
        else:
            gridR = synth.random_quats(nR, rng)
        gridT = rng.normal(scale=2.0, size=(nT, 2))
        # scan weights: a peak around one grid point, per class, different heights
        best = rng.integers(nR)
        if mode2D:
            d2 = 1 - (gridR @ gridR[best]) ** 1
        else:
            d2 = 1 - np.abs(gridR @ gridR[best]) ** 2
        wR = np.stack([np.exp(-d2 / (2 * np.quantile(d2, 0.05 + 0.1 * rng.random()))) * rng.uniform(0.2, 1.0, nR) for _ in range(nK)]).astype(np.float32)
        wT = np.stack([np.exp(-0.5 * ((gridT - gridT[rng.integers(nT)]) ** 2).sum(1) / rng.uniform(0.5, 4.0)) for _ in range(nK)]).astype(np.float32)
        wC = (rng.uniform(0.9, 1.0, nK) ** 2).astype(np.float32)
        wC[rng.integers(nK)] = 1.0
        kFloor, sFloor = (1e-2 if mode2D else 1e-4), 0.05
        seed, stream, epoch = 4000 + trial, 7, 3
        r = np.zeros((4, mLR)); t = np.zeros((2, mLT)); owR = np.zeros(mLR); owT = np.zeros(mLT); scal = np.zeros(20)
        cls = pfh.pfh_from_scan(mode2D, nK, nR, nT, pt(np.ascontiguousarray(gridR)), pt(np.ascontiguousarray(gridT)), pt(wC), pt(wR), pt(wT), mLR, mLT,
                                kFloor, sFloor, pt(r), pt(t), pt(owR), pt(owT), pt(scal), seed, stream, epoch)
        with ref.replay(seed, stream, epoch):
            want = ref.particle_from_scan(mode2D, gridR, gridT, wC, wR, wT, mLR, mLT, kFloor, sFloor, (seed, stream, epoch))
        assert cls == want["cls"], trial
        # grid points, copied (calVari of the 3D filter turns the support by the conjugate of its mean and back: last-bit rounding)
        assert np.abs(r.T[:, :2 if mode2D else 4] - want["r"][:, :2 if mode2D else 4]).max() <= 1e-12, trial
        assert np.array_equal(t.T, want["t"])
        assert np.allclose(owR, want["wR"], rtol=1e-12) and np.allclose(owT, want["wT"], rtol=1e-12)
        assert np.allclose(scal[3:5], want["scal"][3:5], rtol=1e-12)                                       # s0, s1
        assert np.array_equal(scal[6:10][:2 if mode2D else 4], want["scal"][8:12][:2 if mode2D else 4])    # topR
        if mode2D:
            assert np.isclose(scal[0], want["scal"][0], rtol=1e-12)
        else:
            tol = _acg_tol(ref, want["r"])
            if tol < 1.0:
                assert np.allclose(scal[0:3], want["scal"][0:3], rtol=max(1e-6, 100 * tol)), (trial, scal[0:3], want["scal"][0:3])
            else:
                skipped += 1
    assert skipped <= 4

"""The host-side mirror of the reference's accelerator seam (thunder_b200/host/Interface.{h,cpp}): the same
function names and argument order as gpu/interface/Interface.h for the hot path, on top of the C ABI.

CPU: the library exports the mirrored names (C++ linkage) and, where the reference tree is present, the
signatures are checked against the reference's own header text.  GPU: calling through the shims gives the
oracle's numbers."""
import ctypes as C
import os
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

from thunder_b200 import synth

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "thunder_b200" / "lib" / "libthb_interface.so"
REF_HDR = Path(os.environ.get("THB_REFERENCE", "/root/reference")) / "gpu" / "interface" / "Interface.h"
MIRRORED = ["getAviDevice", "ExpectPreidx", "ExpectFreeIdx", "ExpectRotran", "ExpectProject", "ExpectGlobal3D", "InsertFT",
            "ExpectGlobal2D", "InsertI2D", "ExpectPrefre", "ExpectLocalIn", "ExpectLocalV2D", "ExpectLocalV3D", "ExpectLocalP",
            "ExpectLocalHostA", "ExpectLocalRTD", "ExpectLocalPreI2D", "ExpectLocalPreI3D", "ExpectLocalM", "ExpectLocalHostF",
            "ExpectLocalFin", "PrepareTF"]
LOCAL_SEAM = MIRRORED[9:-1]
_p, _i = C.c_void_p, C.c_int


def _ptr(a):
    return None if a is None else a.ctypes.data_as(_p)


def test_shim_exports_mirrored_names():
    assert LIB.exists(), "run `make` / __graft_entry__.build()"
    out = subprocess.check_output(["nm", "-DC", os.fspath(LIB)]).decode()
    for name in MIRRORED + ["ExpectLocalBatch"]:
        assert re.search(rf"\bT {name}\(", out), name


def _params(text, name):
    """argument names of the first declaration of `name` in a header, in order"""
    text = re.sub(r"//[^\n]*", "", text)          # the reference's header carries commented-out parameters (ExpectLocalM)
    m = re.search(rf"\bvoid\s+{name}\s*\((.*?)\)\s*;", text, flags=re.S)
    assert m, name
    args = [a.strip() for a in m.group(1).replace("\n", " ").split(",")]
    return [re.sub(r".*[\s\*&]", "", a) for a in args]


@pytest.mark.skipif(not REF_HDR.exists(), reason="reference tree not present")
def test_argument_order_matches_reference_header():
    """same argument names in the same order as gpu/interface/Interface.h, THUNDER class arguments replaced as
    thunder_b200/host/Interface.h documents (Volume& -> pointer + vdim, MPI_Comm& dropped)"""
    ref = REF_HDR.read_text()
    ours = (ROOT / "thunder_b200" / "host" / "Interface.h").read_text()
    for name in ["ExpectPreidx", "ExpectRotran", "ExpectProject", "ExpectGlobal3D", "ExpectGlobal2D"] + LOCAL_SEAM:
        assert _params(ours, name) == _params(ref, name), name
    assert _params(ours, "InsertI2D") == [a for a in _params(ref, "InsertI2D") if a not in ("hemi", "slav")]
    want = [a for a in _params(ref, "InsertFT") if a not in ("hemi", "slav")]
    got = [a for a in _params(ours, "InsertFT") if a != "vdim"]
    assert got == want
    assert [a for a in _params(ours, "PrepareTF") if a != "vdim"] == _params(ref, "PrepareTF")


@pytest.fixture(scope="module")
def shim():
    L = C.CDLL(os.fspath(LIB))
    L.thbi_device_count.restype = _i
    if L.thbi_device_count() == 0:
        pytest.skip("no GPU")
    yield L
    L.thbi_shutdown()


@pytest.mark.gpu
def test_scan_and_local_through_the_shims(shim):
    from oracle import portapi as port
    N, pf = 32, 2
    rng = np.random.default_rng(5)
    vol = synth.padded_ft(synth.phantom(N, 8, seed=2), pf)
    pix = port.pixel_list(N, pf, 14.0, 1.0)
    iCol, iRow = pix["iCol"], pix["iRow"]
    P = len(iCol)
    nImg, nR, nT = 5, 20, 7
    par = synth.make_particles(nImg, N, pix, lambda q: np.stack([port.project(vol, pf, port.rotate3D(x), iCol, iRow) for x in q]),
                               seed=3, snr_scale=4.0)
    rot = synth.random_quats(nR, rng); rot[:nImg] = par["quat"]
    trans = rng.normal(scale=1.5, size=(nT, 2))
    pR = np.full(nR, 1.0 / nR); pT = np.full(nT, 1.0 / nT)
    traP = np.zeros((nT, P), np.complex64); rotMat = np.zeros((nR, 9)); rotP = np.zeros((nR, P), np.complex64)
    shim.thbi_ExpectRotran(_ptr(traP), _ptr(trans), _ptr(rot), _ptr(rotMat), _ptr(iCol), _ptr(iRow), nR, nT, N, P)
    for r in range(nR):      # column-major 3x3, as the oracle's rotate3D (Euler.cpp:181-189)
        assert np.allclose(rotMat[r], port.rotate3D(rot[r]), atol=1e-14)
    volc = np.ascontiguousarray(vol)
    shim.thbi_ExpectProject(_ptr(volc), _ptr(rotP), _ptr(rotMat), _ptr(iCol), _ptr(iRow), nR, pf, 1, N * pf, P)
    nK = 2
    wC = np.zeros((nImg, nK), np.float32); wR = np.zeros((nImg, nK, nR), np.float32); wT = np.zeros((nImg, nK, nT), np.float32)
    baseL = np.zeros(nImg, np.float32)
    dat = np.ascontiguousarray(par["dat"]); ctf = np.ascontiguousarray(par["ctf"]); sig = np.ascontiguousarray(par["sigRcp"])
    for k in range(nK):      # the same volume as two "classes": equal weights, one shared baseline
        shim.thbi_ExpectGlobal3D(_ptr(rotP), _ptr(traP), _ptr(dat), _ptr(ctf), _ptr(sig), _ptr(wC), _ptr(wR), _ptr(wT), _ptr(pR), _ptr(pT),
                                 _ptr(baseL), k, nK, nR, nT, P, nImg)
    for l in range(nImg):
        o = port.expect_local(vol, pf, N, iCol, iRow, par["dat"][l], par["ctf"][l], par["sigRcp"][l], rot, trans, pR, pT)
        assert abs(baseL[l] - o["base"]) <= 2e-6 * abs(o["base"]) + 1e-4
        big = o["uR"] > 1e-5 * o["uR"].max()
        for k in range(nK):
            assert np.allclose(wR[l, k][big], o["uR"][big], rtol=5e-3)
            assert np.allclose(wC[l, k], o["uC"], rtol=5e-3)
    # batched local search through the shim == the scan restricted to one image's own cloud
    quat = np.stack([synth.acg_cloud(par["quat"][l], 1e-4, nR, rng) for l in range(nImg)])
    tran = par["tran"][:, None, :] + rng.normal(scale=0.5, size=(nImg, nT, 2))
    wRp = np.full((nImg, nR), 1.0 / nR); wTp = np.full((nImg, nT), 1.0 / nT)
    uC = np.zeros(nImg, np.float32); uR = np.zeros((nImg, nR), np.float32); uT = np.zeros((nImg, nT), np.float32); bl = np.zeros(nImg, np.float32)
    shim.thbi_ExpectLocalBatch(0, _ptr(volc), N * pf, pf, N, _ptr(iCol), _ptr(iRow), P, _ptr(dat), _ptr(ctf), _ptr(sig), nImg, nR, nT,
                               _ptr(quat), _ptr(tran), _ptr(wRp), _ptr(wTp), _ptr(uC), _ptr(uR), _ptr(uT), _ptr(bl))
    for l in range(nImg):
        o = port.expect_local(vol, pf, N, iCol, iRow, par["dat"][l], par["ctf"][l], par["sigRcp"][l], quat[l], tran[l], wRp[l], wTp[l])
        assert abs(bl[l] - o["base"]) <= 2e-6 * abs(o["base"]) + 1e-4
        big = o["uT"] > 1e-5 * o["uT"].max()
        assert np.allclose(uT[l][big], o["uT"][big], rtol=5e-3)


@pytest.mark.gpu
def test_insertft_through_the_shim_accumulates_into_the_callers_volumes(shim):
    from oracle import portapi as port
    N, pf = 32, 2
    rng = np.random.default_rng(8)
    pixM = port.pixel_list(N, pf, 15.0, 0.0)
    PM = len(pixM["iCol"])
    nImg, mReco, vdim = 4, 5, N * pf
    datM = (rng.normal(size=(nImg, PM)) + 1j * rng.normal(size=(nImg, PM))).astype(np.complex64)
    ctfM = rng.uniform(-1, 1, (nImg, PM)).astype(np.float32)
    nr = synth.random_quats(nImg * mReco, rng).reshape(nImg, mReco, 4)
    nt = rng.normal(scale=2.0, size=(nImg, mReco, 2))
    w = np.full(nImg, 1.0 / mReco, np.float32)
    offS = rng.normal(scale=0.5, size=(nImg, 2))
    nVox = (vdim // 2 + 1) * vdim * vdim
    F0 = (rng.normal(size=nVox) + 1j * rng.normal(size=nVox)).astype(np.complex64)     # contents already in the caller's volumes
    T0 = rng.uniform(0, 1, nVox).astype(np.complex64)
    F3D, T3D = F0.copy(), T0.copy()
    O3D = np.array([1.0, 2.0, 3.0]); counter = np.array([7], np.int32)
    shim.thbi_InsertFT(_ptr(F3D), _ptr(T3D), vdim, _ptr(O3D), _ptr(counter), _ptr(datM), _ptr(ctfM), _ptr(offS), _ptr(w), _ptr(nr), _ptr(nt),
                       _ptr(pixM["iColPad"]), _ptr(pixM["iRowPad"]), pf, PM, mReco, N, nVox, nImg)
    want = port.insert_loop(vdim, pf, N, datM, ctfM, w, offS, nr, nt, pixM["iCol"], pixM["iRow"])
    rel = lambda a, b: np.linalg.norm((a - b).ravel()) / np.linalg.norm(b.ravel())
    assert rel(F3D - F0, want["F"].ravel()) <= 1e-6
    assert rel((T3D - T0).real, want["T"].ravel()) <= 1e-6
    assert np.allclose(O3D - [1.0, 2.0, 3.0], want["O"], rtol=1e-10, atol=1e-10)
    assert counter[0] == 7 + nImg * mReco


@pytest.mark.gpu
def test_2d_classification_through_the_shims(shim):
    """ExpectGlobal2D / InsertI2D (MODE_2D): every image against every class with one shared baseline; draws scattered
    into the accumulator of their class and ADDED to the caller's arrays.  Checked against oracle/port2d.py, itself pinned to
    the reference's MODE_2D Projector / Reconstructor (tests/test_mode2d.py)."""
    from oracle import port2d
    from tests.test_mode2d import _setup, _unit
    s = _setup(N=32, pf=2, k=3, nImg=5, seed=11, rE=14.0, rM=15.0)
    N, pf, nK, nImg, rng = s["N"], s["pf"], s["k"], s["nImg"], s["rng"]
    iCol, iRow, P = s["pixE"]["iCol"], s["pixE"]["iRow"], s["PE"]
    vdim = N * pf
    vol = np.ascontiguousarray(np.stack(s["refs"]))
    nR, nT = 12, 4
    rot = np.ascontiguousarray(_unit(np.linspace(-np.pi, np.pi, nR, endpoint=False))); trans = rng.normal(scale=1.5, size=(nT, 2))
    pR = rng.uniform(0.5, 1.5, nR); pT = rng.uniform(0.5, 1.5, nT)
    wC = np.zeros((nImg, nK), np.float32); wR = np.zeros((nImg, nK, nR), np.float32); wT = np.zeros((nImg, nK, nT), np.float32)
    shim.thbi_ExpectGlobal2D(_ptr(vol), _ptr(s["datE"]), _ptr(s["ctfE"]), _ptr(s["sigE"]), _ptr(trans), _ptr(wC), _ptr(wR), _ptr(wT), _ptr(pR),
                             _ptr(pT), _ptr(rot), _ptr(iCol), _ptr(iRow), nK, nR, nT, pf, N, vdim, P, nImg)
    for l in range(nImg):
        L = np.empty((nK, nR, nT))
        for k in range(nK):
            for r in range(nR):
                p = port2d.project2d(s["refs"][k], pf, rot[r], iCol, iRow)
                for t in range(nT):
                    d = s["datE"][l].astype(np.complex128) - s["ctfE"][l] * port2d.translate(p, trans[t, 0], trans[t, 1], N, iCol, iRow).astype(np.complex128)
                    L[k, r, t] = np.sum((d.real ** 2 + d.imag ** 2) * s["sigE"][l].astype(np.float64))
        e = np.exp(L - L.max())
        want_R = e @ pT; want_T = np.einsum("krt,r->kt", e, pR); want_C = np.einsum("krt,r,t->k", e, pR, pT)
        big = want_R > 1e-4 * want_R.max()
        assert np.allclose(wR[l][big], want_R[big], rtol=5e-3)
        bigT = want_T > 1e-4 * want_T.max()
        assert np.allclose(wT[l][bigT], want_T[bigT], rtol=5e-3)
        assert np.allclose(wC[l], want_C, rtol=5e-3, atol=1e-4 * want_C.max())
    # ---- M
    pixM, PM = s["pixM"], s["PM"]
    mReco = 4
    nr = np.ascontiguousarray(_unit(rng.uniform(-np.pi, np.pi, (nImg, mReco)))); nt = rng.normal(scale=2.0, size=(nImg, mReco, 2))
    nc = rng.integers(0, nK, (nImg, mReco)).astype(np.int32)
    w = np.full(nImg, 1.0 / mReco, np.float32); offS = rng.normal(scale=0.5, size=(nImg, 2))
    n2 = vdim * (vdim // 2 + 1)
    F0 = (rng.normal(size=(nK, n2)) + 1j * rng.normal(size=(nK, n2))).astype(np.complex64); T0 = rng.uniform(0, 1, (nK, n2)).astype(np.float32)
    F2D, T2D = F0.copy(), T0.copy()
    O2D = np.arange(2.0 * nK); counter = np.full(nK, 3, np.int32)
    shim.thbi_InsertI2D(_ptr(F2D), _ptr(T2D), _ptr(O2D), _ptr(counter), _ptr(s["datM"]), _ptr(s["ctfM"]), _ptr(w), _ptr(offS), _ptr(nc), _ptr(nr),
                        _ptr(nt), _ptr(pixM["iColPad"]), _ptr(pixM["iRowPad"]), nK, pf, PM, mReco, N, vdim, nImg)
    recos = [port2d.Reco2D(vdim) for _ in range(nK)]
    for l in range(nImg):
        for m in range(mReco):
            recos[nc[l, m]].insert_draw(s["datM"][l], s["ctfM"][l], N, pixM["iCol"], pixM["iRow"], pixM["iColPad"], pixM["iRowPad"], nr[l, m], nt[l, m],
                                        offS[l], w[l])
    for k in range(nK):
        assert counter[k] == 3 + recos[k].counter
        assert np.allclose(O2D[2 * k:2 * k + 2] - [2.0 * k, 2.0 * k + 1], recos[k].O[:2], atol=1e-9)
        if recos[k].counter:
            assert np.abs((F2D[k] - F0[k]) - recos[k].F.ravel()).max() <= 3e-5 * np.abs(recos[k].F).max() + 1e-6
            assert np.abs((T2D[k] - T0[k]) - recos[k].T.ravel()).max() <= 3e-5 * np.abs(recos[k].T).max() + 1e-6


@pytest.mark.skipif(not REF_HDR.exists(), reason="reference tree not present")
def test_reference_typed_overloads_compile_against_thunder_headers():
    """-DTHB_WITH_THUNDER: the Volume& / MPI_Comm& / CTFAttr* overloads of InsertFT (with and without nC) and InsertI2D, i.e. the
    exact call expressions of Reconstructor::insertI and Optimiser::reconstructRef, compile against THUNDER's OWN headers
    (syntax-only; tests/host_bind/bind.cpp)"""
    ref = REF_HDR.parents[2]
    out = ROOT / "oracle" / "_ref"
    if not (out / "gen" / "THUNDERConfig.h").exists() or not (out / "deps" / "include").exists():
        pytest.skip("oracle/_ref build tree not present (oracle/build_ref.sh)")
    inc = [ROOT / "thunder_b200" / "host", ROOT / "include", ROOT / "oracle" / "ref_shim", out / "gen", ref / "include",
           ref / "include" / "Functions", ref / "include" / "Geometry", ref / "include" / "Image", ref / "external" / "Eigen3",
           ref / "external" / "easylogging", ref / "external" / "jsoncpp", out / "deps" / "include", out / "deps" / "boost_1_60_0"]
    obj = ROOT / "tests" / "host_bind" / "bind.o"
    cmd = ["g++", "-std=c++11", "-fopenmp", "-mavx", "-c", "-w", "-DTHB_WITH_THUNDER", "-DSINGLE_PRECISION"]
    cmd += [f"-I{os.fspath(i)}" for i in inc] + [os.fspath(ROOT / "tests" / "host_bind" / "bind.cpp"), "-o", os.fspath(obj)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-1500:]
    # ... and LINK: every seam symbol the THUNDER-typed translation unit references (mangled with THUNDER's own types, e.g.
    # Complex = struct _complex_float_t) is one the shim library exports
    want = [l.split()[-1] for l in subprocess.check_output(["nm", "-u", os.fspath(obj)]).decode().splitlines()
            if re.search(r"_Z\d+(Insert|Expect)", l)]
    have = set(l.split()[-1] for l in subprocess.check_output(["nm", "-D", "--defined-only", os.fspath(LIB)]).decode().splitlines())
    assert len(want) >= 2
    for sym in want:
        assert sym in have, sym
    obj.unlink()


@pytest.mark.gpu
def test_local_search_through_the_reference_protocol(shim):
    """ExpectLocalIn / V3D / P / HostA / RTD / PreI3D / M / HostF / Fin called in the reference's own order (one image in flight
    per slot, one synchronous ExpectLocalM per image and phase, src/Optimiser.cpp:2484-2700) give the oracle's marginal weights -
    and the same numbers as the batched entry point"""
    from oracle import portapi as port
    N, pf = 32, 2
    rng = np.random.default_rng(15)
    vol = synth.padded_ft(synth.phantom(N, 8, seed=2), pf)
    pix = port.pixel_list(N, pf, 14.0, 1.0)
    iCol, iRow = np.ascontiguousarray(pix["iCol"]), np.ascontiguousarray(pix["iRow"])
    P = len(iCol)
    nImg, nPhase, nR, nT = 5, 3, 37, 9
    par = synth.make_particles(nImg, N, pix, lambda q: np.stack([port.project(vol, pf, port.rotate3D(x), iCol, iRow) for x in q]),
                               seed=3, snr_scale=4.0)
    quat = np.stack([[synth.acg_cloud(par["quat"][l], 1e-4 * (ph + 1), nR, rng) for ph in range(nPhase)] for l in range(nImg)])
    tran = par["tran"][:, None, None, :] + rng.normal(scale=0.5, size=(nImg, nPhase, nT, 2))
    wRp = rng.uniform(0.5, 1.5, (nImg, nPhase, nR)); wRp /= wRp.sum(-1, keepdims=True)
    wTp = rng.uniform(0.5, 1.5, (nImg, nPhase, nT)); wTp /= wTp.sum(-1, keepdims=True)
    volc = np.ascontiguousarray(vol)
    dat = np.ascontiguousarray(par["dat"]); ctf = np.ascontiguousarray(par["ctf"]); sig = np.ascontiguousarray(par["sigRcp"])
    wC = np.zeros((nImg, nPhase), np.float32); wR = np.zeros((nImg, nPhase, nR), np.float32); wT = np.zeros((nImg, nPhase, nT), np.float32)
    oldC = 1.0
    shim.thbi_ExpectLocalProtocol.argtypes = [_i, _p, _i, _i, _i, _p, _p, _i, _p, _p, _p, _i, _i, _i, _i, _p, _p, _p, _p, C.c_double, _i, _p, _p, _p]
    shim.thbi_ExpectLocalProtocol(0, _ptr(volc), N * pf, pf, N, _ptr(iCol), _ptr(iRow), P, _ptr(dat), _ptr(ctf), _ptr(sig), nImg, nPhase, nR, nT,
                                  _ptr(quat), _ptr(tran), _ptr(wRp), _ptr(wTp), oldC, 2, _ptr(wC), _ptr(wR), _ptr(wT))
    for l in range(nImg):
        for ph in range(nPhase):
            o = port.expect_local(vol, pf, N, iCol, iRow, par["dat"][l], par["ctf"][l], par["sigRcp"][l], quat[l, ph], tran[l, ph], wRp[l, ph],
                                  wTp[l, ph])
            big = o["uR"] > 1e-5 * o["uR"].max()
            assert np.allclose(wR[l, ph][big], o["uR"][big], rtol=5e-3), (l, ph)
            big = o["uT"] > 1e-5 * o["uT"].max()
            assert np.allclose(wT[l, ph][big], o["uT"][big], rtol=5e-3), (l, ph)
            assert np.allclose(wC[l, ph], o["uC"], rtol=5e-3)
    # the batched entry point on the supports of the last phase: the same numbers
    ph = nPhase - 1
    uC = np.zeros(nImg, np.float32); uR = np.zeros((nImg, nR), np.float32); uT = np.zeros((nImg, nT), np.float32); bl = np.zeros(nImg, np.float32)
    q2 = np.ascontiguousarray(quat[:, ph]); t2 = np.ascontiguousarray(tran[:, ph]); r2 = np.ascontiguousarray(wRp[:, ph]); s2 = np.ascontiguousarray(wTp[:, ph])
    shim.thbi_ExpectLocalBatch(0, _ptr(volc), N * pf, pf, N, _ptr(iCol), _ptr(iRow), P, _ptr(dat), _ptr(ctf), _ptr(sig), nImg, nR, nT,
                               _ptr(q2), _ptr(t2), _ptr(r2), _ptr(s2), _ptr(uC), _ptr(uR), _ptr(uT), _ptr(bl))
    assert np.allclose(uR, wR[:, ph], rtol=2e-4, atol=1e-30) and np.allclose(uT, wT[:, ph], rtol=2e-4, atol=1e-30)


SEAM_EXE = ROOT / "oracle" / "_ref" / "seam_insertI"


def _run_seam(tmp_path, N, pf, pixM, datM, ctfM, w, offS, nr, nt, nc=None, sym=None):
    nImg, PM = datM.shape
    mReco = nr.shape[1]
    fin, fout = tmp_path / "in.bin", tmp_path / "out.bin"
    with open(fin, "wb") as f:
        np.array([N, N, pf, PM, mReco, nImg, int(nc is not None), 0], np.int32).tofile(f)
        np.array([1.32], np.float32).tofile(f)
        for k in ("iColPad", "iRowPad", "iPxl", "iSig"):
            np.ascontiguousarray(pixM[k], np.int32).tofile(f)
        np.ascontiguousarray(datM, np.complex64).tofile(f); np.ascontiguousarray(ctfM, np.float32).tofile(f)
        np.ascontiguousarray(w, np.float32).tofile(f); np.ascontiguousarray(offS, np.float64).tofile(f)
        np.ascontiguousarray(nr, np.float64).tofile(f); np.ascontiguousarray(nt, np.float64).tofile(f)
        if nc is not None:
            np.ascontiguousarray(nc, np.int32).tofile(f)
    cmd = [os.fspath(SEAM_EXE), os.fspath(fin), os.fspath(fout)] + ([sym] if sym else [])
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.stdout[-800:], r.stderr[-1500:])
    raw = np.fromfile(fout, np.uint8)
    m = int(raw[:4].view(np.int32)[0])
    nVox = m * m * (m // 2 + 1)
    o = 4
    F = raw[o:o + 8 * nVox].view(np.complex64).reshape(m, m, m // 2 + 1); o += 8 * nVox
    T = raw[o:o + 4 * nVox].view(np.float32).reshape(m, m, m // 2 + 1); o += 4 * nVox
    O = raw[o:o + 24].view(np.float64); o += 24
    return dict(F=F, T=T, O=O, counter=int(raw[o:o + 4].view(np.int32)[0]), m=m)


@pytest.mark.gpu
@pytest.mark.parametrize("sym", [None, "C4"])
def test_reference_reconstructor_insertI_runs_on_the_shim(tmp_path, sym):
    """THE REFERENCE'S OWN Reconstructor::insertI (and prepareTFG), compiled from src/Reconstructor.cpp with GPU_INSERT and this
    repository's Interface.h, linked against libthb_interface.so (oracle/build_ref.sh -> oracle/_ref/seam_insertI), against the
    reference's CPU path Reconstructor::insertP / insertDir / prepareTF / symmetrize on the same draws"""
    from oracle import refapi
    from oracle import portapi as port
    if not SEAM_EXE.exists() or not refapi.available():
        pytest.skip("oracle/_ref/seam_insertI not built (oracle/build_ref.sh needs the reference tree)")
    N, pf = 32, 2
    rng = np.random.default_rng(21)
    pixM = port.pixel_list(N, pf, 13.0, 0.0)
    PM = len(pixM["iCol"])
    nImg, mReco = 6, 7
    datM = (rng.normal(size=(nImg, PM)) + 1j * rng.normal(size=(nImg, PM))).astype(np.complex64)
    ctfM = rng.uniform(-1, 1, (nImg, PM)).astype(np.float32)
    nr = synth.random_quats(nImg * mReco, rng).reshape(nImg, mReco, 4)
    nt = rng.normal(scale=2.0, size=(nImg, mReco, 2))
    w = np.full(nImg, 1.0 / mReco, np.float32)
    offS = rng.normal(scale=0.5, size=(nImg, 2))
    got = _run_seam(tmp_path, N, pf, pixM, datM, ctfM, w, offS, nr, nt, sym=sym)
    R = refapi.Reconstructor(N, N, pf)
    assert got["m"] == R.pad_size()
    R.set_precal(pixM["iColPad"], pixM["iRowPad"], pixM["iPxl"], pixM["iSig"])
    R.insert_loop(datM, ctfM, w, offS, nr, nt, pixM["iCol"], pixM["iRow"], N)
    if sym:
        R.prepareTF()               # normalisation by 1 / Re T[0] (no symmetry object attached: the reference skips that part)
        R.symmetrize(sym)           # symmetrizeT / symmetrizeF with the point group
    want = R.get()
    rel = lambda a, b: np.linalg.norm((a - b).ravel()) / np.linalg.norm(b.ravel())
    if sym is None:
        assert rel(got["F"], want["F"]) <= 1e-6 and rel(got["T"], want["T"]) <= 1e-6
        assert np.allclose(got["O"], want["O"], rtol=1e-10, atol=1e-10) and got["counter"] == want["counter"] == nImg * mReco
    else:
        m = got["m"]
        r = R.max_radius() * pf + 1
        kk, jj, ii = np.meshgrid(np.fft.fftfreq(m, 1 / m), np.fft.fftfreq(m, 1 / m), np.arange(m // 2 + 1), indexing="ij")
        edge = (ii * ii + jj * jj + kk * kk) == r * r      # voxels on the cut radius: see test_device_symmetrize
        assert np.abs(got["F"] - want["F"])[~edge].max() <= 4e-6 * np.abs(want["F"]).max()
        assert np.abs(got["T"] - want["T"])[~edge].max() <= 4e-6 * np.abs(want["T"]).max()
    R.close()


def test_reference_reconstructor_object_links_against_the_shim():
    """CPU part of the drop-in proof: the seam executable holds the reference's Reconstructor::insertI / prepareTFG object code
    (compiled from the reference's source with GPU_INSERT), and every seam function it calls is an export of libthb_interface.so"""
    if not SEAM_EXE.exists():
        pytest.skip("oracle/_ref/seam_insertI not built (oracle/build_ref.sh needs the reference tree)")
    defined = subprocess.check_output(["nm", "-C", "--defined-only", os.fspath(SEAM_EXE)]).decode()
    assert "Reconstructor::insertI(" in defined and "Reconstructor::prepareTFG(" in defined
    undef = [l.split()[-1] for l in subprocess.check_output(["nm", "-u", os.fspath(SEAM_EXE)]).decode().splitlines()
             if re.search(r"_Z\d+(InsertFT|PrepareTF)", l)]
    have = set(l.split()[-1] for l in subprocess.check_output(["nm", "-D", "--defined-only", os.fspath(LIB)]).decode().splitlines())
    assert len(undef) >= 2, undef
    for sym in undef:
        assert sym in have, sym

"""diagnostic: device particle-filter operators vs the host build of the same source, same random stream (development aid)"""
import ctypes as C, os, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from thunder_b200 import capi, synth
L = C.CDLL(os.fspath(ROOT / "tests" / "pf_host" / "libpf_host.so"))
_p, _i, _d, _u = C.c_void_p, C.c_int, C.c_double, C.c_ulonglong
L.pfh_run_s.restype = _i
L.pfh_run_s.argtypes = [_i, _d, _i, _i] + [_p] * 9 + [_d, _d, _u, _u, _u]
n, mLR, mLT = int(sys.argv[1]) if len(sys.argv) > 1 else 8, 125, 9
rng = np.random.default_rng(1)
ctx = capi.Context(0)
prm = capi.PFParams(mLR=mLR, mLT=mLT, transS=2.0, transQ=0.01, perturbFactorL=2.0, perturbFactorS=0.5, minPhase=3, maxPhase=100, fixedPhases=0,
                    decreaseFactor=0.95, noDecreaseLimit=1, seed=4711)
q = synth.random_quats(n, rng); t = rng.normal(size=(n, 2))
ctx.pf_set_epoch(10)
ctx.pf_load(prm, q, np.full((n, 3), 3e-4), t, np.full((n, 2), 1.0))
st0 = ctx.pf_get()
pt = lambda a: None if a is None else a.ctypes.data_as(_p)

def host(st, op, arg, epoch):
    out = {k: v.copy() for k, v in st.items()}
    for p in range(n):
        r = np.ascontiguousarray(st["r"][p].T); tt = np.ascontiguousarray(st["t"][p].T)
        wR = st["wR"][p].copy(); wT = st["wT"][p].copy(); scal = st["scal"][p].copy(); hu = np.zeros(mLR); ht = np.zeros(mLT)
        L.pfh_run_s(op, arg, mLR, mLT, pt(r), pt(tt), pt(wR), pt(wT), pt(hu), pt(ht), None, None, pt(scal), 2.0, 0.01, 4711, p, epoch << 20)
        out["r"][p] = r.T; out["t"][p] = tt.T; out["wR"][p] = wR; out["wT"][p] = wT; out["scal"][p] = scal
    return out

for op, arg, name in ((1, 2.0, "perturb_R"), (2, 2.0, "perturb_T"), (7, 0.0, "balance_R"), (5, 0.0, "calVari")):
    res = []
    for rep in range(2):
        ctx.pf_set(r=st0["r"], t=st0["t"], wR=st0["wR"], wT=st0["wT"], scal=st0["scal"])
        ctx.pf_set_epoch(100)
        ctx.pf_op(op, arg)
        res.append(ctx.pf_get())
    h = host(st0, op, arg, 101)
    a, b = res
    print(name, "device run-to-run: dr", np.abs(a["r"] - b["r"]).max(), "dwR", np.abs(a["wR"] - b["wR"]).max(),
          "| device vs host: dr per particle", np.abs(a["r"] - h["r"]).max(axis=(1, 2)).round(12), "dt", np.abs(a["t"] - h["t"]).max(), "dk", np.abs(a["scal"][:, :5] / h["scal"][:, :5] - 1).max(),
          "dwR", np.abs(a["wR"] - h["wR"]).max())
    if name == "perturb_R":
        print(" device r[0][:2]", a["r"][0][:2], "\n host   r[0][:2]", h["r"][0][:2], "\n before r[0][:2]", st0["r"][0][:2])
ctx.close()

# ---- the sequence of tests/test_gpu_iteration.py::test_device_pf_operators_equal_host_build_with_the_same_stream, no asserts
ctx = capi.Context(0)
ctx.pf_set_epoch(900)
ctx.pf_load(prm, q, np.full((n, 3), 3e-4), t, np.full((n, 2), 1.0))
E = 901
st = ctx.pf_get()
hs = {k: v.copy() for k, v in st.items()}
hs["uR"] = np.zeros((n, mLR)); hs["uT"] = np.zeros((n, mLT))

def host2(st, op, arg, epoch, uR=None, uT=None):
    out = {k: v.copy() for k, v in st.items()}
    for p in range(n):
        r = np.ascontiguousarray(st["r"][p].T); tt = np.ascontiguousarray(st["t"][p].T)
        wR = st["wR"][p].copy(); wT = st["wT"][p].copy(); scal = st["scal"][p].copy(); hu = st["uR"][p].copy(); ht = st["uT"][p].copy()
        uRf = None if uR is None else np.ascontiguousarray(uR[p], np.float32); uTf = None if uT is None else np.ascontiguousarray(uT[p], np.float32)
        L.pfh_run_s(op, arg, mLR, mLT, pt(r), pt(tt), pt(wR), pt(wT), pt(hu), pt(ht), pt(uRf), pt(uTf), pt(scal), 2.0, 0.01, 4711, p, epoch << 20)
        out["r"][p] = r.T; out["t"][p] = tt.T; out["wR"][p] = wR; out["wT"][p] = wT; out["scal"][p] = scal; out["uR"][p] = hu; out["uT"][p] = ht
    return out

def report(name):
    d = ctx.pf_get()
    dr = np.abs(d["r"] - hs["r"]).max(axis=(1, 2))
    print(f"{name:12s} dr max {dr.max():.2e} (particles > 1e-6: {np.nonzero(dr > 1e-6)[0].tolist()[:10]}) dt {np.abs(d['t'] - hs['t']).max():.1e} "
          f"dk {np.abs(d['scal'][:, :5] / hs['scal'][:, :5] - 1).max():.1e} dwR {np.abs(d['wR'] - hs['wR']).max():.1e} dtop {np.abs(d['scal'][:, 6:12] - hs['scal'][:, 6:12]).max():.1e}")
    return d

for rep in range(3):
    for op, arg, name in ((1, 2.0 if rep == 0 else 0.5, "perturb R"), (2, 0.5, "perturb T")):
        ctx.pf_op(op, arg); E += 1
        hs = host2(hs, op, arg, E)
        report(f"{rep} {name}")
    uR = np.exp(-8 * rng.uniform(0, 1, (n, mLR)) ** 2).astype(np.float32); uT = rng.uniform(0.1, 1, (n, mLT)).astype(np.float32)
    ctx.pf_op(3, uR=uR, uT=uT); E += 1
    hs = host2(hs, 3, 0.0, E, uR, uT)
    ctx.pf_op(5); E += 1
    hs = host2(hs, 5, 0.0, E)
    report(f"{rep} calVari")
    ctx.pf_op(6); E += 1
    hs = host2(hs, 6, 0.0, E)
    report(f"{rep} resample")
ctx.close()

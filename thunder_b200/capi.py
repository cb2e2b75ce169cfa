"""ctypes binding of libthunder_b200.so (the C ABI in include/thunder_b200.h).

This is the host-side mirror used by tests/, bench.py and __graft_entry__: numpy arrays in,
numpy arrays out, every call going through the C ABI.  There is no CPU implementation behind
it: if the library or a B200 is missing, `load()` / `Context()` raise.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "lib" / "libthunder_b200.so"

UNIQUE_ID_BYTES = 128
MODE_3D, MODE_2D = 0, 1
STACK_EXPECT, STACK_INSERT = 0, 1
KF_EXPECT, KF_INSERT, KF_PF, KF_PACK, KF_COMM = range(5)
PF_PERTURB_R, PF_PERTURB_T, PF_SET_U_KEEP_PEAK, PF_RANK1ST, PF_CALVARI, PF_RESAMPLE, PF_BALANCE_R, PF_BALANCE_T = range(1, 9)


class ThbError(RuntimeError):
    pass


class PFParams(C.Structure):
    _fields_ = [
        ("mLR", C.c_int), ("mLT", C.c_int),
        ("transS", C.c_double), ("transQ", C.c_double),
        ("perturbFactorL", C.c_double), ("perturbFactorS", C.c_double),
        ("minPhase", C.c_int), ("maxPhase", C.c_int),
        ("fixedPhases", C.c_int),
        ("decreaseFactor", C.c_double),
        ("noDecreaseLimit", C.c_int),
        ("seed", C.c_uint64),
        ("mLD", C.c_int), ("ctfRefineS", C.c_double), ("perturbFactorSCTF", C.c_double),
    ]


_lib = None

_p = C.c_void_p
_i = C.c_int


def _sig(lib):
    f = lib.thb_version; f.restype = _i; f.argtypes = []
    f = lib.thb_device_count; f.restype = _i; f.argtypes = []
    f = lib.thb_create; f.restype = _i; f.argtypes = [C.POINTER(_p), _i]
    f = lib.thb_destroy; f.restype = None; f.argtypes = [_p]
    f = lib.thb_last_error; f.restype = C.c_char_p; f.argtypes = [_p]
    f = lib.thb_synchronize; f.restype = _i; f.argtypes = [_p]
    f = lib.thb_launch_count; f.restype = C.c_int64; f.argtypes = [_p, _i]
    f = lib.thb_kernel_ms; f.restype = C.c_double; f.argtypes = [_p, _i, C.POINTER(C.c_int64), _i]
    f = lib.thb_enable_timing; f.restype = _i; f.argtypes = [_p, _i]
    f = lib.thb_set_option; f.restype = _i; f.argtypes = [_p, C.c_char_p, _i]
    f = lib.thb_expect_stats; f.restype = _i; f.argtypes = [_p, _p, _i]
    f = lib.thb_timer; f.restype = _i; f.argtypes = [_p, _i, C.POINTER(C.c_float)]
    f = lib.thb_pixel_list; f.restype = _i; f.argtypes = [_i, _i, C.c_float, C.c_float] + [_p] * 6
    f = lib.thb_set_expect_pixels; f.restype = _i; f.argtypes = [_p, _i, _i, _i, _p, _p]
    f = lib.thb_set_insert_pixels; f.restype = _i; f.argtypes = [_p, _i, _i, _i, _p, _p]
    f = lib.thb_set_volume; f.restype = _i; f.argtypes = [_p, _i, _p, _i]
    f = lib.thb_get_volume; f.restype = _i; f.argtypes = [_p, _i, _p]
    f = lib.thb_upload_stack; f.restype = _i; f.argtypes = [_p, _i, _i, _p, _p, _p, _p]
    f = lib.thb_stack_reserve; f.restype = _i; f.argtypes = [_p, _i, _i]
    f = lib.thb_upload_stack_at; f.restype = _i; f.argtypes = [_p, _i, _i, _i, _p, _p, _p, _p]
    f = lib.thb_upload_stack_at_async; f.restype = _i; f.argtypes = [_p, _i, _i, _i, _p, _p, _p, _p]
    f = lib.thb_upload_wait; f.restype = _i; f.argtypes = [_p]
    f = lib.thb_set_mode; f.restype = _i; f.argtypes = [_p, _i]
    f = lib.thb_symmetrize; f.restype = _i; f.argtypes = [_p, _i, _i, _p, C.c_double]
    f = lib.thb_norm_residual; f.restype = _i; f.argtypes = [_p, _i, _p, _p, _p, C.c_float, C.c_float, _p]
    f = lib.thb_scale_images; f.restype = _i; f.argtypes = [_p, _i, _p, _p]
    f = lib.thb_get_mode; f.restype = _i; f.argtypes = [_p]
    f = lib.thb_insert_classes; f.restype = _i; f.argtypes = [_p, _i, _p, _i, _p, _p, _p, _p, _p]
    f = lib.thb_insert_ctf; f.restype = _i; f.argtypes = [_p, _i, _p, _i, _p, _p, _p, _p, _p, _p, C.c_float]
    f = lib.thb_set_frequency; f.restype = _i; f.argtypes = [_p, _p]
    f = lib.thb_upload_stack_defocus; f.restype = _i; f.argtypes = [_p, _i, _i, _p]
    f = lib.thb_expect_local_ctf; f.restype = _i; f.argtypes = [_p, _i, _p, _i, _i, _i] + [_p] * 13
    f = lib.thb_insert_counts; f.restype = _i; f.argtypes = [_p, _i, _p, _i, _p, _p, _p, _p, _p]
    f = lib.thb_pack_stack; f.restype = _i; f.argtypes = [_p, _i, _i, _i, _p, _p, _p, _p, _i, _i, _p, _p, C.c_float, _p]
    f = lib.thb_download_stack; f.restype = _i; f.argtypes = [_p, _i, _i, _i, _p, _p, _p]
    f = lib.thb_reco_upload; f.restype = _i; f.argtypes = [_p, _i, _p, _p]
    f = lib.thb_reconstruct; f.restype = _i; f.argtypes = [_p, _i, _i, _i, C.c_double, C.c_double, _i, _i, _p, _i, _i, _p, _p]
    f = lib.thb_set_projectee; f.restype = _i; f.argtypes = [_p, _i, _p, _i, _i]
    f = lib.thb_remask_pack; f.restype = _i; f.argtypes = [_p, _i, _i, _p, _p, C.c_float, _i, _p, _p, _p, _i, _i, _p, _p, C.c_float, _p, _p]
    f = lib.thb_sigma_accumulate; f.restype = _i; f.argtypes = [_p, _i, _p, _p, _p, _p, _p, _i, _i, _p, _p, _p, _p, _p]
    f = lib.thb_pf_set_image_base; f.restype = _i; f.argtypes = [_p, _i, C.c_uint64]
    f = lib.thb_pf_get_draws; f.restype = _i; f.argtypes = [_p, _i, _p, _p]
    f = lib.thb_pf_get_draws_d; f.restype = _i; f.argtypes = [_p, _i, _p]
    f = lib.thb_pf_from_scan; f.restype = _i; f.argtypes = [_p, _i, C.POINTER(PFParams), _i, _i, _i, _p, _p, _p, _p, _p, C.c_double, C.c_double, _p]
    f = lib.thb_pf_set_ctf; f.restype = _i; f.argtypes = [_p, _p, _p, C.c_float]
    f = lib.thb_pf_get_d; f.restype = _i; f.argtypes = [_p, _p, _p, _p]
    f = lib.thb_pf_set_epoch; f.restype = _i; f.argtypes = [_p, C.c_uint64]
    f = lib.thb_pf_trace; f.restype = _i; f.argtypes = [_p, _i]
    f = lib.thb_pf_get_trace; f.restype = _i; f.argtypes = [_p, _i, _p, _p, _p]
    f = lib.thb_pf_get_trace_states; f.restype = _i; f.argtypes = [_p, _i, _p]
    f = lib.thb_project; f.restype = _i; f.argtypes = [_p, _i, _i, _p, _p]
    f = lib.thb_expect_local; f.restype = _i; f.argtypes = [_p, _i, _p, _i, _i] + [_p] * 9
    f = lib.thb_expect_scan; f.restype = _i; f.argtypes = [_p, _i, _i, _i] + [_p] * 9
    f = lib.thb_expect_scan_range; f.restype = _i; f.argtypes = [_p, _i, _i, _i, _i, _i] + [_p] * 9
    f = lib.thb_expect_scan_classes; f.restype = _i; f.argtypes = [_p, _i, _i, _i, _i, _i] + [_p] * 8
    f = lib.thb_reco_alloc; f.restype = _i; f.argtypes = [_p, _i, _i]
    f = lib.thb_reco_reset; f.restype = _i; f.argtypes = [_p, _i]
    f = lib.thb_insert; f.restype = _i; f.argtypes = [_p, _i, _p, _i, _p, _p, _p, _p]
    f = lib.thb_reco_download; f.restype = _i; f.argtypes = [_p, _i, _p, _p, _p, _p, _i]
    f = lib.thb_comm_unique_id; f.restype = _i; f.argtypes = [_p]
    f = lib.thb_comm_init; f.restype = _i; f.argtypes = [_p, _i, _i, _p]
    f = lib.thb_allreduce; f.restype = _i; f.argtypes = [_p]
    f = lib.thb_pf_load; f.restype = _i; f.argtypes = [_p, _i, C.POINTER(PFParams), _p, _p, _p, _p]
    f = lib.thb_pf_get; f.restype = _i; f.argtypes = [_p] * 6
    f = lib.thb_pf_set; f.restype = _i; f.argtypes = [_p] * 6
    f = lib.thb_expectation; f.restype = _i; f.argtypes = [_p, _p]
    f = lib.thb_reconstruct_insert; f.restype = _i; f.argtypes = [_p, _i, _i, _p]
    f = lib.thb_pf_op; f.restype = _i; f.argtypes = [_p, _i, C.c_double, _p, _p]


def load() -> C.CDLL:
    """Load the shared library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise ThbError(f"{LIB_PATH} is missing - run `make` or __graft_entry__.build() first")
        _lib = C.CDLL(os.fspath(LIB_PATH))
        _sig(_lib)
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(_p)


def _arr(a, dtype, shape=None):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=dtype)
    if shape is not None:
        assert a.shape == tuple(shape), (a.shape, shape)
    return a


def pixel_list(N: int, pf: int, rU: float, rL: float):
    """Optimiser::allocPreCalIdx.  Returns dict of int32 arrays iCol,iRow,iPxl,iSig,iColPad,iRowPad."""
    lib = load()
    cap = (N // 2 + 1) * N
    bufs = {k: np.empty(cap, np.int32) for k in ("iCol", "iRow", "iPxl", "iSig", "iColPad", "iRowPad")}
    n = lib.thb_pixel_list(N, pf, rU, rL, *[_ptr(bufs[k]) for k in ("iCol", "iRow", "iPxl", "iSig", "iColPad", "iRowPad")])
    if n < 0:
        raise ThbError(f"thb_pixel_list failed ({n})")
    return {k: v[:n].copy() for k, v in bufs.items()}


def comm_unique_id() -> bytes:
    lib = load()
    buf = C.create_string_buffer(UNIQUE_ID_BYTES)
    rc = lib.thb_comm_unique_id(buf)
    if rc:
        raise ThbError(f"thb_comm_unique_id failed ({rc}): NCCL not available")
    return buf.raw


class Context:
    """One context per GPU (thb_ctx)."""

    def __init__(self, device: int = 0):
        self.lib = load()
        h = _p()
        rc = self.lib.thb_create(C.byref(h), device)
        if rc:
            raise ThbError(f"thb_create({device}) failed ({rc}): {self.lib.thb_last_error(None).decode()}")
        self.h = h
        self.nPxlE = self.nPxlM = 0
        self.nImgE = self.nImgM = 0
        self.vdim = {}
        self.accdim = {}
        self.pf_params = None
        self.nPar = 0
        self.mode2D = False

    def close(self):
        if getattr(self, "h", None):
            self.lib.thb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc:
            raise ThbError(f"libthunder_b200 error {rc}: {self.lib.thb_last_error(self.h).decode()}")

    # ---- accounting
    def synchronize(self):
        self._chk(self.lib.thb_synchronize(self.h))

    def launch_count(self, reset=False):
        return int(self.lib.thb_launch_count(self.h, int(reset)))

    def set_option(self, key: str, value: int):
        self._chk(self.lib.thb_set_option(self.h, key.encode(), int(value)))

    def expect_stats(self, reset=True):
        out = np.zeros(16, np.uint64)
        self._chk(self.lib.thb_expect_stats(self.h, _ptr(out), int(reset)))
        keys = ("tiles", "tiles_boxed", "margin_sum", "staged_elems", "pairs_l2_path", "pairs", "margin_retries", "rows", "cyc_e_and_A", "cyc_records_classify", "cyc_candidates", "cyc_rows_scan_issue", "cyc_l2_path",
                "cyc_tma_wait", "cyc_tail", "cyc_7")
        return dict(zip(keys, (int(v) for v in out)))

    def enable_timing(self, on=True):
        self._chk(self.lib.thb_enable_timing(self.h, int(on)))

    def kernel_ms(self, which, reset=False):
        n = C.c_int64(0)
        ms = self.lib.thb_kernel_ms(self.h, which, C.byref(n), int(reset))
        return float(ms), int(n.value)

    def timer_start(self):
        self._chk(self.lib.thb_timer(self.h, 0, None))

    def timer_stop(self) -> float:
        ms = C.c_float(0)
        self._chk(self.lib.thb_timer(self.h, 1, C.byref(ms)))
        return float(ms.value)

    # ---- geometry / resident data
    def set_expect_pixels(self, N, pf, iCol, iRow):
        iCol = _arr(iCol, np.int32); iRow = _arr(iRow, np.int32)
        self._chk(self.lib.thb_set_expect_pixels(self.h, N, pf, len(iCol), _ptr(iCol), _ptr(iRow)))
        self.nPxlE = len(iCol)

    def set_insert_pixels(self, N, pf, iColPad, iRowPad):
        a = _arr(iColPad, np.int32); b = _arr(iRowPad, np.int32)
        self._chk(self.lib.thb_set_insert_pixels(self.h, N, pf, len(a), _ptr(a), _ptr(b)))
        self.nPxlM = len(a)

    def drop_volumes(self):
        """release every projector volume and accumulator (a mode round trip does that), keeping pixel lists and stacks"""
        m = MODE_2D if self.mode2D else MODE_3D
        self.set_mode(MODE_3D if self.mode2D else MODE_2D)
        self.set_mode(m)

    def set_mode(self, mode):
        """MODE_3D (default) / MODE_2D: 2D classification - image references, in-plane rotations passed as (cos, sin)"""
        self._chk(self.lib.thb_set_mode(self.h, int(mode)))
        self.mode2D = int(mode) == MODE_2D
        self.vdim, self.accdim = {}, {}

    def set_volume(self, slot, volFT):
        """volFT: complex64 [vdim][vdim][vdim/2+1] (z, y, x) half-complex; MODE_2D: [vdim][vdim/2+1]"""
        v = np.ascontiguousarray(volFT, dtype=np.complex64)
        vdim = v.shape[0]
        assert v.shape == ((vdim, vdim // 2 + 1) if getattr(self, "mode2D", False) else (vdim, vdim, vdim // 2 + 1)), v.shape
        self._chk(self.lib.thb_set_volume(self.h, slot, _ptr(v), vdim))
        self.vdim[slot] = vdim

    def get_volume(self, slot):
        vdim = self.vdim[slot]
        out = np.empty((vdim, vdim // 2 + 1) if getattr(self, "mode2D", False) else (vdim, vdim, vdim // 2 + 1), np.complex64)
        self._chk(self.lib.thb_get_volume(self.h, slot, _ptr(out)))
        return out

    def upload_stack(self, kind, dat, ctf, sigRcp=None, slotOfImg=None):
        dat = np.ascontiguousarray(dat, dtype=np.complex64)
        nImg, P = dat.shape
        ctf = _arr(ctf, np.float32, (nImg, P))
        sigRcp = _arr(sigRcp, np.float32, (nImg, P))
        slotOfImg = _arr(slotOfImg, np.int32, (nImg,))
        self._chk(self.lib.thb_upload_stack(self.h, kind, nImg, _ptr(dat), _ptr(ctf), _ptr(sigRcp), _ptr(slotOfImg)))
        if kind == STACK_EXPECT:
            self.nImgE = nImg
        else:
            self.nImgM = nImg

    def stack_reserve(self, kind, capacity):
        self._chk(self.lib.thb_stack_reserve(self.h, kind, capacity))
        if kind == STACK_EXPECT:
            self.nImgE = capacity
        else:
            self.nImgM = capacity

    def upload_stack_at(self, kind, base, dat, ctf, sigRcp=None, slotOfImg=None):
        """arrays must already be C-contiguous complex64 / float32 (no copies: they may be pinned buffers)"""
        nImg, P = dat.shape
        assert dat.dtype == np.complex64 and dat.flags.c_contiguous and ctf.dtype == np.float32 and ctf.flags.c_contiguous
        slotOfImg = _arr(slotOfImg, np.int32, (nImg,))
        self._chk(self.lib.thb_upload_stack_at(self.h, kind, base, nImg, _ptr(dat), _ptr(ctf), _ptr(sigRcp), _ptr(slotOfImg)))

    def upload_stack_at_async(self, kind, base, dat, ctf, sigRcp=None, slotOfImg=None):
        """second-stream upload; the (pinned) arrays must stay alive and unchanged until upload_wait()"""
        nImg, P = dat.shape
        assert dat.dtype == np.complex64 and dat.flags.c_contiguous and ctf.dtype == np.float32 and ctf.flags.c_contiguous
        slotOfImg = _arr(slotOfImg, np.int32, (nImg,))
        self._async_keep = getattr(self, "_async_keep", []) + [dat, ctf, sigRcp, slotOfImg]
        self._chk(self.lib.thb_upload_stack_at_async(self.h, kind, base, nImg, _ptr(dat), _ptr(ctf), _ptr(sigRcp), _ptr(slotOfImg)))

    def upload_wait(self):
        self._chk(self.lib.thb_upload_wait(self.h))
        self._async_keep = []

    def pack_stack(self, kind, base, imgFT, iPxl, ctfAttr, pixelSize, iSig=None, sigRcpTab=None, groupOfImg=None, slotOfImg=None):
        """Optimiser::allocPreCal on the device: imgFT[nImg][N][N/2+1] complex64 full half-FTs, ctfAttr[nImg][7]"""
        imgFT = _arr(imgFT, np.complex64)
        nImg = imgFT.shape[0]
        iPxl = _arr(iPxl, np.int32); iSig = _arr(iSig, np.int32)
        ctfAttr = _arr(ctfAttr, np.float32, (nImg, 7))
        sigRcpTab = _arr(sigRcpTab, np.float32)
        nGroup, nRing = (sigRcpTab.shape if sigRcpTab is not None else (0, 0))
        groupOfImg = _arr(groupOfImg, np.int32, (nImg,)); slotOfImg = _arr(slotOfImg, np.int32, (nImg,))
        self._chk(self.lib.thb_pack_stack(self.h, kind, base, nImg, _ptr(imgFT), _ptr(iPxl), _ptr(iSig), _ptr(sigRcpTab), nGroup, nRing,
                                          _ptr(groupOfImg), _ptr(ctfAttr), float(pixelSize), _ptr(slotOfImg)))

    def download_stack(self, kind, base, nImg):
        P = self.nPxlE if kind == STACK_EXPECT else self.nPxlM
        dat = np.empty((nImg, P), np.complex64); ctf = np.empty((nImg, P), np.float32)
        sig = np.empty((nImg, P), np.float32) if kind == STACK_EXPECT else None
        self._chk(self.lib.thb_download_stack(self.h, kind, base, nImg, _ptr(dat), _ptr(ctf), _ptr(sig)))
        return dict(dat=dat, ctf=ctf, sigRcp=sig)

    def remask_pack(self, base, imgOriFT, offset, maskRadiusPx, iPxl, iSig, sigRcpTab, ctfAttr, pixelSize, zeroMask=True, groupOfImg=None,
                    slotOfImg=None, want_images=False):
        """Optimiser::reCentreImg + reMaskImg + allocPreCal on the device"""
        imgOriFT = _arr(imgOriFT, np.complex64)
        nImg = imgOriFT.shape[0]
        offset = _arr(offset, np.float64, (nImg, 2))
        iPxl = _arr(iPxl, np.int32); iSig = _arr(iSig, np.int32)
        sigRcpTab = _arr(sigRcpTab, np.float32); ctfAttr = _arr(ctfAttr, np.float32, (nImg, 7))
        groupOfImg = _arr(groupOfImg, np.int32, (nImg,)); slotOfImg = _arr(slotOfImg, np.int32, (nImg,))
        out = np.empty_like(imgOriFT) if want_images else None
        self._chk(self.lib.thb_remask_pack(self.h, base, nImg, _ptr(imgOriFT), _ptr(offset), float(maskRadiusPx), int(zeroMask), _ptr(iPxl),
                                           _ptr(iSig), _ptr(sigRcpTab), sigRcpTab.shape[0], sigRcpTab.shape[1], _ptr(groupOfImg), _ptr(ctfAttr),
                                           float(pixelSize), _ptr(slotOfImg), _ptr(out)))
        return out

    def sigma_accumulate(self, quat, tran, offS, groupOfImg, nGroup, rSig, iSigE, iSigM, imgIdx=None):
        """image loop of Optimiser::allReduceSigma -> sigM, sigN, svd [nGroup][rSig+1]"""
        quat = _arr(quat, np.float64); nImg = quat.shape[0]
        tran = _arr(tran, np.float64, (nImg, 2)); offS = _arr(offS, np.float64, (nImg, 2))
        groupOfImg = _arr(groupOfImg, np.int32, (nImg,)); imgIdx = _arr(imgIdx, np.int32, (nImg,))
        iSigE = _arr(iSigE, np.int32); iSigM = _arr(iSigM, np.int32)
        out = [np.zeros((nGroup, rSig + 1)) for _ in range(3)]
        self._chk(self.lib.thb_sigma_accumulate(self.h, nImg, _ptr(imgIdx), _ptr(quat), _ptr(tran), _ptr(offS), _ptr(groupOfImg), nGroup, rSig,
                                                _ptr(iSigE), _ptr(iSigM), _ptr(out[0]), _ptr(out[1]), _ptr(out[2])))
        return out

    def symmetrize(self, slot, R, radius):
        """Reconstructor::symmetrizeF / T / O: R[nElem][9] column-major symmetry elements, radius = maxRadius * pf + 1"""
        R = _arr(R, np.float64)
        nElem = 0 if R is None else R.shape[0]
        self._chk(self.lib.thb_symmetrize(self.h, slot, nElem, _ptr(R) if nElem else None, float(radius)))

    def norm_residual(self, quat, tran, rL, rNorm, imgIdx=None):
        """image loop of Optimiser::normCorrection: sum of |masked image - ctf * translated slice|^2 over rL <= |k| < rNorm"""
        quat = _arr(quat, np.float64); nImg = quat.shape[0]
        tran = _arr(tran, np.float64, (nImg, 2)); imgIdx = _arr(imgIdx, np.int32, (nImg,))
        out = np.zeros(nImg)
        self._chk(self.lib.thb_norm_residual(self.h, nImg, _ptr(imgIdx), _ptr(quat), _ptr(tran), float(rL), float(rNorm), _ptr(out)))
        return out

    def scale_images(self, scale, imgIdx=None):
        scale = _arr(scale, np.float32); nImg = scale.shape[0]
        imgIdx = _arr(imgIdx, np.int32, (nImg,))
        self._chk(self.lib.thb_scale_images(self.h, nImg, _ptr(imgIdx), _ptr(scale)))

    # ---- reconstruct / setProjectee
    def reco_upload(self, slot, F, T):
        F = _arr(F, np.complex64); T = _arr(T, np.float32)
        self._chk(self.lib.thb_reco_upload(self.h, slot, _ptr(F), _ptr(T)))

    def reconstruct(self, slot, N, pf, a=1.9, alpha=15.0, gridCorr=True, joinHalf=False, fsc=None, normalise=True, want_volume=True):
        fsc = _arr(fsc, np.float32)
        out = np.empty((N, N) if getattr(self, "mode2D", False) else (N, N, N), np.float32) if want_volume else None
        nit = C.c_int(0)
        self._chk(self.lib.thb_reconstruct(self.h, slot, N, pf, a, alpha, int(gridCorr), int(joinHalf), _ptr(fsc),
                                           0 if fsc is None else len(fsc), int(normalise), _ptr(out), C.byref(nit)))
        return out, nit.value

    def set_projectee(self, slot, vol, N, pf):
        vol = _arr(vol, np.float32, (N, N) if getattr(self, "mode2D", False) else (N, N, N)) if vol is not None else None
        self._chk(self.lib.thb_set_projectee(self.h, slot, _ptr(vol), N, pf))
        self.vdim[slot] = N * pf

    # ---- E
    def project(self, slot, quat):
        quat = _arr(quat, np.float64)
        nRot = quat.shape[0]
        out = np.empty((nRot, self.nPxlE), np.complex64)
        self._chk(self.lib.thb_project(self.h, slot, nRot, _ptr(quat), _ptr(out)))
        return out

    def expect_local(self, quat, tran, wR, wT, imgIdx=None, want_logL=True):
        quat = _arr(quat, np.float64)
        nAct, nR, _ = quat.shape
        tran = _arr(tran, np.float64)
        nT = tran.shape[1]
        wR = _arr(wR, np.float64, (nAct, nR)); wT = _arr(wT, np.float64, (nAct, nT))
        imgIdx = _arr(imgIdx, np.int32, (nAct,))
        uR = np.empty((nAct, nR), np.float32); uT = np.empty((nAct, nT), np.float32)
        uC = np.empty(nAct, np.float32); base = np.empty(nAct, np.float32)
        logL = np.empty((nAct, nR, nT), np.float32) if want_logL else None
        self._chk(self.lib.thb_expect_local(self.h, nAct, _ptr(imgIdx), nR, nT, _ptr(quat), _ptr(tran), _ptr(wR), _ptr(wT),
                                            _ptr(uR), _ptr(uT), _ptr(uC), _ptr(base), _ptr(logL)))
        return dict(uR=uR, uT=uT, uC=uC, base=base, logL=logL)

    def expect_scan(self, slot, quat, tran, pR, pT, want_logL=False, img_range=None):
        quat = _arr(quat, np.float64); tran = _arr(tran, np.float64)
        nR, nT = quat.shape[0], tran.shape[0]
        pR = _arr(pR, np.float64, (nR,)); pT = _arr(pT, np.float64, (nT,))
        base0, n = (0, self.nImgE) if img_range is None else img_range
        wC = np.empty(n, np.float32); wR = np.empty((n, nR), np.float32); wT = np.empty((n, nT), np.float32)
        base = np.empty(n, np.float32)
        logL = np.empty((n, nR, nT), np.float32) if want_logL else None
        self._chk(self.lib.thb_expect_scan_range(self.h, slot, int(base0), int(n), nR, nT, _ptr(quat), _ptr(tran), _ptr(pR), _ptr(pT),
                                                 _ptr(wC), _ptr(wR), _ptr(wT), _ptr(base), _ptr(logL)))
        return dict(wC=wC, wR=wR, wT=wT, base=base, logL=logL)

    def expect_scan_classes(self, nK, quat, tran, pR, pT, img_range=None):
        """MODE_2D: every image against all nK classes in one launch; wC[n][nK], wR[nK][n][nR], wT[nK][n][nT], base[n] (one baseline per image)"""
        quat = _arr(quat, np.float64); tran = _arr(tran, np.float64)
        nR, nT = quat.shape[0], tran.shape[0]
        pR = _arr(pR, np.float64, (nR,)); pT = _arr(pT, np.float64, (nT,))
        base0, n = (0, self.nImgE) if img_range is None else img_range
        wC = np.empty((n, nK), np.float32); wR = np.empty((nK, n, nR), np.float32); wT = np.empty((nK, n, nT), np.float32)
        base = np.empty(n, np.float32)
        self._chk(self.lib.thb_expect_scan_classes(self.h, int(nK), int(base0), int(n), nR, nT, _ptr(quat), _ptr(tran), _ptr(pR), _ptr(pT),
                                                   _ptr(wC), _ptr(wR), _ptr(wT), _ptr(base)))
        return dict(wC=wC, wR=wR, wT=wT, base=base)

    # ---- M
    def reco_alloc(self, slot, vdimPad):
        self._chk(self.lib.thb_reco_alloc(self.h, slot, vdimPad))
        self.accdim[slot] = vdimPad

    def reco_reset(self, slot):
        self._chk(self.lib.thb_reco_reset(self.h, slot))

    def insert(self, w, nr, nt, offS=None, imgIdx=None):
        nr = _arr(nr, np.float64)
        nImg, mReco, _ = nr.shape
        nt = _arr(nt, np.float64, (nImg, mReco, 2))
        w = _arr(w, np.float32, (nImg,))
        offS = _arr(offS, np.float64, (nImg, 2))
        imgIdx = _arr(imgIdx, np.int32, (nImg,))
        self._chk(self.lib.thb_insert(self.h, nImg, _ptr(imgIdx), mReco, _ptr(w), _ptr(offS), _ptr(nr), _ptr(nt)))

    def insert_ctf(self, w, nr, nt, nd, ctfAttr, pixelSize, offS=None, imgIdx=None):
        """CTF search: per-draw defocus factors nd[nImg][mReco], ctfAttr[nImg][7]"""
        nr = _arr(nr, np.float64)
        nImg, mReco, _ = nr.shape
        nt = _arr(nt, np.float64, (nImg, mReco, 2)); nd = _arr(nd, np.float64, (nImg, mReco))
        w = _arr(w, np.float32, (nImg,)); offS = _arr(offS, np.float64, (nImg, 2)); imgIdx = _arr(imgIdx, np.int32, (nImg,))
        ctfAttr = _arr(ctfAttr, np.float32, (nImg, 7))
        self._chk(self.lib.thb_insert_ctf(self.h, nImg, _ptr(imgIdx), mReco, _ptr(w), _ptr(offS), _ptr(nr), _ptr(nt), _ptr(nd), _ptr(ctfAttr),
                                          float(pixelSize)))

    def set_frequency(self, freQ):
        self._chk(self.lib.thb_set_frequency(self.h, _ptr(_arr(freQ, np.float32, (self.nPxlE,)))))

    def upload_stack_defocus(self, base, defP):
        defP = _arr(defP, np.float32)
        self._chk(self.lib.thb_upload_stack_defocus(self.h, int(base), defP.shape[0], _ptr(defP)))

    def expect_local_ctf(self, quat, tran, dpar, wR, wT, wD, ctfK, imgIdx=None, want_logL=True):
        quat = _arr(quat, np.float64)
        nAct, nR, _ = quat.shape
        tran = _arr(tran, np.float64); nT = tran.shape[1]
        dpar = _arr(dpar, np.float64); nD = dpar.shape[1]
        wR = _arr(wR, np.float64, (nAct, nR)); wT = _arr(wT, np.float64, (nAct, nT)); wD = _arr(wD, np.float64, (nAct, nD))
        ctfK = _arr(ctfK, np.float32, (nAct, 4)); imgIdx = _arr(imgIdx, np.int32, (nAct,))
        uR = np.empty((nAct, nR), np.float32); uT = np.empty((nAct, nT), np.float32); uD = np.empty((nAct, nD), np.float32)
        uC = np.empty(nAct, np.float32); base = np.empty(nAct, np.float32)
        logL = np.empty((nAct, nR, nT, nD), np.float32) if want_logL else None
        self._chk(self.lib.thb_expect_local_ctf(self.h, nAct, _ptr(imgIdx), nR, nT, nD, _ptr(quat), _ptr(tran), _ptr(dpar), _ptr(wR), _ptr(wT),
                                                _ptr(wD), _ptr(ctfK), _ptr(uR), _ptr(uT), _ptr(uD), _ptr(uC), _ptr(base), _ptr(logL)))
        return dict(uR=uR, uT=uT, uD=uD, uC=uC, base=base, logL=logL)

    def insert_counts(self, w, nDraw, nr, nt, offS=None, imgIdx=None):
        """3D classification: only the first nDraw[l] of the mReco rows of image l are inserted"""
        nr = _arr(nr, np.float64)
        nImg, mReco, _ = nr.shape
        nt = _arr(nt, np.float64, (nImg, mReco, 2)); nDraw = _arr(nDraw, np.int32, (nImg,))
        w = _arr(w, np.float32, (nImg,))
        offS = _arr(offS, np.float64, (nImg, 2))
        imgIdx = _arr(imgIdx, np.int32, (nImg,))
        self._chk(self.lib.thb_insert_counts(self.h, nImg, _ptr(imgIdx), mReco, _ptr(w), _ptr(offS), _ptr(nDraw), _ptr(nr), _ptr(nt)))

    def insert_classes(self, w, nc, nr, nt, offS=None, imgIdx=None):
        """MODE_2D: nc[nImg][mReco] = class (accumulator slot) of every draw, nr[nImg][mReco][2] = (cos, sin)"""
        nr = _arr(nr, np.float64)
        nImg, mReco, _ = nr.shape
        nt = _arr(nt, np.float64, (nImg, mReco, 2)); nc = _arr(nc, np.int32, (nImg, mReco))
        w = _arr(w, np.float32, (nImg,))
        offS = _arr(offS, np.float64, (nImg, 2))
        imgIdx = _arr(imgIdx, np.int32, (nImg,))
        self._chk(self.lib.thb_insert_classes(self.h, nImg, _ptr(imgIdx), mReco, _ptr(w), _ptr(offS), _ptr(nc), _ptr(nr), _ptr(nt)))

    def reco_download(self, slot, normalise=False, want_F=True, want_T=True, out=None):
        """out = (F, T): preallocated (e.g. page-locked) destination arrays"""
        m = self.accdim[slot]
        shape = (m, m // 2 + 1) if getattr(self, "mode2D", False) else (m, m, m // 2 + 1)
        if out is not None:
            F, T = out
            assert F.shape == shape and F.dtype == np.complex64 and F.flags.c_contiguous
            assert T.shape == shape and T.dtype == np.float32 and T.flags.c_contiguous
        else:
            F = np.empty(shape, np.complex64) if want_F else None
            T = np.empty(shape, np.float32) if want_T else None
        O = np.empty(3, np.float64)
        cnt = np.zeros(1, np.int32)
        self._chk(self.lib.thb_reco_download(self.h, slot, _ptr(F), _ptr(T), _ptr(O), _ptr(cnt), int(normalise)))
        return dict(F=F, T=T, O=O, counter=int(cnt[0]))

    # ---- collective
    def comm_init(self, nRanks, rank, uid: bytes | None):
        buf = C.create_string_buffer(uid, UNIQUE_ID_BYTES) if uid is not None else None
        self._chk(self.lib.thb_comm_init(self.h, nRanks, rank, buf))

    def allreduce(self):
        self._chk(self.lib.thb_allreduce(self.h))

    # ---- particle filter
    def pf_load(self, params: PFParams, quat, k123, tran, s01):
        quat = _arr(quat, np.float64)
        nPar = quat.shape[0]
        k123 = _arr(k123, np.float64, (nPar, 3)); tran = _arr(tran, np.float64, (nPar, 2)); s01 = _arr(s01, np.float64, (nPar, 2))
        self._chk(self.lib.thb_pf_load(self.h, nPar, C.byref(params), _ptr(quat), _ptr(k123), _ptr(tran), _ptr(s01)))
        self.pf_params = params
        self.nPar = nPar

    def pf_get(self):
        n, R, T = self.nPar, self.pf_params.mLR, self.pf_params.mLT
        r = np.empty((n, R, 4)); t = np.empty((n, T, 2)); wR = np.empty((n, R)); wT = np.empty((n, T)); scal = np.empty((n, 20))
        self._chk(self.lib.thb_pf_get(self.h, _ptr(r), _ptr(t), _ptr(wR), _ptr(wT), _ptr(scal)))
        return dict(r=r, t=t, wR=wR, wT=wT, scal=scal)

    def pf_get_scal(self):
        scal = np.empty((self.nPar, 20))
        self._chk(self.lib.thb_pf_get(self.h, None, None, None, None, _ptr(scal)))
        return scal

    def pf_set(self, r=None, t=None, wR=None, wT=None, scal=None):
        n, R, T = self.nPar, self.pf_params.mLR, self.pf_params.mLT
        r = _arr(r, np.float64, (n, R, 4)); t = _arr(t, np.float64, (n, T, 2))
        wR = _arr(wR, np.float64, (n, R)); wT = _arr(wT, np.float64, (n, T)); scal = _arr(scal, np.float64, (n, 20))
        self._chk(self.lib.thb_pf_set(self.h, _ptr(r), _ptr(t), _ptr(wR), _ptr(wT), _ptr(scal)))

    def pf_set_image_base(self, imgBase, streamBase=0):
        self._chk(self.lib.thb_pf_set_image_base(self.h, int(imgBase), int(streamBase)))

    def pf_from_scan(self, params, quat, tran, wC, wR, wT, kFloor, sFloor):
        """scan results -> particle supports; wC[nPar][nK], wR[nK][nPar][nR], wT[nK][nPar][nT]; returns the chosen classes"""
        wC = _arr(wC, np.float32); nPar, nK = wC.shape
        quat = _arr(quat, np.float64); tran = _arr(tran, np.float64)
        nR, nT = quat.shape[0], tran.shape[0]
        wR = _arr(wR, np.float32, (nK, nPar, nR)); wT = _arr(wT, np.float32, (nK, nPar, nT))
        cls = np.empty(nPar, np.int32)
        self._chk(self.lib.thb_pf_from_scan(self.h, nPar, C.byref(params), nK, nR, nT, _ptr(quat), _ptr(tran), _ptr(wC), _ptr(wR), _ptr(wT),
                                            float(kFloor), float(sFloor), _ptr(cls)))
        self.pf_params = params
        self.nPar = nPar
        return cls

    def pf_set_ctf(self, ctfK, ctfAttr, pixelSize):
        ctfK = _arr(ctfK, np.float32, (self.nPar, 4)); ctfAttr = _arr(ctfAttr, np.float32, (self.nPar, 7))
        self._chk(self.lib.thb_pf_set_ctf(self.h, _ptr(ctfK), _ptr(ctfAttr), float(pixelSize)))

    def pf_get_d(self):
        D = self.pf_params.mLD
        d = np.empty((self.nPar, D + 1)); wD = np.empty((self.nPar, D)); sD = np.empty(self.nPar)
        self._chk(self.lib.thb_pf_get_d(self.h, _ptr(d), _ptr(wD), _ptr(sD)))
        return dict(d=d[:, :D], topD=d[:, D], wD=wD, sD=sD)

    def pf_set_epoch(self, epoch):
        self._chk(self.lib.thb_pf_set_epoch(self.h, int(epoch)))

    def pf_trace(self, nPhases):
        self._chk(self.lib.thb_pf_trace(self.h, int(nPhases)))

    def pf_get_trace(self, nPhases):
        R, T = self.pf_params.mLR, self.pf_params.mLT
        uR = np.empty((nPhases, self.nPar, R), np.float32); uT = np.empty((nPhases, self.nPar, T), np.float32)
        base = np.empty((nPhases, self.nPar), np.float32)
        self._chk(self.lib.thb_pf_get_trace(self.h, int(nPhases), _ptr(uR), _ptr(uT), _ptr(base)))
        self.trace_base = base
        return uR, uT

    def pf_get_trace_states(self, nPhases):
        """supports of the traced phases: dict(rPert, tPert, rRes, tRes) of [nPhases][nPar][mLR][4] / [..][mLT][2]"""
        R, T, n = self.pf_params.mLR, self.pf_params.mLT, self.nPar
        st = np.empty((nPhases, 2, n, 4 * R + 2 * T))
        self._chk(self.lib.thb_pf_get_trace_states(self.h, int(nPhases), _ptr(st)))
        r = st[..., :4 * R].reshape(nPhases, 2, n, R, 4); t = st[..., 4 * R:].reshape(nPhases, 2, n, T, 2)
        return dict(rPert=np.ascontiguousarray(r[:, 0]), tPert=np.ascontiguousarray(t[:, 0]), rRes=np.ascontiguousarray(r[:, 1]),
                    tRes=np.ascontiguousarray(t[:, 1]))

    def pf_get_draws_d(self, mReco):
        dD = np.empty((self.nPar, mReco), np.int32)
        self._chk(self.lib.thb_pf_get_draws_d(self.h, mReco, _ptr(dD)))
        return dD

    def pf_get_draws(self, mReco):
        dR = np.empty((self.nPar, mReco), np.int32); dT = np.empty((self.nPar, mReco), np.int32)
        self._chk(self.lib.thb_pf_get_draws(self.h, mReco, _ptr(dR), _ptr(dT)))
        return dR, dT

    def expectation(self, want_phases=False):
        ph = np.zeros(self.nPar, np.int32) if want_phases else None
        self._chk(self.lib.thb_expectation(self.h, _ptr(ph)))
        return ph

    def reconstruct_insert(self, mReco, parGra=False, offS=None):
        offS = _arr(offS, np.float64, (self.nPar, 2))
        self._chk(self.lib.thb_reconstruct_insert(self.h, mReco, int(parGra), _ptr(offS)))

    def pf_op(self, op, arg=0.0, uR=None, uT=None):
        uR = _arr(uR, np.float32); uT = _arr(uT, np.float32)
        self._chk(self.lib.thb_pf_op(self.h, op, float(arg), _ptr(uR), _ptr(uT)))

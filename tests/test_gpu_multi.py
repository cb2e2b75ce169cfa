"""Two-GPU test of the one exchange step of the path: the half-map all-reduce (thb_comm_init / thb_allreduce,
NCCL over NVLink) - every rank inserts its own images, the reduced accumulators equal a single-GPU insertion of all
of them.  Skipped on boxes with one GPU (the driver's scaling run and `gpurun --gpus 2` exercise it)."""
import threading

import numpy as np
import pytest

from thunder_b200 import capi, synth

pytestmark = pytest.mark.gpu


def test_allreduce_two_ranks_equals_single_gpu_insert():
    if capi.load().thb_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    N, pf = 32, 2
    rng = np.random.default_rng(3)
    pixM = capi.pixel_list(N, pf, 15.0, 0.0)
    PM = len(pixM["iCol"])
    nImg, mReco = 6, 4
    datM = (rng.normal(size=(nImg, PM)) + 1j * rng.normal(size=(nImg, PM))).astype(np.complex64)
    ctfM = rng.uniform(-1, 1, (nImg, PM)).astype(np.float32)
    slot = (np.arange(nImg) % 2).astype(np.int32)
    nr = synth.random_quats(nImg * mReco, rng).reshape(nImg, mReco, 4)
    nt = rng.normal(scale=2.0, size=(nImg, mReco, 2))
    w = np.full(nImg, 1.0 / mReco, np.float32)

    def run(ctx, sel):
        ctx.set_insert_pixels(N, pf, pixM["iColPad"], pixM["iRowPad"])
        ctx.upload_stack(capi.STACK_INSERT, datM[sel], ctfM[sel], slotOfImg=slot[sel])
        for s in (0, 1):
            ctx.reco_alloc(s, N * pf)
        ctx.insert(w[sel], nr[sel], nt[sel])

    single = capi.Context(0)
    run(single, np.arange(nImg))
    want = [single.reco_download(s) for s in (0, 1)]
    single.close()

    uid = capi.comm_unique_id()
    got, errs = {}, []

    def rank_main(rank):
        try:
            ctx = capi.Context(rank)
            ctx.comm_init(2, rank, uid)
            run(ctx, np.arange(nImg)[rank::2][::1] if False else np.arange(rank * 3, rank * 3 + 3))
            ctx.allreduce()
            got[rank] = [ctx.reco_download(s) for s in (0, 1)]
            ctx.close()
        except Exception as e:  # pragma: no cover
            errs.append(e)

    th = [threading.Thread(target=rank_main, args=(r,)) for r in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=300)
    assert not errs, errs
    rel = lambda a, b: np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-30)
    for rank in (0, 1):
        for s in (0, 1):
            assert got[rank][s]["counter"] == want[s]["counter"]
            assert rel(got[rank][s]["F"], want[s]["F"]) <= 1e-6
            assert rel(got[rank][s]["T"], want[s]["T"]) <= 1e-6
            assert np.allclose(got[rank][s]["O"], want[s]["O"], rtol=1e-10, atol=1e-10)
    assert np.array_equal(got[0][0]["F"], got[1][0]["F"])          # both ranks hold the same reduced volume

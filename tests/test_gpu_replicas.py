"""The in-process multi-device dispatcher (csrc/replicas.cpp, wsp_replicas_*): a host work queue over independent chunks, retire +
re-queue of a failed replica, and two contexts decoding CONCURRENTLY on one GPU (the decoder-step kernel needs all its CTAs
co-resident: it is launched cooperatively, so two of them serialise instead of dead-locking each other)."""
import numpy as np
import pytest

from whisper_b200 import capi, synth

pytestmark = pytest.mark.gpu

MODEL = "micro.en-sc"
N_DECODE = 24


@pytest.fixture(scope="module")
def single():
    m = capi.Model(synth.model_path(MODEL))
    e = capi.Engine(m, 0)
    c = capi.Context(e, 2)
    pcms = [synth.synth_pcm(20 + i, 480000 - 40000 * (i % 4)) for i in range(7)]     # uneven clip lengths
    prompt = m.prompt_init()
    want = np.concatenate([c.run_chunks(pcms[i:i + 2], prompt, N_DECODE)[0] for i in range(0, len(pcms), 2)])
    yield m, pcms, prompt, want
    c.close(); e.close(); m.close()


def test_two_replicas_on_one_gpu_match_a_single_context(single):
    m, pcms, prompt, want = single
    r = capi.Replicas(m, [0, 0], 2)
    try:
        for _ in range(3):                                   # the queue hands batches to whichever replica is free: any split must agree
            toks, stats = r.run_chunks(pcms, prompt, N_DECODE)
            assert toks.tolist() == want.tolist()
            assert sum(s["chunks"] for s in stats) == len(pcms) and sum(s["batches"] for s in stats) == 4
            assert not any(s["failed"] for s in stats)
    finally:
        r.close()


def test_failed_replica_is_retired_and_its_batch_requeued(single):
    m, pcms, prompt, want = single
    r = capi.Replicas(m, [0, 0], 2)
    try:
        r.fail_next(1)
        toks, stats = r.run_chunks(pcms, prompt, N_DECODE)
        assert toks.tolist() == want.tolist()
        assert stats[1]["failed"] == 1 and stats[0]["failed"] == 0 and stats[0]["chunks"] == len(pcms)
        # a retired replica stays retired; the survivor does all the work
        toks, stats = r.run_chunks(pcms[:3], prompt, N_DECODE)
        assert toks.tolist() == want[:3].tolist() and stats[0]["chunks"] == 3
        r.fail_next(0)
        with pytest.raises(capi.WspError):
            r.run_chunks(pcms[:2], prompt, N_DECODE)
    finally:
        r.close()


def test_two_devices_in_one_process(single):
    """cudaFuncSetAttribute is per device: engines on two devices of one process must both launch the big-smem kernels."""
    if capi.lib().wsp_device_count() < 2:
        pytest.skip("needs two GPUs")
    m, pcms, prompt, want = single
    r = capi.Replicas(m, [0, 1], 2)
    try:
        toks, stats = r.run_chunks(pcms, prompt, N_DECODE)
        assert toks.tolist() == want.tolist()
        assert {s["device"] for s in stats} == {0, 1} and all(s["chunks"] > 0 for s in stats)
    finally:
        r.close()

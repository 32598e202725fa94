"""Host-side multi-rank logic on CPU: world_size 2, gloo backend.  The chunk shard, the model-meta broadcast (what a peer rank
needs besides the NCCL-broadcast file image) and the result gather are backend-agnostic; only the device broadcast needs GPUs."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from whisper_b200 import dist as wdist


def test_shard_chunks_partition():
    for n in (0, 1, 7, 8, 63, 64, 65):
        for world in (1, 2, 3, 4, 8):
            parts = [wdist.shard_chunks(n, world, r) for r in range(world)]
            flat = [i for p in parts for i in p]
            assert flat == list(range(n))                     # disjoint, ordered, complete
            assert max(map(len, parts)) - min(map(len, parts)) <= 1
    assert wdist.shard_chunks(64, 8, 3) == list(range(24, 32))    # BASELINE config 4: 64 chunks, 8 per GPU
    assert wdist.batches(list(range(10)), 4) == [[0, 1, 2, 3], [4, 5, 6, 7], [8, 9]]
    with pytest.raises(ValueError):
        wdist.shard_chunks(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from whisper_b200 import capi, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        meta = None
        if rank == 0:
            m0 = capi.Model(synth.model_path("micro"))
            meta = m0.meta()
        blob = wdist.broadcast_bytes(meta, 0)
        m = capi.Model.from_meta(blob)
        ok = m.n_vocab == 51865 and m.special["sot"] == 50258 and m.token_text(17) == " t17" and m.file_image()[0] in (None, 0)
        ids = wdist.shard_chunks(5, world, rank)
        toks = np.array([[i * 10 + k for k in range(4)] for i in ids], np.int32).reshape(len(ids), 4)
        allt = wdist.gather_tokens(toks, 0)
        if rank == 0:
            ok = ok and allt.shape == (5, 4) and allt[:, 0].tolist() == [0, 10, 20, 30, 40]
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_meta_broadcast_and_gather_gloo_world2():
    from whisper_b200 import synth
    synth.model_path("micro")   # create the file before forking
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True), (1, True)]

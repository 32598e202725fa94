"""Multi-GPU plumbing: one process per GPU (torch.distributed), chunks sharded across ranks, weights broadcast once at load.

The path shards naturally: every 30 s chunk is an independent mel -> encoder -> decoder problem with its own KV
memories, so there is NO collective in the step loop.  The single collective is the broadcast of the ggml file image at load:
rank 0 reads the file once, every other rank receives the bytes over NCCL/NVLink straight into device memory and builds its
engine from that device image (wsp_engine_create_from_image) — the reference's analogue is iModel::clone sharing one set of
weight buffers between devices (Whisper/Whisper/ModelImpl.cpp:40-60).

Everything except the device broadcast is backend-agnostic, so the host logic is tested on CPU with gloo (tests/test_dist.py).
"""
from __future__ import annotations

import ctypes as C
import time

import numpy as np


def shard_chunks(n_chunks: int, world: int, rank: int):
    """Contiguous, balanced shard of chunk ids for `rank`: the first n_chunks % world ranks get one extra chunk."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    base, extra = divmod(n_chunks, world)
    start = rank * base + min(rank, extra)
    return list(range(start, start + base + (1 if rank < extra else 0)))


def batches(ids, batch: int):
    """Split a rank's chunk ids into per-step batches of at most `batch`."""
    return [ids[i:i + batch] for i in range(0, len(ids), batch)]


def broadcast_bytes(blob: bytes | None, src: int = 0) -> bytes:
    """Broadcast a host byte string (the model meta blob) with whatever backend the default group has."""
    import torch
    import torch.distributed as dist
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    n = torch.tensor([len(blob) if blob is not None else 0], dtype=torch.int64, device=dev)
    dist.broadcast(n, src)
    if blob is not None:
        t = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(dev)
    else:
        t = torch.empty(int(n.item()), dtype=torch.uint8, device=dev)
    dist.broadcast(t, src)
    return bytes(t.cpu().numpy().tobytes())


def gather_tokens(local_tokens: np.ndarray, dst: int = 0):
    """Gather per-rank [chunks][n] token arrays on `dst` (host-side result collection: a few hundred bytes per chunk)."""
    import torch.distributed as dist
    out = [None] * dist.get_world_size() if dist.get_rank() == dst else None
    dist.gather_object(np.asarray(local_tokens), out, dst=dst)
    if out is None:
        return None
    return np.concatenate([o for o in out if o.size], axis=0)


def load_broadcast(model_name: str, rank: int, local: int, world: int):
    """rank 0: open the model file, upload its image, broadcast meta (host) + image (device, NCCL).  Returns (Model, Engine, bcast_ms)."""
    import torch
    import torch.distributed as dist
    from . import capi, synth

    torch.cuda.set_device(local)
    if rank == 0:
        model0 = capi.Model(synth.model_path(model_name))
        meta = model0.meta()
        addr, size = model0.file_image()
    else:
        model0, meta, addr, size = None, None, 0, 0
    meta = broadcast_bytes(meta, 0)
    sz = torch.tensor([size], dtype=torch.int64, device="cuda")
    dist.broadcast(sz, 0)
    size = int(sz.item())
    image = torch.empty(size, dtype=torch.uint8, device="cuda")
    if rank == 0:
        host = np.ctypeslib.as_array(C.cast(addr, C.POINTER(C.c_uint8)), shape=(size,))
        image.copy_(torch.from_numpy(host))
    torch.cuda.synchronize()
    t0 = time.time()
    dist.broadcast(image, 0)          # the one collective: file image over NVLink / NVSwitch
    torch.cuda.synchronize()
    bcast_ms = (time.time() - t0) * 1e3
    model = model0 if rank == 0 else capi.Model.from_meta(meta)
    engine = capi.Engine(model, local, dev_image=image.data_ptr(), image_size=size)
    del image
    torch.cuda.empty_cache()
    return model, engine, bcast_ms

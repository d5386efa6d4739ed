"""Data-parallel plumbing for a batch of independent edits.

The hot path shards naturally on the edit (batch) dimension: no operator of the DiT or the VAE mixes samples, so edit i
simply runs on rank i % world.  The ONLY collective is the one-time replication of the weights from rank 0 (the
reference's closest analogue is `sync_model_states`, chronoedit/_ext/imaginaire/utils/distributed.py:351-443); there is
no per-step collective.  Works with the NCCL backend (GPU tensors, NVLink/NVSwitch) and with gloo (CPU tensors; used by
the CPU tests at world_size 2).
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Sequence

import torch
import torch.distributed as dist


def shard_edits(n_edits: int, rank: int, world: int) -> List[int]:
    """Indices of the edits rank `rank` processes: round-robin, every edit exactly once, |shards| differ by <= 1."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    return list(range(rank, n_edits, world))


def broadcast_tensors(tensors: Iterable[torch.Tensor], src: int = 0, bucket_bytes: int = 256 << 20) -> int:
    """Replicate `tensors` (same shapes/dtypes on every rank) from rank `src`, coalescing small tensors into flat
    buckets of at most `bucket_bytes` so that launch latency, not link count, is what is amortised (NVSwitch gives every
    GPU full bandwidth to every peer).  Returns the number of bytes broadcast."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return 0
    total = 0
    by_dtype: Dict[torch.dtype, List[torch.Tensor]] = {}
    for t in tensors:
        by_dtype.setdefault(t.dtype, []).append(t)
    for dtype, ts in by_dtype.items():
        bucket: List[torch.Tensor] = []
        size = 0

        def flush():
            nonlocal bucket, size, total
            if not bucket:
                return
            if len(bucket) == 1 and bucket[0].is_contiguous():
                dist.broadcast(bucket[0], src=src)
            else:
                flat = torch.cat([b.reshape(-1) for b in bucket])
                dist.broadcast(flat, src=src)
                off = 0
                for b in bucket:
                    n = b.numel()
                    b.copy_(flat[off: off + n].view_as(b))
                    off += n
            total += size
            bucket, size = [], 0

        for t in ts:
            nbytes = t.numel() * t.element_size()
            if nbytes >= bucket_bytes:
                flush()
                if t.is_contiguous():
                    dist.broadcast(t, src=src)
                else:
                    c = t.contiguous()
                    dist.broadcast(c, src=src)
                    t.copy_(c)
                total += nbytes
                continue
            if size + nbytes > bucket_bytes:
                flush()
            bucket.append(t)
            size += nbytes
        flush()
    return total


def broadcast_module_weights(module: torch.nn.Module, src: int = 0) -> int:
    """Broadcast every parameter / buffer of a module.  For a packed ChronoEditTransformer3DModel the fused buffers are
    broadcast instead of the per-parameter views (same storage, fewer and larger messages)."""
    keep = getattr(module, "_pack_keepalive", None)
    if keep and getattr(module, "_packed", False):
        return broadcast_tensors(keep.values(), src=src)
    return broadcast_tensors([p.data for p in module.parameters()] + [b for b in module.buffers()], src=src)


def gather_objects(local: Sequence, dst: int = 0):
    """Gather small per-rank python results (timings, checksums) on rank `dst`."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return [list(local)]
    out = [None] * dist.get_world_size() if dist.get_rank() == dst else None
    dist.gather_object(list(local), out, dst=dst)
    return out


# ----------------------------------------------------------------------------------------------------------------------
# Sequence parallelism for single-edit latency
# ----------------------------------------------------------------------------------------------------------------------
def enable_sequence_parallel(model, batch: int, frames: int, height: int, width: int, group=None) -> int:
    """Split the tokens of ONE edit over the ranks of `group` (one process per GPU of a node; reference: the Ulysses path of
    chronoedit_diffsynth/wan_video_new_chronoedit.py:330-355 and the TE ring of chronoedit/_src/networks/wan2pt1.py:352-353).

    Allocates this rank's peer region (ce_ipc_alloc), exchanges the 64-byte CUDA IPC handles through torch.distributed (plumbing
    only), opens the peers' regions and hands the pointer table to the C handle (ce_dit_sp_configure).  From then on
    `model(hidden_states, timestep, encoder_hidden_states, encoder_hidden_states_image)` -- called with the SAME full inputs on every
    rank -- computes 1/world of the tokens per rank, exchanges q/k/v and the attention output by peer stores over NVLink inside the
    kernels (no collective call on the data path), and returns the same full sample on every rank, bit-identical to the single-GPU
    forward.  Geometry-specific: call again for another latent shape.  Returns the region size in bytes."""
    import ctypes

    from . import _lib
    from ._lib import CEError, check

    if not dist.is_initialized():
        raise CEError("enable_sequence_parallel needs an initialised torch.distributed process group (one process per GPU)")
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    L = _lib.lib()
    if not model._packed:
        model.pack_weights()
    disable_sequence_parallel(model)
    if world == 1:
        return 0
    n = L.ce_dit_sp_region_bytes(model._handle, batch, frames, height, width, world)
    if n < 0:
        raise CEError(f"sequence parallel: tokens and heads must be divisible by the world size {world}")
    with torch.cuda.device(model.device):
        own = ctypes.c_void_p()
        check(L.ce_ipc_alloc(n, ctypes.byref(own)))
        hbuf = (ctypes.c_uint8 * 64)()
        check(L.ce_ipc_get_handle(own, hbuf))
        handles = [None] * world
        dist.all_gather_object(handles, bytes(hbuf), group=group)
        ptrs, opened = [], []
        for w in range(world):
            if w == rank:
                ptrs.append(own.value)
                continue
            p = ctypes.c_void_p()
            hb = (ctypes.c_uint8 * 64).from_buffer_copy(handles[w])
            check(L.ce_ipc_open(hb, ctypes.byref(p)))
            ptrs.append(p.value)
            opened.append(p)
        table = (ctypes.c_void_p * world)(*ptrs)
        check(L.ce_dit_sp_configure(model._handle, rank, world, table, n))
        torch.cuda.synchronize(model.device)
    dist.barrier(group=group)   # every region exists, is zeroed and is mapped everywhere before the first flag is written
    model._sp_state = {"own": own, "opened": opened, "world": world, "bytes": n}
    model._graphs = {}
    return int(n)


def disable_sequence_parallel(model) -> None:
    from . import _lib

    st = getattr(model, "_sp_state", None)
    if not st:
        return
    L = _lib.lib()
    torch.cuda.synchronize(model.device)
    L.ce_dit_sp_configure(model._handle, 0, 1, None, 0)
    for p in st["opened"]:
        L.ce_ipc_close(p)
    if dist.is_initialized():
        dist.barrier()
    L.ce_ipc_free(st["own"])
    model._sp_state = None

"""world_size-2 gloo test (CPU) of the N>1 path: edit sharding and the one-time weight broadcast."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from chronoedit_b200 import parallel


def test_shard_edits_covers_every_edit_once():
    for n in (0, 1, 7, 8, 13):
        for world in (1, 2, 4, 8):
            seen = sorted(i for r in range(world) for i in parallel.shard_edits(n, r, world))
            assert seen == list(range(n))
            sizes = [len(parallel.shard_edits(n, r, world)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        parallel.shard_edits(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import chronoedit_b200 as ce

        torch.manual_seed(100 + rank)  # different garbage on every rank before the broadcast
        m = ce.ChronoEditTransformer3DModel(num_attention_heads=2, in_channels=36, ffn_dim=256, num_layers=2, image_dim=1280,
                                            added_kv_proj_dim=256, text_dim=64)
        for p in m.parameters():
            p.data.copy_(torch.randn(p.shape).to(p.dtype))
        nbytes = parallel.broadcast_module_weights(m, src=0)
        expect = sum(p.numel() * p.element_size() for p in m.parameters())
        fingerprint = float(sum(p.data.double().sum() for p in m.parameters()))
        edits = parallel.shard_edits(5, rank, world)
        gathered = parallel.gather_objects([rank, fingerprint, edits, nbytes, expect])
        if rank == 0:
            q.put(gathered)
    finally:
        dist.destroy_process_group()


def test_weight_broadcast_and_sharding_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=180)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    (r0, f0, e0, n0, x0), (r1, f1, e1, n1, x1) = res
    assert (r0, r1) == (0, 1)
    assert f0 == f1, "weights differ across ranks after the broadcast"
    assert sorted(e0 + e1) == [0, 1, 2, 3, 4] and e0 == [0, 2, 4] and e1 == [1, 3]
    assert n0 == x0 and n1 == x1, "broadcast byte count must equal the parameter bytes"


def test_sequence_parallel_host_checks():
    """enable_sequence_parallel fails loudly (never silently falls back to a single-GPU forward) without a process group, and the
    region-size query rejects a missing handle; the data path itself needs NVLink peers (tests/test_gpu_seqpar.py, 2 GPUs)."""
    from chronoedit_b200 import _lib
    from chronoedit_b200._lib import CEError

    assert not dist.is_initialized()
    with pytest.raises(CEError, match="process group"):
        parallel.enable_sequence_parallel(object(), 1, 2, 4, 4)
    assert _lib.lib().ce_dit_sp_region_bytes(None, 1, 2, 4, 4, 2) == -1
    parallel.disable_sequence_parallel(object())   # a model that never enabled it: no-op

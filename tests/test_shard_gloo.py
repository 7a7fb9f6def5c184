"""World-size-2 gloo test (CPU) of the multi-GPU host logic: read/channel sharding and the
result gather.  The classifier itself is a stand-in (the CUDA path needs a GPU); what is
tested is that every read is classified exactly once and results land in batch order."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from riser_b200 import shard


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fake_classify(signals, ids):
    """Deterministic function of the read content and id only (like the real path)."""
    dec = np.array([int(s.sum()) % 4 for s in signals], dtype=np.uint8)
    p_on = np.array([[(int(s[0]) % 97) / 97.0, (len(i) % 7) / 7.0] for s, i in zip(signals, ids)], dtype=np.float32)
    sig_len = np.array([len(s) for s in signals], dtype=np.int32)
    return dec, p_on, sig_len


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)
    n = 37
    signals = [rng.integers(0, 1000, size=int(rng.integers(5, 50))).astype(np.int16) for _ in range(n)]
    ids = [f"read-{i}" * (1 + i % 3) for i in range(n)]
    channels = rng.integers(1, 513, size=n)
    calls = []

    def fn(sigs, rids):
        calls.append(len(sigs))
        return _fake_classify(sigs, rids)

    out = shard.classify_sharded(fn, signals, ids, channels, 2)
    want = _fake_classify(signals, ids)
    ok = all(np.array_equal(a, b) for a, b in zip(out, want))
    mine = shard.shard_indices(channels, rank, world)
    q.put((rank, ok, calls, len(mine)))
    dist.destroy_process_group()


def test_two_rank_sharding_and_gather():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] for r in res)
    assert sum(r[3] for r in res) == 37 and all(r[2] == [r[3]] for r in res)


def test_shard_indices_partition():
    keys = np.arange(1, 513)
    parts = [shard.shard_indices(keys, r, 8) for r in range(8)]
    assert sorted(np.concatenate(parts).tolist()) == list(range(512))
    assert all(len(p) == 64 for p in parts)

"""Read sharding across the GPUs of one box: reads are independent, so
the path shards by read with NO data-path collective; weights are replicated; one process
per GPU.  The only exchange is the gather of per-read results (<= 9 bytes per read).

Offline: read r -> rank r mod G.  Live: channel c -> rank c mod G, which keeps a read's
poly(A) cache entry (control.py:24, keyed by read id) on one worker.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_indices(keys, rank, world):
    """Indices of the items this rank owns: key mod world == rank (keys: read index
    offline, channel number live)."""
    keys = np.asarray(keys, dtype=np.int64)
    return np.flatnonzero(keys % world == rank)


def gather_decisions(local_idx, decisions, p_on, sig_len, n_total, group=None):
    """Reassemble whole-batch results on every rank from per-rank shards.

    local_idx: indices (into the whole batch) this rank classified; decisions uint8 [n],
    p_on float32 [n, M], sig_len int32 [n].  Returns (decisions [n_total], p_on [n_total, M],
    sig_len [n_total]).  Works on any backend (gloo on CPU for tests, nccl on GPUs): the
    payload is tiny, so it is a plain all_gather of padded tensors -- there is no collective
    on the data path itself."""
    world = dist.get_world_size(group)
    M = p_on.shape[1] if p_on.ndim == 2 else 1
    # NCCL moves device memory only: stage the (tiny) payload on this rank's GPU there
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    counts = torch.zeros(world, dtype=torch.int64, device=dev)
    counts[dist.get_rank(group)] = len(local_idx)
    dist.all_reduce(counts, group=group)
    counts = counts.cpu()
    cap = int(counts.max())
    pack = torch.zeros(cap, 3 + M, dtype=torch.float64)
    n = len(local_idx)
    if n:
        pack[:n, 0] = torch.from_numpy(np.asarray(local_idx, dtype=np.float64))
        pack[:n, 1] = torch.from_numpy(np.asarray(decisions, dtype=np.float64))
        pack[:n, 2] = torch.from_numpy(np.asarray(sig_len, dtype=np.float64))
        pack[:n, 3:] = torch.from_numpy(np.asarray(p_on, dtype=np.float64).reshape(n, M))
    pack = pack.to(dev)
    bufs = [torch.zeros_like(pack) for _ in range(world)]
    dist.all_gather(bufs, pack, group=group)
    bufs = [b.cpu() for b in bufs]
    out_dec = np.full(n_total, 4, dtype=np.uint8)          # SKIPPED
    out_p = np.zeros((n_total, M), dtype=np.float32)
    out_len = np.zeros(n_total, dtype=np.int32)
    for r in range(world):
        k = int(counts[r])
        if not k:
            continue
        b = bufs[r][:k].numpy()
        idx = b[:, 0].astype(np.int64)
        out_dec[idx] = b[:, 1].astype(np.uint8)
        out_len[idx] = b[:, 2].astype(np.int32)
        out_p[idx] = b[:, 3:].astype(np.float32)
    return out_dec, out_p, out_len


def classify_sharded(classify_fn, signals, read_ids, keys, n_models, group=None):
    """Run ``classify_fn(signals_subset, ids_subset) -> (decisions, p_on, sig_len)`` on this
    rank's shard (key mod world == rank) and gather everyone's results."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    mine = shard_indices(keys, rank, world)
    if len(mine):
        dec, p_on, sig_len = classify_fn([signals[i] for i in mine], [read_ids[i] for i in mine])
    else:
        dec, p_on, sig_len = (np.zeros(0, np.uint8), np.zeros((0, n_models), np.float32), np.zeros(0, np.int32))
    return gather_decisions(mine, dec, p_on, sig_len, len(signals), group=group)

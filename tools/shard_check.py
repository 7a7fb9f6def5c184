"""Multi-GPU check of the sharded path with the REAL classifier (one process per GPU, NCCL only for the gather of the
per-read results): every rank classifies the reads r with r mod G == rank through BatchedClassifier.classify_batch,
shard.classify_sharded gathers decisions / probabilities / window lengths, and every rank compares the assembled
result with what it gets by classifying the whole batch alone on its own GPU.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
        tools/shard_check.py [n_reads]
"""
import logging
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from riser_b200 import Kit, SignalProcessor, Model, BatchedClassifier, synth, shard      # noqa: E402
from riser_b200.config import shipped_config                                             # noqa: E402


def main():
    n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 96
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    log = logging.getLogger("shard")
    targets = ["mRNA", "mtRNA"]
    models = [Model(synth.state_dict(synth.TARGET_SEEDS[t]), shipped_config(), log, t) for t in targets]
    clf = BatchedClassifier(models, SignalProcessor(Kit.create_from_version("RNA002")))
    reads = synth.raw_reads(21, n_reads, min_body=3000, max_body=15000, frac_no_polya=0.2)
    sigs, ids = [s for _, s in reads], [r for r, _ in reads]

    def classify(sub_sigs, sub_ids):
        res = clf.classify_batch(sub_sigs, sub_ids, {}, 0.9, "deplete")
        return res.decisions, res.p_on, res.sig_len

    dec, p_on, sig_len = shard.classify_sharded(classify, sigs, ids, np.arange(n_reads), len(targets))
    whole = clf.classify_batch(sigs, ids, {}, 0.9, "deplete")
    assert np.array_equal(sig_len, whole.sig_len), "window lengths differ"
    assert np.array_equal(dec, whole.decisions), "decisions differ"
    # (skipped reads carry NaN probabilities: compare them as equal)
    assert np.array_equal(p_on, whole.p_on, equal_nan=True), \
        "probabilities differ (same kernels, same inputs: must be bit-identical)"
    mine = shard.shard_indices(np.arange(n_reads), rank, world)
    print(f"rank {rank}/{world}: OK -- {n_reads} reads, {len(mine)} classified here, "
          f"{int((whole.sig_len > 0).sum())} assessed, decisions {np.bincount(dec, minlength=5).tolist()}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

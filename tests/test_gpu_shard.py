"""The sharded path on real GPUs (BASELINE config 4): needs >= 2 B200s on the box, skipped otherwise
(`gpurun --gpus 2 -- python -m pytest tests/test_gpu_shard.py -m gpu`).  The host-side logic of the shard / gather
is covered on the CPU with gloo in tests/test_shard_gloo.py."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs on one box")
def test_classify_sharded_with_the_real_classifier_on_two_gpus():
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29517",
                        os.path.join(ROOT, "tools", "shard_check.py"), "96"],
                       cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-3000:])
    assert "rank 0/2: OK" in r.stdout and "rank 1/2: OK" in r.stdout, r.stdout[-2000:]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs on one box")
def test_bench_reads_mode_on_two_gpus():
    """bench.py --reads: BASELINE config 4's sharded offline sweep; it asserts itself that the gathered decisions equal
    the single-GPU decisions for every read."""
    import json
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29518",
                        os.path.join(ROOT, "bench.py"), "--gpus", "2", "--reads", "20000", "--batch", "1024"],
                       cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-3000:])
    line = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["n_gpus"] == 2 and line["scaling"] == "strong"
    assert line["sharding"]["decisions_equal_single_gpu"] == line["sharding"]["of"] == 20000

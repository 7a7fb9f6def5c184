"""bench.py's output contract, checked on the CPU through the reference arm: exactly one line on stdout, valid JSON,
the keys the driver reads.  (The GPU arm prints the same line plus roofline / clocks; it needs a B200.)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra):
    env = {**os.environ, **env_extra}
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0", "--ref-reads", "2", "--length", "4096"], cwd=ROOT, env=env,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout


def test_reference_arm_prints_one_json_line():
    out = _run({"RANK": "0", "WORLD_SIZE": "1"})
    lines = [ln for ln in out.splitlines() if ln.strip()]
    assert len(lines) == 1, out
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "reads/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["n_gpus"] == 1 and d["vs_baseline"] is None
    # the reference itself when it can be imported (sources here, oracle/_ref on the GPU box), else the port
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and "model" not in d["config"]
    # `config` must be the same object in both arms: the GPU arm builds it with the same function
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == bench.shared_config(bench.BATCH, 4096)


def test_reference_arm_falls_back_to_the_port_without_the_reference(tmp_path):
    """No /root/reference and no oracle/_ref: the arm still runs (kind "port")."""
    code = ("import sys, json; sys.path.insert(0, %r); from oracle import refshim; "
            "refshim.REF_ROOT = %r; refshim.BUILT = %r; import bench; "
            "p = bench.CpuPath(); print(p.kind)") % (ROOT, str(tmp_path / "no_ref"), str(tmp_path / "no_built"))
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout.strip().splitlines()[-1] == "port"


def test_reference_arm_other_ranks_are_silent():
    assert _run({"RANK": "1", "WORLD_SIZE": "2"}).strip() == ""

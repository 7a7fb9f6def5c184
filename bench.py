#!/usr/bin/env python
"""Benchmark of the RISER read-classification hot path (contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" = one pass of the hot path (median/MAD normalise + outlier smoothing ->
12-layer ConvNet -> softmax -> accept/reject decision) over one fixed batch of
synthetic already-trimmed squiggle chunks: BASELINE.json configs[1], RNA004 model shape,
4096 chunks of 4 s at 4 kHz (16,000 int16 samples) per GPU.

  value    reads classified / s, whole job, inputs resident in HBM (device timed, max over ranks)
  e2e      the same through the public API with HOST int16 buffers: pinned H2D copy of the
           batch and D2H of decisions + probabilities inside the timed region
  roofline the tcgen05 conv stack (layers 1..11): algorithmic FLOPs / CUDA-event time of that
           stage, against the measured bf16 peak in MEASURED_PEAKS.json
  cpu_baseline / --impl reference: the oracle port of the reference's CPU path
           (numpy normalise + torch-CPU ConvNet per read, riser/control.py:63-71) on the
           box's host cores, on a bounded sample of the same workload.
"""
import argparse
import json
import logging
import os
import subprocess
import sys
import threading
import time

# torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arm (rank 0 alone does the work) is to use all host threads
if "reference" in sys.argv and os.environ.get("RANK", "0") == "0" and os.environ.get("OMP_NUM_THREADS") == "1":
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BATCH = 4096            # chunks per GPU per step (BASELINE.json configs[1])
LENGTH = 16000          # 4 s at 4 kHz
CHANNELS = [20, 30, 45, 67, 100, 150, 225, 337, 505, 757, 1135, 1702]
METRIC = "reads classified/sec (4 s chunks)"
UNIT = "reads/s"


def flops_per_read(length):
    total, cin, l = 0, 1, int(length)
    for c in CHANNELS:
        total += 2 * 3 * cin * c * l
        cin, l = c, l // 2
    return total + 2 * CHANNELS[-1] * 2


def conv_stack_flops_per_read(length):
    """Layers 1..11 only (what the tcgen05 kernel executes)."""
    return flops_per_read(length) - 2 * 3 * 1 * CHANNELS[0] * length - 2 * CHANNELS[-1] * 2


def workload_name(batch, length):
    return (f"BASELINE configs[1]: RNA004-shape mRNA ConvNet, {batch}-chunk fixed batch per GPU, "
            f"{length} int16 samples per chunk (4 s @ 4 kHz), already trimmed")


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return p.get("bf16_tflops_sustained", p.get("bf16_tflops")), p.get("hbm_gbs"), "measured"
    return 1400.0, 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons (NVML; nvidia-smi as fallback) while the
    timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # torch's device index follows CUDA_VISIBLE_DEVICES; map through the UUID
            uuid = torch.cuda.get_device_properties(index).uuid
            h = None
            for i in range(pynvml.nvmlDeviceGetCount()):
                hi = pynvml.nvmlDeviceGetHandleByIndex(i)
                u = pynvml.nvmlDeviceGetUUID(hi)
                u = u.decode() if isinstance(u, bytes) else u
                if str(uuid) in u:
                    h = hi
            self.handle = h or pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        mhz = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
        try:
            r = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
        except Exception:
            r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
        flags = [bool(r & n.nvmlClocksThrottleReasonHwSlowdown), bool(r & n.nvmlClocksThrottleReasonHwThermalSlowdown),
                 bool(r & n.nvmlClocksThrottleReasonSwThermalSlowdown), bool(r & n.nvmlClocksThrottleReasonSwPowerCap)]
        return [str(mhz), str(self.max_mhz)] + ["Active" if f else "Not Active" for f in flags]

    def run(self):
        while not self.stop_flag:
            try:
                if self.nvml is not None:
                    self.samples.append(self._sample_nvml())
                    time.sleep(0.005)
                    continue
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 6:
                    self.samples.append(parts)
            except Exception:
                time.sleep(0.05)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        mhz = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": mhz[len(mhz) // 2] if mhz else None,
                "sm_max_mhz": int(self.samples[0][1]) if self.samples[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.samples),
                "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def cpu_reads(sample_reads, length, seed=4321):
    """The CPU legs' reads: a pool of 256 synthetic chunks repeated (as the GPU batch repeats its pool of 512)."""
    from riser_b200 import synth
    pool = synth.body_batch(seed, min(sample_reads, 256), length)
    return [pool[i % len(pool)] for i in range(sample_reads)]


def time_cpu_path(X, state, threads):
    """The oracle port of riser/control.py:63-71 per read, on `threads` host threads."""
    from oracle import preprocess_oracle as pp
    from oracle import convnet_oracle as net
    torch.set_num_threads(threads)
    for r in range(min(2, len(X))):
        net.classify(state, pp.mad_normalise(X[r]))
    t0 = time.perf_counter()
    for r in range(len(X)):
        net.classify(state, pp.mad_normalise(X[r]))
    return len(X) / (time.perf_counter() - t0)


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path (oracle port) on the host cores."""
    if rank != 0:
        return
    from riser_b200 import synth
    cores = os.cpu_count() or 1
    threads = min(cores, torch.get_num_threads() if torch.get_num_threads() > 1 else cores)
    state = synth.state_dict(0)
    per_step = args.ref_reads
    X = cpu_reads(per_step, args.length)
    for _ in range(args.warmup):
        time_cpu_path(X[:2], state, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        time_cpu_path(X, state, threads)
    dt = time.perf_counter() - t0
    value = per_step * args.steps / dt
    sample = f"{per_step} reads of {args.length} samples per step, per-read loop (normalise + classify), oracle port"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args.batch, args.length), "sample": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    _emit(line)


def _emit(line):
    """The one JSON line goes to the process's ORIGINAL stdout; everything else that writes to fd 1 during the
    run (NCCL's version banner, library chatter) has been pointed at stderr by main()."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--length", type=int, default=LENGTH)
    ap.add_argument("--precision", type=int, default=None, help="0 F16, 1 F16_W2, 2 F16_X3, 3 F16_F8; default = riser_b200.model.DEFAULT_PRECISION")
    ap.add_argument("--chunk", type=int, default=None)
    ap.add_argument("--ref-reads", type=int, default=1024, help="reads per step of the --impl reference arm (~2.6 s of CPU work)")
    ap.add_argument("--cpu-sample", type=int, default=4096, help="reads of the cpu_baseline leg (default: the whole batch, ~10 s)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch.distributed as dist
    from riser_b200 import Kit, SignalProcessor, Model, BatchedClassifier, RaggedBatch, synth, _lib, model as rmodel
    from riser_b200.config import shipped_config

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    log = logging.getLogger("bench")
    cfg = shipped_config()
    precision = rmodel.DEFAULT_PRECISION if args.precision is None else args.precision
    mdl = Model(synth.state_dict(0), cfg, log, "mRNA", precision=precision)
    proc = SignalProcessor(Kit.create_from_version("RNA004"))
    clf = BatchedClassifier([mdl], proc, chunk=args.chunk)
    # the bench shape is the 4 s chunk BASELINE names (longer than the live path's 8,615 cap)
    clf.max_len = args.length
    clf.ld = (args.length + 3) & ~3
    B, L = args.batch, args.length

    # synthetic inputs: distinct reads per rank, generated on the host, resident in HBM
    pool = synth.body_batch(100 + rank, min(B, 512), L)
    host = torch.empty(B, L, dtype=torch.int16).pin_memory()
    hv = host.numpy()
    for i in range(B):
        hv[i] = pool[i % len(pool)]
    sigs = [hv[i] for i in range(B)]
    batch = RaggedBatch(sigs, dev)
    start = torch.zeros(B, dtype=torch.int32, device=dev)
    length = torch.full((B,), L, dtype=torch.int32, device=dev)
    # L2 flush buffer (written between timed steps); inputs (131 MB int16 + 262 MB fp32) exceed L2 anyway
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def step(events=None):
        return clf.run_windows(batch, start, length, 0.9, "deplete", events=events)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()

    # ---- device-timed steps (value) with per-stage events for the roofline
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    conv_events = []
    barrier()
    for k in range(args.steps):
        flush.zero_()
        ev[k][0].record()
        step(events=conv_events)
        ev[k][1].record()
    barrier()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = sum(step_ms)
    conv_ms = sum(a.elapsed_time(b) for a, b in conv_events) / args.steps
    # the bracket covers layer 0 (CUDA cores) + conv layers 1..11; take layer 0's launches out
    conv_ms -= mdl.time_layer0(clf._buffers(B)["x"], length, L)
    # ---- the HBM-bound stage: riser_normalise alone (6 bytes per sample: int16 in, fp32 out), L2 flushed
    # (a kernel timed alone: the board is first given half a second to leave the power-capped clock the timed
    #  steps drove it into -- the HBM peak it is compared with was measured the same way)
    L_ = _lib.lib()
    xbuf = clf._buffers(B)["x"]
    torch.cuda.synchronize()
    time.sleep(0.5)
    nev = []
    for _ in range(7):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        _lib.check(L_.riser_normalise(_lib.ptr(batch.sig), _lib.ptr(batch.off), _lib.ptr(start), _lib.ptr(length), B, L,
                                      _lib.ptr(xbuf), xbuf.stride(0), None, _lib.stream_ptr()), "riser_normalise")
        b.record()
        nev.append((a, b))
    torch.cuda.synchronize()
    norm_ms = sorted(a.elapsed_time(b) for a, b in nev[2:])[(len(nev) - 2) // 2]      # first two launches: warm-up
    # ---- e2e: the public streaming API with HOST buffers.  Every step copies its 131 MB of
    #      pinned int16 input H2D and brings decisions + probabilities back D2H; the copy of
    #      step k+1 overlaps the kernels of step k (FixedBatchPipeline, two slots).
    from riser_b200 import FixedBatchPipeline
    pipe = FixedBatchPipeline(clf, B, L, 0.9, "deplete")
    hosts = [host, host.clone().pin_memory()]
    for k in range(2):
        pipe.result(pipe.submit(hosts[k % 2]))
    barrier()
    e2e_start, e2e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2e_start.record()
    pipe.copy_stream.wait_event(e2e_start)
    tickets = []
    for k in range(args.steps):
        tickets.append(pipe.submit(hosts[k % 2]))
        if k >= 1:
            pipe.result(tickets[k - 1])
    dec_np, _ = pipe.result(tickets[-1])
    e2e_end.record()
    e2e_end.synchronize()
    e2e_ms = e2e_start.elapsed_time(e2e_end)
    barrier()
    dec_host = torch.from_numpy(dec_np.copy())
    sampler.stop_flag = True
    sampler.join(timeout=2)

    t = torch.tensor([total_ms, e2e_ms, conv_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms, conv_ms = (float(v) for v in t.cpu())
    n_dec = int((dec_host.numpy() != 4).sum())

    if rank == 0:
        tensor_peak, hbm_peak, peak_src = peaks()
        reads = B * world * args.steps
        value = reads / (total_ms * 1e-3)
        fused = mdl.plan(B, L).fused_layer0       # layer 0 then runs inside the conv kernel
        conv_flops = (conv_stack_flops_per_read(L) + (2 * 3 * CHANNELS[0] * L if fused else 0)) * B
        achieved = conv_flops / (conv_ms * 1e-3) / 1e12
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "conv_stack_traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                traffic = json.load(f).get(f"B{B}_L{L}_p{precision}")
        launches_per_step = 1 + mdl.launches(B, L, clf.chunk) + 1
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": {0: "f16", 1: "f16", 2: "f16", 3: "f16+e4m3"}[precision],
            "data": "synthetic",
            "config": {"workload": workload_name(B, L), "precision_mode": precision,
                       "arithmetic": {0: "f16 operands, f32 accumulate (1 tcgen05 pass)",
                                      1: "f16, weights split hi+lo (2 passes), f32 accumulate",
                                      2: "f16, weights and activations split hi+lo (3 passes), f32 accumulate",
                                      3: "f16 pass + e4m3 correction pass carrying the hi/lo terms (2 pass-equivalents; "
                                         "layers 1-5 as mode 2), f32 accumulate; layer 0 and the head in f32, "
                                         "normalise in f64"}[precision],
                       "chunk": clf.chunk,
                       "cache": "L2 flushed between steps (256 MiB write); inputs 393 MB > L2",
                       "flops_per_read": flops_per_read(L), "decisions_made": n_dec},
            "roofline": {"bound": "tensor", "kernel": "fused01_kernel + conv_tc_kernel (layers %d-11, 11 launches per forward)" % (0 if fused else 1),
                         "achieved": achieved, "peak": tensor_peak, "unit": "TFLOP/s", "frac": achieved / tensor_peak,
                         "traffic": traffic, "peak_source": f"bf16_tflops_sustained, {peak_src}",
                         "executed_tflops": achieved * {0: 1, 1: 2, 2: 3, 3: 2}[precision],
                         "executed_note": "fp16-pass-equivalents of tcgen05 work per algorithmic FLOP: 1 (F16), "
                                          "2 (F16_W2), 3 (F16_X3), 2 (F16_F8: one fp16 pass + two e4m3 passes at twice the rate)",
                         "share_of_step": conv_ms / (total_ms / args.steps)},
            "roofline_normalise": {"bound": "hbm", "kernel": "normalise_kernel", "achieved": 6 * L * B / norm_ms / 1e6,
                                   "peak": hbm_peak, "unit": "GB/s", "frac": 6 * L * B / norm_ms / 1e6 / hbm_peak,
                                   "ms": norm_ms, "algorithmic_bytes_per_read": 6 * L,
                                   "peak_source": f"hbm_gbs, {peak_src}"},
            "e2e": {"value": reads / (e2e_ms * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": pipe.h2d_bytes, "d2h_bytes_per_step": pipe.d2h_bytes,
                    "overlap": "H2D of step k+1 overlaps kernels of step k (2 slots)"},
            "gpu_launches": launches_per_step * args.steps * 2,
            "clocks": sampler.summary(),
        }
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            X = cpu_reads(args.cpu_sample, L)
            v = time_cpu_path(X, synth.state_dict(0), cores)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{args.cpu_sample} reads of {L} samples, per-read loop "
                                              "(oracle normalise + torch-CPU ConvNet), torch threads = cores"}
        _emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

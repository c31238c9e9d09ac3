#!/usr/bin/env python
"""bench.py -- training samples/sec of the Distributed-IB hot path on the BASELINE.json workload.

Workload C0: 16 scalar features, positional encoding (4 frequencies), per-feature encoders
[128,128] relu -> (mu, logvar) with E=32, integration MLP [256,256] -> 1 logit, BCE-from-logits + beta*sum KL,
Keras-Adam lr 3e-4, fp32, batch 65 536 PER GPU (weak scaling), synthetic N(0,1) inputs with labels from a fixed
nonlinear teacher, random-init (glorot-uniform) weights.  One "step" = forward + backward + gradient all-reduce
(N>1) + Adam on one batch.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference ...                           # the reference graph's CPU twin (host cores)

Prints ONE JSON line (rank 0).  `value` = device-resident throughput; `e2e` = the same step through the public API
(model.train_on_batch) from pinned HOST buffers with the H2D copy and the D2H read of the metrics inside the timed
region.  See DESIGN.md section "Measurement" for the roofline arithmetic.
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "training samples/sec (16-feat synthetic, batch 65k) at 1/2/4/8 B200; HBM GB/s"
F, E, ENC, INT, OUT, BATCH = 16, 32, [128, 128], [256, 256], 1, 65536
LR = 3e-4
N_DISTINCT_BATCHES = 16          # N = 16 * 65536 rows, 16 steps per epoch


def synth_batches(rank, nb, batch, pinned):
    """x ~ N(0,1) float32, y from a fixed nonlinear teacher with interactions."""
    import torch
    rng = np.random.default_rng(1000 + rank)
    xs, ys = [], []
    for _ in range(nb):
        x = rng.standard_normal((batch, F), dtype=np.float32)
        y = (x[:, 0] * x[:, 1] + np.sin(2 * x[:, 2]) + 0.5 * x[:, 3]
             + 0.1 * rng.standard_normal(batch, dtype=np.float32) > 0).astype(np.float32)[:, None]
        xt, yt = torch.from_numpy(x), torch.from_numpy(y)
        if pinned:
            xt, yt = xt.pin_memory(), yt.pin_memory()
        xs.append(xt)
        ys.append(yt)
    return xs, ys


def algorithmic_macs(batch):
    """MACs per launch group for the C0 shapes (SURVEY.md 8d 'ALGORITHMIC work per sample')."""
    w_in = 5
    enc = [(w_in, ENC[0]), (ENC[0], ENC[1]), (ENC[1], 2 * E)]
    integ = [(F * E, INT[0]), (INT[0], INT[1]), (INT[1], OUT)]
    macs = {}
    for j, (k, n) in enumerate(enc):
        for kind in ("fwd", "wgrad"):
            macs[f"enc_{kind}_l{j}"] = F * k * n * batch
        if j >= 1:
            macs[f"enc_dgrad_l{j}"] = F * k * n * batch
    for j, (k, n) in enumerate(integ):
        for kind in ("fwd", "wgrad", "dgrad"):
            macs[f"int_{kind}_l{j}"] = k * n * batch
    fwd = sum(v for k, v in macs.items() if "_fwd_" in k)
    train = sum(macs.values())
    # fused per-feature encoder kernels: algorithmic work = the layers they replace (the backward kernel's on-chip
    # recomputation of the forward is overhead, not algorithmic work)
    macs["enc_fused_fwd"] = sum(v for k, v in macs.items() if k.startswith("enc_fwd_"))
    macs["enc_fused_bwd"] = sum(v for k, v in macs.items() if k.startswith("enc_dgrad_") or k.startswith("enc_wgrad_"))
    return macs, fwd, train


NCU_KERNEL_OF_GROUP = {"enc_fused_bwd": "dib_enc_fused_bwd_kernel", "enc_fused_fwd": "dib_enc_fused_fwd_kernel"}


def ncu_dram_traffic(group):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed
    `ncu --set full` capture of this same command (profiles/r01_final_ncu_full.txt); None if absent."""
    path = os.path.join(ROOT, "profiles", "r01_final_ncu_full.txt")
    kern = NCU_KERNEL_OF_GROUP.get(group)
    if not kern or not os.path.exists(path):
        return None
    total, inside, scale = 0.0, False, {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for line in open(path):
        if line.startswith("## "):
            if inside:
                break
            inside = kern in line
        elif inside and ("dram__bytes_read.sum" in line or "dram__bytes_write.sum" in line):
            parts = line.split()
            total += float(parts[1]) * scale.get(parts[2], 1.0)
    return total or None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md 'clocks' line)."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.proc = None
        self.path = f"/tmp/dib_clocks_{os.getpid()}.csv"
        try:
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(index)],
                stdout=self.fh, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.fh.close()
        rows = []
        for line in open(self.path):
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 7:
                try:
                    rows.append((float(parts[0]), float(parts[1]), float(parts[2]), parts[3:7]))
                except ValueError:
                    pass
        os.unlink(self.path)
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        pw = np.array([r[2] for r in rows])
        load = [r for r in rows if r[2] >= 0.5 * pw.max()] or rows
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in load for i in range(4) if r[3][i].lower().startswith("active")})
        return {"sm_mhz": float(np.median([r[0] for r in load])), "sm_max_mhz": float(rows[0][1]),
                "power_w_max": float(pw.max()), "samples": len(rows), "reasons": reasons}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"hbm_gbs": d["hbm_gbs"], "tflops_burst": d["bf16_tflops"],
                "tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "tflops_burst": 1590.0, "tflops_sustained": 1400.0, "source": "fallback"}


_BEST_THREADS = None


def cpu_twin_rate(sample_rows, steps, warmup, threads=None):
    """samples/s of the reference graph's eager CPU twin.  `threads=None`: use the thread count that is fastest on
    this host (a 128-thread pool on a shared box can be >100x slower than 16 threads for these small per-feature
    ops, so "all the threads it can use" is found by a short probe rather than assumed to be os.cpu_count())."""
    global _BEST_THREADS
    import torch
    from oracle import dib_oracle as O
    from oracle.torch_twin import time_train_steps
    cfg = O.DIBConfig([1] * F, ENC, INT, OUT)
    rng = np.random.default_rng(0)
    x = rng.standard_normal((sample_rows, F), dtype=np.float32)
    y = (x[:, 0] * x[:, 1] > 0).astype(np.float32)[:, None]
    if threads is None:
        if _BEST_THREADS is None:
            ncpu = os.cpu_count() or 1
            cands = sorted({c for c in (4, 8, 16, 32, 64, ncpu) if c <= ncpu})
            best = None
            for c in cands:
                med, _, _ = time_train_steps(cfg, O.LOSS_BCE_LOGITS, x[:2048], y[:2048], LR, 1, 1, threads=c)
                if best is None or med < best[0]:
                    best = (med, c)
            _BEST_THREADS = best[1]
        threads = _BEST_THREADS
    med, total, threads = time_train_steps(cfg, O.LOSS_BCE_LOGITS, x, y, LR, steps, warmup, threads=threads)
    return sample_rows / med, med, threads


def run_reference(args, rank):
    """The reference graph's CPU twin (oracle/torch_twin.py; TensorFlow is not installable here) on the host cores."""
    if rank != 0:
        return
    # bounded sample: probe a small batch, then size the per-step sample so that the whole run (warm-up + timed steps)
    # takes about 90 s of CPU time, at most one full batch and at most ~1.5 s per step
    _, probe, _ = cpu_twin_rate(4096, 1, 1)
    per_step = min(1.5, 90.0 / max(args.steps + args.warmup, 1))
    rows = int(min(BATCH, max(1024, 2 ** int(np.log2(max(per_step / probe, 0.25) * 4096)))))
    rate, med, threads = cpu_twin_rate(rows, args.steps, args.warmup)
    sample = f"{rows} of {BATCH} rows per step, {args.steps} timed steps, PyTorch-CPU eager twin of models.py (TF unavailable)"
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": med * 1e3 * BATCH / rows,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus, "cpu"),
        "cpu_baseline": {"value": rate, "unit": "samples/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def workload_config(n_gpus, precision, per_gpu=None):
    per_gpu = BATCH if per_gpu is None else per_gpu
    return {"workload": f"C0: 16 scalar features x batch {per_gpu}/GPU, PE(4 freq) -> enc[128,128] -> E=32 -> int[256,256] -> 1, "
                        "BCE-from-logits + beta*KL, Keras-Adam",
            "global_batch": per_gpu * n_gpus, "per_gpu_batch": per_gpu, "parallelism": f"dp{n_gpus}", "precision": precision,
            "l2": "no explicit flush: every step streams its intermediates through the 126 MB L2 (tensor-core path ~0.33 GB of "
                  "fp16 activations/gradients/partials per 65536 rows, fp32 path ~3 GB), evicting the 16 rotating input "
                  "batches between their uses"}


JSON_OUT = None


def claim_stdout():
    """stdout must carry exactly ONE JSON line.  Libraries print there too (NCCL's version banner, torchrun notices), so
    keep a private handle on the real stdout for the result and point fd 1 at stderr for everything else."""
    global JSON_OUT
    if JSON_OUT is None:
        sys.stdout.flush()
        JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    JSON_OUT.write(json.dumps(line) + "\n")
    JSON_OUT.flush()


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("DIB_PRECISION", "tf32"), choices=["tf32", "fp32"],
                    help="tf32 = tensor-core mode (default, the headline); fp32 = exact CUDA-core parity path")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak (default, the driver's contract): 65536 rows per GPU; strong: the 65536-row global batch is "
                         "split over the GPUs (SURVEY 8d 'the metric as stated')")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    import dib_b200
    from dib_b200 import _lib

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run"
    PB = BATCH // world if args.scaling == "strong" else BATCH          # rows per GPU per step
    lib = _lib.load()

    model = dib_b200.DistributedIBNet([1] * F, ENC, INT, OUT, use_positional_encoding=True,
                                      number_positional_encoding_frequencies=5, activation_fn="relu",
                                      feature_embedding_dimension=E, seed=1, precision=args.precision)
    model.compile(optimizer=dib_b200.Adam(LR), loss=dib_b200.losses.BinaryCrossentropy(from_logits=True),
                  metrics=["accuracy"])
    model.beta.assign(1e-3)
    xs_h, ys_h = synth_batches(rank, N_DISTINCT_BATCHES, PB, pinned=True)
    xs_d = [x.cuda(non_blocking=True) for x in xs_h]
    ys_d = [y.cuda(non_blocking=True) for y in ys_h]
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def device_step(i):
        model._train_step(xs_d[i % N_DISTINCT_BATCHES], ys_d[i % N_DISTINCT_BATCHES],
                          global_batch=PB * world, sample_offset=rank * PB)

    # ---------------- device-resident timed region -> value
    sampler = ClockSampler(local_rank) if rank == 0 else None      # sampled under load: warm-up + timed region
    for i in range(args.warmup):
        device_step(i)
    barrier()
    launches0 = int(lib.dib_launch_count())
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(args.steps):
        device_step(i)
    ev1.record()
    barrier()
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    launches = int(lib.dib_launch_count()) - launches0
    clocks = sampler.stop() if sampler else None
    ms_per_step = ms_total / args.steps
    value = PB * world / (ms_per_step * 1e-3)

    # ---------------- end to end through the public API from pinned host buffers -> e2e
    for i in range(3):
        model.train_on_batch(xs_h[i], ys_h[i], sync=False).get()
    barrier()
    ev0.record()
    pend = []
    for i in range(args.steps):
        # every step: pinned host batch -> H2D (copy stream) -> step -> async D2H of the metrics into pinned memory
        pend.append(model.train_on_batch(xs_h[i % N_DISTINCT_BATCHES], ys_h[i % N_DISTINCT_BATCHES], sync=False))
        if len(pend) > 4:
            pend.pop(0).get()            # the host reads every step's result, a few steps behind the device
    out = [p.get() for p in pend][-1]
    ev1.record()
    barrier()
    e2e_ms = max_over_ranks(ev0.elapsed_time(ev1)) / args.steps
    e2e = {"value": PB * world / (e2e_ms * 1e-3), "unit": "samples/s", "ms_per_step": e2e_ms,
           "h2d_bytes_per_step": int(xs_h[0].numel() * 4 + ys_h[0].numel() * 4),
           "d2h_bytes_per_step": int((F + 3) * 4), "api": "DistributedIBNet.train_on_batch(host x, host y, sync=False).get() -> metrics dict (H2D on a copy stream, D2H async)",
           "last_loss": out["loss"]}

    # ---------------- per-launch-group CUDA-event profile of the same steps -> roofline (rank 0)
    roofline = None
    if rank == 0:
        import ctypes
        model._ensure_handle(PB)
        _lib.check(lib.dib_profile_enable(model._handle, 1))
        nprof = min(args.steps, 5)
        for i in range(nprof):
            model._backward(xs_d[i], ys_d[i], global_batch=PB * world, sample_offset=rank * PB)
        cap = 4096
        ms = (ctypes.c_float * cap)()
        labels = ctypes.create_string_buffer(1 << 16)
        n = lib.dib_profile_read(model._handle, labels, len(labels), ms, cap)
        lib.dib_profile_enable(model._handle, 0)
        names = labels.value.decode().split("\n")[:n]
        groups = {}
        for nm, t in zip(names, list(ms)[:n]):
            groups.setdefault(nm, []).append(float(t))
        avg = {k: float(np.mean(v)) for k, v in groups.items()}
        macs, fwd_macs, train_macs = algorithmic_macs(PB)
        top = max((k for k in avg if k in macs), key=lambda k: avg[k])
        peaks = measured_peaks()
        ach = 2 * macs[top] / (avg[top] * 1e-3) / 1e12
        step_ach = 2 * train_macs / (ms_per_step * 1e-3) / 1e12
        roofline = {"bound": "tensor", "kernel": top, "achieved": ach, "peak": peaks["tflops_sustained"],
                    "unit": "TFLOP/s", "frac": ach / peaks["tflops_sustained"], "traffic": ncu_dram_traffic(top),
                    "traffic_unit": "bytes per launch (ncu dram read+write, profiles/r01_final_ncu_full.txt)",
                    "peak_source": f"{peaks['source']} bf16 dense sustained (MEASURED_PEAKS.json); math runs as {args.precision}",
                    "kernel_ms": avg[top], "kernel_share_of_step": avg[top] / sum(avg.values()),
                    "step_achieved_tflops": step_ach, "step_frac": step_ach / peaks["tflops_sustained"],
                    "algorithmic_gflop_per_step": 2 * train_macs / 1e9,
                    "hbm_algorithmic_gbs": (68 * PB + 7 * model.count_params() * 4) / (ms_per_step * 1e-3) / 1e9,
                    "hbm_peak_gbs": peaks["hbm_gbs"],
                    "group_ms": {k: round(v, 4) for k, v in sorted(avg.items(), key=lambda kv: -kv[1])}}
    barrier()

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rows = 16384
        rate, med, threads = cpu_twin_rate(rows, 3, 1)
        cpu_baseline = {"value": rate, "unit": "samples/s", "cores": threads, "kind": "port",
                        "sample": f"{rows} of {BATCH} rows per step, 1 warm-up + 3 timed steps of oracle/torch_twin.py "
                                  f"(PyTorch-CPU eager twin of models.py; TensorFlow is not installable here)"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": {"fp32": "f32", "tf32": "tf32", "bf16": "bf16"}[args.precision],
            "data": "synthetic", "config": workload_config(world, args.precision, PB), "clocks": clocks, "e2e": e2e,
            "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu_baseline,
            "library": os.path.relpath(_lib.library_path(), ROOT), "build": lib.dib_build_info().decode(),
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

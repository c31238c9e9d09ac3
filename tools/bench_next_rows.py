"""Measurement for the "next" rows of SURVEY.md section 8f (f1-f4): device time of each entry point on the C0 model
beside the CPU oracle (numpy float64 restatement, or the reference's own build for f4) on a bounded sample.
Prints one JSON object per row and writes them to gpurun_out/next_rows.json.

    python tools/bench_next_rows.py            # on the GPU box
    python tools/bench_next_rows.py --cpu-only # f4 only (no GPU needed)
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def gpu_ms(fn, iters=20, warm=3):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def cpu_s(fn, reps=1):
    t = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cpu-only", action="store_true")
    args = ap.parse_args()
    rows = []
    import dib_b200
    from oracle import ctw_oracle
    from oracle import dib_oracle as O

    if not args.cpu_only:
        import torch
        from dib_b200 import utils
        rng = np.random.default_rng(0)
        F, E = 16, 32
        cfg = O.DIBConfig([1] * F, [128, 128], [256, 256], 1)
        m = dib_b200.DistributedIBNet([1] * F, [128, 128], [256, 256], 1, seed=1, precision="tf32")
        m.compile(optimizer=dib_b200.Adam(3e-4), loss="bce_logits", metrics=["accuracy"])
        x = rng.standard_normal((65536, F)).astype(np.float32)
        xd = torch.from_numpy(x).cuda()
        p = m.get_flat_weights()

        # ---- f1: MI sandwich bounds, all 16 features x 8 batches x 1024 rows (InfoPerFeatureCallback defaults)
        def f1():
            for i in range(F):
                utils.estimate_mi_sandwich_bounds(m.feature_encoders[i], xd[:, i:i + 1], 1024, 8, seed=3)
        t_gpu = gpu_ms(f1, iters=5, warm=2)
        idx = [rng.integers(0, 65536, 1024)]
        eps = [rng.standard_normal((1024, E))]
        t_cpu = cpu_s(lambda: O.estimate_mi_sandwich_bounds(cfg, p, 0, x[:, 0:1], idx, eps))      # ONE feature, ONE batch
        rows.append(dict(row="f1", what="estimate_mi_sandwich_bounds, 16 features x 8 batches x 1024 rows, E=32",
                         gpu_ms=t_gpu, cpu_oracle_s_extrapolated=t_cpu * F * 8,
                         cpu_sample="numpy float64 oracle, 1 feature x 1 batch, x128", speedup=t_cpu * F * 8 * 1e3 / t_gpu))

        # ---- f2: compression matrices of all 16 features, 128 display rows each (SaveCompressionMatricesCallback)
        ridx = np.stack([rng.choice(65536, 128) for _ in range(F)])
        ridx_d = torch.from_numpy(ridx.astype(np.int32)).cuda()
        t_gpu = gpu_ms(lambda: m.compression_matrices(xd, ridx_d, want=("dist", "comp")))
        t_cpu = cpu_s(lambda: O.compression_matrices(cfg, p, x, ridx))
        rows.append(dict(row="f2", what="compression matrices, 16 features x 128 rows (encoders + Bhattacharyya + exp)",
                         gpu_ms=t_gpu, cpu_oracle_s=t_cpu, cpu_sample="numpy float64 oracle, closed form O(n^2 E)",
                         speedup=t_cpu * 1e3 / t_gpu))
        mu, lv = rng.standard_normal((1024, E)).astype(np.float32), rng.standard_normal((1024, E)).astype(np.float32) * 0.5
        mud, lvd = torch.from_numpy(mu).cuda(), torch.from_numpy(lv).cuda()
        t_gpu = gpu_ms(lambda: utils.kl_divergence_mat(mud, lvd))
        t_cpu = cpu_s(lambda: O.kl_divergence_mat(mu, lv, mu, lv))
        rows.append(dict(row="f2", what="kl_divergence_mat 1024 x 1024, E=32", gpu_ms=t_gpu, cpu_oracle_s=t_cpu,
                         speedup=t_cpu * 1e3 / t_gpu))

        # ---- f3: InfoNCE head (loss + both gradients), n = 2048, d = 64; and an external-loss train step at B = 65536
        n, d = 2048, 64
        a, b = rng.standard_normal((n, d)).astype(np.float32), rng.standard_normal((n, d)).astype(np.float32)
        ad, bd = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
        for kind in ("l2sq", "cosine", "linf"):
            t_gpu = gpu_ms(lambda: utils.infonce_loss_and_grads(ad, bd, kind, 8.0))
            t_cpu = cpu_s(lambda: O.infonce_loss_and_grads(a[:512], b[:512], kind, 8.0))
            rows.append(dict(row="f3", what=f"infonce head '{kind}' n={n} d={d} (loss + d_e1 + d_e2)", gpu_ms=t_gpu,
                             cpu_oracle_s_extrapolated=t_cpu * (n / 512) ** 2, cpu_sample="numpy float64 oracle at n=512, x16",
                             speedup=t_cpu * (n / 512) ** 2 * 1e3 / t_gpu))
        me = dib_b200.DistributedIBNet([1] * F, [128, 128], [256, 256], 16, seed=1, precision="tf32")
        me.compile(optimizer=dib_b200.Adam(3e-4), loss="external")
        dpred = torch.from_numpy((rng.standard_normal((65536, 16)) / 65536).astype(np.float32)).cuda()
        me.compute_gradients(xd, dpred)
        t_ext = gpu_ms(lambda: me.compute_gradients(xd, dpred), iters=20)
        y = torch.from_numpy((x[:, :1] > 0).astype(np.float32)).cuda()
        m.compute_gradients(xd, y)
        t_cmp = gpu_ms(lambda: m.compute_gradients(xd, y), iters=20)
        rows.append(dict(row="f3", what="train step (fwd+bwd) B=65536 with caller-owned loss, out=16 (TF32 integration kernels)",
                         gpu_ms=t_ext, compiled_loss_fast_path_ms=t_cmp))

    # ---- f4: CTW (host).  nb-chaos cell 3 evaluates 75 sub-sequences per partition
    rng = np.random.default_rng(1)
    x_, seqs = 0.3, []
    for _ in range(75):
        s = []
        for _ in range(20000):
            x_ = 3.9 * x_ * (1 - x_)
            s.append(int(x_ > 0.5) + 2 * int(rng.random() < 0.5))
        seqs.append(np.array(s, dtype=np.int8))
    from dib_b200 import ctw
    t_one = cpu_s(lambda: ctw.estimate_entropy(seqs[0], 4), reps=3)
    t_batch = cpu_s(lambda: ctw.estimate_entropy_batch(seqs, 4), reps=2)
    rec = dict(row="f4", what="CTW entropy rate, 75 sequences x 20000 symbols, A=4 (host)", one_sequence_s=t_one,
               batch_75_s=t_batch, threads=os.cpu_count())
    if ctw_oracle.reference_available():
        t_ref = cpu_s(lambda: ctw_oracle.reference_estimate_entropy(seqs[0], 4), reps=3)
        same = ctw.estimate_entropy(seqs[0], 4) == ctw_oracle.reference_estimate_entropy(seqs[0], 4)
        rec.update(reference_one_sequence_s=t_ref, reference_75_sequential_s_extrapolated=75 * t_ref, bit_identical=bool(same),
                   speedup_one=t_ref / t_one, speedup_batch=75 * t_ref / t_batch, cpu_kind="reference (oracle/_ref/libctw_ref.so)")
    rows.append(rec)

    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "next_rows.json"), "w") as fh:
        json.dump(rows, fh, indent=1)
    for r in rows:
        print(json.dumps(r))


if __name__ == "__main__":
    main()

"""A/B of the fused encoder backward kernels on a B200: gradient parity of each variant against the fp32 path, and the
CUDA-event time of the backward launch group.  Usage: python tools/bwd_ab.py <variant> [quick]   (run each variant in its
own process under `timeout`, a hang then costs one variant, not the whole call)."""
import ctypes, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dib_b200
from dib_b200 import _lib
from oracle import dib_oracle as O

variant = int(sys.argv[1])
rb = int(sys.argv[2]) if len(sys.argv) > 2 else 1
lib = _lib.load()
_lib.check(lib.dib_debug_set_variant(0, variant))
_lib.check(lib.dib_debug_set_variant(1, rb))
for _k, _a in ((3, 3), (4, 4), (5, 5)):  # optional: fused tail on/off, CTA-pair GEMMs on/off, measurement-only epilogue switch
    if len(sys.argv) > _a:
        _lib.check(lib.dib_debug_set_variant(_k, int(sys.argv[_a])))
out = {"variant": variant, "int16_resident_b": rb}


def model(cfgargs, prec, loss="bce_logits"):
    m = dib_b200.DistributedIBNet(*cfgargs[0], precision=prec, seed=1, **cfgargs[1])
    m.compile(optimizer=dib_b200.Adam(3e-4), loss=loss, metrics=["accuracy"])
    return m


def rel(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())


cases = {
    "c0": ((([1] * 16, [128, 128], [256, 256], 1), {}), "bce_logits", 16, 1),
    "hetero_tanh": ((([2, 1, 2, 1], [128, 128], [256, 256], 6), {"activation_fn": "tanh"}), "mse", 6, 6),
}
rng = np.random.default_rng(0)
for name, (ca, loss, D, outd) in cases.items():
    for B in ((128 * 3 + 17, 4096 + 64, 65536) if name == "c0" else (128 * 3 + 17, 4096 + 64)):
        x = rng.standard_normal((B, D)).astype(np.float32)
        y = (x[:, :1] * x[:, 1:2] > 0).astype(np.float32) if outd == 1 else rng.standard_normal((B, outd)).astype(np.float32)
        ref = model(ca, "fp32", loss)
        p = ref.get_flat_weights()
        p = p + (p == 0) * (0.05 * rng.standard_normal(p.size)).astype(np.float32)
        ref.set_flat_weights(p); ref.beta.assign(0.02)
        g32, s32 = ref.compute_gradients(x, y, step=5)
        m = model(ca, "fp16", loss)
        m.set_flat_weights(p); m.beta.assign(0.02)
        g16, s16 = m.compute_gradients(x, y, step=5)
        g16b, _ = m.compute_gradients(x, y, step=5)
        out[f"{name}_B{B}"] = {"grad_rel": rel(g16.cpu().numpy(), g32.cpu().numpy()), "stats_rel": rel(s16.cpu().numpy(), s32.cpu().numpy()),
                               "deterministic": bool(torch.equal(g16, g16b)), "finite": bool(torch.isfinite(g16).all())}
        del ref, m
# timing at C0 / 65536
m = model(cases["c0"][0], "fp16")
m.beta.assign(1e-3)
B = 65536
xs = [torch.randn(B, 16, device="cuda") for _ in range(8)]
ys = [(x[:, :1] * x[:, 1:2] > 0).float() for x in xs]
for i in range(5):
    m._backward(xs[i % 8], ys[i % 8], global_batch=B)
torch.cuda.synchronize()
_lib.check(lib.dib_profile_enable(m._handle, 1))
for i in range(10):
    m._backward(xs[i % 8], ys[i % 8], global_batch=B)
cap = 4096
ms = (ctypes.c_float * cap)(); labels = ctypes.create_string_buffer(1 << 16)
n = lib.dib_profile_read(m._handle, labels, len(labels), ms, cap)
lib.dib_profile_enable(m._handle, 0)
g = {}
for nm, t in zip(labels.value.decode().split("\n")[:n], list(ms)[:n]):
    g.setdefault(nm, []).append(float(t))
out["group_ms"] = {k: round(float(np.mean(v)), 4) for k, v in sorted(g.items(), key=lambda kv: -np.mean(kv[1]))}
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for i in range(50):
    m._backward(xs[i % 8], ys[i % 8], global_batch=B)
ev1.record(); torch.cuda.synchronize()
out["backward_ms_per_step"] = ev0.elapsed_time(ev1) / 50
print(json.dumps(out))

"""Per-quantity comparison of the fused tensor-core path against the unfused TF32 path and the fp32 path."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import dib_oracle as O
from tests.test_gpu_parity import build_model, rel_err

cfg = O.DIBConfig([1] * 16, [128, 128], [256, 256], 1)
rng = np.random.default_rng(3)
p = O.glorot_uniform_params(cfg, rng)
p = p + (p == 0) * (0.05 * rng.standard_normal(p.size)).astype(np.float32)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
x = rng.standard_normal((B, 16)).astype(np.float32)
y = (x[:, 0] * x[:, 1] > 0).astype(np.float32)[:, None]
res = {}
for tag, prec, unfused in (("fp32", "fp32", False), ("unfused", "tf32", True), ("fused", "tf32", False)):
    m = build_model(cfg, precision=prec)
    m.debug_force_unfused(unfused)
    m.set_flat_weights(p)
    m.beta.assign(0.02)
    with torch.cuda.device(m.device):
        xd = m._to_device(x, 16)
        pred, emb, st0 = m._forward(xd, None, None, 5, 0, want_emb=True)
    g, st = m.compute_gradients(x, y, step=5)
    res[tag] = dict(pred=pred.cpu().numpy(), emb=emb.cpu().numpy(), g=g.cpu().numpy(), st=st.cpu().numpy())
names = []
for f in range(16):
    names += [f"f{f}.W0", f"f{f}.b0", f"f{f}.W1", f"f{f}.b1", f"f{f}.W2", f"f{f}.b2"]
names += ["I.W0", "I.b0", "I.W1", "I.b1", "I.W2", "I.b2"]
for tag in ("unfused", "fused"):
    a, r = res[tag], res["fp32"]
    print(f"== {tag} vs fp32 (B={B}): pred {rel_err(a['pred'], r['pred']):.2e} emb {rel_err(a['emb'], r['emb']):.2e} "
          f"g {rel_err(a['g'], r['g']):.2e}")
    print("   stats rel:", np.abs(a['st'] / r['st'] - 1).max(), " KL0", a['st'][0], r['st'][0], "loss", a['st'][16], r['st'][16])
    off = 0
    worst = []
    for nm, s in zip(names, cfg.param_shapes()):
        n = int(np.prod(s))
        worst.append((rel_err(a['g'][off:off + n], r['g'][off:off + n]), nm, float(np.abs(r['g'][off:off+n]).max())))
        off += n
    for e, nm, mag in worst[:6] + worst[-6:]:
        print(f"   {nm:8s} rel {e:.2e}  |ref|max {mag:.3e}")
    print("   worst:", sorted(worst, reverse=True)[:5])
    d = a['emb'] - r['emb']
    print("   emb err by feature:", [f"{np.abs(d[:, f*32:(f+1)*32]).max():.1e}" for f in range(0, 16, 3)],
          " by row block:", [f"{np.abs(d[i:i+128]).max():.1e}" for i in range(0, min(B, 1024), 128)])

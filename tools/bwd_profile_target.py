"""A few C0 backward launches for `ncu -k regex:enc_fused_bwd` (kernel-level profile of the fused encoder backward)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dib_b200
m = dib_b200.DistributedIBNet([1] * 16, [128, 128], [256, 256], 1, precision=os.environ.get("DIB_PRECISION", "fp16"), seed=1)
m.compile(optimizer=dib_b200.Adam(3e-4), loss=dib_b200.losses.BinaryCrossentropy(from_logits=True), metrics=["accuracy"])
m.beta.assign(1e-3)
B = 65536
x = torch.randn(B, 16, device="cuda")
y = (x[:, :1] * x[:, 1:2] > 0).float()
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    m._backward(x, y, global_batch=B)
torch.cuda.synchronize()
print("done")

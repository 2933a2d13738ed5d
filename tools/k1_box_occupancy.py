"""Experiment: K1 time with BOX=22 (4 CTAs/SM) vs BOX=24 (3 CTAs/SM) on transforms that fit both."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from torchio_b200 import ops
B, S = 32, 256
rng = np.random.default_rng(0)
x = torch.rand((B, 1, S, S, S), device="cuda")
mats = []
for b in range(B):
    ang = np.deg2rad(rng.uniform(-6, 6, 3)); sc = rng.uniform(0.95, 1.05)
    cx, sx, cy, sy, cz, sz = np.cos(ang[0]), np.sin(ang[0]), np.cos(ang[1]), np.sin(ang[1]), np.cos(ang[2]), np.sin(ang[2])
    r = (np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]]) @ np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
         @ np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])) * sc
    c = np.full(3, (S - 1) / 2); m = np.eye(4); m[:3, :3] = r; m[:3, 3] = c - r @ c
    mats.append(m.astype(np.float32)[:3].reshape(12))
mat = torch.tensor(np.stack(mats)).cuda()
ext = max((np.abs(m.reshape(3, 4)[:, :3]).sum(axis=1) * 15 + 2).max() for m in mats)
print("largest extent", ext)
cp = torch.tensor(rng.uniform(-3, 3, (B, 7, 7, 7, 3)).astype(np.float32)).cuda()
flags = torch.full((B,), 2, dtype=torch.uint8).cuda()
fill = torch.tensor([0.0]).cuda() + 0.5
one = (1.0, 1.0, 1.0)
for label, c, f in (("affine", None, None), ("elastic+affine", cp, flags)):
    outs = {}
    for hint in (24, 22):
        for _ in range(2):
            y = ops.resample(x, mat, c, f, one, one, affine_first=True, mode=ops.LINEAR, fill=fill, box_hint=hint)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            y = ops.resample(x, mat, c, f, one, one, affine_first=True, mode=ops.LINEAR, fill=fill, box_hint=hint)
        e1.record(); torch.cuda.synchronize()
        outs[hint] = y
        print(label, "box", hint, f"{e0.elapsed_time(e1)/5:.3f} ms")
    print(label, "max diff", float((outs[22] - outs[24]).abs().max()))

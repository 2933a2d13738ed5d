"""Locate label-map mismatches at 256^3: tile path vs general path vs C oracle, per transform."""
import copy, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from golden_cases import CASES_BY_NAME, build_inputs
from util import GOLDEN, product_batch, blank_transform
from oracle import c_port
from torchio_b200 import ops

name = "full256_config2_b2"
z = np.load(GOLDEN / f"{name}.npz")
history = json.loads(bytes(z["history"]).decode())
inputs = build_inputs(CASES_BY_NAME[name])
images = {}
for n in inputs["subjects"][0]:
    images[n] = {"kind": inputs["subjects"][0][n][0], "data": torch.stack([s[n][1] for s in inputs["subjects"]]),
                 "affines": [np.array(s[n][2], dtype=np.float64) for s in inputs["subjects"]]}
pass
orc = copy.deepcopy(images)
batch = product_batch(images, device="cuda")
orig = ops.resample
mode = {"hint": None}
def patched(*a, **k):
    if mode["hint"] is not None:
        k["box_hint"] = mode["hint"]
    return orig(*a, **k)
ops.resample = patched
import torchio_b200.transforms.spatial as sp
for step in history:
    prev = {n: ib.data.clone() for n, ib in batch.images.items()}
    prev_aff = {n: list(ib.affines) for n, ib in batch.images.items()}
    c_port.replay(orc, [step])
    outs = {}
    for hint in (None, -1):
        mode["hint"] = hint
        for n, ib in batch.images.items():
            ib.data = prev[n].clone(); ib.affines[:] = prev_aff[n]
        blank_transform(step["name"]).apply_transform(batch, step["params"])
        outs[hint] = batch.images["seg"].data.cpu()
    want = orc["seg"]["data"]
    for hint, got in outs.items():
        bad = (got != want)
        print(step["name"], "hint", hint, "mismatch vs oracle", int(bad.sum()))
        if bad.any():
            idx = bad.nonzero()[:5]
            print(idx.tolist(), got[bad][:5].tolist(), want[bad][:5].tolist())
    print(step["name"], "tile vs general", int((outs[None] != outs[-1]).sum()))
    # continue from the oracle's state so that steps are judged independently
    batch.images["seg"].data = want.cuda()

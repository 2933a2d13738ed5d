"""The boundary is executable: the header, the product's own ctypes binding and the stub printed in
INTEGRATION.md (what a TorchIO maintainer would paste) must agree argument for argument, and the
stub — run verbatim — must reproduce `ops.resample` on a golden case."""

import ctypes
import re
from pathlib import Path

import numpy as np
import pytest
import torch

from torchio_b200 import _native

ROOT = Path(__file__).resolve().parent.parent
C2CTYPES = {"float": ctypes.c_float, "int": ctypes.c_int, "int64_t": ctypes.c_int64, "uint64_t": ctypes.c_uint64, "size_t": ctypes.c_size_t}


def header_prototypes():
    """name -> list of ctypes argument types, parsed from include/tio_b200.h."""
    text = re.sub(r"/\*.*?\*/", "", (ROOT / "include" / "tio_b200.h").read_text(), flags=re.S)
    out = {}
    for m in re.finditer(r"(?:const\s+char\s*\*|size_t|int)\s+(tio_\w+)\s*\(([^)]*)\)\s*;", text):
        name, args = m.group(1), m.group(2).strip()
        types = []
        if args not in ("", "void"):
            for a in args.split(","):
                ctype = re.sub(r"\s+\w+$", "", a.strip())
                types.append(ctypes.c_void_p if "*" in ctype else C2CTYPES[ctype])
        out[name] = types
    return out


def integration_blocks():
    text = (ROOT / "INTEGRATION.md").read_text()
    section = text[text.index("## B."):]
    return re.findall(r"```python\n(.*?)```", section, flags=re.S)


class _Recorder:
    """Stands in for the CDLL while the INTEGRATION.md argtypes lines are executed."""

    def __init__(self):
        object.__setattr__(self, "fns", {})

    def __getattr__(self, name):
        return self.fns.setdefault(name, type("F", (), {})())


def test_product_binding_matches_the_header_argument_for_argument():
    protos = header_prototypes()
    assert set(protos) - {"tio_last_error"} == set(_native._SIGNATURES)
    for name, argtypes in _native._SIGNATURES.items():
        assert argtypes == protos[name], name


def test_integration_md_argtypes_match_the_header():
    protos = header_prototypes()
    first, second = integration_blocks()[:2]
    rec = _Recorder()
    env = {"ctypes": ctypes, "_lib": rec, "P": ctypes.c_void_p, "I32": ctypes.c_int, "I64": ctypes.c_int64,
           "U64": ctypes.c_uint64, "SZ": ctypes.c_size_t}
    # the argtypes assignments of the stub (they may wrap over two lines) + the table below it
    src = "\n".join(l for l in first.splitlines() if re.match(r"(_lib\.\w+\.(argtypes|restype)\s*=|\s+ctypes\.c_size_t, P\])", l))
    exec(src, env)
    exec(second, env)
    seen = {name: fn.argtypes for name, fn in rec.fns.items() if hasattr(fn, "argtypes")}
    assert len(seen) >= 12
    for name, argtypes in seen.items():
        assert list(argtypes) == protos[name], name
    assert rec.fns["tio_resample_workspace_bytes"].restype is ctypes.c_size_t


@pytest.mark.gpu
def test_integration_md_stub_runs_and_reproduces_ops_resample():
    """Execute the §B stub verbatim (only the library path is pointed at the in-tree build) and
    call it the way the patched reference would."""
    from golden_cases import CASES_BY_NAME
    from oracle import c_port
    from torchio_b200 import ops
    from util import load_golden

    stub = integration_blocks()[0].replace('ctypes.CDLL("libtio_b200.so")', f'ctypes.CDLL("{_native.LIB_PATH}")')
    env: dict = {}
    exec(stub, env)
    name = "config1_affine_deg10_64"
    _, images, history, expected, _ = load_golden(name)
    params = history[0]["params"]
    data = images["t1"]["data"].cuda()
    mat, cp, flags, _ = c_port.spatial_tables(params, data.shape[0], tuple(data.shape[2:]), images["t1"]["affines"][0])
    mat = mat.cuda()
    fill = ops.min_sample0(data)
    got = env["resample_f32"](data, mat, None, None, (1.0, 1.0, 1.0), (1.0, 1.0, 1.0), True, True, fill)
    want = ops.resample(data, mat, None, None, (1.0, 1.0, 1.0), (1.0, 1.0, 1.0), affine_first=True,
                        mode=ops.LINEAR, fill=fill)
    assert torch.equal(got, want)
    ref = expected["t1"]
    rng = float(ref.max() - ref.min())
    assert float((got.cpu() - ref).abs().max()) <= 1e-4 * rng
    assert CASES_BY_NAME[name]["batch"] == data.shape[0]

"""CPU: the C-ABI library loads, exports every symbol include/grl_b200.h declares, and the ctypes mirror of the
descriptor structs has the C compiler's sizes / field offsets.  No compute calls (no GPU here)."""
import ctypes as C
import os
import re
import subprocess
import sys
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "grl_b200.h")


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(grl_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    from geometry_rl_b200.csrc import build as b
    b.build()
    from geometry_rl_b200 import _lib
    return _lib


def test_every_declared_symbol_is_exported_and_bound(lib):
    handle = lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(handle, name), f"{name} declared in grl_b200.h but not exported"
        assert name in lib.SIGNATURES, f"{name} has no ctypes signature in _lib.py"
    assert handle.grl_abi_version() == 1
    assert handle.grl_last_error() is not None


def test_struct_layouts_match_the_c_compiler(lib):
    structs = {"GrlEmbedDesc": lib.GrlEmbedDesc, "GrlBasisDesc": lib.GrlBasisDesc, "GrlConvDesc": lib.GrlConvDesc,
               "GrlProjDesc": lib.GrlProjDesc, "GrlLossDesc": lib.GrlLossDesc, "GrlReadoutDesc": lib.GrlReadoutDesc,
               "GrlFusedEdgeDesc": lib.GrlFusedEdgeDesc, "GrlCriticDesc": lib.GrlCriticDesc,
               "GrlEncoderDesc": lib.GrlEncoderDesc}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', "int main(void) {"]
    for name, cls in structs.items():
        lines.append(f'printf("{name} %zu\\n", sizeof({name}));')
        for field, _ in cls._fields_:
            lines.append(f'printf("{name}.{field} %zu\\n", offsetof({name}, {field}));')
    lines += ["return 0; }"]
    with tempfile.TemporaryDirectory() as d:
        src, exe = os.path.join(d, "t.c"), os.path.join(d, "t")
        open(src, "w").write("\n".join(lines))
        subprocess.run(["gcc", src, "-o", exe], check=True)
        out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    c_layout = dict(l.split() for l in out.strip().splitlines())
    for name, cls in structs.items():
        assert int(c_layout[name]) == C.sizeof(cls), name
        for field, _ in cls._fields_:
            assert int(c_layout[f"{name}.{field}"]) == getattr(cls, field).offset, f"{name}.{field}"


def test_product_has_no_cpu_path():
    import torch
    from geometry_rl_b200 import ops
    with pytest.raises(RuntimeError):
        ops.gae(torch.zeros(2, 3), torch.zeros(2, 4), torch.zeros(2, 3, dtype=torch.bool),
                torch.zeros(2, 3, dtype=torch.bool), 0.99, 0.95)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "geometry_rl_b200")
    offenders = []
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(dirpath, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M):
                    offenders.append(os.path.join(dirpath, f))
    assert not offenders, offenders


def test_loss_scalar_indices_match_the_header_enum(lib):
    """_lib.LOSS_SCALAR_INDEX (used by TRPLLoss to pick scalars out of grl_trpl_loss_fwd's output) mirrors the
    GRL_LS_* enum of include/grl_b200.h, and the size macros agree."""
    src = open(HEADER).read()
    body = re.search(r"enum\s*\{(.*?)\};", src, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = [n.strip().split("=")[0].strip() for n in body.split(",") if n.strip()]
    assert names[0] == "GRL_LS_LOSS_OBJECTIVE"
    header_index = {n[len("GRL_LS_"):].lower(): i for i, n in enumerate(names)}
    mine = {k.lower(): v for k, v in lib.LOSS_SCALAR_INDEX.items()}
    assert mine == header_index

    def macro(name):
        return int(re.search(rf"#define\s+{name}\s+(\d+)", src).group(1))

    assert macro("GRL_LOSS_TERMS") == lib.LOSS_TERMS and macro("GRL_LOSS_SCALARS") == lib.LOSS_SCALARS
    assert macro("GRL_LOSS_STATS") == lib.LOSS_STATS and macro("GRL_LOSS_SUMS") == lib.LOSS_SUMS
    assert macro("GRL_READOUT_MAX_OUT") == lib.READOUT_MAX_OUT
    assert len(names) <= lib.LOSS_SCALARS


def test_sm_reservation_policy_is_reflected_in_the_reported_sm_count():
    """grl_reserve_sms(n): persistent grids are sized for (SMs - n) (DataParallel(side_group=True) leaves 4 to the critic
    stream's collectives).  No GPU needed: without a device the library reports 148 SMs."""
    from geometry_rl_b200 import _lib as L
    L.reserve_sms(0)
    full = L.sm_count()
    try:
        L.reserve_sms(4)
        assert L.sm_count() == full - 4
        with pytest.raises(RuntimeError):
            L.reserve_sms(-1)
    finally:
        L.reserve_sms(0)
    assert L.sm_count() == full


def test_encoder_kernel_gate_declines_what_the_kernel_does_not_cover():
    """ops.encoder_layer_supported: the K5 kernels take CUDA tensors of the shipped transformer shape only; anything else
    (CPU tensors, other widths / head counts, active dropout, too many tokens) stays on nn.TransformerEncoder."""
    import torch
    from geometry_rl_b200 import ops
    ok = torch.nn.TransformerEncoderLayer(d_model=64, nhead=2, dim_feedforward=64, dropout=0.0)
    x = torch.zeros(2, 50, 64)
    assert not ops.encoder_layer_supported(x, ok)  # CPU tensor
    meta = torch.zeros(2, 50, 64, device="meta")
    assert not ops.encoder_layer_supported(meta, ok)
    # the shape / option checks, evaluated on a stand-in that claims to live on a CUDA device
    class _Cuda:
        is_cuda = True
        def __init__(self, shape):
            self.shape = shape
        def dim(self):
            return len(self.shape)
    assert ops.encoder_layer_supported(_Cuda((2, 50, 64)), ok)
    assert not ops.encoder_layer_supported(_Cuda((2, ops.ENCODER_MAX_TOKENS + 1, 64)), ok)
    assert not ops.encoder_layer_supported(_Cuda((2, 50, 32)), ok)
    drop = torch.nn.TransformerEncoderLayer(d_model=64, nhead=2, dim_feedforward=64, dropout=0.1)
    assert not ops.encoder_layer_supported(_Cuda((2, 50, 64)), drop)      # training mode, p > 0
    assert ops.encoder_layer_supported(_Cuda((2, 50, 64)), drop.eval())    # inactive dropout is fine
    for bad in (torch.nn.TransformerEncoderLayer(64, 4, 64, 0.0), torch.nn.TransformerEncoderLayer(64, 2, 128, 0.0),
                torch.nn.TransformerEncoderLayer(64, 2, 64, 0.0, activation="gelu"),
                torch.nn.TransformerEncoderLayer(64, 2, 64, 0.0, norm_first=True)):
        assert not ops.encoder_layer_supported(_Cuda((2, 50, 64)), bad)

"""CPU-side checks: the C-ABI library loads and exports every symbol include/picklebot_b200.h declares, the
drop-in modules keep the reference's state_dict layout, and the product path refuses to run without CUDA."""
import json
import os
import re
import subprocess

import pytest
import torch

from _util import GOLDEN, MODEL_NAMES, ROOT, synthetic_checkpoint


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "picklebot_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from picklebot_b200 import _lib
    lib = _lib.lib()
    declared = _header_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert sorted(_lib.EXPORTS) == declared, set(_lib.EXPORTS) ^ set(declared)
    assert lib.pb_abi_version() == 1
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (pb_[a-z0-9_]+)", out))
    assert exported == set(declared), exported ^ set(declared)


def test_signature_table_matches_header_arity():
    from picklebot_b200 import _lib
    text = open(os.path.join(ROOT, "include", "picklebot_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    for name, sig in _lib.SIGNATURES.items():
        m = re.search(r"\b%s\s*\((.*?)\)\s*;" % name, text, flags=re.S)
        assert m, name
        args = [a.strip() for a in m.group(1).split(",")]
        assert len(args) == len(sig), (name, len(args), len(sig))
        for a, c in zip(args, sig):
            if "*" in a or "pb_stream_t" in a:
                assert c == "p", (name, a, c)
            elif a.startswith("long long"):
                assert c == "l", (name, a, c)
            elif a.startswith("float"):
                assert c == "f", (name, a, c)
            else:
                assert c == "i", (name, a, c)


@pytest.mark.parametrize("model", MODEL_NAMES)
def test_state_dict_layout_is_the_references(model):
    import picklebot_b200 as pb
    keys = json.load(open(os.path.join(GOLDEN, "statedict_keys.json")))[model]
    nc = 13 if model == "MoViNetA2" else 2
    m = pb.valid_models[model](num_classes=nc)
    sd = m.state_dict()
    assert {k: list(v.shape) for k, v in sd.items()} == keys
    for k, v in sd.items():
        assert v.dtype == (torch.int64 if k.endswith("num_batches_tracked") else torch.float32)
    m.load_state_dict(synthetic_checkpoint(model))          # strict
    assert sum(p.numel() for p in m.parameters()) == {"MobileNetLarge3D": 4191584, "MobileNetSmall3D": 1672816,
                                                      "MoViNetA2": 3992289}[model]


def test_constructor_signatures():
    import inspect
    import picklebot_b200 as pb
    assert list(inspect.signature(pb.Bottleneck3D.__init__).parameters)[1:] == [
        "in_channels", "out_channels", "expanded_channels", "stride", "use_se", "kernel_size", "nonlinearity",
        "batchnorm", "dropout", "bias"]
    assert list(inspect.signature(pb.MoviNetBottleneck.__init__).parameters)[1:] == [
        "in_channels", "out_channels", "expanded_channels", "kernel_size", "stride", "use_se", "batchnorm",
        "nonlinearity", "bias", "dropout", "padding", "dilation"]
    assert list(inspect.signature(pb.MoViNetA2.__init__).parameters)[1:] == ["num_classes", "buffer_size"]
    assert list(inspect.signature(pb.CausalConv3d.__init__).parameters)[1:] == [
        "in_channels", "out_channels", "kernel_size", "stride", "dilation", "stream_buffer", "kwargs"]
    assert pb.MobileNetLarge3D().num_classes == 2


def test_no_cpu_fallback():
    import picklebot_b200 as pb
    m = pb.MobileNetSmall3D()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 3, 4, 16, 16))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pb.Bottleneck3D(16, 16, 16)(torch.zeros(1, 16, 2, 4, 4))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "picklebot_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src.replace("oracle's", "").replace("the oracle", "") or "import" not in [
                ln for ln in src.splitlines() if "oracle" in ln and "import" in ln] or True
            for ln in src.splitlines():
                assert not re.match(r"\s*(from|import)\s+oracle", ln), (fn, ln)


def test_checkpoint_prefix_conversion_round_trip(tmp_path):
    """Reference checkpoints saved from torch.compile / DDP wrappers (train.py:38-44,316-318,338) load strictly."""
    import torch
    import picklebot_b200 as pb
    from picklebot_b200.checkpoint import load_reference_checkpoint, state_dict_converter
    m = pb.MobileNetSmall3D(num_classes=2)
    sd = {("module._orig_mod." if i % 2 else "_orig_mod.") + k: v.clone() for i, (k, v) in enumerate(m.state_dict().items())}
    path = str(tmp_path / "ckpt.pth")
    torch.save(sd, path)
    m2 = load_reference_checkpoint(pb.MobileNetSmall3D(num_classes=2), path)
    for (k1, v1), (k2, v2) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert k1 == k2 and torch.equal(v1, v2)
    assert list(state_dict_converter({"a.b": 1, "_orig_mod.c": 2})) == ["a.b", "c"]

"""The drop-in boundary of SURVEY 8(b1): the reference's callers import the hot path with cwd = image/ as
``from models.sit import SiT_models`` / ``from loss import SILoss`` / ``from samplers import ...``.  These tests execute those
LITERAL lines (train.py:21-22,25; generate.py:9,17-18) in a fresh interpreter whose working directory is this repo's
``image/`` and check they resolve to the B200-native classes."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
IMAGE = os.path.join(ROOT, "image")

REFERENCE_IMPORT_LINES = """
from models.sit import SiT_models
from loss import SILoss
from dataset import CustomDataset
from samplers import euler_sampler, euler_maruyama_sampler
from utils import load_legacy_checkpoints
"""

CHECK = """
import inspect, sys
import reed_b200.image.models.sit as impl
assert SiT_models is impl.SiT_models and sorted(SiT_models) == sorted(f"SiT-{f}/{p}" for f in ("S", "B", "L", "XL") for p in (2, 4, 8))
assert SILoss.__module__ == "reed_b200.image.loss"
assert euler_sampler.__module__ == euler_maruyama_sampler.__module__ == "reed_b200.image.samplers"
assert CustomDataset.__module__ == "reed_b200.image.dataset"
assert list(inspect.signature(euler_sampler).parameters) == ["model", "latents", "y", "num_steps", "heun", "cfg_scale",
                                                             "guidance_low", "guidance_high", "path_type"]
m = SiT_models["SiT-S/2"](input_size=16, decoder_hidden_size=384, num_classes=10, z_dims=[32], encoder_depth=2, qk_norm=False,
                          fused_attn=True, use_cfg=True)
assert type(m).__name__ == "SiT" and m.in_channels == 4 and len(m.projectors) == 1
from models.sit import SiT_XL_2, SiT_S_8
assert SiT_XL_2 is SiT_models["SiT-XL/2"] and SiT_S_8 is SiT_models["SiT-S/8"]
assert load_legacy_checkpoints({"decoder_blocks.1.x": 1, "blocks.0.y": 2}, 8) == {"blocks.9.x": 1, "blocks.0.y": 2}
print("DROPIN-OK")
"""


def test_reference_import_lines_resolve_with_cwd_image():
    env = dict(os.environ)
    env.pop("PYTHONPATH", None)                     # nothing but the shim's own bootstrap may put the repo on the path
    res = subprocess.run([sys.executable, "-c", REFERENCE_IMPORT_LINES + CHECK], cwd=IMAGE, env=env, capture_output=True,
                         text=True, timeout=300)
    assert res.returncode == 0 and "DROPIN-OK" in res.stdout, res.stderr[-2000:]


def test_import_lines_are_the_references():
    """The lines above are the reference's own (skipped where /root/reference is absent, i.e. on the GPU box)."""
    ref = "/root/reference/image"
    if not os.path.isdir(ref):
        import pytest
        pytest.skip("reference tree not present")
    train, gen = open(os.path.join(ref, "train.py")).read(), open(os.path.join(ref, "generate.py")).read()
    for line in ("from models.sit import SiT_models", "from loss import SILoss", "from dataset import CustomDataset"):
        assert line in train
    for line in ("from models.sit import SiT_models", "from samplers import euler_sampler, euler_maruyama_sampler"):
        assert line in gen
    assert "load_legacy_checkpoints" in gen

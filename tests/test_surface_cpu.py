"""CPU-side checks of the drop-in surface: parameter tree identical to the reference's (via the golden fixtures),
constructor validation, and that the C-ABI library exports every symbol include/sdm_b200.h declares."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT, load_golden


@pytest.mark.parametrize("name", ["tiny", "tiny_cond", "gpu_small", "gpu_cond", "default64"])
def test_state_dict_matches_reference(name):
    from models.U_Net import U_Net
    fx = load_golden(f"unet_{name}.pt")
    net = U_Net(**fx["kwargs"])
    own = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    assert own == {k: tuple(v) for k, v in fx["shapes"].items()}
    assert list(net.state_dict().keys()) == list(fx["shapes"].keys())      # same registration order too


def test_ctor_validation_matches_reference():
    from models.U_Net import U_Net
    with pytest.raises(TypeError):
        U_Net(num_layers=2.0)
    with pytest.raises(TypeError):
        U_Net(attn_layers=(1, 2))
    with pytest.raises(ValueError):
        U_Net(num_layers=0)
    with pytest.raises(ValueError):
        U_Net(num_layers=2, attn_layers=[2])
    with pytest.raises(ValueError):
        U_Net(num_layers=2, attn_layers=[0.5])


def test_no_cpu_fallback():
    from b200 import B200Error
    from models.U_Net import U_Net
    net = U_Net(num_resnet_blocks=1, num_layers=1, attn_layers=[], min_channel=128, max_channel=128)
    with pytest.raises(B200Error):
        with torch.no_grad():
            net(torch.zeros(1, 3, 8, 8), torch.tensor([5]))


def test_custom_load_state_dict_skips_mismatches(capsys):
    from models.U_Net import U_Net
    net = U_Net(num_resnet_blocks=1, num_layers=1, attn_layers=[], min_channel=128, max_channel=128)
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    key = "in_layer.0.conv_layer.0.bias"
    sd[key] = torch.full_like(sd[key], 0.25)
    sd["bogus.weight"] = torch.zeros(3)
    sd["in_layer.1.conv_layer.0.bias"] = torch.zeros(7)
    net.custom_load_state_dict(sd)
    out = capsys.readouterr().out
    assert "No Layer found: bogus.weight" in out and "Skipped: in_layer.1.conv_layer.0.bias" in out
    assert torch.all(net.state_dict()[key] == 0.25)


def test_c_abi_exports_every_declared_symbol():
    from b200._lib import LIB_PATH, SIGNATURES
    header = open(os.path.join(ROOT, "include", "sdm_b200.h")).read()
    declared = set(re.findall(r"\b(b2_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert os.path.exists(LIB_PATH), "build the library first: make"
    handle = ctypes.CDLL(LIB_PATH)
    for sym in sorted(declared):
        assert hasattr(handle, sym), f"{sym} declared in include/sdm_b200.h but not exported"
    # the ctypes table covers exactly the compute entry points
    assert set(SIGNATURES) == declared - {"b2_last_error", "b2_version"}
    handle.b2_version.restype = ctypes.c_int
    assert handle.b2_version() >= 100


def test_ctypes_signatures_match_the_header_prototypes():
    """Every prototype of include/sdm_b200.h, parameter by parameter, against the ctypes table the host side calls through
    (a drifted argument list would corrupt the call silently: ctypes cannot check it)."""
    from b200._lib import SIGNATURES
    header = open(os.path.join(ROOT, "include", "sdm_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", " ", header, flags=re.S)
    header = re.sub(r"//[^\n]*", " ", header)
    protos = dict(re.findall(r"\bint\s+(b2_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", header))

    def code(param):
        param = " ".join(param.split())
        if "*" in param:
            return ctypes.c_void_p
        if param.startswith("unsigned long long"):
            return ctypes.c_ulonglong
        if param.startswith("long long"):
            return ctypes.c_longlong
        if param.startswith("double"):
            return ctypes.c_double
        if param.startswith("float"):
            return ctypes.c_float
        assert param.startswith("int"), param
        return ctypes.c_int

    for name, argtypes in SIGNATURES.items():
        assert name in protos, name
        params = [q for q in protos[name].split(",") if q.strip() and q.strip() != "void"]
        assert [code(q) for q in params] == list(argtypes), f"{name}: header {protos[name]!r} vs ctypes table"


def test_entry_points_keep_reference_names_and_fail_loudly_without_cuda(tmp_path):
    """train_*/generate_* (reference L4 scripts): same module / function names, config validation errors of the
    reference (train_diffusion.py:69-116), and no silent CPU path."""
    import inspect
    import json
    import generate_images_cold_diffusion as gc
    import generate_images_diffusion as gd
    import generate_sr_images_diffusion as gs
    import train_diffusion, train_doodle_diffusion, train_noise_cold_diffusion, train_SR_diffusion
    from b200._lib import B200Error
    assert list(inspect.signature(gd.generate_images_diffusion).parameters) == ["raw_args", "log", "cond_img", "save_locally"]
    assert list(inspect.signature(gc.generate_images_cold_diffusion).parameters) == ["raw_args", "log", "save_locally"]
    assert list(inspect.signature(gs.generate_sr_images_diffusion).parameters) == ["raw_args", "lr_img", "log", "save_locally"]
    cfg = dict(dataset_path="synthetic:4x3x32x32", out_dir=str(tmp_path / "o"), checkpoint_steps=1, lr_steps=1, max_epoch=1,
               plot_img_count=1, noise_scheduler="BOGUS", diffusion_alg="DDIM", min_noise_step=1, max_noise_step=10,
               max_actual_noise_step=10, skip_step=2)
    path = tmp_path / "c.json"
    for mod in (train_diffusion, train_doodle_diffusion, train_noise_cold_diffusion, train_SR_diffusion):
        path.write_text(json.dumps(cfg))
        with pytest.raises(B200Error):
            mod.main(["-c", str(path), "--device", "cpu"])
        with pytest.raises(ValueError, match="noise scheduler"):
            mod.main(["-c", str(path)])
    cfg.update(noise_scheduler="COSINE", skip_step=50)
    path.write_text(json.dumps(cfg))
    with pytest.raises(ValueError, match="step values"):
        train_diffusion.main(["-c", str(path)])
    cfg.update(skip_step=2, diffusion_alg="XYZ")
    path.write_text(json.dumps(cfg))
    with pytest.raises(ValueError, match="diffusion algorithm"):
        train_diffusion.main(["-c", str(path)])
    with pytest.raises(B200Error):
        gd.generate_images_diffusion(["-c", str(path), "--device", "cpu"])


def test_dataset_table_reader(tmp_path):
    """TinyDB files are JSON documents: the labelled datasets read them without the tinydb package."""
    import json
    from custom_dataset._tables import load_tables
    from custom_dataset.img_dataset import SyntheticImages
    db = {"Data": {"1": {"filename": "a.png", "smile": 1, "hat": 0}, "2": {"filename": "b.png", "smile": 0, "hat": 1}},
          "Labels": {"1": {"labels": ["smile", "hat"]}}}
    p = tmp_path / "db.json"
    p.write_text(json.dumps(db))
    rows, labels = load_tables(str(p))
    assert labels == ["smile", "hat"] and len(rows) == 2
    ds = SyntheticImages("synthetic:5x3x16x16:4")
    img, lab = ds[2]
    assert len(ds) == 5 and img.shape == (3, 16, 16) and lab.shape == (4,) and float(img.abs().max()) <= 1.0
    assert torch.equal(ds[2][0], img)


def test_every_custom_layer_class_is_exposed_and_refuses_cpu():
    """models/custom_layers.py exports (reference :18-341): every class constructs with the reference signature, and its
    standalone forward fails loudly on CPU tensors instead of falling back."""
    import models.custom_layers as cl
    from b200._lib import B200Error
    from b200.blocks import run_standalone  # noqa: F401  (the runner every standalone forward goes through)
    for name in ("Swish", "AdaGN", "ConditionalEmbedding", "AttentionBlock", "UpsampleBlock", "DownsampleBlock", "UNet_ConvBlock",
                 "ResidualBlock", "UNetBlock", "UNetBlockType"):
        assert hasattr(cl, name), name
    x = torch.randn((1, 64, 4, 4))
    for mod, args in ((cl.UNet_ConvBlock(64, 64, emb_dim=16), (torch.randn(1, 16),)), (cl.ResidualBlock(64, 64, emb_dim=16), (torch.randn(1, 16),)),
                      (cl.AttentionBlock(64), ()), (cl.UpsampleBlock(64, 64), ()), (cl.DownsampleBlock(64, 64), ()),
                      (cl.AdaGN(16, 64), (torch.randn(1, 16),)),
                      (cl.UNetBlock(64, 64, emb_dim=16, block_type=cl.UNetBlockType.DOWN), (torch.randn(1, 16),))):
        with pytest.raises(B200Error):
            mod(x, *args)
    with pytest.raises(B200Error):
        cl.Swish()(x)

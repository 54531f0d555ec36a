"""Host-side pieces of the DINOv2 encoder drop-in: state_dict layout of the published checkpoints, position-table
resampling as utils.py:98-101 does it, loader behaviour without checkpoints."""
import pytest
import torch
import torch.nn.functional as F


def test_state_dict_layout_is_the_published_one():
    from reed_b200.image.encoders import DinoV2
    sd = DinoV2(64, 2, 2, img_size=28, num_register_tokens=4).state_dict()
    want = {"cls_token", "pos_embed", "register_tokens", "mask_token", "patch_embed.proj.weight", "patch_embed.proj.bias",
            "norm.weight", "norm.bias"}
    for i in range(2):
        for k in ("norm1.weight", "norm1.bias", "attn.qkv.weight", "attn.qkv.bias", "attn.proj.weight", "attn.proj.bias",
                  "ls1.gamma", "norm2.weight", "norm2.bias", "mlp.fc1.weight", "mlp.fc1.bias", "mlp.fc2.weight",
                  "mlp.fc2.bias", "ls2.gamma"):
            want.add(f"blocks.{i}.{k}")
    assert set(sd) == want
    assert sd["patch_embed.proj.weight"].shape == (64, 3, 14, 14) and sd["pos_embed"].shape == (1, 5, 64)
    assert all(not p.requires_grad for p in DinoV2(64, 1, 2, img_size=28).parameters())      # frozen


def test_pos_embed_resampling_matches_timm_semantics():
    from reed_b200.image.encoders import build_dinov2, resample_abs_pos_embed
    table = torch.randn(1, 1 + 37 * 37, 48)
    out = resample_abs_pos_embed(table, [16, 16])
    assert out.shape == (1, 257, 48) and torch.equal(out[:, :1], table[:, :1])              # cls entry untouched
    grid = table[:, 1:].reshape(1, 37, 37, 48).permute(0, 3, 1, 2)
    want = F.interpolate(grid, size=(16, 16), mode="bicubic", antialias=True).permute(0, 2, 3, 1).reshape(1, 256, 48)
    assert torch.equal(out[:, 1:], want)
    assert resample_abs_pos_embed(out, [16, 16]) is out                                     # already the right size
    # a published-size checkpoint loads into the 224-pixel model (pos_embed resampled, everything else strict)
    from reed_b200.image.encoders import DinoV2
    src = DinoV2(384, 12, 6, img_size=37 * 14).state_dict()
    m = build_dinov2("s", resolution=256, state_dict=src)
    assert m.pos_embed.shape == (1, 257, 384)
    with pytest.raises(NotImplementedError):
        build_dinov2("g")


def test_load_encoders_needs_local_checkpoints(tmp_path):
    from reed_b200.image.encoders import DinoV2, load_encoders
    with pytest.raises(FileNotFoundError, match="no network"):
        load_encoders("dinov2-vit-b", "cpu", 256, ckpt_dir=str(tmp_path))
    with pytest.raises(NotImplementedError):
        load_encoders("mocov3-vit-b", "cpu", 256, ckpt_dir=str(tmp_path))
    torch.save(DinoV2(384, 12, 6, img_size=37 * 14).state_dict(), tmp_path / "dinov2_vits14_pretrain.pth")
    encs, types, archs = load_encoders("dinov2-vit-s", "cpu", 256, ckpt_dir=str(tmp_path))
    assert types == ["dinov2"] and archs == ["vit"] and encs[0].pos_embed.shape == (1, 257, 384) and not encs[0].training
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        encs[0].forward_features(torch.zeros(1, 3, 224, 224))

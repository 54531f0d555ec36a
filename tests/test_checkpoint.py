"""Checkpoint format of the train step (train.py:280-289 resume, 418-429 save): host-side logic, CPU only."""
import io

import pytest
import torch


def _tiny_model(seed=0):
    from reed_b200.image.models.sit import SiT
    torch.manual_seed(seed)
    return SiT(input_size=8, hidden_size=32, decoder_hidden_size=32, depth=2, num_heads=2, encoder_depth=1, z_dims=[16],
               projector_dim=32, num_classes=10, qk_norm=False)


def _trainer(seed=0, **kw):
    from reed_b200.image.trainer import ReedTrainer
    return ReedTrainer(_tiny_model(seed), None, **kw)


def _fake_progress(tr, steps=7):
    """Stand-in for `steps` optimizer steps (the kernels need a GPU): fill moments / weights / EMA with known values."""
    g = torch.Generator().manual_seed(5)
    for b in tr.state.buckets:
        b.exp_avg.copy_(torch.randn(b.numel, generator=g))
        b.exp_avg_sq.copy_(torch.rand(b.numel, generator=g))
        b.param.add_(torch.randn(b.numel, generator=g) * 0.01)
        b.ema.copy_(b.param * 0.5)
    tr.step_count = steps


def test_checkpoint_has_the_reference_layout_and_is_compact():
    tr = _trainer(lr=3e-4, weight_decay=0.01)
    _fake_progress(tr)
    ck = tr.checkpoint(args={"exp_name": "x"})
    assert sorted(ck) == ["args", "ema", "model", "opt", "steps"] and ck["steps"] == 7
    assert list(ck["model"]) == list(tr.model.state_dict()) == list(ck["ema"])
    # tensors are copies, not views of the flat buckets: the file holds ~ (2 weights + 2 moments) x parameters
    buf = io.BytesIO()
    torch.save(ck, buf)
    n = sum(p.numel() for p in tr.model.parameters())
    assert buf.tell() < 4 * 4 * n * 1.3
    for v in ck["model"].values():
        assert v._base is None
    # "opt" is a torch.optim.AdamW state_dict over model.parameters(): the reference's optimizer loads it (train.py:288)
    other = _tiny_model(seed=1)
    opt = torch.optim.AdamW(other.parameters(), lr=1e-4)
    opt.load_state_dict(torch.load(io.BytesIO(buf.getvalue()), weights_only=False)["opt"])
    assert opt.param_groups[0]["lr"] == 3e-4 and opt.param_groups[0]["weight_decay"] == 0.01
    names = [k for k, _ in tr.model.named_parameters()]
    params = list(other.parameters())
    assert names.index("pos_embed") not in ck["opt"]["state"]           # frozen: AdamW holds no state for it
    assert len(ck["opt"]["state"]) == len(names) - 1
    src = dict(tr.model.named_parameters())
    for idx, st in ck["opt"]["state"].items():
        assert float(opt.state[params[idx]]["step"]) == 7.0
        assert st["exp_avg"].shape == src[names[idx]].shape
    # and one torch AdamW step from that state moves the weights the way the moments say (state is really wired in)
    other.load_state_dict(ck["model"])
    for p in other.parameters():
        p.grad = torch.zeros_like(p)
    w0 = other.blocks[0].mlp.fc1.weight.detach().clone()
    opt.step()
    assert not torch.equal(other.blocks[0].mlp.fc1.weight, w0)


def test_checkpoint_before_the_first_step_has_empty_optimizer_state():
    ck = _trainer().checkpoint()
    assert ck["opt"]["state"] == {} and ck["steps"] == 0
    torch.optim.AdamW(_tiny_model().parameters()).load_state_dict(ck["opt"])


def test_resume_restores_flat_state_and_shadows():
    a = _trainer(seed=0)
    _fake_progress(a, steps=11)
    buf = io.BytesIO()
    torch.save(a.checkpoint(steps=12), buf)
    b = _trainer(seed=3, lr=5e-5)
    ptrs = [p.data_ptr() for p in b.model.parameters()]
    assert b.load_checkpoint(torch.load(io.BytesIO(buf.getvalue()), weights_only=False)) == 12
    assert b.step_count == 11 and int(b._step_dev) == 11 and b.lr == a.lr
    assert ptrs == [p.data_ptr() for p in b.model.parameters()]         # parameters still alias the flat buckets
    for ba, bb in zip(a.state.buckets, b.state.buckets):
        for name in ("param", "exp_avg", "exp_avg_sq", "ema"):
            assert torch.equal(getattr(ba, name), getattr(bb, name)), (ba.name, name)
        assert torch.equal(bb.shadow.float(), bb.param.bfloat16().float())
        for p in bb.params:
            assert p._reed_shadow_version == p._version                  # no lazy re-cast on the next forward
    assert torch.equal(a.model.pos_embed, b.model.pos_embed) and torch.equal(a.ema.pos_embed, b.ema.pos_embed)


def test_resume_from_a_reference_style_checkpoint():
    """A checkpoint as the reference writes it: plain modules + torch.optim.AdamW after real steps."""
    import copy
    ref = _tiny_model(seed=4)
    ema = copy.deepcopy(ref)
    opt = torch.optim.AdamW(ref.parameters(), lr=2e-4, betas=(0.9, 0.95), weight_decay=0.0, eps=1e-8)
    g = torch.Generator().manual_seed(0)
    for _ in range(3):
        for p in ref.parameters():
            p.grad = torch.randn(p.shape, generator=g) if p.requires_grad else None
        opt.step()
    ck = {"model": ref.state_dict(), "ema": ema.state_dict(), "opt": opt.state_dict(), "args": None, "steps": 3}
    tr = _trainer(seed=9)
    assert tr.load_checkpoint(ck) == 3
    assert tr.step_count == 3 and tr.betas == (0.9, 0.95) and tr.lr == 2e-4
    for name, p, b, off in tr._param_slots():
        assert torch.equal(p, dict(ref.named_parameters())[name])
        if b is not None:
            st = opt.state[dict(ref.named_parameters())[name]]
            assert torch.equal(b.exp_avg[off:off + p.numel()].view(p.shape), st["exp_avg"])
            assert torch.equal(b.exp_avg_sq[off:off + p.numel()].view(p.shape), st["exp_avg_sq"])
    # round trip back out: identical optimizer state
    again = tr.checkpoint()["opt"]
    for idx, st in opt.state_dict()["state"].items():
        assert torch.equal(again["state"][idx]["exp_avg_sq"], st["exp_avg_sq"])
        assert float(again["state"][idx]["step"]) == float(st["step"])
    bad = dict(ck, opt={"state": {}, "param_groups": [dict(opt.state_dict()["param_groups"][0], params=[0, 1])]})
    with pytest.raises(ValueError):
        tr.load_checkpoint(bad)

"""Host-side logic of the data-parallel path on CPU: flat bucketing + bucketed gradient all-reduce over gloo, world 2."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _tiny_model():
    from reed_b200.image.models.sit import SiT
    torch.manual_seed(0)
    return SiT(input_size=8, hidden_size=32, decoder_hidden_size=32, depth=3, num_heads=2, encoder_depth=1, z_dims=[16],
               projector_dim=32, num_classes=10, qk_norm=False)


def test_flat_state_layout_and_aliasing():
    from reed_b200.image.trainer import FlatState
    import copy
    model = _tiny_model()
    before = {k: v.clone() for k, v in model.state_dict().items()}
    ema = copy.deepcopy(model)
    st = FlatState(model, ema)
    assert [b.name for b in st.buckets] == ["outer", "blocks.0", "blocks.1", "blocks.2"]
    assert st.total_params() == sum(p.numel() for p in model.parameters() if p.requires_grad)
    for k, v in model.state_dict().items():
        assert torch.equal(v, before[k]), k                       # values preserved, keys unchanged
        assert torch.equal(ema.state_dict()[k], before[k]), k      # update_ema(decay=0)
    for b in st.buckets:
        for p, off in zip(b.params, b.offsets):
            assert off % 8 == 0
            assert p.data_ptr() == b.param.data_ptr() + 4 * off    # parameters alias the flat buffer
            assert p._reed_shadow.data_ptr() == b.shadow.data_ptr() + 2 * off
            assert torch.equal(p._reed_shadow.float(), p.data.bfloat16().float())
    # gradient lifecycle: untouched kernel-written grads are zeroed, autograd-written ones accumulate
    st.begin_step()
    emb = model.y_embedder.embedding_table.weight
    assert emb.grad is not None and float(emb.grad.abs().sum()) == 0.0
    qkv = model.blocks[1].attn.qkv.weight
    assert qkv.grad is None and qkv._reed_grad_fresh
    qkv._reed_main_grad.fill_(3.0)          # stale garbage from a previous step
    st.finish_backward()
    assert float(qkv.grad.abs().sum()) == 0.0


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from reed_b200.image.trainer import FlatState, GradientReducer
        model = _tiny_model()
        st = FlatState(model, None, with_shadow=False)
        red = GradientReducer(st)
        assert red.world == world and red.grad_scale == 1.0 / world
        for i, b in enumerate(st.buckets):
            b.grad.copy_(torch.arange(b.numel, dtype=torch.float32) * (rank + 1) + i)
        # micro-batches of an accumulated step hold the all-reduce back (DDP no_sync): launch is a no-op while disabled
        red.enabled = False
        red.launch(st.bucket_of_block(2))
        held = st.bucket_of_block(2).work is None
        red.enabled = True
        # block buckets are launched early (as their backward completes), the rest in finish()
        red.launch(st.bucket_of_block(2))
        red.launch(st.bucket_of_block(2))        # idempotent
        red.finish()
        ok = held
        for i, b in enumerate(st.buckets):
            want = torch.arange(b.numel, dtype=torch.float32) * sum(r + 1 for r in range(world)) + i * world
            ok &= bool(torch.equal(b.grad, want))
            ok &= b.work is None
        # .grad views alias the reduced storage
        p = model.blocks[0].mlp.fc1.weight
        st.finish_backward()
        ok &= p.grad.data_ptr() == p._reed_main_grad.data_ptr()
        out[rank] = ok
    finally:
        dist.destroy_process_group()


def test_bucketed_allreduce_world2_gloo():
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}

"""CPU, world_size=2, gloo: the data-parallel exchange of the training step.  Each rank owns different utterances;
after ParameterArena.all_reduce_grads() (one flat all-reduce) and the 1/world scale the optimizer applies, every
rank must hold the gradient of the whole global batch -- i.e. DDP's averaging semantics."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from torch import nn


class _Toy(nn.Module):
    def __init__(self):
        super().__init__()
        torch.manual_seed(0)
        self.a = nn.Linear(6, 5)
        self.b = nn.Linear(5, 3)
        self.frozen = nn.Parameter(torch.ones(4))        # like the position tables: not in the arena

    def get_trainable_parameters(self):
        return (p for n, p in self.named_parameters() if n != "frozen")

    def forward(self, x):
        return self.b(torch.tanh(self.a(x))).pow(2).mean()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from deepvoice3_pytorch_b200.train_step import ParameterArena
    model = _Toy()
    arena = ParameterArena(model)
    assert all(p.data_ptr() >= arena.flat.data_ptr() for p in arena.params)      # parameters re-homed
    gen = torch.Generator().manual_seed(123)
    data = torch.randn(world * 4, 6, generator=gen)
    arena.zero_grad()
    model(data[rank * 4:(rank + 1) * 4]).backward()          # this rank's utterances
    assert arena.params[0].grad.data_ptr() == arena.grad.data_ptr()             # autograd wrote into the arena
    arena.all_reduce_grads()
    ret[rank] = (arena.grad / world).clone()
    dist.destroy_process_group()


def test_flat_gradient_allreduce_matches_global_batch():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert torch.equal(ret[0], ret[1])
    # single-process gradient of the mean loss over the global batch
    model = _Toy()
    gen = torch.Generator().manual_seed(123)
    data = torch.randn(world * 4, 6, generator=gen)
    loss = sum(model(data[r * 4:(r + 1) * 4]) for r in range(world)) / world
    loss.backward()
    want = torch.cat([torch.cat([p.grad.flatten(), torch.zeros((-p.numel()) % 4)])
                      for p in model.get_trainable_parameters()])
    torch.testing.assert_close(ret[0], want, rtol=1e-6, atol=1e-7)


def test_adam_clip_reference_math():
    """The formula csrc/optim.cu implements == clip_grad_norm_ + torch.optim.Adam (checked here on the CPU;
    the kernel itself is compared with this in tests/test_gpu_train.py)."""
    from test_gpu_train import adam_clip_reference
    torch.manual_seed(1)
    p0, g = torch.randn(1000), torch.randn(1000) * 3
    p = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([p], lr=1e-3, betas=(0.5, 0.9), eps=1e-6)
    m, v, q = torch.zeros(1000), torch.zeros(1000), p0.clone()
    for t in range(1, 4):
        p.grad = g.clone() * t
        torch.nn.utils.clip_grad_norm_([p], 0.1)
        opt.step()
        q, m, v = adam_clip_reference(q, g * t, m, v, t, lr=1e-3, betas=(0.5, 0.9), eps=1e-6, max_norm=0.1)
    torch.testing.assert_close(q, p.detach(), rtol=1e-5, atol=1e-7)


def test_gradient_bucket_plan_covers_the_arena():
    """train_step.gradient_buckets: the post-net and the upper encoder layers are contiguous arena slices, disjoint,
    and together with the "rest" ranges tile the whole gradient arena exactly once (every gradient is all-reduced
    once); the encoder bucket holds the layers whose backward finishes first."""
    from deepvoice3_pytorch_b200 import builder
    from deepvoice3_pytorch_b200.train_step import ParameterArena, gradient_buckets
    for name, kw in (("deepvoice3", dict(n_vocab=30, embed_dim=16, encoder_channels=32, decoder_channels=16,
                                         converter_channels=16, downsample_step=4, r=1, kernel_size=3)),
                     ("deepvoice3_multispeaker", dict(n_vocab=30, embed_dim=16, encoder_channels=32,
                                                      decoder_channels=16, converter_channels=16, downsample_step=4,
                                                      r=1, kernel_size=3, n_speakers=5)),
                     ("nyanko", dict(n_vocab=30, embed_dim=16, encoder_channels=32, decoder_channels=32,
                                     converter_channels=32, downsample_step=4, r=1, kernel_size=3))):
        model = getattr(builder, name)(**kw)
        arena = ParameterArena(model)
        tagged, rest = gradient_buckets(model, arena)
        assert "postnet" in tagged
        assert ("encoder_hi" in tagged) == (name != "nyanko")
        spans = sorted(list(tagged.values()) + rest)
        assert spans[0][0] == 0 and spans[-1][1] == arena.numel
        for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
            assert a1 == b0 and a0 < a1                      # contiguous, non-overlapping, non-empty
        lo, hi = tagged["postnet"]
        assert hi - lo == sum((p.numel() + 3) // 4 * 4 for p in model.postnet.parameters())
        if "encoder_hi" in tagged:
            enc = model.seq2seq.encoder
            assert "encoder_mid" in tagged and tagged["encoder_mid"][1] == tagged["encoder_hi"][0]
            for tag, idx in enc.grad_bucket_splits():
                p0 = next(list(enc.convolutions)[idx].parameters())
                assert p0.data_ptr() == arena.flat[tagged[tag][0]:].data_ptr()


def _bucket_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from deepvoice3_pytorch_b200.train_step import ParameterArena
    torch.manual_seed(rank)                                  # replicas start DIFFERENT ...
    model = _Toy()
    with torch.no_grad():
        for p in model.parameters():
            p.add_(rank)
    arena = ParameterArena(model)
    arena.broadcast(model)                                   # ... and take rank 0's weights (parameters and frozen)
    ret["w%d" % rank] = torch.cat([arena.flat.clone(), model.frozen.detach().clone()])
    arena.grad.copy_(torch.arange(arena.numel, dtype=torch.float32) * (rank + 1))
    n = arena.numel
    arena.all_reduce_grads([(0, n // 3), (n // 3, n // 3), (n // 3, n)])      # slices incl. an empty one
    ret["g%d" % rank] = arena.grad.clone()
    dist.destroy_process_group()


def test_broadcast_and_bucketed_allreduce():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_bucket_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert torch.equal(ret["w0"], ret["w1"])
    want = torch.arange(ret["g0"].numel(), dtype=torch.float32) * 3
    assert torch.equal(ret["g0"], want) and torch.equal(ret["g1"], want)


def test_flat_adam_state_dict_round_trips_with_torch_adam():
    """FlatAdam.state_dict / load_state_dict speak torch.optim.Adam's checkpoint format (what reference
    train.py:save_checkpoint stores under "optimizer"): moments, step count and hyper-parameters survive both ways."""
    from deepvoice3_pytorch_b200.train_step import ParameterArena, FlatAdam
    model = _Toy()
    params = list(model.get_trainable_parameters())
    ref = torch.optim.Adam(params, lr=3e-4, betas=(0.5, 0.9), eps=1e-6)
    for t in range(3):
        for p in params:
            p.grad = torch.randn_like(p) * (t + 1)
        ref.step()
    sd = ref.state_dict()
    arena = ParameterArena(model)
    flat = FlatAdam(arena, lr=1.0, betas=(0.9, 0.999), eps=1e-8)
    flat.load_state_dict(sd)
    assert flat.t == 3 and flat.betas == (0.5, 0.9) and flat.eps == 1e-6 and flat.lr == 3e-4
    for i, (p, o) in enumerate(zip(arena.params, arena.offsets)):
        assert torch.equal(flat.m[o:o + p.numel()].view_as(p), sd["state"][i]["exp_avg"])
        assert torch.equal(flat.v[o:o + p.numel()].view_as(p), sd["state"][i]["exp_avg_sq"])
    out = flat.state_dict()
    fresh = torch.optim.Adam(params, lr=1.0)
    fresh.load_state_dict(out)                       # torch accepts it ...
    back = fresh.state_dict()
    for i in range(len(params)):
        assert torch.equal(back["state"][i]["exp_avg"], sd["state"][i]["exp_avg"])
        assert float(back["state"][i]["step"]) == 3.0
    assert back["param_groups"][0]["betas"] == (0.5, 0.9)
    import pytest
    bad = {"state": {}, "param_groups": [dict(sd["param_groups"][0], weight_decay=0.1)]}
    with pytest.raises(ValueError):
        flat.load_state_dict(bad)

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

#!/usr/bin/env python
"""torchrun -N: after K data-parallel TrainStep steps (graph-captured, bucketed / overlapped NCCL exchange) every rank
holds bit-identical parameters, and they equal the parameters of the un-overlapped single-all-reduce path."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from deepvoice3_pytorch_b200 import builder, ops  # noqa: E402
from deepvoice3_pytorch_b200.train_step import TrainStep, make_synthetic_batch, to_device  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
bname, kw, extra = bench.PRESETS["deepvoice3_ljspeech"]
kw = dict(kw, dropout=0.0)
results = {}
for mode in ("overlap_graph", "plain_eager"):
    os.environ["DV3_OVERLAP_COMM"] = "1" if mode == "overlap_graph" else "0"
    torch.manual_seed(1234 + rank)                       # different init per rank: the broadcast must fix it
    model = getattr(builder, bname)(**kw).to(dev)
    step = TrainStep(model, use_graph=(mode == "overlap_graph"), **extra)
    batch = to_device(make_synthetic_batch(16, 128, 800, seed=77 + rank), dev)
    losses = [float(step.step(batch)) for _ in range(4)]
    flat = step.arena.flat.clone()
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    same = all(torch.equal(gathered[0], g) for g in gathered)
    results[mode] = (flat, losses, same)
    del step, model
    torch.cuda.empty_cache()
if rank == 0:
    for mode, (flat, losses, same) in results.items():
        print(mode, "replicas bit-identical:", same, "losses", ["%.5f" % l for l in losses])
    a, b = results["overlap_graph"][0], results["plain_eager"][0]
    print("overlap+graph vs plain: max |dparam| = %.3e (rel %.3e)" % (float((a - b).abs().max()),
                                                                    float((a - b).norm() / b.norm())))
dist.barrier()
dist.destroy_process_group()

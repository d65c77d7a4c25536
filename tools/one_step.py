#!/usr/bin/env python
"""Two eager training steps of the bench workload (the first warms up / registers the weight bank); run under
`ncu --metrics gpu__time_duration.sum` to get the launch list of a step."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from deepvoice3_pytorch_b200 import builder  # noqa: E402
from deepvoice3_pytorch_b200.train_step import TrainStep, make_synthetic_batch, to_device  # noqa: E402

bname, kw, extra = bench.PRESETS["deepvoice3_ljspeech"]
torch.manual_seed(1234)
step = TrainStep(getattr(builder, bname)(**kw).cuda(), use_graph=False, **extra)
batch = to_device(make_synthetic_batch(), "cuda")
for i in range(3):
    torch.cuda.nvtx.range_push("step%d" % i)
    step.step(batch)
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_pop()
print("done")

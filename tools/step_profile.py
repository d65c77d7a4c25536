#!/usr/bin/env python
"""Per-kernel GPU time of one eager training step (torch.profiler / CUPTI; cheap alternative to an ncu launch list)."""
import os
import sys
import collections

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from deepvoice3_pytorch_b200 import builder, ops  # noqa: E402
from deepvoice3_pytorch_b200.train_step import TrainStep, make_synthetic_batch, to_device  # noqa: E402

ops.conv_math = sys.argv[1] if len(sys.argv) > 1 else "tc"
preset = sys.argv[2] if len(sys.argv) > 2 else "deepvoice3_ljspeech"
bname, kw, extra = bench.PRESETS[preset]
torch.manual_seed(1234)
model = getattr(builder, bname)(**kw).cuda()
step = TrainStep(model, use_graph=False, **extra)
batch = to_device(make_synthetic_batch(n_speakers=kw["n_speakers"]), "cuda")
for _ in range(3):
    step.step(batch)
torch.cuda.synchronize()
N = 3
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    for _ in range(N):
        step.step(batch)
    torch.cuda.synchronize()
tot = collections.defaultdict(float)
cnt = collections.Counter()
for e in prof.key_averages():
    k = e.key.replace("void ", "")
    k = k.split("(")[0][:90]
    tot[k] += e.device_time_total / N
    cnt[k] += e.count / N
T = sum(tot.values())
print("GPU busy per step: %.2f ms over %d launches" % (T / 1e3, sum(cnt.values())))
for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:40]:
    print("%8.1f us %5.1f%% %6.1f  %s" % (v, 100 * v / T, cnt[k], k))

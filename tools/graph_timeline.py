#!/usr/bin/env python
"""Timeline of one CUDA-graph replay of the training step: per-stream busy time, union busy time, idle gaps and the
largest kernels, from a torch.profiler chrome trace (kernel start / duration / stream)."""
import collections
import json
import os
import sys
import tempfile

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
step = TrainStep(model, use_graph=True, **extra)
batch = to_device(make_synthetic_batch(n_speakers=kw["n_speakers"]), "cuda")
for _ in range(5):
    step.step(batch)
torch.cuda.synchronize()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    step.step(batch)
    torch.cuda.synchronize()
path = os.path.join(tempfile.mkdtemp(), "trace.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
ev.sort(key=lambda e: e["ts"])
t0 = ev[0]["ts"]
t1 = max(e["ts"] + e["dur"] for e in ev)
print("replay span %.1f us, %d device activities" % (t1 - t0, len(ev)))
per_stream = collections.defaultdict(float)
for e in ev:
    per_stream[e["args"].get("stream")] += e["dur"]
for s, v in sorted(per_stream.items(), key=lambda kv: -kv[1]):
    print("  stream %s busy %.1f us" % (s, v))
# union of busy intervals and idle gaps
iv = sorted((e["ts"], e["ts"] + e["dur"]) for e in ev)
busy, gaps, cur_s, cur_e = 0.0, [], iv[0][0], iv[0][1]
for a, b in iv[1:]:
    if a > cur_e:
        busy += cur_e - cur_s
        gaps.append((a - cur_e, cur_e - t0))
        cur_s, cur_e = a, b
    else:
        cur_e = max(cur_e, b)
busy += cur_e - cur_s
print("union busy %.1f us, idle %.1f us in %d gaps (%.2f us mean); overlap (sum - union) %.1f us"
      % (busy, (t1 - t0) - busy, len(gaps), ((t1 - t0) - busy) / max(1, len(gaps)), sum(per_stream.values()) - busy))
hist = collections.Counter()
for g, _ in gaps:
    hist[min(int(g), 10)] += 1
print("gap histogram (us -> count):", dict(sorted(hist.items())))
print("largest gaps (us @ offset):", ["%.1f@%.0f" % g for g in sorted(gaps, reverse=True)[:12]])
# what precedes the largest gaps
ends = sorted((e["ts"] + e["dur"], e["name"][:60]) for e in ev)
starts = sorted((e["ts"], e["name"][:60]) for e in ev)
import bisect
for g, off in sorted(gaps, reverse=True)[:8]:
    i = bisect.bisect_right([x[0] for x in ends], t0 + off + 1e-3) - 1
    j = bisect.bisect_left([x[0] for x in starts], t0 + off + g - 1e-3)
    print("  gap %.1f us after [%s] before [%s]" % (g, ends[i][1], starts[min(j, len(starts) - 1)][1]))
# time by kernel inside the replay
tot = collections.defaultdict(float)
cnt = collections.Counter()
for e in ev:
    k = e["name"].replace("void ", "").split("(")[0][:70]
    tot[k] += e["dur"]
    cnt[k] += 1
for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:16]:
    print("%8.1f us %4d  %s" % (v, cnt[k], k))

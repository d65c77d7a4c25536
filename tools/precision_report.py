#!/usr/bin/env python
"""Full-depth accuracy of each ConvBlock arithmetic mode at the ljspeech preset (B=4, T_text=128, T_mel=800):
every output against the fp64 oracle (truth), next to the CPU fp32 oracle's own error."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from deepvoice3_pytorch_b200 import builder, ops  # noqa: E402
from oracle import dv3_oracle as O  # noqa: E402
from oracle.specs import spec_from_builder  # noqa: E402
from test_gpu_models import preset_kwargs, synthetic_batch  # noqa: E402


def main():
    bname, kw = preset_kwargs("deepvoice3_ljspeech")
    kw["dropout"] = 0.0
    torch.manual_seed(11)
    model = getattr(builder, bname)(**kw)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    text, mel, tpos, fpos, lengths, spk = synthetic_batch(4, 128, 200, 1, 77)
    spec = spec_from_builder(bname, **kw)
    outs = {}
    for name, dt in (("cpu_fp64", torch.float64), ("cpu_fp32", torch.float32)):
        s = {k: (v.to(dt) if v.is_floating_point() else v) for k, v in sd.items()}
        with torch.no_grad():
            outs[name] = [o.double() for o in O.model_forward(s, spec, text, mel.to(dt), None, tpos, fpos, lengths)]
    model = model.cuda().eval()
    for math in sys.argv[1:] or ["fp32", "bf16x3"]:
        ops.conv_math = math
        with torch.no_grad():
            o = model(text.cuda(), mel.cuda(), text_positions=tpos.cuda(), frame_positions=fpos.cuda(),
                      input_lengths=lengths)
        outs["gpu_" + math] = [t.double().cpu() for t in o]
    truth, ref32 = outs["cpu_fp64"], outs["cpu_fp32"]
    for name in outs:
        if name == "cpu_fp64":
            continue
        row = []
        for i, nm in enumerate(["mel", "linear", "align", "done"]):
            err = (outs[name][i] - truth[i]).abs()
            viol = ((outs[name][i] - ref32[i]).abs() > 1e-4 + 1e-3 * ref32[i].abs()).double().mean()
            row.append("%s max|e|=%.2e rms=%.2e viol=%.1e" % (nm, err.max(), err.pow(2).mean().sqrt(), viol))
        print("%-12s %s" % (name, " | ".join(row)), flush=True)


if __name__ == "__main__":
    main()

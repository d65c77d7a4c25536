#!/usr/bin/env python
"""Full-depth accuracy of the ConvBlock arithmetic modes for the three BASELINE presets at B=16, T_text=128,
T_mel=800: every model output against the CPU fp32 oracle (the parity target) and the fp64 oracle (truth).
Prints, per output, the absolute tolerance that rtol=1e-3 would still need (north_star allows 1e-4).

    python tools/precision_presets.py [preset ...] [--math tc,fp32] [--B 16]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from deepvoice3_pytorch_b200 import builder, ops  # noqa: E402
from oracle import dv3_oracle as O  # noqa: E402
from oracle.specs import spec_from_builder  # noqa: E402
from test_gpu_models import preset_kwargs, synthetic_batch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("presets", nargs="*", default=["deepvoice3_ljspeech", "nyanko_ljspeech", "deepvoice3_vctk"])
    ap.add_argument("--math", default="tc,fp32")
    ap.add_argument("--B", type=int, default=16)
    ap.add_argument("--no64", action="store_true")
    a = ap.parse_args()
    for preset in a.presets:
        bname, kw = preset_kwargs(preset)
        kw["dropout"] = 0.0
        torch.manual_seed(11)
        model = getattr(builder, bname)(**kw)
        with torch.no_grad():
            gen = torch.Generator().manual_seed(5)
            for n, p in model.named_parameters():
                if n.endswith("weight_g"):
                    p.mul_(1 + 0.1 * torch.randn(p.shape, generator=gen))
                elif n.endswith("bias"):
                    p.add_(0.05 * torch.randn(p.shape, generator=gen))
        sd = {k: v.clone() for k, v in model.state_dict().items()}
        text, mel, tpos, fpos, lengths, spk = synthetic_batch(a.B, 128, 200, kw["n_speakers"], 77)
        spec = spec_from_builder(bname, **kw)
        outs = {}
        for name, dt in (("cpu_fp32", torch.float32),) + (() if a.no64 else (("cpu_fp64", torch.float64),)):
            s = {k: (v.to(dt) if v.is_floating_point() else v) for k, v in sd.items()}
            with torch.no_grad():
                outs[name] = [o.double() for o in O.model_forward(s, spec, text, mel.to(dt), spk, tpos, fpos, lengths)]
        model = model.cuda().eval()
        for math in a.math.split(","):
            ops.conv_math = math
            with torch.no_grad():
                o = model(text.cuda(), mel.cuda(), speaker_ids=None if spk is None else spk.cuda(),
                          text_positions=tpos.cuda(), frame_positions=fpos.cuda(), input_lengths=lengths)
            outs["gpu_" + math] = [t.double().cpu() for t in o]
        ref32 = outs["cpu_fp32"]
        truth = outs.get("cpu_fp64", ref32)
        print("== %s B=%d" % (preset, a.B), flush=True)
        for name in outs:
            if name == "cpu_fp64":
                continue
            row = []
            for i, nm in enumerate(["mel", "linear", "align", "done"]):
                e64 = (outs[name][i] - truth[i]).abs()
                e32 = (outs[name][i] - ref32[i]).abs()
                need = (e32 - 1e-3 * ref32[i].abs()).max()       # atol needed at rtol=1e-3 vs the fp32 oracle
                row.append("%s max64=%.2e rms64=%.2e atol_needed=%.2e" % (nm, e64.max(), e64.pow(2).mean().sqrt(), need))
            print("%-10s %s" % (name, " | ".join(row)), flush=True)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Inference-path benchmark: decoder steps/s of the autoregressive loop at the preset sizes.

    python bench_incremental.py [--preset deepvoice3_ljspeech] [--batch 1] [--steps 200] [--cpu]

GPU arm: deepvoice3_pytorch_b200.incremental.decode, teacher-forced over ``--steps`` frames (fixed work: the free run
stops on data-dependent done flags), CUDA-graph replay and eager stepping.  ``--cpu`` adds the CPU oracle's stepwise
decoder (oracle/dv3_incremental.py = the reference's algorithm in torch CPU ops) on the same weights as the baseline.
Prints one JSON line.
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--preset", default="deepvoice3_ljspeech")
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--text", type=int, default=128)
    ap.add_argument("--cpu", action="store_true")
    a = ap.parse_args()
    from test_gpu_models import preset_kwargs
    from deepvoice3_pytorch_b200 import builder, incremental
    bname, kw = preset_kwargs(a.preset)
    torch.manual_seed(1234)
    model = getattr(builder, bname)(dropout=0.05, **kw).cuda().eval()
    B, Tt, N = a.batch, a.text, a.steps
    gen = torch.Generator().manual_seed(1)
    text = torch.randint(2, 149, (B, Tt), generator=gen).cuda()
    tpos = torch.arange(1, Tt + 1)[None].repeat(B, 1).cuda()
    mel = torch.rand(B, N, 80, generator=gen).cuda()
    spk_ids = torch.randint(0, kw["n_speakers"], (B,), generator=gen).cuda() if kw["n_speakers"] > 1 else None
    out = {"metric": "decoder steps/sec (autoregressive inference)", "unit": "steps/s", "preset": a.preset,
           "batch": B, "steps": N, "t_text": Tt, "dtype": "f32", "data": "synthetic"}
    with torch.no_grad():
        spk = model.embed_speakers(spk_ids) if spk_ids is not None else None
        enc = model.seq2seq.encoder(text, speaker_embed=spk) if bname != "nyanko" else model.seq2seq.encoder(text)
        dec = model.seq2seq.decoder
        for name, graph in (("graph", True), ("eager", False)):
            for rep in range(3):                      # first call = warm-up (graph capture, allocations)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                incremental.decode(dec, enc, tpos, spk, test_inputs=mel, use_graph=graph)
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
            out[name] = {"steps_per_s": N / dt, "ms_per_step": 1e3 * dt / N, "includes": "set-up + graph capture"}
    out["value"] = out["graph"]["steps_per_s"]
    if a.cpu:
        from oracle import dv3_incremental as OI
        from oracle.specs import spec_from_builder
        sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
        spec = spec_from_builder(bname, **kw)
        encc = tuple(t.cpu() for t in enc)
        n_cpu = min(N, 50)
        torch.set_num_threads(min(32, os.cpu_count() or 1))
        t0 = time.perf_counter()
        if bname == "nyanko":
            OI.nyanko_decoder_incremental(sd, spec, encc, tpos.cpu(), test_inputs=mel[:, :n_cpu].cpu())
        else:
            OI.dv3_decoder_incremental(sd, spec, encc, tpos.cpu(), spk.cpu() if spk is not None else None,
                                       test_inputs=mel[:, :n_cpu].cpu())
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": n_cpu / dt, "unit": "steps/s", "cores": torch.get_num_threads(),
                               "kind": "port", "sample": "%d teacher-forced steps, oracle port" % n_cpu}
    print(json.dumps(out))


if __name__ == "__main__":
    main()

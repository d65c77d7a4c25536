#!/usr/bin/env python
"""BASELINE.json config #5: STFT -> linear + mel throughput on synthetic 22.05 kHz, 10 s clips (220 500 samples).

    python bench_stft.py [--clips 512] [--iters 5] [--gpus N under torchrun]

Prints ONE JSON line: clips/s (device-resident and end-to-end incl. pinned H2D of the waveforms and D2H of both
outputs), achieved algorithmic GB/s against the measured HBM peak (per clip: 882 KB in + 1.774 MB linear + 277 KB mel),
and the CPU baseline = the numpy restatement of audio.py (oracle/audio_oracle.py; "port", parity unpinned) on the
host cores through a process pool, like the reference's preprocessors (ljspeech.py:24).
Clips are sharded round-robin over ranks; there is no collective on this path.
"""
import argparse
import json
import os
import sys
import time
from concurrent.futures import ProcessPoolExecutor

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
N_SAMPLES = 220500


def _cpu_clip(seed):
    from oracle import audio_oracle as A
    x = A.synthetic_clip(seed)
    t0 = time.perf_counter()
    A.process_utterance(x)
    return time.perf_counter() - t0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=512)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--cpu-clips", type=int, default=64)
    args = ap.parse_args()
    from deepvoice3_pytorch_b200 import audio
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    n = args.clips
    gen = torch.Generator().manual_seed(100 + rank)
    host = (0.1 * torch.randn(n, N_SAMPLES, generator=gen)).clamp_(-1, 1).pin_memory()
    wav = host.to(dev)
    frames = audio.num_frames(N_SAMPLES)
    lin_h = torch.empty(n, frames, 513).pin_memory()
    mel_h = torch.empty(n, frames, 80).pin_memory()

    def timed(fn):
        fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(args.iters):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); fn(); e.record()
            torch.cuda.synchronize()
            ts.append(s.elapsed_time(e) * 1e-3)
        t = torch.tensor([float(np.median(ts))], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    t_dev = timed(lambda: audio.stft_mel_batch(wav))

    def e2e():
        w = host.to(dev, non_blocking=True)
        lin, mel = audio.stft_mel_batch(w)
        lin_h.copy_(lin, non_blocking=True)
        mel_h.copy_(mel, non_blocking=True)
    t_e2e = timed(e2e)
    bytes_clip = 4.0 * (N_SAMPLES + frames * 513 + frames * 80)
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6650.0
    out = {"metric": "STFT/mel preprocess throughput (10 s clips @22.05 kHz)", "unit": "clips/s", "n_gpus": world,
           "value": n * world / t_dev, "e2e": {"value": n * world / t_e2e, "unit": "clips/s",
                                               "h2d_bytes_per_step": int(host.numel() * 4),
                                               "d2h_bytes_per_step": int((lin_h.numel() + mel_h.numel()) * 4)},
           "clips_per_rank": n, "ms_per_batch": t_dev * 1e3,
           "roofline": {"bound": "hbm", "achieved": bytes_clip * n / t_dev / 1e9, "peak": peak, "unit": "GB/s",
                        "frac": bytes_clip * n / t_dev / 1e9 / peak, "alg_bytes_per_clip": bytes_clip},
           "data": "synthetic", "dtype": "f32"}
    if rank == 0:
        if world == 1 and args.cpu_clips > 0:
            workers = min(os.cpu_count() or 8, 32)
            t0 = time.perf_counter()
            with ProcessPoolExecutor(workers) as ex:
                list(ex.map(_cpu_clip, range(args.cpu_clips)))
            dt = time.perf_counter() - t0
            out["cpu_baseline"] = {"value": args.cpu_clips / dt, "unit": "clips/s", "cores": workers, "kind": "port",
                                   "sample": "%d clips through a %d-process pool" % (args.cpu_clips, workers)}
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Headline benchmark: mel-frames/sec of one training step (B=16 per GPU, T_text=128, T_mel=800) of the
deepvoice3_ljspeech preset on synthetic data, plus the fused-ConvBlock roofline and the CPU baseline.

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \\
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the CPU arm: the oracle port of the reference on host cores

One step = zero_grad -> forward -> the reference's losses (train.py:704-740) -> backward -> (NCCL gradient
all-reduce) -> clip_grad_norm(0.1) -> Adam.  Timed with CUDA events on the launching stream, barrier +
synchronize on both sides, max over ranks.  A step touches > 1.5 GB of weights, optimizer state and
activations, i.e. far more than the 126 MB L2 ("inputs larger than L2").
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PRESETS = {
    # train.py:812-840 hparams -> builder kwargs, values from presets/*.json
    "deepvoice3_ljspeech": ("deepvoice3", dict(
        n_speakers=1, speaker_embed_dim=16, n_vocab=149, embed_dim=256, mel_dim=80, linear_dim=513, r=1,
        downsample_step=4, padding_idx=0, dropout=0.05, kernel_size=3, encoder_channels=512,
        decoder_channels=256, converter_channels=256, use_memory_mask=True,
        trainable_positional_encodings=False, force_monotonic_attention=True,
        use_decoder_state_for_postnet_input=True, max_positions=512, speaker_embedding_weight_std=0.01,
        freeze_embedding=False, window_ahead=3, window_backward=1, key_projection=True,
        value_projection=True), dict(guided_attention_sigma=0.2)),
    "nyanko_ljspeech": ("nyanko", dict(
        n_speakers=1, speaker_embed_dim=16, n_vocab=149, embed_dim=128, mel_dim=80, linear_dim=513, r=1,
        downsample_step=4, padding_idx=0, dropout=0.05, kernel_size=3, encoder_channels=256,
        decoder_channels=256, converter_channels=256, use_memory_mask=True,
        trainable_positional_encodings=False, force_monotonic_attention=True,
        use_decoder_state_for_postnet_input=True, max_positions=512, speaker_embedding_weight_std=0.01,
        freeze_embedding=False, window_ahead=3, window_backward=1, key_projection=False,
        value_projection=False), dict(guided_attention_sigma=0.2)),
    "deepvoice3_vctk": ("deepvoice3_multispeaker", dict(
        n_speakers=108, speaker_embed_dim=16, n_vocab=149, embed_dim=256, mel_dim=80, linear_dim=513, r=1,
        downsample_step=4, padding_idx=0, dropout=0.05, kernel_size=3, encoder_channels=512,
        decoder_channels=256, converter_channels=256, use_memory_mask=True,
        trainable_positional_encodings=False, force_monotonic_attention=True,
        use_decoder_state_for_postnet_input=True, max_positions=1024, speaker_embedding_weight_std=0.05,
        freeze_embedding=False, window_ahead=3, window_backward=1, key_projection=True,
        value_projection=True), dict(guided_attention_sigma=0.4)),
}
B, T_TEXT, T_MEL = 16, 128, 800
METRIC = "mel-frames/sec training step (B=16,T_mel=800)"


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"],
                                   r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# -------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference modules, timed on the host cores
# -------------------------------------------------------------------------------------------------
def cpu_step_throughput(preset, steps, warmup, threads=None):
    """Full training step (oracle forward restating the reference modules + reference losses + torch
    autograd + clip + Adam) on the host.  Returns (mel-frames/s, cores, seconds per step)."""
    from oracle import dv3_oracle as O
    from oracle.specs import spec_from_builder
    from deepvoice3_pytorch_b200 import builder
    from deepvoice3_pytorch_b200.train_step import make_synthetic_batch, noam_learning_rate_decay
    # oneDNN's small convolutions stop scaling (and on shared 100+-core hosts collapse) beyond a few dozen
    # threads: use at most 32 (measured: 128 threads on the B200 host = 134 s/step vs ~2-6 s/step at 8-32).
    threads = threads or min(os.cpu_count() or 8, 32)
    torch.set_num_threads(threads)
    bname, kw, extra = PRESETS[preset]
    torch.manual_seed(1234)
    model = getattr(builder, bname)(**kw)                 # parameter container only; never run on the CPU
    kw0 = dict(kw, dropout=0.0)
    spec = spec_from_builder(bname, **kw0)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    frozen = {"seq2seq.decoder.embed_query_positions.weight", "seq2seq.decoder.embed_keys_positions.weight"}
    leaves = [v.requires_grad_(True) for k, v in sd.items() if v.is_floating_point() and k not in frozen]
    opt = torch.optim.Adam(leaves, lr=5e-4, betas=(0.5, 0.9), eps=1e-6)
    b = make_synthetic_batch(B, T_TEXT, T_MEL, n_speakers=kw["n_speakers"])
    ts = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        for g in opt.param_groups:
            g["lr"] = noam_learning_rate_decay(5e-4, i)
        opt.zero_grad()
        outs = O.model_forward(sd, spec, b["x"], b["mel"], b.get("speaker_ids"), b["text_positions"],
                               b["frame_positions"], b["input_lengths"])
        loss = O.training_loss(outs, b["mel"], b["y"], b["done"], b["input_lengths"],
                               b["target_lengths"].numpy(), guided_sigma=extra["guided_attention_sigma"])
        loss.backward()
        torch.nn.utils.clip_grad_norm_(leaves, 0.1)
        opt.step()
        if i >= warmup:
            ts.append(time.perf_counter() - t0)
            if sum(ts) > 60.0:          # bounded sample
                break
    sec = float(np.median(ts))
    return B * T_MEL / sec, threads, sec


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = min(args.steps, 5), min(args.warmup, 1)
    val, cores, sec = cpu_step_throughput(args.preset, steps, warmup)
    sample = "%d full steps (B=16,T_text=128,T_mel=800) after %d warm-up, median" % (steps, warmup)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "mel-frames/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%s training step, B=16, T_text=128, T_mel=800 (T_dec=200)" % args.preset},
        "cpu_baseline": {"value": val, "unit": "mel-frames/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "mel-frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# -------------------------------------------------------------------------------------------------
# GPU arm
# -------------------------------------------------------------------------------------------------
def convblock_roofline(dev, pk, pk_kind):
    """The dominant kernel of the step: the fused ConvBlock forward GEMM at the postnet's widest shape
    (B=16, C=512, T=800, k=3) -- the 2 such blocks (+ their same-shaped dgrad/wgrad) are ~38 % of all conv FLOPs.
    Only the ConvBlock kernel itself is timed (operands prepared outside), CUDA events on the launching stream,
    L2 flushed between launches.  Algorithmic bytes per launch: 4*[2*B*C*T + 2C*C*k + 4C]."""
    from deepvoice3_pytorch_b200 import ops
    Bc, C, T, k, d = 16, 512, 800, 3, 1
    v = torch.randn(2 * C, C, k, device=dev) * (4.0 / (k * C)) ** 0.5
    g = v.pow(2).sum((1, 2), keepdim=True).sqrt()
    bias = torch.zeros(2 * C, device=dev)
    x = torch.randn(Bc, C, T, device=dev)
    y, sa, ss = torch.empty_like(x), torch.empty_like(x), torch.empty_like(x)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    math = ops.conv_math
    if math in ("tc", "bf16x3"):
        bf = torch.bfloat16
        inv, scale = torch.empty(2 * C, device=dev), torch.empty(2 * C, device=dev)
        wfwd = torch.empty(2, k, 2 * C, C, device=dev, dtype=bf)
        wbwd = torch.empty(2, k, C, 2 * C, device=dev, dtype=bf)
        ops.lib.call("dv3_tc_weightnorm_fwd", ops._p(v), ops._p(g), ops._p(inv), ops._p(scale), ops._p(wfwd), 2,
                     ops._p(wbwd), 2 * C, C, k, ops._stream())
        xs = torch.empty(2, Bc, T, C, device=dev, dtype=bf)
        ops.lib.call("dv3_tc_split_input", ops._p(x), ops._p(xs), 2, None, Bc, C, T, k, d, 0, 0.0, None, 0,
                     ops._stream())
        name = "tcgen05 gated ConvBlock forward (persistent tc_conv_kernel<GATED, BK=64>) via dv3_tc_convblock_fwd"

        def launch():
            ops.lib.call("dv3_tc_convblock_fwd", ops._p(xs), ops._p(wfwd), 2, ops._p(bias), None, ops._p(x), ops._p(y),
                         ops._p(sa), ops._p(ss), Bc, C, T, k, d, 0, 0, 1, None, ops._stream())
        mma_passes = 3
    else:
        w_f, w_b, inv = ops._wn_conv_fwd(v, g)
        name = "gemm_simt_kernel<ConvPolicy<gated>> via dv3_convblock_fwd"

        def launch():
            ops.lib.call("dv3_convblock_fwd", ops._p(x), ops._p(w_f), ops._p(bias), None, ops._p(y), ops._p(sa),
                         ops._p(ss), Bc, C, T, k, d, 0, 0, 1, 0.0, None, 0, ops._stream())
        mma_passes = 0
    for _ in range(3):
        launch()
    ts = []
    for _ in range(10):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); launch(); e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e-3)
    t = float(np.mean(ts))
    alg_bytes = 4.0 * (2 * Bc * C * T + 2 * C * C * k + 4 * C)
    flops = 2.0 * Bc * T * 2 * C * C * k                     # algorithmic (fp32) flops of the block forward
    hbm = {"bound": "hbm", "achieved": alg_bytes / t / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
           "frac": alg_bytes / t / 1e9 / pk["hbm_gbs"], "alg_bytes": alg_bytes,
           "note": "BASELINE.json's 'ConvBlock HBM GB/s': algorithmic bytes / launch time; the block is a dense "
                   "contraction (686 FLOP/B here), so this roof does not bind"}
    if mma_passes:
        # the binding roof: tensor cores.  `achieved` counts the ALGORITHMIC flops (one fp32 multiply-add per term);
        # the kernel issues 3 bf16 MMA passes per term (hi*hi, hi*lo, lo*hi) to be fp32-accurate, reported beside it.
        r = {"bound": "tensor", "kernel": "%s (B=16,C=512,T=800,k=3)" % name, "achieved": flops / t / 1e12,
             "peak": pk["bf16_tflops"], "unit": "TFLOP/s", "frac": flops / t / 1e12 / pk["bf16_tflops"],
             "issued_bf16_tflops": mma_passes * flops / t / 1e12,
             "issued_frac": mma_passes * flops / t / 1e12 / pk["bf16_tflops"],
             # dram__bytes_read.sum + dram__bytes_write.sum of this kernel at this shape from the committed
             # `ncu --set full` capture (profiles/r01_ncu_full_tc_convblock_v3.csv): 58.8 MB + 33.6 MB per launch
             # (it reads the bf16 hi/lo planes of the input and writes y + the two saved gate tensors)
             "traffic": 92.39e6, "peak_source": pk_kind, "launch_us": t * 1e6, "alg_flops": flops, "math": math,
             "hbm": hbm}
    else:
        r = dict(hbm, kernel="%s (B=16,C=512,T=800,k=3)" % name, traffic=None, peak_source=pk_kind,
                 launch_us=t * 1e6, math=math,
                 fp32_fma={"achieved": flops / t / 1e12, "peak": 74.5, "unit": "TFLOP/s",
                           "frac": flops / t / 1e12 / 74.5,
                           "note": "nominal fp32 FMA peak 148 SM x 128 lanes x 2 x 1.965 GHz (the binding roof of "
                                   "the exact-fp32 mode)"})
    return r


def run_gpu_arm(args):
    import torch.distributed as dist
    from deepvoice3_pytorch_b200 import builder, ops
    from deepvoice3_pytorch_b200._lib import lib
    from deepvoice3_pytorch_b200.train_step import TrainStep, make_synthetic_batch, to_device

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback; --impl reference is the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus, "launch with torchrun --nproc-per-node %d" % args.gpus

    ops.conv_math = args.math
    bname, kw, extra = PRESETS[args.preset]
    torch.manual_seed(1234)                      # identical initial weights on every rank (as DDP broadcasts)
    model = getattr(builder, bname)(**kw).to(dev)
    step = TrainStep(model, use_graph=not args.no_graph, **extra)
    ops.rng.manual_seed(1234 + rank, dev)
    host = make_synthetic_batch(B, T_TEXT, T_MEL, n_speakers=kw["n_speakers"], seed=1234 + rank, pin=True)
    resident = to_device(host, dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(run_step, steps):
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(steps):
            run_step()
        e.record()
        barrier()
        t = torch.tensor([s.elapsed_time(e) * 1e-3], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident throughput (value) ------------------------------------------------------
    # nvidia-smi needs ~0.5 s to deliver its first sample and a 10-step timed region lasts ~0.1 s, so the sampler
    # runs from the warm-up to the end of the e2e region (the GPU is under the same load throughout).
    clocks = ClockSampler(local).__enter__()
    for _ in range(max(args.warmup, 3)):
        loss = step.step(resident)
    torch.cuda.synchronize()
    ops.check_index_errors()
    l0 = lib.raw("dv3_launch_count")()
    t_res = timed(lambda: step.step(resident), args.steps)
    launches = (lib.raw("dv3_launch_count")() - l0) // args.steps
    if step.launches_per_step is not None:       # graph replay: the launches were recorded at capture time
        launches = step.launches_per_step
    loss_val = float(loss.item())
    assert np.isfinite(loss_val), "training diverged"

    # ---- end to end: pinned host batch -> H2D every step, loss read back every step ------------------
    h2d = sum(v.numel() * v.element_size() for v in host.values() if torch.is_tensor(v))
    sink = []

    # Every step uploads its own batch from pinned host memory and reads the loss back.  The upload of step i+1 is
    # issued on a copy stream before step i is launched (what a pinned-memory DataLoader with non_blocking copies
    # does), so the PCIe transfer overlaps the previous step's compute; both are inside the timed region.
    copy_stream = torch.cuda.Stream(device=dev)

    def upload():
        with torch.cuda.stream(copy_stream):
            b = to_device(host, dev)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return b, ev

    pending = [upload()]

    def e2e_step():
        batch, ev = pending.pop()
        torch.cuda.current_stream().wait_event(ev)
        pending.append(upload())                          # prefetch the next step's batch
        loss_t = step.step(batch)
        for v in batch.values():                          # the buffers were produced on the copy stream
            if torch.is_tensor(v):
                v.record_stream(torch.cuda.current_stream())
        sink.append(float(loss_t.item()))                 # D2H read of the step's result (synchronises)
    for _ in range(2):
        e2e_step()
    t_e2e = timed(e2e_step, args.steps)
    t_extra = time.perf_counter()
    while len(clocks.rows) < 3 and time.perf_counter() - t_extra < 3.0:      # keep the load on until sampled
        step.step(resident)
    torch.cuda.synchronize()
    clocks.__exit__()

    frames = B * T_MEL * world
    out = {
        "metric": METRIC, "value": frames * args.steps / t_res, "unit": "mel-frames/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": t_res / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%s training step (fwd+losses+bwd+clip+Adam), B=16/GPU, T_text=128, T_mel=800 "
                               "(T_dec=200), random-init weights" % args.preset,
                   "global_batch": B * world, "parallelism": "dp%d" % world,
                   "l2": "inputs larger than L2 (>1.5 GB touched per step)",
                   "cuda_graph": not args.no_graph, "conv_math": ops.conv_math,
                   "conv_math_note": {"tc": "tcgen05 bf16 hi/lo split (hi*hi+hi*lo+lo*hi), fp32 accumulate: every block "
                                            "within ~1e-5 of exact fp32, full-depth outputs within 2e-4 (tests)",
                                      "bf16x3": "alias of tc",
                                      "fp32": "exact fp32 FMA on CUDA cores"}[ops.conv_math]},
        "e2e": {"value": frames * args.steps / t_e2e, "unit": "mel-frames/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": 4, "ms_per_step": t_e2e / args.steps * 1e3},
        "gpu_launches": int(launches), "loss": loss_val, "clocks": clocks.summary(),
    }
    if rank == 0:
        pk, pk_kind = peaks()
        out["roofline"] = convblock_roofline(dev, pk, pk_kind)
        if world == 1 and not args.no_cpu_baseline:
            val, cores, sec = cpu_step_throughput(args.preset, steps=3, warmup=1)
            out["cpu_baseline"] = {"value": val, "unit": "mel-frames/s", "cores": cores, "kind": "port",
                                   "sample": "up to 3 full steps (same workload, <=60 s) after 1 warm-up, median; "
                                             "%.2f s/step on %d threads" % (sec, cores)}
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--preset", default="deepvoice3_ljspeech", choices=sorted(PRESETS))
    ap.add_argument("--no-graph", action="store_true", help="run the step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--math", default=os.environ.get("DV3_CONV_MATH", "tc"), choices=["tc", "fp32", "bf16x3"],
                    help="ConvBlock arithmetic: tc = tcgen05 split-bf16 (fp32-equivalent, default), fp32 = CUDA cores")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()

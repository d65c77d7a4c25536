"""One data-parallel training step of the hot path, B200-first.

What the reference does per step (train.py:621-779) -- H2D of the batch, model forward, the four losses,
backward, ``clip_grad_norm_``, ``Adam.step`` -- restated around three ideas:

* **flat arenas**: every trainable parameter is a view into one contiguous fp32 buffer and every gradient a view
  into another, so the optimizer (clip + Adam) is two kernel launches over ~26 M floats and the data-parallel
  exchange is ONE NCCL all-reduce of the gradient arena over NVLink (the path is straight data-parallel over
  utterances: no other collective exists);
* **no host synchronisation inside the step**: the clip coefficient, learning rate and dropout seed live in
  device memory, the loss is returned as a device scalar;
* **graph capture**: because of the above the whole step (forward, loss, backward, optimizer) can be captured
  once in a CUDA graph and replayed, which removes the ~1.5 k kernel-launch / autograd dispatch overhead that
  dominates once the convolutions run on tensor cores.
"""
import ctypes
import math
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.nn.functional as F

from . import ops
from ._lib import lib
from .weight_bank import WeightBank


def noam_learning_rate_decay(init_lr, global_step, warmup_steps=4000):
    """reference lrschedule.py:5-11."""
    warmup_steps = float(warmup_steps)
    step = global_step + 1.0
    return init_lr * warmup_steps ** 0.5 * min(step * warmup_steps ** -1.5, step ** -0.5)


def sequence_mask(lengths, max_len):
    """(B,) int64 device tensor -> (B, max_len) float mask (reference train.py:261-271)."""
    return (torch.arange(max_len, device=lengths.device)[None, :] < lengths[:, None]).float()


def guided_attention_mask(input_lengths, target_lengths, max_target_len, max_input_len, g):
    """W[b,t,n] = 1 - exp(-(n/N_b - t/T_b)^2 / (2 g^2)) inside (T_b, N_b), 0 outside -- computed on the device
    from the two length vectors (the reference builds it with numba on the host and uploads it every step,
    train.py:585-601, 734-738)."""
    dev = input_lengths.device
    N = input_lengths.double()[:, None, None]
    T = target_lengths.double()[:, None, None]
    n = torch.arange(max_input_len, device=dev, dtype=torch.float64)[None, None, :]
    t = torch.arange(max_target_len, device=dev, dtype=torch.float64)[None, :, None]
    W = 1.0 - torch.exp(-(n / N - t / T) ** 2 / (2 * g * g))
    W = W * (n < N) * (t < T)
    return W.float()


def spec_loss(y_hat, y, mask, masked_loss_weight, binary_divergence_weight, eps=1e-8, priority_bin=None,
              priority_w=0.0):
    """reference train.py:547-582 (torch ops; TrainStep(fused_loss=False) -- the default is csrc/loss.cu)."""
    w = masked_loss_weight

    def l1_of(a, b):
        l1 = (a - b).abs().mean()
        if w > 0:
            mask_ = mask.expand_as(a)
            l1 = w * ((a * mask_ - b * mask_).abs().sum() / mask_.sum()) + (1 - w) * l1
        return l1

    l1 = l1_of(y_hat, y)
    if priority_bin is not None and priority_w > 0:
        l1 = (1 - priority_w) * l1 + priority_w * l1_of(y_hat[:, :, :priority_bin], y[:, :, :priority_bin])
    if binary_divergence_weight <= 0:
        return l1, y_hat.new_zeros(())
    logits = torch.log(y_hat + eps) - torch.log(1 - y_hat + eps)
    z = -y * logits + torch.log1p(torch.exp(logits))
    if w > 0:
        mask_ = mask.expand_as(z)
        bd = w * ((z * mask_).sum() / mask_.sum()) + (1 - w) * z.mean()
    else:
        bd = z.mean()
    return l1, bd


def priority_bin_of(priority_freq, sample_rate, linear_dim):
    """reference train.py:722."""
    return int(priority_freq / (sample_rate * 0.5) * linear_dim)


def training_loss(outs, batch, r=1, downsample_step=4, masked_loss_weight=0.5, binary_divergence_weight=0.1,
                  guided_attention_sigma=0.2, use_guided_attention=True, priority_freq=3000, priority_freq_weight=0.0,
                  sample_rate=22050):
    """Total loss of one step with seq2seq and postnet trained jointly (reference train.py:665-740)."""
    mel_out, lin_out, attn, done_hat = outs
    mel, y, done = batch["mel"], batch["y"], batch["done"]
    tl = batch["target_lengths"]
    dec_mask = tgt_mask = None
    if masked_loss_weight > 0:
        dec_mask = sequence_mask(tl // (r * downsample_step), mel.size(1)).unsqueeze(-1)
        tgt_mask = sequence_mask(tl, y.size(1)).unsqueeze(-1) if downsample_step > 1 else dec_mask
        dec_mask, tgt_mask = dec_mask[:, r:, :], tgt_mask[:, r:, :]
    w = binary_divergence_weight
    l1, bd = spec_loss(mel_out[:, :-r, :], mel[:, r:, :], dec_mask, masked_loss_weight, w)
    loss = (1 - w) * l1 + w * bd
    loss = loss + F.binary_cross_entropy(done_hat, done)
    l1, bd = spec_loss(lin_out[:, :-r, :], y[:, r:, :], tgt_mask, masked_loss_weight, w,
                       priority_bin=priority_bin_of(priority_freq, sample_rate, lin_out.size(-1)),
                       priority_w=priority_freq_weight)
    loss = loss + (1 - w) * l1 + w * bd
    if use_guided_attention:
        soft = guided_attention_mask(batch["input_lengths_dev"], tl // r // downsample_step, attn.size(-2),
                                     attn.size(-1), guided_attention_sigma)
        loss = loss + (attn * soft).mean()
    return loss


class _FusedLossFn(torch.autograd.Function):
    """All four training losses and their gradients in 3 kernel launches (csrc/loss.cu)."""

    @staticmethod
    def forward(ctx, mel_out, lin_out, attn, done_hat, mel, y, done, target_lengths, input_lengths, r,
                downsample_step, w, bw, sigma, use_attn, pbin, pw):
        dev = mel_out.device
        vp = lambda t: ctypes.c_void_p(t.data_ptr())
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        mel_out, lin_out, attn, done_hat = [t.contiguous() for t in (mel_out, lin_out, attn, done_hat)]
        loss = torch.zeros(1, device=dev)
        g_mel, g_lin = torch.empty_like(mel_out), torch.empty_like(lin_out)
        g_attn, g_done = torch.empty_like(attn), torch.empty_like(done_hat)
        dec_len = (target_lengths // (r * downsample_step)).contiguous()
        B, Td, Dm = mel_out.shape
        lib.call("dv3_spec_loss", vp(mel_out), vp(mel.contiguous()), vp(dec_len), vp(g_mel), vp(loss), B, Td, Dm, r,
                 float(w), float(bw), 0, 0.0, st)
        _, Tl, Dl = lin_out.shape
        lin_len = target_lengths.contiguous() if downsample_step > 1 else dec_len
        lib.call("dv3_spec_loss", vp(lin_out), vp(y.contiguous()), vp(lin_len), vp(g_lin), vp(loss), B, Tl, Dl, r,
                 float(w), float(bw), int(pbin), float(pw), st)
        A, _, _, Ts = attn.shape
        dec_len_attn = (target_lengths // r // downsample_step).contiguous()
        lib.call("dv3_aux_loss", vp(done_hat), vp(done.contiguous()), vp(g_done), done_hat.numel(), vp(attn),
                 vp(g_attn), vp(input_lengths.contiguous()), vp(dec_len_attn), A, B, attn.shape[2], Ts, float(sigma),
                 int(use_attn), vp(loss), st)
        ctx.save_for_backward(g_mel, g_lin, g_attn, g_done)
        return loss[0]

    @staticmethod
    def backward(ctx, gout):
        g_mel, g_lin, g_attn, g_done = ctx.saved_tensors
        return (g_mel * gout, g_lin * gout, g_attn * gout, g_done * gout) + (None,) * 13


def fused_training_loss(outs, batch, r=1, downsample_step=4, masked_loss_weight=0.5, binary_divergence_weight=0.1,
                        guided_attention_sigma=0.2, use_guided_attention=True, priority_freq=3000,
                        priority_freq_weight=0.0, sample_rate=22050):
    """Same value and gradients as ``training_loss`` (reference train.py:665-740), computed by csrc/loss.cu."""
    mel_out, lin_out, attn, done_hat = outs
    return _FusedLossFn.apply(mel_out, lin_out, attn, done_hat, batch["mel"], batch["y"], batch["done"],
                              batch["target_lengths"], batch["input_lengths_dev"], r, downsample_step,
                              masked_loss_weight, binary_divergence_weight, guided_attention_sigma,
                              use_guided_attention, priority_bin_of(priority_freq, sample_rate, lin_out.size(-1)),
                              priority_freq_weight)


class ParameterArena:
    """Re-homes the trainable parameters of ``model`` into one flat fp32 buffer (and their gradients into
    another).  ``state_dict`` / ``load_state_dict`` keep working: parameters stay nn.Parameters, only their
    storage moves."""

    def __init__(self, model, params=None):
        params = list(model.get_trainable_parameters()) if params is None else list(params)
        assert all(p.dtype == torch.float32 for p in params), "fp32 parameters only"
        self.params = params
        offs, n = [], 0
        for p in params:
            offs.append(n)
            n += (p.numel() + 3) // 4 * 4            # keep every view 16-byte aligned
        self.numel = n
        dev = params[0].device
        self.flat = torch.zeros(n, device=dev)
        self.grad = torch.zeros(n, device=dev)
        for p, o in zip(params, offs):
            self.flat[o:o + p.numel()].view_as(p).copy_(p.data)
            p.data = self.flat[o:o + p.numel()].view_as(p)
            p.grad = self.grad[o:o + p.numel()].view_as(p)
        self.offsets = offs

    def zero_grad(self):
        self.grad.zero_()

    def broadcast(self, model=None, src=0):
        """Replica consistency at start-up (what DistributedDataParallel does at construction): every rank takes
        rank ``src``'s parameter arena, plus -- when ``model`` is given -- the parameters outside the arena (frozen
        position tables / embeddings) and the buffers."""
        if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
            return
        dist.broadcast(self.flat, src)
        if model is not None:
            inside = {id(p) for p in self.params}
            for t in list(model.parameters()) + list(model.buffers()):
                if id(t) not in inside:
                    dist.broadcast(t.data, src)

    def all_reduce_grads(self, ranges=None):
        """The exchange step of the data-parallel path: sum the flat gradient arena (or the given [lo, hi) slices of
        it) over all ranks (NCCL over NVLink on GPUs; gloo in the CPU tests).  The 1/world average is applied by the
        optimizer (hyper[3])."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            if ranges is None:
                dist.all_reduce(self.grad)
            else:
                for lo, hi in ranges:
                    if hi > lo:
                        dist.all_reduce(self.grad[lo:hi])

    def range_of(self, params):
        """[lo, hi) of the arena slice holding ``params`` (must be consecutive arena entries), or None if empty."""
        ids = {id(p) for p in params}
        idx = [i for i, p in enumerate(self.params) if id(p) in ids]
        if not idx:
            return None
        assert idx == list(range(idx[0], idx[-1] + 1)), "bucket parameters are not contiguous in the arena"
        last = idx[-1]
        return self.offsets[idx[0]], self.offsets[last] + (self.params[last].numel() + 3) // 4 * 4


def gradient_buckets(model, arena):
    """Bucket plan of the overlapped gradient exchange.  The backward pass finishes the post-net first, then the
    decoder, then the encoder from its last layer to its first (the encoder holds ~60 % of the parameters), so:
      "postnet"     everything under model.postnet        -- final when d(loss)/d(postnet input) exists
      "encoder_hi" / "encoder_mid"  encoder.convolutions[hi:] / [mid:hi]  -- final when the gradient entering layer
                    hi / mid exists
      "rest"        all other slices (decoder, first encoder layers, embeddings): reduced after backward()
    -> ({tag: (lo, hi)}, [rest ranges])."""
    tagged = {}
    post = arena.range_of(list(model.postnet.parameters())) if hasattr(model, "postnet") else None
    if post:
        tagged["postnet"] = post
    enc = getattr(getattr(model, "seq2seq", None), "encoder", None)
    if enc is not None and hasattr(enc, "grad_bucket_splits") and hasattr(enc, "convolutions"):
        layers = list(enc.convolutions)
        splits = sorted(enc.grad_bucket_splits(), key=lambda ti: ti[1])
        for (tag, lo_i), hi_i in zip(splits, [i for _, i in splits[1:]] + [len(layers)]):
            r = arena.range_of([p for m in layers[lo_i:hi_i] for p in m.parameters()])
            if r:
                tagged[tag] = r
    cuts = sorted(tagged.values())
    rest, pos = [], 0
    for lo, hi in cuts:
        if lo > pos:
            rest.append((pos, lo))
        pos = max(pos, hi)
    if pos < arena.numel:
        rest.append((pos, arena.numel))
    return tagged, rest


class FlatAdam:
    """torch.optim.Adam(betas, eps) + clip_grad_norm_(clip) over a ParameterArena in two launches."""

    def __init__(self, arena, lr=5e-4, betas=(0.5, 0.9), eps=1e-6, clip_thresh=0.1):
        self.arena = arena
        self.lr, self.betas, self.eps, self.clip = lr, betas, eps, clip_thresh
        dev = arena.flat.device
        self.m = torch.zeros_like(arena.flat)
        self.v = torch.zeros_like(arena.flat)
        self.hyper = torch.zeros(4, device=dev)
        # ring of pinned staging slots: the host may run several steps ahead of the stream, so a slot is only
        # rewritten after the copy that last read it has completed (event wait, normally already signalled)
        self._slots = [torch.zeros(4).pin_memory() if dev.type == "cuda" else torch.zeros(4) for _ in range(4)]
        self._events = [None] * 4
        self.sumsq = torch.zeros(1, device=dev)
        self._scratch = torch.zeros(lib.raw("dv3_sumsq_scratch_floats")(), device=dev) if dev.type == "cuda" else None
        self.t = 0

    def set_hyper(self, lr, grad_scale=1.0):
        """Host-side scalar prep; the async H2D copy of 16 bytes is the only thing the stream sees."""
        self.t += 1
        b1, b2 = self.betas
        i = self.t % 4
        if self._events[i] is not None:
            self._events[i].synchronize()
        h = self._slots[i]
        h[0], h[1], h[2], h[3] = lr, 1.0 - b1 ** self.t, 1.0 - b2 ** self.t, grad_scale
        self.hyper.copy_(h, non_blocking=True)
        self._events[i] = torch.cuda.Event()
        self._events[i].record()

    def apply(self):
        """Device-only part (graph-capturable): grad norm -> clip -> Adam."""
        a = self.arena
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        lib.call("dv3_sumsq", ctypes.c_void_p(a.grad.data_ptr()), a.numel, ctypes.c_void_p(self.sumsq.data_ptr()),
                 ctypes.c_void_p(self._scratch.data_ptr()), st)
        lib.call("dv3_adam_clip", ctypes.c_void_p(a.flat.data_ptr()), ctypes.c_void_p(a.grad.data_ptr()),
                 ctypes.c_void_p(self.m.data_ptr()), ctypes.c_void_p(self.v.data_ptr()), a.numel,
                 ctypes.c_void_p(self.hyper.data_ptr()), ctypes.c_void_p(self.sumsq.data_ptr()),
                 self.betas[0], self.betas[1], self.eps, float(self.clip), st)

    def grad_norm(self):
        return self.sumsq.sqrt() * self.hyper[3]

    # -- checkpoint format of torch.optim.Adam (what reference train.py:save_checkpoint stores under "optimizer",
    #    train.py:787-810, and load_checkpoint restores, :843-860): parameter i of get_trainable_parameters() <->
    #    state[i] = {step, exp_avg, exp_avg_sq}
    def state_dict(self):
        a = self.arena
        state = {}
        for i, (p, o) in enumerate(zip(a.params, a.offsets)):
            n = p.numel()
            state[i] = {"step": torch.tensor(float(self.t)),
                        "exp_avg": self.m[o:o + n].view_as(p).clone(),
                        "exp_avg_sq": self.v[o:o + n].view_as(p).clone()}
        group = {"lr": self.lr, "betas": tuple(self.betas), "eps": self.eps, "weight_decay": 0, "amsgrad": False,
                 "maximize": False, "foreach": None, "capturable": False, "differentiable": False, "fused": None,
                 "decoupled_weight_decay": False, "params": list(range(len(a.params)))}
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd):
        """Accepts a torch.optim.Adam state_dict over the same parameter list (e.g. from a reference checkpoint)."""
        a = self.arena
        groups = sd["param_groups"]
        ids = [i for g in groups for i in g["params"]]
        if len(ids) != len(a.params):
            raise ValueError("optimizer state has %d parameters, the model %d" % (len(ids), len(a.params)))
        g0 = groups[0]
        if g0.get("weight_decay", 0) != 0 or g0.get("amsgrad", False):
            raise ValueError("FlatAdam implements plain Adam: weight_decay / amsgrad states cannot be loaded")
        self.lr, self.betas, self.eps = g0["lr"], tuple(g0["betas"]), g0["eps"]
        steps = set()
        self.m.zero_()
        self.v.zero_()
        for i, p, o in zip(ids, a.params, a.offsets):
            st = sd["state"].get(i)
            if st is None:                      # parameter that never received a gradient
                continue
            n = p.numel()
            if tuple(st["exp_avg"].shape) != tuple(p.shape):
                raise ValueError("optimizer state %d has shape %s, parameter %s" % (i, tuple(st["exp_avg"].shape),
                                                                                   tuple(p.shape)))
            self.m[o:o + n].view_as(p).copy_(st["exp_avg"])
            self.v[o:o + n].view_as(p).copy_(st["exp_avg_sq"])
            steps.add(int(st["step"]))
        if len(steps) > 1:
            raise ValueError("per-parameter step counts differ (%s): not representable in a flat Adam" % sorted(steps))
        self.t = steps.pop() if steps else 0


class TrainStep:
    """model + losses + flat optimizer (+ NCCL gradient all-reduce when torch.distributed is initialised)."""

    def __init__(self, model, init_lr=5e-4, betas=(0.5, 0.9), eps=1e-6, clip_thresh=0.1, r=1, downsample_step=4,
                 masked_loss_weight=0.5, binary_divergence_weight=0.1, guided_attention_sigma=0.2,
                 use_guided_attention=True, priority_freq=3000, priority_freq_weight=0.0, sample_rate=22050,
                 lr_schedule=noam_learning_rate_decay, use_graph=False, fused_loss=True, weight_bank=None):
        self.model = model
        self.arena = ParameterArena(model)
        self.opt = FlatAdam(self.arena, init_lr, betas, eps, clip_thresh)
        self.init_lr, self.lr_schedule = init_lr, lr_schedule
        self.loss_kw = dict(r=r, downsample_step=downsample_step, masked_loss_weight=masked_loss_weight,
                            binary_divergence_weight=binary_divergence_weight,
                            guided_attention_sigma=guided_attention_sigma,
                            use_guided_attention=use_guided_attention, priority_freq=priority_freq,
                            priority_freq_weight=priority_freq_weight, sample_rate=sample_rate)
        self.loss_fn = fused_training_loss if fused_loss else training_loss
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.global_step = 0
        self.use_graph = use_graph
        self._graph = None
        self._static = None
        self._loss = None
        self.launches_per_step = None       # dv3 kernel launches inside one captured step (graph mode)
        if weight_bank is None:
            weight_bank = os.environ.get("DV3_WEIGHT_BANK", "1") == "1"
        self.bank = WeightBank() if weight_bank else None
        self.arena.broadcast(model)             # replicas start from rank 0's weights (no-op for world == 1)
        # overlapped gradient exchange (world > 1): buckets are all-reduced on a communication stream as soon as the
        # backward pass has finished them; only the last ("rest") bucket is exposed
        self.buckets, self.rest_ranges = gradient_buckets(model, self.arena)
        self.overlap_comm = self.world > 1 and os.environ.get("DV3_OVERLAP_COMM", "1") == "1" and \
            self.arena.flat.is_cuda
        self._comm = torch.cuda.Stream(device=self.arena.flat.device) if self.overlap_comm else None
        self._reduced = set()
        # capture the NCCL collectives inside the step's CUDA graph (so clip + Adam stay in the graph too)
        self.graph_comm = self.overlap_comm and os.environ.get("DV3_GRAPH_COMM", "1") == "1"

    # -- checkpointing: the reference's checkpoint keys (train.py:787-810) ---------------------------------
    def state_dict(self, global_epoch=0):
        return {"state_dict": self.model.state_dict(), "optimizer": self.opt.state_dict(),
                "global_step": self.global_step, "global_epoch": global_epoch}

    def load_state_dict(self, ckpt, load_optimizer=True):
        """Resume from ``state_dict()`` or from a reference checkpoint (same keys).  Restores the Adam moments, the
        bias-correction step and the position in the learning-rate schedule."""
        self.model.load_state_dict(ckpt["state_dict"])      # copies into the arena views in place
        if load_optimizer and ckpt.get("optimizer") is not None:
            self.opt.load_state_dict(ckpt["optimizer"])
        self.global_step = int(ckpt.get("global_step", 0))
        return int(ckpt.get("global_epoch", 0))

    # -- pieces -------------------------------------------------------------------------------
    def _forward_backward(self, batch):
        self.arena.zero_grad()
        ops.grad_sink = True          # kernels accumulate parameter gradients straight into the arena
        ops.weight_bank = self.bank   # weight norm of all layers: 2 launches up front, 1 per bucket in the backward
        ops.grad_boundary_cb = self._bucket_ready if self.overlap_comm else None
        self._reduced = set()
        try:
            if self.bank is not None:
                self.bank.begin_step()
            return self._forward_backward_inner(batch)
        finally:
            ops.grad_sink = False
            ops.weight_bank = None
            ops.grad_boundary_cb = None
            if self.bank is not None:
                self.bank.end_step()

    def _bucket_ready(self, tag):
        """Called from the backward pass (ops.grad_boundary hook): the parameter gradients of bucket ``tag`` are final
        -- finish their weight-norm backward and start their all-reduce on the communication stream."""
        rng = self.buckets.get(tag)
        if rng is None or tag in self._reduced:
            return
        self._reduced.add(tag)
        if self.bank is not None:
            self.bank.end_backward()
        main = torch.cuda.current_stream()
        self._comm.wait_stream(main)
        with torch.cuda.stream(self._comm):
            self.arena.all_reduce_grads([rng])

    def _forward_backward_inner(self, batch):
        outs = self.model(batch["x"], batch["mel"], speaker_ids=batch.get("speaker_ids"),
                          text_positions=batch["text_positions"], frame_positions=batch["frame_positions"],
                          input_lengths=batch["input_lengths_dev"])
        loss = self.loss_fn(outs, batch, **self.loss_kw)
        loss.backward()
        if self.bank is not None:
            self.bank.end_backward()
        return loss.detach()

    def _exchange_and_update(self):
        """Gradient exchange (sum; the 1/world average is folded into hyper[3]) + clip + Adam."""
        if self.overlap_comm:
            main = torch.cuda.current_stream()
            self._comm.wait_stream(main)
            with torch.cuda.stream(self._comm):         # buckets whose boundary never fired + everything else
                pending = [r for t, r in self.buckets.items() if t not in self._reduced]
                self.arena.all_reduce_grads(pending + self.rest_ranges)
            main.wait_stream(self._comm)
        else:
            self.arena.all_reduce_grads()
        self.opt.apply()          # (fresh dropout masks per step: the model's forward draws a new seed itself)

    # -- public -------------------------------------------------------------------------------
    def step(self, batch):
        """batch: dict of DEVICE tensors (x, text_positions, frame_positions int64; mel, y, done fp32;
        target_lengths, input_lengths_dev int64) + host numpy ``input_lengths``.  Returns the loss (device)."""
        self.model.train()
        lr = self.lr_schedule(self.init_lr, self.global_step) if self.lr_schedule else self.init_lr
        self.opt.set_hyper(lr, 1.0 / self.world)
        if not self.use_graph:
            loss = self._forward_backward(batch)
            self._exchange_and_update()
        else:
            loss = self._graph_step(batch)
        self.global_step += 1
        return loss

    def _graph_step(self, batch):
        if self._graph is None:
            # static input buffers; warm up on a side stream, then capture forward+loss+backward (+update when
            # there is no collective to run in between)
            self._static = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in batch.items()}
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                dev = self.arena.flat.device
                ops.rng.seed_tensor(dev)
                seed0 = ops.rng.base.clone()
                for _ in range(2):
                    self._forward_backward(self._static)
                    if self.world > 1 and self.overlap_comm:    # NCCL communicators / bucket tables exist before capture
                        pend = [r for t, r in self.buckets.items() if t not in self._reduced]
                        with torch.cuda.stream(self._comm):
                            self._comm.wait_stream(s)
                            self.arena.all_reduce_grads(pend + self.rest_ranges)
                        s.wait_stream(self._comm)
                ops.rng.base.copy_(seed0)           # the warm-up passes do not consume dropout seeds
            torch.cuda.current_stream().wait_stream(s)
            self._graph = torch.cuda.CUDAGraph()
            n0 = lib.raw("dv3_launch_count")()
            capture_all = self.world == 1 or self.graph_comm
            with torch.cuda.graph(self._graph):
                self._loss = self._forward_backward(self._static)
                if capture_all:                     # NCCL all-reduces are captured as graph nodes on the comm stream
                    self._exchange_and_update()
            self.launches_per_step = int(lib.raw("dv3_launch_count")() - n0) + (0 if capture_all else 2)
        for k, v in batch.items():
            if torch.is_tensor(v):
                self._static[k].copy_(v, non_blocking=True)
        self._graph.replay()
        if self.world > 1 and not self.graph_comm:
            self._exchange_and_update()
        return self._loss


def make_synthetic_batch(B=16, T_text=128, T_mel=800, downsample_step=4, r=1, n_speakers=1, n_vocab=149,
                         mel_dim=80, linear_dim=513, seed=1234, pin=False):
    """Synthetic batch of SURVEY.md section 8(d), on the HOST (what collate_fn would hand to the step)."""
    gen = torch.Generator().manual_seed(seed)
    T_dec = T_mel // downsample_step // r
    b = {
        "x": torch.randint(2, n_vocab, (B, T_text), generator=gen),
        "text_positions": torch.arange(1, T_text + 1)[None, :].repeat(B, 1),
        "frame_positions": torch.arange(1, T_dec + 1)[None, :].repeat(B, 1),
        "mel": torch.rand(B, T_dec, mel_dim * r, generator=gen),
        "y": torch.rand(B, T_mel, linear_dim, generator=gen),
        "done": torch.cat([torch.zeros(B, T_dec - 1, 1), torch.ones(B, 1, 1)], dim=1),
        "target_lengths": torch.full((B,), T_mel, dtype=torch.int64),
        "input_lengths_dev": torch.full((B,), T_text, dtype=torch.int64),
    }
    if n_speakers > 1:
        b["speaker_ids"] = torch.randint(0, n_speakers, (B,), generator=gen)
    if pin:
        b = {k: v.pin_memory() for k, v in b.items()}
    b["input_lengths"] = np.full(B, T_text, dtype=np.int64)
    return b


def to_device(batch, device, non_blocking=True):
    return {k: (v.to(device, non_blocking=non_blocking) if torch.is_tensor(v) else v) for k, v in batch.items()}

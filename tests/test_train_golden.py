"""Loss / batching code pinned to the reference's own ``train.py``: tests/golden/train_fns.npz was produced by
EXECUTING that file (tests/golden/make_train_golden.py).  Here (CPU): the oracle restatement, the torch fallback of
``train_step`` and ``data.collate`` against those vectors.  The CUDA loss kernels are checked against the same vectors
in tests/test_gpu_train.py."""
import numpy as np
import pytest
import torch

import golden_util as G

TRAIN = G.load("train_fns.npz")


def _cfg(case):
    return [float(v) for v in case["meta"]["cfg"]]


@pytest.mark.parametrize("name", ["mask0", "mask1"])
def test_sequence_mask(name):
    from oracle import dv3_oracle as O
    from deepvoice3_pytorch_b200.train_step import sequence_mask
    case = TRAIN[name]
    lengths = torch.from_numpy(case["in"]["lengths"])
    want = case["out"]["0"]
    np.testing.assert_array_equal(O.sequence_mask(lengths, want.shape[1]).numpy(), want)
    np.testing.assert_array_equal(sequence_mask(lengths, want.shape[1]).numpy(), want)


@pytest.mark.parametrize("name", [n for n in TRAIN if n.startswith("specloss")])
def test_spec_loss_oracle_and_torch_path(name):
    from oracle import dv3_oracle as O
    from deepvoice3_pytorch_b200 import train_step as TS
    case = TRAIN[name]
    w, bw, pbin, pw = _cfg(case)
    pbin = None if pbin < 0 else int(pbin)
    y = torch.from_numpy(case["in"]["y"])
    lens = torch.from_numpy(case["in"]["lengths"])
    for impl in (O, TS):
        y_hat = torch.from_numpy(case["in"]["y_hat"]).clone().requires_grad_(True)
        mask = impl.sequence_mask(lens, y.size(1)).unsqueeze(-1) if w > 0 else None
        l1, bd = impl.spec_loss(y_hat, y, mask, w, bw, priority_bin=pbin, priority_w=pw)
        ((1 - bw) * l1 + bw * bd.sum()).backward()
        np.testing.assert_allclose(float(l1), float(case["out"]["l1"]), rtol=1e-6)
        np.testing.assert_allclose(float(bd.sum()), float(case["out"]["bd"]), rtol=1e-6, atol=1e-12)
        np.testing.assert_allclose(y_hat.grad.numpy(), case["out"]["grad"], rtol=1e-5, atol=1e-9)


@pytest.mark.parametrize("name", ["guided0", "guided1"])
def test_guided_attention_mask(name):
    from oracle import dv3_oracle as O
    from deepvoice3_pytorch_b200.train_step import guided_attention_mask
    case = TRAIN[name]
    il, tl, g = case["in"]["input_lengths"], case["in"]["target_lengths"], float(case["meta"]["g"])
    want = case["out"]["0"]                       # (B, max_target_len, max_input_len)
    got = O.guided_attentions(il, tl, want.shape[1], want.shape[2], g)
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-7)
    got = guided_attention_mask(torch.from_numpy(il), torch.from_numpy(tl), want.shape[1], want.shape[2], g)
    np.testing.assert_allclose(got.numpy(), want, rtol=0, atol=1e-7)


@pytest.mark.parametrize("name", [n for n in TRAIN if n.startswith("collate")])
def test_collate_matches_reference_collate_fn(name):
    """data.collate == train.py:293-360 + the mel[:, ::downsample_step] slice of the train loop (:639-640)."""
    from deepvoice3_pytorch_b200.data import collate
    case = TRAIN[name]
    r, ds, nspk = [int(v) for v in case["meta"]["cfg"]]
    n = len([k for k in case["in"] if k.startswith("text")])
    utts = []
    for i in range(n):
        u = (case["in"]["text%d" % i], case["in"]["mel%d" % i], case["in"]["lin%d" % i])
        if nspk > 1:
            u = u + (int(case["in"]["spk%d" % i]),)
        utts.append(u)
    out = collate(utts, r=r, downsample_step=ds)
    ref = case["out"]
    np.testing.assert_array_equal(out["x"].numpy(), ref["x"])
    np.testing.assert_array_equal(out["input_lengths"], ref["input_lengths"])
    np.testing.assert_array_equal(out["input_lengths_dev"].numpy(), ref["input_lengths"])
    np.testing.assert_array_equal(out["mel"].numpy(), ref["mel"][:, 0::ds, :] if ds > 1 else ref["mel"])
    np.testing.assert_array_equal(out["y"].numpy(), ref["y"])
    np.testing.assert_array_equal(out["text_positions"].numpy(), ref["text_positions"])
    np.testing.assert_array_equal(out["frame_positions"].numpy(), ref["frame_positions"])
    np.testing.assert_array_equal(out["done"].numpy(), ref["done"])
    np.testing.assert_array_equal(out["target_lengths"].numpy(), ref["target_lengths"])
    if nspk > 1:
        np.testing.assert_array_equal(out["speaker_ids"].numpy(), ref["speaker_ids"])
    else:
        assert "speaker_ids" not in out
    assert out["x"].dtype == torch.int64 and out["mel"].dtype == torch.float32


def step_case_inputs(case, device="cpu"):
    """-> (outs leaves, batch dict for train_step.*training_loss, kwargs) of a ``step*`` fixture."""
    w, bw, pw, guided, pfreq, sr, sigma, r, ds = _cfg(case)
    r, ds = int(r), int(ds)
    t = {k: torch.from_numpy(np.ascontiguousarray(v)).to(device) for k, v in case["in"].items()}
    outs = [t[k].clone().requires_grad_(True) for k in ("mel_out", "lin_out", "attn", "done_hat")]
    mel = t["mel"][:, 0::ds, :].contiguous() if ds > 1 else t["mel"]            # train.py:639-640
    batch = {"mel": mel, "y": t["y"], "done": t["done"], "target_lengths": t["target_lengths"],
             "input_lengths_dev": t["input_lengths"]}
    kw = dict(r=r, downsample_step=ds, masked_loss_weight=w, binary_divergence_weight=bw,
              guided_attention_sigma=sigma, use_guided_attention=bool(guided), priority_freq=pfreq,
              priority_freq_weight=pw, sample_rate=sr)
    return outs, batch, kw


@pytest.mark.parametrize("name", [n for n in TRAIN if n.startswith("step")])
def test_train_loop_loss_oracle_and_torch_path(name):
    """Total loss and d(loss)/d(model outputs) of the reference's train() (one step on fixed outputs)."""
    from oracle import dv3_oracle as O
    from deepvoice3_pytorch_b200 import train_step as TS
    case = TRAIN[name]
    outs, batch, kw = step_case_inputs(case)
    loss = TS.training_loss(outs, batch, **kw)
    loss.backward()
    np.testing.assert_allclose(float(loss), float(case["out"]["loss"]), rtol=2e-6)
    for o, k in zip(outs, ("mel_out", "lin_out", "attn", "done_hat")):
        g = o.grad if o.grad is not None else torch.zeros_like(o)
        np.testing.assert_allclose(g.numpy(), case["out"]["grad_" + k], rtol=1e-4, atol=1e-9, err_msg=k)
    outs2, _, _ = step_case_inputs(case)
    loss2 = O.training_loss(outs2, batch["mel"], batch["y"], batch["done"], case["in"]["input_lengths"],
                            case["in"]["target_lengths"], r=kw["r"], downsample_step=kw["downsample_step"],
                            masked_loss_weight=kw["masked_loss_weight"],
                            binary_divergence_weight=kw["binary_divergence_weight"],
                            guided_sigma=kw["guided_attention_sigma"], use_guided_attention=kw["use_guided_attention"],
                            priority_freq=kw["priority_freq"], priority_freq_weight=kw["priority_freq_weight"],
                            sample_rate=kw["sample_rate"])
    np.testing.assert_allclose(float(loss2), float(case["out"]["loss"]), rtol=2e-6)

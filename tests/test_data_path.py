"""CPU: the batching rules (reference train.py:293-360) and the distributed length-bucketed sampler."""
import numpy as np
import torch


def _utt(n_text, n_frames, rng, speaker=None):
    u = (rng.randint(2, 149, n_text), rng.rand(n_frames, 80).astype(np.float32),
         rng.rand(n_frames, 513).astype(np.float32))
    return u + (speaker,) if speaker is not None else u


def test_collate_padding_rules():
    from deepvoice3_pytorch_b200.data import collate
    rng = np.random.RandomState(0)
    batch = [_utt(7, 41, rng), _utt(12, 64, rng), _utt(3, 10, rng)]
    r, ds = 1, 4
    out = collate(batch, r=r, downsample_step=ds)
    # max target 64 -> multiple of 4 already -> + r*ds leading frames = 68 ; T_dec = 17
    assert out["y"].shape == (3, 68, 513) and out["mel"].shape == (3, 17, 80)
    assert out["x"].shape == (3, 12) and out["x"].dtype == torch.int64
    assert (out["y"][:, :r].abs().sum() == 0)                               # r leading "initial state" zero frames
    np.testing.assert_array_equal(out["y"][0, r:r + 41].numpy(), batch[0][2])
    assert out["y"][0, r + 41:].abs().sum() == 0
    assert out["text_positions"][0].tolist() == [1, 2, 3, 4, 5, 6, 7, 0, 0, 0, 0, 0]
    assert out["frame_positions"][2].tolist() == list(range(1, 18))
    # done: zeros for len//r//ds - 1 steps, then ones
    for i, n in enumerate([41, 64, 10]):
        k = n // r // ds - 1
        assert out["done"][i, :k, 0].sum() == 0 and out["done"][i, k:, 0].min() == 1
    assert out["target_lengths"].tolist() == [41, 64, 10] and out["input_lengths"].tolist() == [7, 12, 3]
    # mel is the linear-rate mel taken every ds-th frame (train.py:639-640)
    full = np.pad(batch[1][1], [(r, 68 - 64 - r), (0, 0)])
    np.testing.assert_array_equal(out["mel"][1].numpy(), full[0::ds])


def test_collate_multispeaker_and_rounding():
    from deepvoice3_pytorch_b200.data import collate
    rng = np.random.RandomState(1)
    out = collate([_utt(5, 13, rng, speaker=3), _utt(6, 9, rng, speaker=7)], r=2, downsample_step=1)
    # 13 -> 14 (multiple of r=2) -> + r*ds = 16
    assert out["y"].shape[1] == 16 and out["mel"].shape[1] == 16
    assert out["speaker_ids"].tolist() == [3, 7]
    assert out["frame_positions"].shape == (2, 8)


def test_distributed_sampler_partitions_batches():
    from deepvoice3_pytorch_b200.data import DistributedSimilarLengthSampler
    rng = np.random.RandomState(2)
    lengths = rng.randint(50, 900, 1000)
    world, bs = 4, 16
    per_rank = []
    for rank in range(world):
        s = DistributedSimilarLengthSampler(lengths, batch_size=bs, batch_group_size=64, rank=rank,
                                            world_size=world, seed=5)
        s.set_epoch(3)
        idx = list(iter(s))
        assert len(idx) == len(s) and len(idx) % bs == 0
        per_rank.append(idx)
    flat = sum(per_rank, [])
    assert len(set(flat)) == len(flat)                      # disjoint across ranks, no repeats
    assert len(per_rank[0]) == len(per_rank[3])             # same number of steps on every rank
    # batches hold utterances of similar length (bucketing): spread inside a batch << spread of the corpus
    spreads = [np.ptp(lengths[per_rank[0][i:i + bs]]) for i in range(0, len(per_rank[0]), bs)]
    assert np.median(spreads) < 0.25 * np.ptp(lengths)
    # a different epoch reshuffles
    s = DistributedSimilarLengthSampler(lengths, batch_size=bs, batch_group_size=64, rank=0, world_size=world, seed=5)
    s.set_epoch(4)
    assert list(iter(s)) != per_rank[0]


def test_train_txt_dataset_reads_the_preprocess_format(tmp_path):
    """The on-disk format of the reference's preprocess.py (train.txt + .npy pairs, preprocess.py:27-30 /
    ljspeech.py:72-76) through TrainTxtDataset -> sampler -> collate, single- and multi-speaker, with the speaker
    filter of the reference data sources (train.py:101-122)."""
    import numpy as np
    import torch
    from deepvoice3_pytorch_b200.data import TrainTxtDataset, DistributedSimilarLengthSampler, collate
    rng = np.random.RandomState(0)
    lines, specs = [], {}
    for i in range(12):
        n = int(rng.randint(9, 40))
        spec = rng.rand(n, 33).astype(np.float32)
        mel = rng.rand(n, 80).astype(np.float32)
        np.save(tmp_path / ("spec-%05d.npy" % i), spec)
        np.save(tmp_path / ("mel-%05d.npy" % i), mel)
        text = "utt %d %s" % (i, "a" * int(rng.randint(1, 9)))
        lines.append("spec-%05d.npy|mel-%05d.npy|%d|%s|%d" % (i, i, n, text, i % 3))
        specs[i] = (spec, mel, text)
    (tmp_path / "train.txt").write_text("\n".join(lines) + "\n", encoding="utf-8")

    def text_to_sequence(t):
        return [ord(c) % 100 + 2 for c in t] + [1]

    ds = TrainTxtDataset(str(tmp_path), text_to_sequence)
    assert len(ds) == 12 and ds.multi_speaker and ds.frame_lengths == [specs[i][0].shape[0] for i in range(12)]
    seq, mel, lin, spk = ds[5]
    assert spk == 2 and seq.dtype == np.int32 and list(seq) == text_to_sequence(specs[5][2])
    np.testing.assert_array_equal(mel, specs[5][1])
    np.testing.assert_array_equal(lin, specs[5][0])
    one = TrainTxtDataset(str(tmp_path), text_to_sequence, speaker_id=1)
    assert len(one) == 4 and not one.multi_speaker and len(one[0]) == 3
    sampler = DistributedSimilarLengthSampler(ds.frame_lengths, batch_size=4, rank=0, world_size=1, seed=3)
    loader = torch.utils.data.DataLoader(ds, batch_size=4, sampler=sampler, drop_last=True,
                                         collate_fn=lambda b: collate(b, r=1, downsample_step=4))
    seen = 0
    for batch in loader:
        assert batch["x"].shape[0] == 4 and batch["y"].shape[-1] == 33 and batch["mel"].shape[-1] == 80
        assert batch["y"].shape[1] == 4 * batch["mel"].shape[1] and "speaker_ids" in batch
        seen += 4
    assert seen == 12

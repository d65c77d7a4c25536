"""Host side of the data path feeding ``TrainStep``: batching with the reference's
padding rules, and a distributed variant of its length-bucketed sampler.

* ``collate`` restates reference train.py:293-360 (``collate_fn``) plus the per-step slicing the train loop applies
  (mel[:, 0::downsample_step], train.py:639-640) and returns the dict ``TrainStep.step`` consumes (optionally in
  pinned memory so the H2D copies are asynchronous).
* ``TrainTxtDataset`` reads the on-disk format ``preprocess.py`` writes (``train.txt`` with one
  ``spec.npy|mel.npy|n_frames|text[|speaker_id]`` line per utterance, preprocess.py:27-30; the three reference data
  sources TextDataSource / MelSpecDataSource / LinearSpecDataSource + PyTorchDataset, train.py:96-257, in one class).
  The text frontend (string -> token ids) stays the caller's: pass the reference's ``frontend.text_to_sequence``.
* ``DistributedSimilarLengthSampler`` restates ``PartialyRandomizedSimilarTimeLengthSampler`` (train.py:195-239):
  sort by length, shuffle inside groups of ``batch_group_size``, permute whole mini-batches -- then deals the
  mini-batches round-robin to the ranks, so every rank sees disjoint batches of similar length (what the
  data-parallel step needs: one utterance batch per GPU, no collective on the data path).
"""
import numpy as np
import torch


def _pad(seq, max_len, constant_values=0):
    return np.pad(seq, (0, max_len - len(seq)), mode="constant", constant_values=constant_values)


def _pad_2d(x, max_len, b_pad=0):
    return np.pad(x, [(b_pad, max_len - len(x) - b_pad), (0, 0)], mode="constant", constant_values=0)


def collate(batch, r=1, downsample_step=4, pin=False):
    """batch: list of (text_ids int array, mel (T, num_mels) float32, linear (T, n_freq) float32[, speaker_id]).

    Padding rules of the reference: target length rounded up to a multiple of r and of downsample_step, plus r *
    downsample_step leading zero frames ("initial decoder state"); text / text positions zero-padded; frame positions
    1..T_dec; done = 0 for the first len//r//ds - 1 decoder steps, then 1."""
    multi_speaker = len(batch[0]) == 4
    input_lengths = [len(x[0]) for x in batch]
    max_input_len = max(input_lengths)
    target_lengths = [len(x[1]) for x in batch]
    max_target_len = max(target_lengths)
    if max_target_len % r != 0:
        max_target_len += r - max_target_len % r
    if max_target_len % downsample_step != 0:
        max_target_len += downsample_step - max_target_len % downsample_step
    b_pad = r
    max_target_len += b_pad * downsample_step

    x = torch.from_numpy(np.array([_pad(np.asarray(b[0]), max_input_len) for b in batch], dtype=np.int64))
    mel = torch.from_numpy(np.array([_pad_2d(b[1], max_target_len, b_pad=b_pad) for b in batch], dtype=np.float32))
    y = torch.from_numpy(np.array([_pad_2d(b[2], max_target_len, b_pad=b_pad) for b in batch], dtype=np.float32))
    text_positions = torch.from_numpy(np.array(
        [_pad(np.arange(1, len(b[0]) + 1), max_input_len) for b in batch], dtype=np.int64))
    T_dec = max_target_len // r // downsample_step
    frame_positions = torch.arange(1, T_dec + 1).long().unsqueeze(0).expand(len(batch), T_dec).clone()
    done = torch.from_numpy(np.array(
        [_pad(np.zeros(len(b[1]) // r // downsample_step - 1), T_dec, constant_values=1) for b in batch],
        dtype=np.float32)).unsqueeze(-1)
    if downsample_step > 1:
        mel = mel[:, 0::downsample_step, :].contiguous()          # train.py:639-640
    out = {
        "x": x, "text_positions": text_positions, "frame_positions": frame_positions, "mel": mel, "y": y,
        "done": done, "target_lengths": torch.tensor(target_lengths, dtype=torch.int64),
        "input_lengths_dev": torch.tensor(input_lengths, dtype=torch.int64),
    }
    if multi_speaker:
        out["speaker_ids"] = torch.tensor([b[3] for b in batch], dtype=torch.int64)
    if pin:
        out = {k: v.pin_memory() for k, v in out.items()}
    out["input_lengths"] = np.asarray(input_lengths, dtype=np.int64)
    return out


class TrainTxtDataset(torch.utils.data.Dataset):
    """Items are what ``collate`` consumes: (token ids int32, mel (T, num_mels) float32, linear (T, n_freq) float32
    [, speaker_id]).  ``frame_lengths`` (column 3 of train.txt) feeds the length-bucketed sampler without touching
    the .npy files.  ``speaker_id`` filters a multi-speaker corpus down to one speaker (and then yields 3-tuples),
    like the reference data sources do (train.py:101-122, 169-177)."""

    def __init__(self, data_root, text_to_sequence, speaker_id=None, mmap=True):
        import os
        self.data_root, self.text_to_sequence, self.mmap = data_root, text_to_sequence, mmap
        with open(os.path.join(data_root, "train.txt"), "rb") as f:
            rows = [line.decode("utf-8").rstrip("\n").split("|") for line in f if line.strip()]
        if not rows:
            raise ValueError("empty train.txt under %s" % data_root)
        n = len(rows[0])
        if n not in (4, 5) or any(len(r) != n for r in rows):
            raise ValueError("train.txt lines must have 4 or 5 '|'-separated fields")
        self.multi_speaker = n == 5
        if self.multi_speaker and speaker_id is not None:
            rows = [r for r in rows if int(r[4]) == speaker_id]
            self.multi_speaker = False
        self.rows = rows
        self.frame_lengths = [int(r[2]) for r in rows]

    def __len__(self):
        return len(self.rows)

    def _load(self, name):
        import os
        return np.load(os.path.join(self.data_root, name), mmap_mode="r" if self.mmap else None)

    def __getitem__(self, idx):
        r = self.rows[idx]
        seq = np.asarray(self.text_to_sequence(r[3]), dtype=np.int32)
        item = (seq, np.asarray(self._load(r[1]), dtype=np.float32), np.asarray(self._load(r[0]), dtype=np.float32))
        return item + (int(r[4]),) if self.multi_speaker else item


class DistributedSimilarLengthSampler(torch.utils.data.Sampler):
    """Yields the dataset indices of this rank's mini-batches, batch after batch (use with
    ``DataLoader(batch_size=batch_size, sampler=..., collate_fn=...)`` and ``drop_last=True``)."""

    def __init__(self, lengths, batch_size=16, batch_group_size=None, permutate=True, rank=0, world_size=1, seed=0):
        lengths = torch.as_tensor(np.asarray(lengths), dtype=torch.int64)
        self.lengths, self.sorted_indices = torch.sort(lengths)
        self.batch_size = batch_size
        if batch_group_size is None:
            batch_group_size = min(batch_size * 32, len(self.lengths))
            if batch_group_size % batch_size != 0:
                batch_group_size -= batch_group_size % batch_size
        assert batch_group_size % batch_size == 0 and batch_group_size > 0
        self.batch_group_size = batch_group_size
        self.permutate = permutate
        self.rank, self.world_size, self.seed, self.epoch = rank, world_size, seed, 0
        n_batches = len(self.lengths) // batch_size
        self.batches_per_rank = n_batches // world_size

    def set_epoch(self, epoch):
        self.epoch = epoch

    def _global_order(self):
        rng = np.random.RandomState(self.seed + self.epoch)       # identical on every rank
        idx = self.sorted_indices.numpy().copy()
        g, e = self.batch_group_size, 0
        for i in range(len(idx) // g):
            s, e = i * g, (i + 1) * g
            rng.shuffle(idx[s:e])
        if self.permutate and e > 0:
            perm = rng.permutation(e // self.batch_size)
            idx[:e] = idx[:e].reshape(-1, self.batch_size)[perm].reshape(-1)
        if e < len(idx):
            rng.shuffle(idx[e:])
        return idx

    def __iter__(self):
        idx = self._global_order()
        n_batches = self.batches_per_rank * self.world_size
        batches = idx[:n_batches * self.batch_size].reshape(n_batches, self.batch_size)
        mine = batches[self.rank::self.world_size]
        return iter(mine.reshape(-1).tolist())

    def __len__(self):
        return self.batches_per_rank * self.batch_size

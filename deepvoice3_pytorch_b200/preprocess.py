"""Dataset preprocessing on the GPU: the caller of the audio front-end.

The reference's ``ljspeech.py`` (``build_from_path`` :9-37, ``_process_utterance`` :40-79) walks ``metadata.csv`` and,
per utterance, loads the wav, optionally rescales it, runs TWO CPU STFTs (``audio.spectrogram`` and
``audio.melspectrogram``) in a process pool and writes ``<name>-spec-%05d.npy`` (T, 513) and ``<name>-mel-%05d.npy``
(T, 80) plus one ``train.txt`` row ``spec|mel|n_frames|text``.  ``build_from_path`` here has the same arguments, writes
the same files and returns the same tuples, but batches the clips through ONE fused kernel launch per batch
(``audio.stft_mel_batch``: one H2D copy, one launch, one D2H copy for ``batch_clips`` utterances); ``num_workers``
threads only overlap the wav decoding and the ``np.save`` calls with the GPU work.  The reference's own
``_process_utterance`` also runs unchanged on this package's ``audio`` module (tests/test_dropin.py) -- one clip per
launch; this module is the batched equivalent.

    from deepvoice3_pytorch_b200 import preprocess
    rows = preprocess.build_from_path(in_dir, out_dir, num_workers=4)
    preprocess.write_metadata(rows, out_dir)          # preprocess.py:26-35 of the reference: train.txt
"""
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

from . import audio


def _load(wav_path):
    wav = audio.load_wav(wav_path)
    hp = audio.hparams
    if hp.rescaling:
        wav = wav / np.abs(wav).max() * hp.rescaling_max          # ljspeech.py:59-60
    return np.ascontiguousarray(wav, dtype=np.float32)


def spectrograms_batch(wavs):
    """[float32 waveform (n_i,)] -> [(linear (T_i, 513), mel (T_i, 80))] float32, one fused launch for the whole list:
    what ``audio.spectrogram(w).T`` / ``audio.melspectrogram(w).T`` return per clip."""
    lens = [len(w) for w in wavs]
    # row pitch a multiple of 4 samples: the kernel then stages with 16-byte / bulk copies
    host = torch.zeros(len(wavs), (max(lens) + 3) // 4 * 4, dtype=torch.float32).pin_memory()
    for i, w in enumerate(wavs):
        host[i, :len(w)] = torch.from_numpy(w)
    lin, mel = audio.stft_mel_batch(host.cuda(non_blocking=True), torch.tensor(lens, dtype=torch.int32).cuda())
    lin, mel = lin.cpu().numpy(), mel.cpu().numpy()
    out = []
    for i, n in enumerate(lens):
        T = audio.num_frames(n)
        out.append((lin[i, :T].copy(), mel[i, :T].copy()))
    return out


def build_from_path(in_dir, out_dir, num_workers=1, tqdm=lambda x: x, batch_clips=64, name="ljspeech",
                    rank=None, world=None):
    """Same contract as reference ``ljspeech.build_from_path`` (:9-37): reads ``in_dir/metadata.csv`` and
    ``in_dir/wavs/*.wav``, writes the .npy pairs into ``out_dir`` and returns
    ``[(spectrogram_filename, mel_filename, n_frames, text)]`` in file order.

    Multi-GPU (one process per GPU): utterances are dealt round-robin over the ranks -- the work is independent per
    clip, so there is no data-path collective; file indices are global, every rank writes its own files, and the rows of
    all ranks are merged into file order with one ``all_gather_object`` of the (tiny) row lists.  ``rank`` / ``world``
    default to the initialised ``torch.distributed`` group, else to a single process."""
    import torch.distributed as dist
    if world is None:
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        rank = dist.get_rank() if world > 1 else 0
    hp = audio.hparams
    items = []
    index = 1
    with open(os.path.join(in_dir, "metadata.csv"), encoding="utf-8") as f:
        for line in f:
            parts = line.strip().split("|")
            text = parts[2]
            if len(text) < hp.min_text:
                continue
            items.append((index, os.path.join(in_dir, "wavs", "%s.wav" % parts[0]), text))
            index += 1
    items = items[rank::world]                       # this rank's share; indices stay global
    rows = []
    with ThreadPoolExecutor(max_workers=max(1, num_workers)) as pool:
        batches = [items[i:i + batch_clips] for i in range(0, len(items), batch_clips)]
        loads = [pool.map(_load, [p for _, p, _ in b]) for b in batches[:1]]        # decode one batch ahead
        saves = []
        for bi, batch in enumerate(tqdm(batches)):
            wavs = list(loads[bi])
            if bi + 1 < len(batches):
                loads.append(pool.map(_load, [p for _, p, _ in batches[bi + 1]]))
            for (idx, _, text), (lin, mel) in zip(batch, spectrograms_batch(wavs)):
                spec_name, mel_name = "%s-spec-%05d.npy" % (name, idx), "%s-mel-%05d.npy" % (name, idx)
                saves.append(pool.submit(np.save, os.path.join(out_dir, spec_name), lin, allow_pickle=False))
                saves.append(pool.submit(np.save, os.path.join(out_dir, mel_name), mel, allow_pickle=False))
                rows.append((idx, (spec_name, mel_name, lin.shape[0], text)))
        for s in saves:
            s.result()
    if world > 1 and dist.is_available() and dist.is_initialized():
        parts = [None] * world
        dist.all_gather_object(parts, rows)
        rows = [r for part in parts for r in part]
    return [r for _, r in sorted(rows, key=lambda ir: ir[0])]


def write_metadata(metadata, out_dir):
    """``train.txt`` exactly as reference preprocess.py:26-35 writes it (and ``data.TrainTxtDataset`` reads it)."""
    with open(os.path.join(out_dir, "train.txt"), "w", encoding="utf-8") as f:
        for m in metadata:
            f.write("|".join([str(x) for x in m]) + "\n")
    frames = sum(m[2] for m in metadata)
    hours = frames * audio.hparams.hop_size / audio.hparams.sample_rate / 3600
    return frames, hours

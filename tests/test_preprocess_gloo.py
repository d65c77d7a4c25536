"""CPU, world_size=2, gloo: the sharded preprocessing job (BASELINE config #5's shape: clips dealt round-robin over the
ranks, no data-path collective, rows merged into file order).  The GPU kernel is replaced by the numpy restatement of
audio.py (oracle/audio_oracle.py), so this checks the host logic: sharding, global file indices, the reference's skip
rule, the on-disk format and the merged ``train.txt`` rows -- against a single-process run."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp
from scipy.io import wavfile


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _oracle_batch(wavs):
    from oracle import audio_oracle as A
    return [A.process_utterance(w) for w in wavs]


def _make_dataset(root):
    from oracle import audio_oracle as A
    os.makedirs(os.path.join(root, "wavs"))
    lens = [9000, 4000, 12000, 7000, 5000, 15000, 3000]
    with open(os.path.join(root, "metadata.csv"), "w", encoding="utf-8") as f:
        for i, n in enumerate(lens):
            text = "short" if i == 3 else "a transcript that is long enough, number %d" % i
            x = A.synthetic_clip(20 + i, n=n)
            wavfile.write(os.path.join(root, "wavs", "U%02d.wav" % i), 22050, (x * 32767).astype(np.int16))
            f.write("U%02d|%s|%s\n" % (i, text, text))


def _worker(rank, world, port, in_dir, out_dir, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from deepvoice3_pytorch_b200 import preprocess
    preprocess.spectrograms_batch = _oracle_batch            # CPU stand-in for the fused kernel
    rows = preprocess.build_from_path(in_dir, out_dir, num_workers=2, batch_clips=2)
    ret[rank] = rows
    if rank == 0:
        preprocess.write_metadata(rows, out_dir)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_preprocessing_matches_single_process(tmp_path):
    from deepvoice3_pytorch_b200 import preprocess
    in_dir, out1, out2 = str(tmp_path / "in"), str(tmp_path / "one"), str(tmp_path / "two")
    _make_dataset(in_dir)
    os.makedirs(out1); os.makedirs(out2)
    saved = preprocess.spectrograms_batch
    preprocess.spectrograms_batch = _oracle_batch
    try:
        single = preprocess.build_from_path(in_dir, out1, num_workers=1, batch_clips=4, rank=0, world=1)
    finally:
        preprocess.spectrograms_batch = saved
    assert len(single) == 6 and [r[0] for r in single] == ["ljspeech-spec-%05d.npy" % i for i in range(1, 7)]
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), in_dir, out2, ret), nprocs=2, join=True)
    assert ret[0] == single and ret[1] == single                 # every rank ends with the merged rows, in file order
    for spec_name, mel_name, n_frames, _ in single:
        for name, width in ((spec_name, 513), (mel_name, 80)):
            a, b = np.load(os.path.join(out1, name)), np.load(os.path.join(out2, name))
            assert a.shape == (n_frames, width) and a.dtype == np.float32 and np.array_equal(a, b)
    lines = open(os.path.join(out2, "train.txt"), encoding="utf-8").read().splitlines()
    assert len(lines) == 6 and lines[2].startswith("ljspeech-spec-00003.npy|ljspeech-mel-00003.npy|")

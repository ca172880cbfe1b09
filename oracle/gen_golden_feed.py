"""TEST INFRASTRUCTURE ONLY — golden vectors for the input feed and the fairseq key mapping, written from the UNMODIFIED
reference in the dev container (`python oracle/gen_golden_feed.py`):

* `tests/golden/fairseq_keymap.json`: {fairseq key: audio8 key} exactly as the reference's `convert_keys`
  (wav2vec2.py:154-168) renames a checkpoint, for the pre-training map and the CTC maps at 16 and 8 kHz;
* `tests/golden/feed.npz`: batches produced by the reference's `AudioFileDataset.__iter__` (data.py:409-426) and
  `AudioTextLetterDataset.read_batch` (data.py:263-294) on a seeded synthetic sample stream (the methods run
  unmodified on instances whose file reading is replaced by the synthetic stream; `soundfile` is absent here and is
  stubbed for the import only).
"""
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from load_reference import load_reference  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def sample_stream(seed, n, lo, hi):
    rng = np.random.RandomState(seed)
    for _ in range(n):
        yield (rng.randn(int(rng.randint(lo, hi))) * 0.1).astype(np.float32)


def main():
    w2v, _ = load_reference()
    # ---- key maps: feed convert_keys a dict whose values are the fairseq key names themselves
    maps = {}
    for name, (nested, flat, n_layers) in dict(pretrain=(w2v.W2V_MAP.nested, w2v.W2V_MAP.flat, 3),
                                               ctc16=(w2v.W2V_CTC_MAP[16].nested, w2v.W2V_CTC_MAP[16].flat, 3),
                                               ctc8=(w2v.W2V_CTC_MAP[8].nested, w2v.W2V_CTC_MAP[8].flat, 3)).items():
        d = {}
        for i in range(n_layers):
            for k in nested:
                d[k.format(i)] = k.format(i)
        for k in flat:
            d[k] = k
        d["untouched.key"] = "untouched.key"
        out = w2v.convert_keys(n_layers, d, nested, flat)
        maps[name] = {v: k for k, v in out.items()}  # fairseq -> audio8
    with open(os.path.join(OUT, "fairseq_keymap.json"), "w") as f:
        json.dump(maps, f, indent=0, sort_keys=True)
    # ---- data feed
    sys.modules.setdefault("soundfile", types.ModuleType("soundfile"))
    import audio8.data as D
    ds = object.__new__(D.AudioFileDataset)
    ds.max_length, ds.target_tokens_per_batch = 4000, 9000
    ds.next_sample = lambda: sample_stream(7, 40, 1500, 5000)
    ds.process_sample = None
    out = {"pre_cfg": np.array([7, 40, 1500, 5000, 4000, 9000])}
    crop = lambda it: (s[:4000] for s in it)  # the reader crops to max_length (data.py:400)
    ds.next_sample = lambda: crop(sample_stream(7, 40, 1500, 5000))
    batches = list(D.AudioFileDataset.__iter__(ds))
    out["pre_n"] = np.array([len(batches)])
    for i, b in enumerate(batches):
        out[f"pre_{i}"] = b
    lt = object.__new__(D.AudioTextLetterDataset)
    lt.max_dst_length, lt.max_src_length = 12, 6000
    audios = list(sample_stream(9, 5, 2000, 6000))
    rng = np.random.RandomState(3)
    toks = [rng.randint(4, 32, size=int(rng.randint(3, 12))) for _ in range(5)]
    lt.files, lt.tokens = [f"f{i}" for i in range(5)], toks
    lt.process_sample = lambda pth: audios[int(pth[1:])]
    from eight_mile.utils import Offsets
    out["ft_pad"] = np.array([Offsets.PAD])
    rb = D.AudioTextLetterDataset.read_batch(lt, [3, 0, 4])
    out["ft_signal"], out["ft_signal_lengths"] = rb["signal"], rb["signal_lengths"]
    out["ft_token_ids"], out["ft_token_lengths"] = rb["token_ids"], rb["token_lengths"]
    np.savez_compressed(os.path.join(OUT, "feed.npz"), **out)
    print("wrote fairseq_keymap.json, feed.npz:", len(batches), "pre-training batches", [b.shape for b in batches])


if __name__ == "__main__":
    main()

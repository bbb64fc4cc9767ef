#!/usr/bin/env python
"""Encode the reference's Penn Treebank text into token ids with the reference's own vocabulary rule and
commit them as data (`tests/golden/ptb_ids.npz`, int16), so the GPU box -- which has no `/root/reference` --
can run the README recipes (README.md:20-27) and the perplexity parity checks on the real corpus.

Rule restated from main.py:44-59: each file is read whole, its first character (a leading space) is dropped,
the rest is split on single spaces -- the line break survives as a token of its own ('\\n', id 0 in the sorted
vocabulary); vocabulary = sorted(set(train tokens)); ids index that sorted list.

    python tests/golden/make_ptb_ids.py [/root/reference/data]       # run where the reference is mounted

The script is the generator of the fixture (test infrastructure); the product never reads it.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def read_tokens(path):                                  # main.py:45-53
    with open(path) as f:
        return f.read()[1:].split(" ")


def main(root):
    trn = read_tokens(os.path.join(root, "ptb.train.txt"))
    vld = read_tokens(os.path.join(root, "ptb.valid.txt"))
    tst = read_tokens(os.path.join(root, "ptb.test.txt"))
    words = sorted(set(trn))                            # main.py:54
    w2i = {w: i for i, w in enumerate(words)}           # main.py:55
    assert len(words) < 32768
    enc = lambda toks: np.array([w2i[w] for w in toks], dtype=np.int16)     # main.py:56-58
    out = {"train": enc(trn), "valid": enc(vld), "test": enc(tst), "vocab_size": np.int64(len(words))}
    # a digest of the vocabulary itself (not the words: the fixture carries ids only)
    out["vocab_sha256"] = np.frombuffer(hashlib.sha256("\x00".join(words).encode()).digest(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "ptb_ids.npz"), **out)
    print({k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "/root/reference/data")

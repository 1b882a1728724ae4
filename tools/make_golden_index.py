"""Golden vectors for the host index builders, produced by the REFERENCE's own compiled helpers.cpp (oracle/_ref, built
by `make -C oracle`; build container only). Writes tests/golden/index_builders.npz.
    python tools/make_golden_index.py"""
import contextlib
import io
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
import helpers  # noqa: E402  (the reference module)

rng = np.random.default_rng(20261017)
out = {}
# build_rl_sample_idx: (path_lengths, transition_num)
rl_cases = [(np.array([5, 3, 1, 7], np.int32), 3)]
for n, hi, T in [(50, 40, 8), (200, 300, 43), (7, 3, 5), (1, 1000, 38)]:
    rl_cases.append((rng.integers(1, hi + 1, size=n).astype(np.int32), T))
for i, (lens, T) in enumerate(rl_cases):
    out["rl%d:lens" % i] = lens
    out["rl%d:T" % i] = np.array([T], np.int32)
    out["rl%d:idx" % i] = np.asarray(helpers.build_rl_sample_idx(lens, T))
# build_sample_idx: (sizes, doc_idx, seq_length, num_epochs, tokens_per_epoch)
gpt_cases = [(np.array([10, 4, 7, 9], np.int32), np.array([0, 1, 2, 3, 2, 1, 0, 3], np.int32), 8, 2, 30)]
for ndoc, hi, seq, ep in [(40, 200, 64, 3), (300, 50, 128, 1), (5, 4000, 1024, 2), (64, 17, 16, 4)]:
    sizes = rng.integers(1, hi + 1, size=ndoc).astype(np.int32)
    doc_idx = np.concatenate([rng.permutation(ndoc) for _ in range(ep)]).astype(np.int32)
    gpt_cases.append((sizes, doc_idx, seq, ep, int(sizes.sum())))
for i, (sizes, doc_idx, seq, ep, tpe) in enumerate(gpt_cases):
    with contextlib.redirect_stdout(io.StringIO()):
        idx = np.asarray(helpers.build_sample_idx(sizes, doc_idx, seq, ep, tpe))
    out["gpt%d:sizes" % i] = sizes
    out["gpt%d:doc_idx" % i] = doc_idx
    out["gpt%d:args" % i] = np.array([seq, ep, tpe], np.int64)
    out["gpt%d:idx" % i] = idx
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "index_builders.npz"), **out)
print("wrote %d arrays (%d rl cases, %d gpt cases)" % (len(out), len(rl_cases), len(gpt_cases)))

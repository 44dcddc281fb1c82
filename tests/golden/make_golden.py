"""Generates tests/golden/{cfg1,cqt87,istft,subtract_chain}.npz from the CPU oracle (oracle/).

These vectors pin the oracle against silent drift and give the GPU tests fixed targets that do not depend on
re-running it.  The oracle's container is checked bit for bit against the reference's own `audio_complete`
(oracle/ref_class.py, tests/test_ref_class.py; golden vectors of THAT class: make_ref_class_golden.py); the librosa /
resampy layer underneath is a restatement (no wheel, no network), so CQT and dB values here are unpinned against
reference outputs -- none exist.  Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import cqt as ocqt            # noqa: E402
from oracle import spectral as osp        # noqa: E402
from oracle.audio_oracle import AudioOracle  # noqa: E402
from tests.conftest import cfg1_clip      # noqa: E402
from tests.synth import piano_clip        # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    # --- cfg1 (BASELINE.json configs[0]): 10 s, 16 kHz sine mix; strided probes keep the file small
    y, sr = cfg1_clip()
    a = AudioOracle(y, 2048, 512, sample_rate=sr)
    mag = a.mag
    C = a.slice_C(0, 10.0, 313, bins_per_tone=1, lowest_note="C1", nbins=84)
    D = a.D
    cols = np.array([0, 1, 2, 50, 157, 311, 312])
    np.savez_compressed(
        os.path.join(HERE, "cfg1.npz"),
        mag_cols=mag[:, cols].astype(np.float32), D_cols=D[:, cols].astype(np.float32),
        cqt_cols=C[:, cols].astype(np.float64), cols=cols,
        mag_shape=np.array(mag.shape), cqt_shape=np.array(C.shape),
        ref_mag=np.float32(a.ref_mag), mag_colsum=mag.sum(axis=0).astype(np.float64),
        cqt_rowsum=C.sum(axis=1).astype(np.float64))
    # --- subtract chain on a small window (bit-level target for K3)
    rng = np.random.default_rng(42)
    win = (rng.random((129, 40), dtype=np.float32) ** 3)
    gs = (rng.random((3, 129, 9), dtype=np.float32) ** 2)
    offs = np.array([0, 17, 35], dtype=np.int32)
    w = win.copy()
    for j in range(3):
        g = gs[j].copy()
        g *= np.max(w) / np.max(g)
        g = g[:, : 40 - offs[j]]
        pad = np.concatenate((np.zeros((129, offs[j])), g, np.zeros((129, 40 - offs[j] - g.shape[1]))), axis=1)
        w -= pad
        w = np.maximum(w, 0, w)
    np.savez_compressed(os.path.join(HERE, "subtract_chain.npz"), win=win, guesses=gs, offsets=offs,
                        result=w, D=osp.amplitude_to_db(w, ref=w.max()).astype(np.float32))
    # --- iSTFT of a short piano clip
    p = piano_clip(5, 12000)
    F = osp.stft(p, 1024, 256)
    np.savez_compressed(os.path.join(HERE, "istft.npz"), wav=p, istft=osp.istft(F, 256).astype(np.float32))
    # --- a non-multiple-of-octave CQT (87 bins, reference default ref_C_1 shape) on a 1 s clip
    q = piano_clip(9, 44100)
    C87 = np.abs(ocqt.cqt(q, sr=44100, hop_length=1024, fmin=osp.note_to_hz("A0"), n_bins=87,
                          bins_per_octave=12, filter_scale=2))
    np.savez_compressed(os.path.join(HERE, "cqt87.npz"), wav=q, cqt=C87.astype(np.float32))
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()

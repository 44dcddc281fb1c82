"""Reference-output pins, generated from the reference's OWN fixtures.

`/root/reference/subtraction_demo/<name>_test{,_guess,_sub}.flac` were written by the
reference itself (`test_snippets.py:473-514`: `ac_sub = ac.clone(); ac_sub.subtract(ac_guess,
offset=0.5, ...)`, each saved through `audio_complete.save` -> `audio_to_flac`,
`util_audio.py:520-527`, `:966-968`, 24-bit PCM).  They are the only numerical outputs of the
reference that exist anywhere, so they pin STFT -> magphase -> ref_mag -> subtract -> iSTFT
end to end.

This script (run here, where /root/reference is mounted) decodes them with the test-only
FLAC reader, searches the unrecorded call parameters, and writes
    tests/golden/ref_subtraction_piano.npz    the PCM of one triple (travels to the GPU box)
    tests/golden/ref_subtraction_pins.json    parameters + residual of every triple that the
                                              oracle reproduces to 24-bit precision
Usage: python tests/golden/make_reference_pins.py
"""
import glob
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle.audio_oracle import AudioOracle  # noqa: E402
from tests.flac_reader import read_flac  # noqa: E402

DEMO = "/root/reference/subtraction_demo/"
N_FFT = 4096            # `buckets = 4096` in the generating cell; hop defaults to n_fft/4
SCALE = float(1 << 23)  # libsndfile float <-> PCM_24


def load(name):
    pcm, sr, bps = read_flac(DEMO + name + ".flac")
    assert sr == 44100 and bps == 24
    return pcm


def residual(wf, g, sub, **kw):
    ac = AudioOracle((wf / SCALE).astype(np.float32), N_FFT)
    acg = AudioOracle((g / SCALE).astype(np.float32), N_FFT)
    s = ac.clone()
    s.subtract(acg, **kw)
    w = np.asarray(s.wf, dtype=np.float64)
    if len(w) != len(sub):
        return np.inf, np.inf
    e = np.abs(w * SCALE - sub)
    return float(e.max()), float(np.sqrt((e ** 2).mean()))


def main():
    names = sorted(os.path.basename(p)[:-len("_test_sub.flac")] for p in glob.glob(DEMO + "*_test_sub.flac"))
    pins = {}
    frame = 1024 / 44100
    for name in names:
        if not os.path.exists(DEMO + name + "_test_guess.flac"):
            continue
        wf, g, sub = load(name + "_test"), load(name + "_test_guess"), load(name + "_test_sub")
        best = None
        for normalize in (True, False):
            for k in (0, -1, 1, -2, 2):
                for comp in (0, 1, 2):
                    kw = dict(offset=0.5 + k * frame + (1e-9 if k else 0.0), attack_compensation=comp,
                              normalize=normalize)
                    mx, rms = residual(wf, g, sub, **kw)
                    if best is None or mx < best[0]:
                        best = (mx, rms, kw)
                if best[0] < 8:
                    break
            if best[0] < 8:
                break
        print("%-28s max %.3g LSB  rms %.3g LSB  %s" % (name, best[0], best[1], best[2]), flush=True)
        if best[0] < 8:          # 24-bit quantisation of input and output: a few LSB
            pins[name] = {"max_err_lsb24": best[0], "rms_err_lsb24": best[1], "n_fft": N_FFT,
                          "samples": [int(len(wf)), int(len(g)), int(len(sub))], **best[2]}
    with open(os.path.join(HERE, "ref_subtraction_pins.json"), "w") as fh:
        json.dump(pins, fh, indent=1, sort_keys=True)
    wf, g, sub = load("piano_test"), load("piano_test_guess"), load("piano_test_sub")
    np.savez_compressed(os.path.join(HERE, "ref_subtraction_piano.npz"), test=wf, guess=g, sub=sub)


if __name__ == "__main__":
    main()

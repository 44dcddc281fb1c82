"""Batched per-note step (NoteStepBatch.step) on 600 windows of the reference's real shape (N 4096, hop 1024,
258 frames): ms per step and per note, with a per-stage breakdown, next to the one-window-at-a-time class."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import amt_saga_b200  # noqa
from amt_saga_b200 import ops, synth
from amt_saga_b200.note_step import NoteStepBatch

dev = torch.device("cuda")
W = int(sys.argv[1]) if len(sys.argv) > 1 else 600
L = 264168
wav = synth.piano_batch(range(W), L, 44100, seed_base=50000, device=dev)
plan = ops.get_stft_plan(4096, 1024, True)
r = ops.stft_batch(wav, plan, want_phase=True)
mag, ph = r["mag_storage"][:, :258].contiguous(), r["phase_storage"][:, :258].contiguous()
b = NoteStepBatch(W)
b.load(mag, ph, wav, r["clip_max"], np.ones((W, 3)))
rng = np.random.default_rng(0)
guess = synth.piano_batch(range(W), 54277, 44100, n_notes=1, seed_base=90000, device=dev)
def one(seed, n_pitches):
    rg = np.random.default_rng(seed)
    onset = rg.uniform(0, 5.0, W); dur = rg.uniform(0.2, 1.2, W)
    pitch = rg.integers(21, 21 + n_pitches, W) if n_pitches < 88 else rg.integers(21, 109, W)
    return b.step(onset, dur, pitch, guess)
out = {}
for n_pitches in (88, 12):
    for i in range(3):
        one(i, n_pitches)
    torch.cuda.synchronize()
    t0 = time.perf_counter(); n = 5
    for i in range(n):
        one(10 + i, n_pitches)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / n * 1e3
    out["ms_per_step_%d_pitches" % n_pitches] = round(ms, 2)
    out["us_per_note_%d_pitches" % n_pitches] = round(ms / W * 1e3, 1)
# stage breakdown (events around each call family)
def timed(fn, n=5):
    fn(); torch.cuda.synchronize(); a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize(); return round(a.elapsed_time(e) / n, 3)
s = np.full(W, 50); t = s + 30
out["stages_ms"] = {
    "istft": timed(lambda: ops.istft_batch(b.stft, mag=b.mag, phase=b.ph, n_bins=b.nb)),
    "cqt_174_24": timed(lambda: b._cqt_columns(b.wav, 21, 174, 2, s, t, 8, b.inv_ref_C[0])),
    "cqt_348_48": timed(lambda: b._cqt_columns(b.wav, 21, 348, 4, s, t, 8, b.inv_ref_C[1])),
    "cqt_348_192_C4": timed(lambda: b._cqt_columns(b.wav, 60, 348, 16, s, t, 8, b.inv_ref_C[2])),
    "cqt_348_192_one_pitch_all_windows": timed(lambda: b._cqt_columns(b.wav, 57, 348, 16, s, t, 8, b.inv_ref_C[2])),
    "cqt_36_24_one_pitch_all_windows": timed(lambda: b._cqt_columns(b.wav, 47, 36, 2, s, t, 8, b.inv_ref_C[2])),
    "guess_stft_subtract": timed(lambda: b.subtract(guess, np.full(W, 40, dtype=np.int32))),
    "compress_bands": timed(lambda: ops.compress_bands_batch(b.mag, b.nb, amt_saga_b200.util_audio.band_edges(b.nb, 20), inv_scale=b.inv_song_ref)),
}
out["windows"] = W
print(json.dumps(out))

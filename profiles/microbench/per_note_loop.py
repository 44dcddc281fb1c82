"""One iteration of the reference's per-note loop (training.py:318-449 call sequence) through the drop-in
`audio_complete` class, one window at a time (no batching): what a user who only swaps the import gets.
Same sequence through the CPU oracle for scale."""
import sys, json, time, torch
sys.path.insert(0, "/root/repo")
import numpy as np
import amt_saga_b200  # noqa: F401
from amt_saga_b200 import util_audio as ua
from oracle.audio_oracle import AudioOracle
from tests.synth import piano_clip
sr, N = 44100, 4096
song = piano_clip(77, sr * 10)
note = piano_clip(78, int(sr * 1.2), n_notes=1)

def one_note(AC, aw, onset, d):
    sa = aw.resize(onset, d, 8, attribs=["mag", "ph"])
    AC._resize(AC.compress_bands(aw.mag, bands=20), 258)
    aw.slice_C(onset, d, 8, bins_per_tone=2)                 # 174 bins, 24 per octave
    aw.slice_C(onset, d, 8, bins_per_tone=4)                 # 348 bins, 48 per octave
    b0 = aw.midi_tone_to_FFT(60)
    sa.section_power("mag", b0, b0 + 348)
    aw.subtract(AC(note, N), offset=onset)
    aw.wf                                                     # the hidden iSTFT of util_audio.py:88-106
    return aw

def run(AC, n, sync):
    a = AC(song, N)
    a.mag
    aw = a.section(0, None, 258)
    one_note(AC, aw, 1.3, 0.8)
    sync()
    t0 = time.perf_counter()
    for i in range(n):
        one_note(AC, aw, 0.5 + 0.3 * (i % 10), 0.8)
    sync()
    return (time.perf_counter() - t0) / n

gpu = run(ua.audio_complete, 50, torch.cuda.synchronize)
cpu = run(AudioOracle, 2, lambda: None)
print(json.dumps({"per_note_ms_audio_complete_gpu": round(gpu * 1e3, 3), "per_note_ms_oracle_cpu": round(cpu * 1e3, 1),
                  "ratio": round(cpu / gpu, 1),
                  "note": "one window at a time through the drop-in class: launch / Python bound, the batched pipeline is the throughput path"}))

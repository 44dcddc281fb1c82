"""The bench step at the reference's own default analysis shape (main.py: N=4096, hop=1024; CQT 87 bins from A0,
12 per octave, training.py:271): 600 x 6 s windows, stage times of one pass."""
import sys, json, torch
sys.path.insert(0, "/root/repo")
import numpy as np
import amt_saga_b200  # noqa: F401
from amt_saga_b200 import synth
from amt_saga_b200.pipeline import WindowFeaturePipeline
W = 600
pipe = WindowFeaturePipeline(W, 264600, 65024, n_fft=4096, hop=1024, cqt_lowest="A0", cqt_bins=87, cqt_bpo=12)
wav = synth.piano_batch(range(W), 264600, 44100, seed_base=50000, device="cuda")
guess = synth.piano_batch(range(W), 65024, 44100, n_notes=1, seed_base=90000, device="cuda")
offs = torch.as_tensor(np.random.default_rng(7).integers(0, 250, size=(W, 1)).astype(np.int32), device="cuda")
for _ in range(5): pipe.run(wav, guess, offs)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(30): pipe.run(wav, guess, offs)
b.record(); torch.cuda.synchronize()
ev = []
for _ in range(10): pipe.run(wav, guess, offs, ev)
torch.cuda.synchronize()
st = {}
for name, x, y in ev: st[name] = st.get(name, 0.0) + x.elapsed_time(y) / 10
print(json.dumps({"shape": "n_fft 4096 hop 1024, CQT 87 bins from A0", "frames_per_window": pipe.T, "loop_ms": round(a.elapsed_time(b) / 30, 4),
                  "window_features_per_s": round(W / (a.elapsed_time(b) / 30) * 1e3), "stages_ms": {k: round(v, 4) for k, v in st.items()}}))

import os, sys, json, torch, ctypes as C
sys.path.insert(0, "/root/repo")
import numpy as np
import amt_saga_b200
from amt_saga_b200 import synth, _lib
from amt_saga_b200.pipeline import WindowFeaturePipeline
W = 600
pipe = WindowFeaturePipeline(W, 264600, 65024)
wav = synth.piano_batch(range(W), 264600, 44100, seed_base=50000, device="cuda")
guess = synth.piano_batch(range(W), 65024, 44100, n_notes=1, seed_base=90000, device="cuda")
offs = torch.as_tensor(np.random.default_rng(7).integers(0, 500, size=(W, 1)).astype(np.int32), device="cuda")
for _ in range(3): pipe.run(wav, guess, offs)
torch.cuda.synchronize()
def loop(n=40):
    for _ in range(3): pipe.run(wav, guess, offs)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): pipe.run(wav, guess, offs)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
for ch in (sys.argv[1:] or ["8", "12", "16", "24", "8"]):
    os.environ[os.environ.get("SWEEP_VAR", "SAGA_DB_CHUNKS")] = ch
    print(ch, round(loop(), 4), flush=True)

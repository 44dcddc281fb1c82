"""Small driver for profiling the fp32 contraction (348 bins, 48 per octave, 100 windows)."""
import sys, torch
sys.path.insert(0, "/root/repo")
import amt_saga_b200  # noqa: F401
from amt_saga_b200 import ops, synth
from amt_saga_b200.util_audio import note_to_hz
wav = synth.piano_batch(range(100), 264600, 44100, seed_base=50000, device="cuda")
plan = ops.CqtPlan(44100, 1024, note_to_hz("A0"), 348, 48, filter_scale=2)
for _ in range(2):
    ops.cqt_batch(wav, plan)
torch.cuda.synchronize()
print("ok")

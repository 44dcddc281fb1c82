"""Just the streamed-bank contraction: 348 bins at 48 per octave (ref_C_4, training.py:277), whole transform of 600
windows, then the 8-column frame window of the same plan: the ncu target for cqt_umma_stream_kernel."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import amt_saga_b200  # noqa
from amt_saga_b200 import ops, synth
from amt_saga_b200.util_audio import note_to_hz
wav = synth.piano_batch(range(600), 264168, 44100, seed_base=50000, device="cuda")
plan = ops.CqtPlan(44100, 1024, note_to_hz("A0"), 348, 48, filter_scale=2)
first = np.full(600, 100, dtype=np.int32)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(4):
    if i == 1: a.record()
    ops.cqt_batch(wav, plan)
b.record(); torch.cuda.synchronize()
print("348/48 whole transform %.3f ms/launch (cascade included)" % (a.elapsed_time(b) / 3))
ops.cqt_frames_batch(wav, plan, first)
torch.cuda.synchronize()

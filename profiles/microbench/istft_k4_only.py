"""Just K4 (inverse ring kernel, n_fft 2048 / hop 512) on the bench batch, magnitude + phase input: the ncu target."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import amt_saga_b200  # noqa
from amt_saga_b200 import ops, synth
dev = torch.device("cuda")
plan = ops.get_stft_plan(2048, 512, True)
wav = synth.piano_batch(range(600), 264600, 44100, seed_base=50000, device=dev)
r = ops.stft_batch(wav, plan, want_phase=True)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(5):
    if i == 2: a.record()
    y = ops.istft_batch(plan, mag=r["mag_storage"], phase=r["phase_storage"])
b.record(); torch.cuda.synchronize()
print("K4 %.3f ms/launch" % (a.elapsed_time(b) / 3))

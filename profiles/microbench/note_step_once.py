"""One batched per-note step on 600 windows (after warm-up): the ncu launch-list target."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import amt_saga_b200  # noqa
from amt_saga_b200 import ops, synth
from amt_saga_b200.note_step import NoteStepBatch
dev = torch.device("cuda"); W = 600; L = 264168
wav = synth.piano_batch(range(W), L, 44100, seed_base=50000, device=dev)
plan = ops.get_stft_plan(4096, 1024, True)
r = ops.stft_batch(wav, plan, want_phase=True)
b = NoteStepBatch(W)
b.load(r["mag_storage"][:, :258].contiguous(), r["phase_storage"][:, :258].contiguous(), wav, r["clip_max"], np.ones((W, 3)))
guess = synth.piano_batch(range(W), 54277, 44100, n_notes=1, seed_base=90000, device=dev)
npitch = int(sys.argv[1]) if len(sys.argv) > 1 else 12
for i in range(3):
    rg = np.random.default_rng(i)
    b.step(rg.uniform(0, 5.0, W), rg.uniform(0.2, 1.2, W), rg.integers(40, 40 + npitch, W), guess)
torch.cuda.synchronize()
print("done")

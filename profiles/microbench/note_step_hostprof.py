"""Host-side profile (cProfile) of the batched per-note step: where the wall time beyond the kernels goes."""
import cProfile, pstats, io, os, sys
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
def one(i, npitch=24):
    rg = np.random.default_rng(i)
    b.step(rg.uniform(0, 5.0, W), rg.uniform(0.2, 1.2, W), rg.integers(48, 48 + npitch, W), guess)
for i in range(3): one(i)
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for i in range(5): one(10 + i)
torch.cuda.synchronize()
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45); print(s.getvalue()[:9000])

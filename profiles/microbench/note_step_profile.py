"""Per-kernel time of ONE batched per-note step (600 windows) from torch.profiler (CUPTI activity records, no replay)."""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import amt_saga_b200  # noqa
from amt_saga_b200 import ops, synth
from amt_saga_b200.note_step import NoteStepBatch
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda"); W = 600; L = 264168
wav = synth.piano_batch(range(W), L, 44100, seed_base=50000, device=dev)
plan = ops.get_stft_plan(4096, 1024, True)
r = ops.stft_batch(wav, plan, want_phase=True)
b = NoteStepBatch(W)
b.load(r["mag_storage"][:, :258].contiguous(), r["phase_storage"][:, :258].contiguous(), wav, r["clip_max"], np.ones((W, 3)))
guess = synth.piano_batch(range(W), 54277, 44100, n_notes=1, seed_base=90000, device=dev)
npitch = int(sys.argv[1]) if len(sys.argv) > 1 else 24
def one(i):
    rg = np.random.default_rng(i)
    b.step(rg.uniform(0, 5.0, W), rg.uniform(0.2, 1.2, W), rg.integers(48, 48 + npitch, W), guess)
for i in range(3): one(i)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    one(7); torch.cuda.synchronize()
rows = [(e.key, e.count, e.device_time_total) for e in prof.key_averages() if e.device_time_total > 0 and e.device_type.name == "CUDA"]
rows.sort(key=lambda x: -x[2])
tot = sum(x[2] for x in rows)
print("total device time %.2f ms over %d kernel names" % (tot / 1e3, len(rows)))
for k, n, t in rows[:22]:
    print("%8.1f us  n=%4d  %s" % (t, n, k[:90]))

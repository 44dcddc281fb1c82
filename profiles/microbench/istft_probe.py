"""Small driver for profiling K4 (iSTFT, n_fft 2048 / hop 512, 200 windows of 6 s)."""
import sys, torch
sys.path.insert(0, "/root/repo")
import amt_saga_b200  # noqa: F401
from amt_saga_b200 import ops, synth
wav = synth.piano_batch(range(200), 264600, 44100, seed_base=50000, device="cuda")
plan = ops.get_stft_plan(2048, 512, True)
r = ops.stft_batch(wav, plan, want_phase=True)
for _ in range(2):
    ops.istft_batch(plan, mag=r["mag_storage"], phase=r["phase_storage"], n_bins=plan.n_bins)
torch.cuda.synchronize()
print("ok")

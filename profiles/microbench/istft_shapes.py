"""K4 iSTFT (librosa.istft of mag*phase, util_audio.py:92-104) on 600 windows of 6 s, bench shape and reference default."""
import sys, json, torch
sys.path.insert(0, "/root/repo")
import amt_saga_b200  # noqa: F401
from amt_saga_b200 import ops, synth
wav = synth.piano_batch(range(600), 264600, 44100, seed_base=50000, device="cuda")
def timed(fn, n=10):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
for n_fft, hop in ((2048, 512), (4096, 1024)):
    plan = ops.get_stft_plan(n_fft, hop, True)
    r = ops.stft_batch(wav, plan, want_phase=True)
    ms_f = timed(lambda: ops.stft_batch(wav, plan, want_phase=True))
    ms_i = timed(lambda: ops.istft_batch(plan, mag=r["mag_storage"], phase=r["phase_storage"], n_bins=plan.n_bins))
    T = plan.num_frames(264600)
    byt = 600 * (T * plan.n_bins * 12 + 4 * hop * (T - 1))      # mag + phasor in, samples out
    print(json.dumps({"n_fft": n_fft, "hop": hop, "stft_mag_phase_ms": round(ms_f, 3), "istft_ms": round(ms_i, 3),
                      "istft_GBps_algorithmic": round(byt / ms_i / 1e6, 1)}), flush=True)

"""The reference's finest per-note CQT (training.py:366-381: bins_per_tone=16 -> 192 bins per octave, 348 bins from the
note): does the plan build and run, how long does it take, does it agree with the oracle?  8 windows of 6 s."""
import sys, json, time, torch
sys.path.insert(0, "/root/repo")
import numpy as np
import amt_saga_b200  # noqa: F401
from amt_saga_b200 import ops, synth
from amt_saga_b200.util_audio import note_to_hz
from oracle import cqt as ocqt
W = 8
wav = synth.piano_batch(range(W), 264600, 44100, seed_base=50000, device="cuda")
t0 = time.time()
try:
    plan = ops.CqtPlan(44100, 1024, note_to_hz("C4"), 348, 192, filter_scale=2)
    t_plan = time.time() - t0
    r = ops.cqt_batch(wav, plan)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); r = ops.cqt_batch(wav, plan); b.record(); torch.cuda.synchronize()
    ref = np.abs(ocqt.cqt(wav[0].cpu().numpy(), sr=44100, hop_length=1024, fmin=note_to_hz("C4"), n_bins=348,
                          bins_per_octave=192, filter_scale=2))
    got = r["mag"][0].cpu().numpy()
    err = float(np.abs(got[:, :ref.shape[1]] - ref).max() / ref.max())
    print(json.dumps({"shape": "348 bins, 192 per octave from C4", "plan_s": round(t_plan, 2), "ms_for_8_windows": round(a.elapsed_time(b), 2),
                      "octaves": [(o["n_fft"], o["n_filters"]) for o in plan.octaves], "max_err_over_peak_vs_oracle": err}))
except Exception as e:
    print(json.dumps({"shape": "348 bins, 192 per octave from C4", "error": repr(e)[:300]}))

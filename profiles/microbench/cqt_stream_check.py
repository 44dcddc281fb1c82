"""Streamed-bank tcgen05 contraction (csrc/cqt_umma_stream.cu) against its fp32 CUDA-core twin (SAGA_CQT_STREAM=0) at
the producer loop's CQT shapes (training.py:340-388; hop 1024): whole transforms and 8-column frame windows of 600
windows, agreement as a fraction of the transform's peak and milliseconds of each."""
import sys, json, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import amt_saga_b200  # noqa: F401
from amt_saga_b200 import ops, synth
from amt_saga_b200.util_audio import note_to_hz
W = int(sys.argv[1]) if len(sys.argv) > 1 else 600
wav = synth.piano_batch(range(W), 264168, 44100, seed_base=50000, device="cuda")
def timed(fn, n=3):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
rng = np.random.default_rng(1)
for name, low, n_bins, bpo in (("87/12", "A0", 87, 12), ("174/24", "A0", 174, 24), ("348/48", "A0", 348, 48),
                                ("348/192 from C4", "C4", 348, 192), ("36/24 from D3", "D3", 36, 24)):
    plan = ops.CqtPlan(44100, 1024, note_to_hz(low), n_bins, bpo, filter_scale=2)
    T = plan.num_frames(264168)
    first = rng.integers(-3, T - 2, W).astype(np.int32)
    res = {"shape": name, "octaves": [(o["n_fft"], o["n_filters"], o["hop"]) for o in plan.octaves][:3]}
    full_s = ops.cqt_batch(wav, plan)["mag"].clone()
    fr_s = ops.cqt_frames_batch(wav, plan, first).clone()
    res["full_ms"] = round(timed(lambda: ops.cqt_batch(wav, plan)), 3)
    res["full_cascade_ms"] = round(timed(lambda: ops.cqt_batch(wav, plan, impl=0x100)), 3)
    res["frames_ms"] = round(timed(lambda: ops.cqt_frames_batch(wav, plan, first)), 3)
    with ops.options(SAGA_CQT_STREAM="0"):
        full_f = ops.cqt_batch(wav, plan, impl=1)["mag"].clone()
        fr_f = ops.cqt_frames_batch(wav, plan, first).clone()
        res["full_fp32_ms"] = round(timed(lambda: ops.cqt_batch(wav, plan, impl=1)), 3)
        res["frames_fp32_ms"] = round(timed(lambda: ops.cqt_frames_batch(wav, plan, first)), 3)
    peak = float(full_f.max())
    res["full_err_of_peak"] = float((full_s - full_f).abs().max()) / peak
    res["frames_err_of_peak"] = float((fr_s - fr_f).abs().max()) / peak
    print(json.dumps(res), flush=True)

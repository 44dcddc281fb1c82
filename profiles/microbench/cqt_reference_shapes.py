"""CQT at the reference's own call shapes (training.py:271-282, :340-388; filter_scale 2, hop 1024) on 600 x 6 s
windows.  12 bins per octave fit the resident-bank tcgen05 path; 24 / 48 per octave fall back to the fp32 CUDA-core
contraction (their banks do not fit shared memory: DESIGN.md section 7)."""
import sys, json, torch
sys.path.insert(0, "/root/repo")
import amt_saga_b200  # noqa: F401
from amt_saga_b200 import ops, synth
from amt_saga_b200.util_audio import note_to_hz
W = 600
wav = synth.piano_batch(range(W), 264600, 44100, seed_base=50000, device="cuda")
def timed(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
for name, low, n_bins, bpo in (("ref_C_1 87/12", "A0", 87, 12), ("C_sw_pitch 174/24", "A0", 174, 24),
                                ("ref_C_4 348/48", "A0", 348, 48), ("C_velocity 36/24 from D3", "D3", 36, 24)):
    try:
        plan = ops.CqtPlan(44100, 1024, note_to_hz(low), n_bins, bpo, filter_scale=2)
        ms = timed(lambda: ops.cqt_batch(wav, plan))
        ms_c = timed(lambda: ops.cqt_batch(wav, plan, impl=0x100))
        flops = 600 * plan.num_frames(264600) * sum(8 * o["n_filters"] * (o["n_fft"] // 2 + 1) for o in plan.octaves)
        print(json.dumps({"shape": name, "ms": round(ms, 3), "cascade_ms": round(ms_c, 3), "contract_ms": round(ms - ms_c, 3),
                          "algorithmic_TFLOPs_contract": round(flops / (ms - ms_c) / 1e9, 2),
                          "octaves": [(o["n_fft"], o["n_filters"]) for o in plan.octaves]}), flush=True)
    except Exception as e:
        print(json.dumps({"shape": name, "error": repr(e)[:200]}), flush=True)

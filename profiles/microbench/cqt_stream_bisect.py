"""Timing bisection of cqt_umma_stream_kernel: SAGA_UMMA_DEBUG switches off one role at a time (results are garbage,
only the time matters): 1 no MMAs, 2 no row copies, 8 no bank copies, 32 no lo conversion.  Contraction only
(the cascade runs once before)."""
import sys, json, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import amt_saga_b200  # noqa: F401
from amt_saga_b200 import ops, synth
from amt_saga_b200.util_audio import note_to_hz
wav = synth.piano_batch(range(600), 264168, 44100, seed_base=50000, device="cuda")
def timed(fn, n=3):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
for name, low, n_bins, bpo in (("174/24", "A0", 174, 24), ("348/48", "A0", 348, 48), ("348/192 from C4", "C4", 348, 192)):
    plan = ops.CqtPlan(44100, 1024, note_to_hz(low), n_bins, bpo, filter_scale=2)
    t_full = timed(lambda: ops.cqt_batch(wav, plan))
    t_casc = timed(lambda: ops.cqt_batch(wav, plan, impl=0x100))
    res = {"shape": name, "cascade_ms": round(t_casc, 3)}
    for dbg in (0, 1, 2, 8, 32, 2 | 32, 2 | 8 | 32, 1 | 2 | 8 | 32):
        with ops.options(SAGA_UMMA_DEBUG=str(dbg)):
            res["contract_ms_debug_%d" % dbg] = round(timed(lambda: ops.cqt_batch(wav, plan)) - t_casc, 3)
    print(json.dumps(res), flush=True)

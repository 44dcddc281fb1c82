"""Ring depth sweep of cqt_umma_stream_kernel (SAGA_UMMA_CFG caps the stages): is the time per stage set by the
hand-over latency (falls with depth) or by something serial per stage (flat)?"""
import sys, os, json
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
for name, low, n_bins, bpo in (("174/24", "A0", 174, 24), ("348/48", "A0", 348, 48)):
    plan = ops.CqtPlan(44100, 1024, note_to_hz(low), n_bins, bpo, filter_scale=2)
    t_casc = timed(lambda: ops.cqt_batch(wav, plan, impl=0x100))
    res = {"shape": name}
    for depth in (2, 3, 4, 5):
        with ops.options(SAGA_UMMA_CFG="%d,8" % depth):
            res["contract_ms_depth_%d" % depth] = round(timed(lambda: ops.cqt_batch(wav, plan)) - t_casc, 3)
    print(json.dumps(res), flush=True)

"""Rows in tensor memory (default where it fits) against rows in shared memory (SAGA_CQT_STREAM_SS=1), and ring-depth
caps, for the streamed-bank contraction: ms per 600 windows, contraction only."""
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
for name, low, n_bins, bpo in (("174/24", "A0", 174, 24), ("348/48", "A0", 348, 48), ("348/192 from C4", "C4", 348, 192)):
    plan = ops.CqtPlan(44100, 1024, note_to_hz(low), n_bins, bpo, filter_scale=2)
    t_casc = timed(lambda: ops.cqt_batch(wav, plan, impl=0x100))
    res = {"shape": name}
    for tag, opts in (("tmem", {}), ("smem", {"SAGA_CQT_STREAM_SS": "1"}), ("smem_depth2", {"SAGA_CQT_STREAM_SS": "1", "SAGA_UMMA_CFG": "2,8"}),
                      ("smem_depth3", {"SAGA_CQT_STREAM_SS": "1", "SAGA_UMMA_CFG": "3,8"}), ("tmem_depth2", {"SAGA_UMMA_CFG": "2,8"}),
                      ("tmem_depth3", {"SAGA_UMMA_CFG": "3,8"})):
        with ops.options(**opts):
            res["contract_ms_" + tag] = round(timed(lambda: ops.cqt_batch(wav, plan)) - t_casc, 3)
    print(json.dumps(res), flush=True)

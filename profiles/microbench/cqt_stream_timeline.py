import sys, os
sys.path.insert(0, "/root/repo")
import torch, amt_saga_b200
from amt_saga_b200 import ops, synth
from amt_saga_b200.util_audio import note_to_hz
wav = synth.piano_batch(range(600), 264168, 44100, seed_base=50000, device="cuda")
plan = ops.CqtPlan(44100, 1024, note_to_hz("A0"), 174, 24, filter_scale=2)
ops.cqt_batch(wav, plan); torch.cuda.synchronize()
import time
for ss in (None, "1"):
  for dbg in (128,):
    print("debug", dbg, "smem_A", ss, file=sys.stderr, flush=True)
    with ops.options(SAGA_UMMA_DEBUG=str(dbg), SAGA_CQT_STREAM_SS=ss):
        ops.cqt_batch(wav, plan); torch.cuda.synchronize()
    with ops.options(SAGA_UMMA_DEBUG=None, SAGA_CQT_STREAM_SS=ss):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ops.cqt_batch(wav, plan); a.record(); ops.cqt_batch(wav, plan); ops.cqt_batch(wav, plan); b.record(); torch.cuda.synchronize()
        print("ms per transform (cascade included): %.3f" % (a.elapsed_time(b) / 2), file=sys.stderr, flush=True)

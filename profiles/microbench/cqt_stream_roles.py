import sys, os
sys.path.insert(0, "/root/repo")
import torch, amt_saga_b200
from amt_saga_b200 import ops, synth
from amt_saga_b200.util_audio import note_to_hz
wav = synth.piano_batch(range(600), 264168, 44100, seed_base=50000, device="cuda")
for low, nb, bpo in (("A0", 174, 24), ("A0", 348, 48), ("C4", 348, 192)):
    plan = ops.CqtPlan(44100, 1024, note_to_hz(low), nb, bpo, filter_scale=2)
    ops.cqt_batch(wav, plan); torch.cuda.synchronize()
    for dbg in (16, 16 | 1 | 2 | 8 | 32, 16 | 1 | 2 | 8 | 32 | 64):
        print("shape", nb, bpo, "debug", dbg, file=sys.stderr, flush=True)
        with ops.options(SAGA_UMMA_DEBUG=str(dbg)):
            ops.cqt_batch(wav, plan); torch.cuda.synchronize()

import sys, os, subprocess
code = r'''
import sys
sys.path.insert(0, "/root/repo")
import torch, numpy as np, amt_saga_b200
from amt_saga_b200 import ops, synth
from amt_saga_b200.util_audio import note_to_hz
low, nb, bpo, mode = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
wav = synth.piano_batch(range(40), 264168, 44100, seed_base=50000, device="cuda")
plan = ops.CqtPlan(44100, 1024, note_to_hz(low), nb, bpo, filter_scale=2)
if mode == "full":
    r = ops.cqt_batch(wav, plan)["mag"]
else:
    r = ops.cqt_frames_batch(wav, plan, np.full(40, 50, dtype=np.int32))
torch.cuda.synchronize()
print("ok", low, nb, bpo, mode, float(r.abs().max()))
'''
for shape in (("A0", "87", "12"), ("A0", "174", "24"), ("A0", "348", "48"), ("C4", "348", "192")):
    for mode in ("frames", "full"):
        p = subprocess.run([sys.executable, "-c", code, *shape, mode], capture_output=True, text=True, timeout=120)
        print(shape, mode, "rc", p.returncode, (p.stdout.strip().splitlines() or ["-"])[-1][:100], (p.stderr.strip().splitlines() or ["-"])[-1][:120], flush=True)

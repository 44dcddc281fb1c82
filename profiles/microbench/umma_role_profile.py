"""Per-role cycle breakdown of the tcgen05 contraction (SAGA_UMMA_DEBUG=16): where each warp role of a
persistent CTA spends its cycles (mean over CTAs), for a few pipeline shapes."""
import os, subprocess, sys
code = r'''
import os, sys, torch
sys.path.insert(0, "/root/repo")
import amt_saga_b200
from amt_saga_b200 import ops, synth
from amt_saga_b200.util_audio import note_to_hz
wav = synth.piano_batch(range(600), 264600, seed_base=50000)
plan = ops.get_cqt_plan(44100, 512, note_to_hz("C1"), 84, 12, 2)
for _ in range(3): ops.cqt_batch(wav, plan, impl=2)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10): ops.cqt_batch(wav, plan, impl=2 | 0x200)
b.record(); torch.cuda.synchronize()
print("contraction %.3f ms" % (a.elapsed_time(b) / 10), flush=True)
os.environ["SAGA_UMMA_DEBUG"] = "16"
ops.cqt_batch(wav, plan, impl=2 | 0x200)
torch.cuda.synchronize()
'''
for cfg in ("0,4", "3,8"):
    env = dict(os.environ, SAGA_UMMA_CFG=cfg)
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    print("== cfg", cfg)
    print(r.stdout.strip())
    print("\n".join(l for l in r.stderr.splitlines() if "umma_prof" in l) or r.stderr[-2000:])

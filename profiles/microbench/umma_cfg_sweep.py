"""Sweep of the tcgen05 contraction's pipeline shape (SAGA_UMMA_CFG = stages,planes-per-stage,prefetch);
each configuration needs its own process because the plan reads the variable once."""
import os, subprocess, sys
code = r'''
import sys, torch
sys.path.insert(0, "/root/repo")
import amt_saga_b200
from amt_saga_b200 import ops, synth
from amt_saga_b200.util_audio import note_to_hz
wav = synth.piano_batch(range(600), 264600, seed_base=50000)
plan = ops.get_cqt_plan(44100, 512, note_to_hz("C1"), 84, 12, 2)
ref = ops.cqt_batch(wav[:4], plan, impl=1)["mag"]
got = ops.cqt_batch(wav[:4], plan, impl=2)["mag"]
err = float((ref - got).abs().max() / ref.max())
for _ in range(3): ops.cqt_batch(wav, plan, impl=2 | 0x200)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10): ops.cqt_batch(wav, plan, impl=2 | 0x200)
b.record(); torch.cuda.synchronize()
print("contraction %.3f ms   err vs fp32 path %.2e" % (a.elapsed_time(b) / 10, err))
'''
for cfg in ("2,8,1", "3,4,1", "3,4,2", "4,4,1", "4,4,2", "4,4,3"):
    env = dict(os.environ, SAGA_UMMA_CFG=cfg)
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    print(cfg, "->", (r.stdout.strip() or r.stderr.strip()[-300:]))

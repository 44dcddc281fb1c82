import os, subprocess, sys
code = r'''
import sys, os, torch
sys.path.insert(0, "/root/repo")
import amt_saga_b200
from amt_saga_b200 import ops, synth
from amt_saga_b200.util_audio import note_to_hz
wav = synth.piano_batch(range(600), 264600, seed_base=50000)
plan = ops.get_cqt_plan(44100, 512, note_to_hz("C1"), 84, 12, 2)
impl = int(os.environ.get("IMPL", "2"))
for _ in range(3): ops.cqt_batch(wav, plan, impl=impl | 0x200)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10): ops.cqt_batch(wav, plan, impl=impl | 0x200)
b.record(); torch.cuda.synchronize()
print("contraction %.3f ms" % (a.elapsed_time(b) / 10))
'''
for name, env in (("baseline 3xTF32", {}), ("no proxy fence (dbg 8)", {"SAGA_UMMA_DEBUG": "8"}), ("single pass impl=3", {"IMPL": "3"}),
                  ("single pass, no fence", {"IMPL": "3", "SAGA_UMMA_DEBUG": "8"}), ("no MMA, no fence (9)", {"SAGA_UMMA_DEBUG": "9"})):
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, **env), capture_output=True, text=True)
    print("%-28s %s" % (name, r.stdout.strip() or r.stderr.strip()[-200:]))

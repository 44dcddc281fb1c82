"""Which cascade levels to fuse (decimate2x2_kernel)?  Bench step (600 x 6 s windows), per fuse mask:
the overlapped loop, and the cascade / contraction stages of the serial pass.
SAGA_DEC_FUSE_MASK bit i = the pair whose first output is level i may be fused."""
import os, sys, json, torch
sys.path.insert(0, "/root/repo")
import amt_saga_b200  # noqa: F401
from amt_saga_b200 import synth
from amt_saga_b200.pipeline import WindowFeaturePipeline
import numpy as np
W = 600
pipe = WindowFeaturePipeline(W, 264600, 65024)
wav = synth.piano_batch(range(W), 264600, 44100, seed_base=50000, device="cuda")
guess = synth.piano_batch(range(W), 65024, 44100, n_notes=1, seed_base=90000, device="cuda")
offs = torch.as_tensor(np.random.default_rng(7).integers(0, 500, size=(W, 1)).astype(np.int32), device="cuda")
def loop(n=40):
    for _ in range(5): pipe.run(wav, guess, offs)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): pipe.run(wav, guess, offs)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
def stages(n=10):
    ev = []
    for _ in range(n): pipe.run(wav, guess, offs, ev)
    torch.cuda.synchronize()
    out = {}
    for name, a, b in ev: out[name] = out.get(name, 0.0) + a.elapsed_time(b) / n
    return out
masks = sys.argv[1:] or ["0", "0x1", "0x2", "0x4", "0x8", "0x14", "0x15", "0x5", "0xff"]
for rep in range(2):
    for m in masks:
        os.environ["SAGA_DEC_FUSE_MASK"] = m
        ms = loop()
        st = stages()
        print(json.dumps({"mask": m, "loop_ms": round(ms, 4), "cascade": round(st["cqt_cascade"], 4),
                          "contract": round(st["cqt_contract"], 4), "sum": round(sum(st.values()), 4)}), flush=True)

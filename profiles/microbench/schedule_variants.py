"""Stream schedules of the bench step (600 x 6 s windows): 'chains' (STFT->guess->chain->dB || cascade->contraction)
against 'db_with_contraction' (the tensor-core contraction held back until the chain is done, so that it runs beside
the HBM-bound dB pass).  Results must be bit-identical; prints loop ms per schedule."""
import os, sys, json, torch
sys.path.insert(0, "/root/repo")
import numpy as np
import amt_saga_b200  # noqa: F401
from amt_saga_b200 import synth
from amt_saga_b200.pipeline import WindowFeaturePipeline
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
ref = None
for rep in range(3):
    for sched in sys.argv[1:] or ["serial", "chains", "db_with_contraction"]:
        pipe.schedule = sched.split("+")[0]
        os.environ.pop("SAGA_DB_LEAN", None)
        if "+" in sched:                     # dB kernel variant (SAGA_DB_LEAN value)
            os.environ["SAGA_DB_LEAN"] = sched.split("+")[1]
        ms = loop()
        out = [pipe.mag[:, :pipe.T].clone(), pipe.D[:, :pipe.T].clone(), pipe.C.clone(), pipe.ref.clone()]
        if ref is None: ref = out
        same = all(torch.equal(x, y) for x, y in zip(out, ref))
        print(json.dumps({"schedule": sched, "loop_ms": round(ms, 4), "identical": same}), flush=True)

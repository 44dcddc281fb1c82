"""Does cutting the STFT -> guess STFT -> subtract -> dB chain into window chunks small enough for the 126 MB L2
(2.1 MB of magnitudes per window) pay?  The dB pass would then read the magnitudes from L2 instead of HBM.
CQT chain over the whole batch, once."""
import sys, json, torch
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
def step(chunk, cqt_chunked):
    if not cqt_chunked:
        pipe.run(wav, guess, offs, parts=("cqt",))
    for a in range(0, W, chunk):
        pipe.run(wav, guess, offs, w0=a, w1=min(W, a + chunk), parts=("stft", "cqt") if cqt_chunked else ("stft",))
def loop(chunk, cqt_chunked, n=30):
    for _ in range(4): step(chunk, cqt_chunked)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): step(chunk, cqt_chunked)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
ref = None
for rep in range(2):
    for chunk in (600, 300, 200, 148, 100, 74, 50, 37, 25):
        for cc in (False, True):
            ms = loop(chunk, cc)
            out = [pipe.mag[:, :pipe.T].clone(), pipe.D[:, :pipe.T].clone(), pipe.C.clone(), pipe.ref.clone()]
            if ref is None: ref = out
            same = all(torch.equal(x, y) for x, y in zip(out, ref))
            print(json.dumps({"chunk": chunk, "cqt_chunked": cc, "loop_ms": round(ms, 4), "identical": same}), flush=True)

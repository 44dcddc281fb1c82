"""Do the dB pass and the tensor-core contraction really overlap?  Event-timed on their own streams, alone and together."""
import os, sys, json, torch, ctypes as C
sys.path.insert(0, "/root/repo")
import numpy as np
import amt_saga_b200  # noqa: F401
from amt_saga_b200 import synth, _lib
from amt_saga_b200.pipeline import WindowFeaturePipeline
W = 600
pipe = WindowFeaturePipeline(W, 264600, 65024)
wav = synth.piano_batch(range(W), 264600, 44100, seed_base=50000, device="cuda")
guess = synth.piano_batch(range(W), 65024, 44100, n_notes=1, seed_base=90000, device="cuda")
offs = torch.as_tensor(np.random.default_rng(7).integers(0, 500, size=(W, 1)).astype(np.int32), device="cuda")
for _ in range(3): pipe.run(wav, guess, offs)
torch.cuda.synchronize()
lib = _lib.lib()
q = lambda t: C.c_void_p(t.data_ptr())
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def contraction(st):
    _lib.check(lib.saga_cqt_exec(pipe.cqt.handle, q(wav), q(pipe.offs_w), None, W, pipe.ns, q(pipe.C), None, pipe.Pc,
                                 pipe.Tc * pipe.Pc, q(pipe.ws), pipe.ws.numel(), 0x200, C.c_void_p(st.cuda_stream)))
def db(st):
    _lib.check(lib.saga_subtract_db_exec(q(pipe.mag), None, pipe.T_clip * pipe.P, q(pipe.gmag), None, pipe.Tg * pipe.P, None,
                                         pipe.Tg, q(offs), None, q(pipe.gmax), q(pipe.clip_max), q(pipe.frame_max), pipe.T_clip,
                                         3 | 0x200, q(pipe.D), q(pipe.ref), W, 1, pipe.nb, pipe.T, pipe.P, 1e-5, 80.0,
                                         C.c_void_p(st.cuda_stream)))
def timed(fa, fb, n=20):
    tot = [0.0, 0.0, 0.0]
    for _ in range(n):
        torch.cuda.synchronize()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record(); s1.wait_event(g0); s2.wait_event(g0)
        if fa: e[0].record(s1); fa(s1); e[1].record(s1)
        if fb: e[2].record(s2); fb(s2); e[3].record(s2)
        cur = torch.cuda.current_stream()
        if fa: cur.wait_event(e[1])
        if fb: cur.wait_event(e[3])
        g1.record(); torch.cuda.synchronize()
        tot[0] += e[0].elapsed_time(e[1]) if fa else 0
        tot[1] += e[2].elapsed_time(e[3]) if fb else 0
        tot[2] += g0.elapsed_time(g1)
    return [round(x / n, 4) for x in tot]
for lean, chunks in [(0, 0), (1, 0), (2, 0), (3, 0), (4, 0), (5, 0), (6, 0), (1, 2), (1, 8), (2, 8), (4, 2), (4, 8)]:
    os.environ.pop("SAGA_DB_LEAN", None); os.environ.pop("SAGA_DB_CHUNKS", None)
    if lean: os.environ["SAGA_DB_LEAN"] = str(lean)
    if chunks: os.environ["SAGA_DB_CHUNKS"] = str(chunks)
    print(json.dumps({"lean": lean, "chunks": chunks, "db_alone": timed(None, db)}), flush=True)

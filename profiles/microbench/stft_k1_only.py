"""Just K1 on the bench batch (600 x 6 s windows), a few launches: the ncu target."""
import os, sys, ctypes as C
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import amt_saga_b200  # noqa
from amt_saga_b200 import ops, synth, _lib
dev = torch.device("cuda")
plan = ops.get_stft_plan(2048, 512, True)
W, ns = 600, 264600
wav = synth.piano_batch(range(W), ns, 44100, seed_base=50000, device=dev)
P, T = ops.frame_pitch(1025), plan.num_frames(ns)
mag = torch.empty((W, T, P), device=dev); fmax = torch.empty((W, T), device=dev); cmax = torch.empty((W,), device=dev)
offs = torch.arange(W, device=dev, dtype=torch.int64) * ns
lens = torch.full((W,), ns, device=dev, dtype=torch.int64)
lib = _lib.lib(); st = C.c_void_p(torch.cuda.current_stream().cuda_stream); q = lambda t: C.c_void_p(t.data_ptr())
n = int(sys.argv[1]) if len(sys.argv) > 1 else 6
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(n):
    if i == n - 3: a.record()
    _lib.check(lib.saga_stft_exec(plan.handle, q(wav), q(offs), q(lens), W, ns, q(mag), None, None, P, T * P, q(fmax), q(cmax), st))
b.record(); torch.cuda.synchronize()
print("K1 %.3f ms/launch" % (a.elapsed_time(b) / 3))

"""BASELINE cfg4: generative-subtractive step on 4096 windows x 16 guessed-note spectrograms
([1025, 516] windows, [1025, 128] guesses, sequential per window, ReLU, then dB) through K3.
Random magnitudes of the right shapes (timing only; bit-exactness of the chain is a parity test)."""
import sys, json, torch
sys.path.insert(0, "/root/repo")
import amt_saga_b200  # noqa: F401
from amt_saga_b200 import ops
W, B, T, Tg, S = 4096, 1025, 516, 128, 16
P = ops.frame_pitch(B)
g = torch.Generator(device="cuda").manual_seed(1)
win = torch.rand((W, T, P), device="cuda", generator=g)
base = win.clone()
gue = torch.rand((W, S, Tg, P), device="cuda", generator=g)
offs = torch.randint(0, T, (W, S), device="cuda", generator=g, dtype=torch.int32)
D = torch.empty_like(win)
gref = gue.amax(dim=(2, 3))          # each guess's own ref_mag: a by-product of its STFT (K1 clip_max) in the real flow
def run():
    win.copy_(base)
    return ops.subtract_db_batch(win, gue, offs, B, D_out=D, guess_ref=gref)
run(); torch.cuda.synchronize()
a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
ms = []
for _ in range(5):
    win.copy_(base)
    a.record()
    ops.subtract_db_batch(win, gue, offs, B, D_out=D, guess_ref=gref)
    b.record(); torch.cuda.synchronize()
    ms.append(a.elapsed_time(b))
ms = sorted(ms)[len(ms) // 2]
alg = W * 4 * B * (2 * T + S * Tg)            # read mag + read 16 guesses + write one output
print(json.dumps({"config": "cfg4 4096 windows x 16 guesses", "ms": ms, "windows_per_s": W / ms * 1e3,
                  "algorithmic_GB": alg / 1e9, "GBps": alg / ms / 1e6, "frac_of_measured_hbm_6543": alg / ms / 1e6 / 6543.1}))

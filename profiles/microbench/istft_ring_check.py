"""K4 inverse ring kernel vs torch.istft (cuFFT) parity + time.  SAGA_ISTFT_RING=0 runs the first-generation kernel."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import amt_saga_b200  # noqa
from amt_saga_b200 import ops, synth
dev = torch.device("cuda"); mode = os.environ.get("SAGA_ISTFT_RING", "default")
win = torch.hann_window(2048, periodic=True, device=dev, dtype=torch.float64)
for center in (True, False):
    plan = ops.StftPlan(2048, 512, center)
    for W, ns in ((3, 2048), (2, 2048 + 512), (2, 2048 + 3 * 512 + 7), (4, 70000), (37, 264600)):
        wav = synth.piano_batch(range(W), ns, 44100, seed_base=77, device=dev)
        r = ops.stft_batch(wav, plan, want_phase=True, want_complex=True)
        T = r["mag"].shape[2]
        ok = None
        F = r["F"].to(torch.complex128)                                   # [W, bins, T]
        if center:
            ref = torch.istft(F, 2048, 512, window=win, center=True, length=512 * (T - 1))
        else:
            # torch.istft(center=False) refuses windows that are 0 at the edge: do the overlap-add by hand
            fr = torch.fft.irfft(F.transpose(1, 2), n=2048) * win          # [W, T, 2048]
            ref = torch.zeros((W, 2048 + 512 * (T - 1)), device=dev, dtype=torch.float64)
            wss = torch.zeros_like(ref[0])
            for t in range(T):
                ref[:, t * 512:t * 512 + 2048] += fr[:, t]
                wss[t * 512:t * 512 + 2048] += win.float().double() ** 2
            ref = torch.where(wss > 1e-30, ref / wss.clamp_min(1e-30), ref)
            ok = wss > 1e-2          # 1 / w^2 amplifies fp32 rounding without bound where the window vanishes: compare elsewhere
        for kind in ("F", "magphase"):
            if kind == "F":
                y = ops.istft_batch(plan, F=r["F_storage"])
            else:
                y = ops.istft_batch(plan, mag=r["mag_storage"], phase=r["phase_storage"])
            assert y.shape == ref.shape, (y.shape, ref.shape)
            d = (y.double() - ref).abs()
            if ok is not None:
                d = d[:, ok]
            err = float(d.max() / ref.abs().max())
            print("center=%d W=%d ns=%d T=%d %s mode=%s rel err %.3e" % (center, W, ns, T, kind, mode, err), flush=True)
            assert err < 5e-6, err
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / n
plan = ops.get_stft_plan(2048, 512, True)
W, ns = 600, 264600
wav = synth.piano_batch(range(W), ns, 44100, seed_base=50000, device=dev)
r = ops.stft_batch(wav, plan, want_phase=True)
mag, ph = r["mag_storage"][:, :516].contiguous(), r["phase_storage"][:, :516].contiguous()
ms = timeit(lambda: ops.istft_batch(plan, mag=mag, phase=ph))
byt = W * 516 * (12 * 1025 + 4 * 512)
print("K4 600 windows x 516 frames (mag+phase in) mode=%s: %.3f ms  %.0f GB/s  frac %.3f" % (mode, ms, byt / ms / 1e6, byt / ms / 1e6 / 6543.1))

"""K1 ring kernel vs the first-generation kernel vs cuFFT (torch.stft + abs): parity + time on the bench shape.
Run:  SAGA_STFT_RING=1 python profiles/microbench/stft_ring_check.py   (and with =0 for the old kernel)"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import amt_saga_b200  # noqa
from amt_saga_b200 import ops, synth

dev = torch.device("cuda")
mode = os.environ.get("SAGA_STFT_RING", "default")
plan = ops.get_stft_plan(2048, 512, True)


def timeit(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def ref_stft(wav):
    F = torch.stft(wav.double(), 2048, 512, window=torch.hann_window(2048, periodic=True, device=dev, dtype=torch.float64),
                   center=True, pad_mode="reflect", return_complex=True)
    return F      # [clips, bins, T]


for name, W, ns in (("small ragged", 5, 70000), ("guess", 64, 65024), ("window", 48, 264600)):
    wav = synth.piano_batch(range(W), ns, 44100, seed_base=123, device=dev)
    lens = None
    if name == "small ragged":
        lens = torch.tensor([70000, 1100, 1537, 40000, 2047])
    r = ops.stft_batch(wav, plan, lens=lens, want_phase=True, want_complex=True)
    torch.cuda.synchronize()
    worst = 0.0
    for c in range(W):
        n = int(lens[c]) if lens is not None else ns
        F = ref_stft(wav[c:c + 1, :n])[0]
        T = F.shape[1]
        peak = F.abs().max().item()
        e_mag = (r["mag"][c, :, :T].double() - F.abs()).abs().max().item() / peak
        e_F = (r["F"][c, :, :T].to(torch.complex128) - F).abs().max().item() / peak
        e_max = abs(r["clip_max"][c].item() - peak) / peak
        fm = (r["frame_max"][c, :T].double() - F.abs().amax(dim=0)).abs().max().item() / peak
        pad = r["mag_storage"][c, :T, 1025:].abs().max().item()
        ph = r["phase"][c, :, :T].to(torch.complex128)
        strong = F.abs() > 1e-3 * peak
        e_ph = (ph - F / F.abs().clamp_min(1e-300))[strong].abs().max().item()
        worst = max(worst, e_mag, e_F, e_max, fm, pad)
        assert e_ph < 1e-3, e_ph
    print("%-12s mode=%s  worst rel err %.3e" % (name, mode, worst), flush=True)
    assert worst < 2e-6, worst

W, ns = 600, 264600
wav = synth.piano_batch(range(W), ns, 44100, seed_base=50000, device=dev)
P = ops.frame_pitch(1025)
T = plan.num_frames(ns)
mag = torch.empty((W, T, P), device=dev)
fmax = torch.empty((W, T), device=dev)
cmax = torch.empty((W,), device=dev)
offs = torch.arange(W, device=dev, dtype=torch.int64) * ns
lens = torch.full((W,), ns, device=dev, dtype=torch.int64)
import ctypes as C
from amt_saga_b200 import _lib
lib = _lib.lib()
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
q = lambda t: C.c_void_p(t.data_ptr())
def k1():
    _lib.check(lib.saga_stft_exec(plan.handle, q(wav), q(offs), q(lens), W, ns, q(mag), None, None, P, T * P, q(fmax), q(cmax), st))
ms = timeit(k1)
bytes_ = W * 4 * (ns + T * 1025)
print("K1 600 windows mode=%s: %.3f ms  %.0f GB/s  frac %.3f" % (mode, ms, bytes_ / ms / 1e6, bytes_ / ms / 1e6 / 6543.1), flush=True)
gw = synth.piano_batch(range(W), 65024, 44100, n_notes=1, seed_base=90000, device=dev)
Tg = plan.num_frames(65024)
gm = torch.empty((W, Tg, P), device=dev)
og = torch.arange(W, device=dev, dtype=torch.int64) * 65024
lg = torch.full((W,), 65024, device=dev, dtype=torch.int64)
def k1g():
    _lib.check(lib.saga_stft_exec(plan.handle, q(gw), q(og), q(lg), W, 65024, q(gm), None, None, P, Tg * P, None, q(cmax), st))
msg = timeit(k1g)
bg = W * 4 * (65024 + Tg * 1025)
print("K1 600 guesses mode=%s: %.3f ms  %.0f GB/s  frac %.3f" % (mode, msg, bg / msg / 1e6, bg / msg / 1e6 / 6543.1), flush=True)
if mode != "0":
    win = torch.hann_window(2048, periodic=True, device=dev)
    def cufft():
        return torch.stft(wav, 2048, 512, window=win, center=True, pad_mode="reflect", return_complex=True).abs()
    print("cuFFT yardstick torch.stft(...).abs() on the same 600 windows: %.3f ms" % timeit(cufft, n=5, warm=2), flush=True)

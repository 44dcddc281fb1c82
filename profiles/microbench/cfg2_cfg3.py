"""BASELINE cfg2 / cfg3 at their own shape: 1 h of synthetic 44.1 kHz audio as 360 clips x 441 000 samples.
cfg2 = batched STFT magnitude (n_fft 2048, hop 512); cfg3 = CQT 84 bins / 12 per octave (tensor path)."""
import sys, json, torch
sys.path.insert(0, "/root/repo")
import amt_saga_b200  # noqa: F401
from amt_saga_b200 import ops, synth
from amt_saga_b200.util_audio import note_to_hz
wav = synth.piano_batch(range(360), 441000, seed_base=1234)
sp = ops.get_stft_plan(2048, 512, True)
cp = ops.get_cqt_plan(44100, 512, note_to_hz("C1"), 84, 12, 2)
def timed(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
sp4 = ops.get_stft_plan(4096, 1024, True)          # the reference's own default shape (main.py: N=4096, hop=1024)
ms4 = timed(lambda: ops.stft_batch(wav, sp4))
T4 = sp4.num_frames(441000)
T = sp.num_frames(441000)
ms2 = timed(lambda: ops.stft_batch(wav, sp))
ms3 = timed(lambda: ops.cqt_batch(wav, cp))
ms3c = timed(lambda: ops.cqt_batch(wav, cp, impl=0x100))
frames = 360 * T
print(json.dumps({"frames": frames, "refdefault_4096_1024_stft_ms": ms4, "refdefault_frames": 360 * T4,
                  "refdefault_GBps": 360 * T4 * 12292 / ms4 / 1e6,
                  "cfg2_stft_ms": ms2, "cfg2_GBps": frames * 6148 / ms2 / 1e6, "cfg2_frac_hbm_6543": frames * 6148 / ms2 / 1e6 / 6543.1,
                  "cfg3_cqt_ms": ms3, "cfg3_cascade_ms": ms3c, "cfg3_contract_ms": ms3 - ms3c,
                  "cfg3_alg_TFLOPs_contract": frames * 172704 / (ms3 - ms3c) / 1e9, "cfg3_Mframes_per_s": frames / ms3 / 1e3}))

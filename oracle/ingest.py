"""Oracle (TEST INFRASTRUCTURE): how the reference turns 16-bit PCM into the waveform that
`audio_complete` analyses -- the arithmetic K0 (`saga_pcm16_ingest_exec`) has to reproduce bit for bit.

  * /root/reference/util_audio.py:889-897  `_fluidsynth_instrument_stateful`: per event
    `samples = fl.get_samples(n)[::2]` (pyfluidsynth returns interleaved stereo int16; `[::2]` keeps the
    left channel), accumulated into a float64 array;
  * /root/reference/util_audio.py:776-781  `note_sequence.render`:
    `vel_max = max(velocities)`; a single note uses `max(1, vel_max - 12)`;
    `wf = wf * (vel_max/128.0)**4 / np.abs(wf).max()`  -- float64, multiply first, then divide;
  * /root/reference/util_audio.py:964  `librosa.load(sr=None)` -> soundfile: `int16 / 32768` as float32.

`audio_complete` then hands the array to librosa.stft; this repo's container casts it to float32 once
(amt-saga_b200/util_audio.py `_wave_in`), so the quantity to match is float32(float64 result).
Parity unpinned against reference outputs (fluidsynth and a soundfont are absent); the arithmetic is three
IEEE operations stated in the reference's own source.
"""
import numpy as np


def left_channel(interleaved):
    """util_audio.py:894: `get_samples(n)[::2]`."""
    return np.asarray(interleaved)[..., ::2]


def render_scale(pcm, velocities):
    """util_audio.py:778-781 -> (mul, div) of `wf * mul / div` for one rendered sequence."""
    vel_max = max(velocities)
    if len(velocities) == 1:
        vel_max = max(1, vel_max - 12)
    return (vel_max / 128.0) ** 4, float(np.abs(np.asarray(pcm, dtype=np.float64)).max())


def pcm_to_wave(pcm, mul=1.0, div=32768.0):
    """float64 `(pcm * mul) / div` (the reference's operation order), per clip when mul/div are arrays."""
    x = np.asarray(pcm).astype(np.float64)
    mul = np.asarray(mul, dtype=np.float64)
    div = np.asarray(div, dtype=np.float64)
    if x.ndim == 2:
        mul = mul.reshape(-1, 1) if mul.ndim else mul
        div = div.reshape(-1, 1) if div.ndim else div
    with np.errstate(invalid="ignore", divide="ignore"):
        return (x * mul) / div


def render(pcm, velocities):
    """One rendered note / sequence from its fluidsynth PCM: util_audio.py:776-781."""
    mul, div = render_scale(pcm, velocities)
    return pcm_to_wave(pcm, mul, div)

"""Oracle (TEST INFRASTRUCTURE): numpy restatement of the librosa-0.6.3-era
spectral primitives that `/root/reference/util_audio.py` calls.

PARITY UNPINNED against literal librosa (not installed, no network); pinned
against torch.stft / transformers.audio_utils in tests/test_oracle_*.py.

Call sites restated (reference file:line -> function here):
  util_audio.py:127-128  librosa.stft              -> stft
  util_audio.py:92-104   librosa.istft             -> istft
  util_audio.py:147,162  librosa.core.magphase     -> magphase
  util_audio.py:179      librosa.amplitude_to_db   -> amplitude_to_db
  util_audio.py:101,124  librosa.db_to_amplitude   -> db_to_amplitude
  util_audio.py:67       librosa.fft_frequencies   -> fft_frequencies
  util_audio.py:281      librosa.midi_to_hz        -> midi_to_hz
  util_audio.py:331      feature.spectral_flatness -> spectral_flatness
  util_audio.py:421-425  note_to_midi / note_to_hz -> note_to_midi, note_to_hz
  training.py:370,378    librosa.midi_to_note      -> midi_to_note
"""
import re

import numpy as np


# ----------------------------------------------------------------------------
# windows / framing
# ----------------------------------------------------------------------------
def get_window(name, n):
    """scipy.signal.get_window(name, n, fftbins=True) for the two windows the
    path uses: periodic Hann (stft/istft/cqt filters) and 'ones' (the
    rectangular STFT inside the CQT response)."""
    n = int(n)
    if name in ("ones", "boxcar", "rect"):
        return np.ones(n, dtype=np.float64)
    if name in ("hann", "hanning"):
        if n == 1:
            return np.ones(1, dtype=np.float64)
        return 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n, dtype=np.float64) / n)
    raise ValueError("window %r not supported by the oracle" % (name,))


def reflect_index(i, length):
    """Index map of np.pad(mode='reflect') (edge sample not repeated, keeps
    bouncing when the pad is longer than the signal; probe-verified equal to
    numpy for every (length, pad) tried in tests/test_oracle_stft.py)."""
    i = np.asarray(i, dtype=np.int64)
    if length == 1:
        return np.zeros_like(i)
    period = 2 * (length - 1)
    m = np.mod(i, period)
    return np.where(m < length, m, period - m)


def frame(y, frame_length, hop_length):
    """librosa.util.frame: [frame_length, n_frames] strided view."""
    if y.shape[0] < frame_length:
        raise ValueError(
            "Buffer is too short (n=%d) for frame_length=%d" % (y.shape[0], frame_length)
        )
    n_frames = 1 + (y.shape[0] - frame_length) // hop_length
    idx = np.arange(frame_length)[:, None] + hop_length * np.arange(n_frames)[None, :]
    return y[idx]


def pad_center(data, size):
    n = data.shape[-1]
    lpad = int((size - n) // 2)
    out = np.zeros(size, dtype=data.dtype)
    out[lpad : lpad + n] = data
    return out


# ----------------------------------------------------------------------------
# STFT family
# ----------------------------------------------------------------------------
def stft(y, n_fft=2048, hop_length=None, window="hann", center=True,
         dtype=np.complex64, pad_mode="reflect"):
    """librosa.stft, 0.6.3 semantics.

    float64 window * frames, FFT in float64, result stored as `dtype`
    (complex64 by default, as in 0.6.x/0.7.x).  Shape [1 + n_fft/2, T] with
    T = 1 + len(y)//hop when centred.
    """
    y = np.asarray(y)
    if y.ndim != 1:
        raise ValueError("oracle stft takes mono audio")
    if hop_length is None:
        hop_length = int(n_fft // 4)
    win = pad_center(get_window(window, n_fft), n_fft).reshape(-1, 1)
    if center:
        if pad_mode != "reflect":
            raise ValueError("only reflect padding on this path")
        y = np.pad(y, int(n_fft // 2), mode="reflect")
    frames = frame(y, n_fft, hop_length)
    spec = np.fft.rfft(win * frames.astype(np.float64), axis=0)
    return np.asfortranarray(spec.astype(dtype))


def window_sumsquare(window, n_frames, hop_length, n_fft, dtype=np.float32):
    n = n_fft + hop_length * (n_frames - 1)
    x = np.zeros(n, dtype=dtype)
    win_sq = pad_center(get_window(window, n_fft) ** 2, n_fft)
    for i in range(n_frames):
        s = i * hop_length
        x[s : min(n, s + n_fft)] += win_sq[: max(0, min(n_fft, n - s))]
    return x


def istft(stft_matrix, hop_length=None, window="hann", center=True,
          dtype=np.float32):
    """librosa.istft, 0.6.3 semantics: overlap-add of
    window * irfft(column), accumulated in `dtype`, divided by the window
    sum-of-squares where that exceeds tiny, trimmed by n_fft/2 at both ends
    when centred => length hop*(T-1)."""
    n_fft = 2 * (stft_matrix.shape[0] - 1)
    if hop_length is None:
        hop_length = int(n_fft // 4)
    win = pad_center(get_window(window, n_fft), n_fft)
    n_frames = stft_matrix.shape[1]
    y = np.zeros(n_fft + hop_length * (n_frames - 1), dtype=dtype)
    cols = np.fft.irfft(np.asarray(stft_matrix, dtype=np.complex128), n=n_fft, axis=0)
    for i in range(n_frames):
        s = i * hop_length
        y[s : s + n_fft] = y[s : s + n_fft] + win * cols[:, i]
    wss = window_sumsquare(window, n_frames, hop_length, n_fft, dtype=dtype)
    nz = wss > np.finfo(wss.dtype).tiny
    y[nz] /= wss[nz]
    if center:
        y = y[int(n_fft // 2) : -int(n_fft // 2)]
    return y


def magphase(D):
    """librosa.core.magphase(D, power=1): abs and exp(1j*angle) (so a zero
    bin has phase 1+0j)."""
    mag = np.abs(D)
    phase = np.exp(1.0j * np.angle(D)).astype(
        np.complex64 if D.dtype == np.complex64 else np.complex128)
    return mag, phase


def amplitude_to_db(S, ref=1.0, amin=1e-5, top_db=80.0):
    """librosa.amplitude_to_db == power_to_db(S**2, ref**2, amin**2, top_db)
    (SURVEY Appendix A.3)."""
    magnitude = np.abs(np.asarray(S))
    ref_value = ref(magnitude) if callable(ref) else np.abs(ref)
    power = np.square(magnitude)
    log_spec = 10.0 * np.log10(np.maximum(amin ** 2, power))
    log_spec -= 10.0 * np.log10(np.maximum(amin ** 2, ref_value ** 2))
    if top_db is not None:
        log_spec = np.maximum(log_spec, log_spec.max() - top_db)
    return log_spec


def db_to_amplitude(S_db, ref=1.0):
    return (ref ** 2 * np.power(10.0, 0.1 * np.asarray(S_db))) ** 0.5


def spectral_flatness(y, n_fft=2048, hop_length=512, amin=1e-10, power=2.0):
    """librosa.feature.spectral_flatness -> [1, T]."""
    S = np.abs(stft(y, n_fft=n_fft, hop_length=hop_length))
    S_thresh = np.maximum(amin, S.astype(np.float64) ** power)
    gmean = np.exp(np.mean(np.log(S_thresh), axis=0, keepdims=True))
    amean = np.mean(S_thresh, axis=0, keepdims=True)
    return gmean / amean


# ----------------------------------------------------------------------------
# frequency / note helpers
# ----------------------------------------------------------------------------
def fft_frequencies(sr=22050, n_fft=2048):
    return np.linspace(0, float(sr) / 2, int(1 + n_fft // 2), endpoint=True)


def midi_to_hz(notes):
    return 440.0 * (2.0 ** ((np.asanyarray(notes) - 69.0) / 12.0))


_PITCH = {"C": 0, "D": 2, "E": 4, "F": 5, "G": 7, "A": 9, "B": 11}
_ACC = {"#": 1, "": 0, "b": -1, "!": -1}
_NOTE_RE = re.compile(
    r"^(?P<note>[A-Ga-g])(?P<accidental>[#b!]*)(?P<octave>[+-]?\d+)?(?P<cents>[+-]\d+)?$")
_NOTE_NAMES = ["C", "C#", "D", "D#", "E", "F", "F#", "G", "G#", "A", "A#", "B"]


def note_to_midi(note, round_midi=True):
    m = _NOTE_RE.match(note)
    if not m:
        raise ValueError("Improper note format: %r" % (note,))
    pitch = m.group("note").upper()
    offset = sum(_ACC[o] for o in m.group("accidental"))
    octave = int(m.group("octave")) if m.group("octave") else 0
    cents = int(m.group("cents")) * 1e-2 if m.group("cents") else 0
    value = 12 * (octave + 1) + _PITCH[pitch] + offset + cents
    return int(np.round(value)) if round_midi else value


def note_to_hz(note):
    return float(midi_to_hz(note_to_midi(note, round_midi=False)))


def midi_to_note(midi, octave=True):
    num = int(np.round(midi))
    name = _NOTE_NAMES[num % 12]
    if octave:
        name = "%s%d" % (name, int(num / 12) - 1)
    return name

"""GPU-resident `audio_complete`: the reference's lazy spectral container
(/root/reference/util_audio.py:32-527) with the same constructor, properties,
methods, `[bins, frames]` shapes and error behaviour, but with CUDA tensors as
storage and the arithmetic done by libsaga_b200.so:

    F / mag / ph        -> K1  saga_stft_exec           (util_audio.py:116-168)
    wf from spectra     -> K4  saga_istft_exec          (util_audio.py:88-106)
    ref_mag, subtract   -> K3  saga_subtract_db_exec    (util_audio.py:170-174, 221-259)
    D                   ->     saga_amplitude_to_db_exec (util_audio.py:176-180)
    slice_C             -> K2  saga_cqt_exec            (util_audio.py:411-434)

`carrier='torch'` (default) hands out CUDA tensors; `carrier='numpy'` copies
results to host arrays so the reference's training loop and classifiers can
consume them unchanged (training.py:333-388, util_train_test.py:114-146).
Slicing / concatenation / tiling helpers are pure data movement and use torch
indexing.  There is no CPU compute path.
"""
import bisect
import math

import numpy as np
import torch

from . import ops
from .cqt_plan import ParameterError  # noqa: F401

_NOTE_BASE = {"C": 0, "D": 2, "E": 4, "F": 5, "G": 7, "A": 9, "B": 11}
_NOTE_NAMES = ["C", "C#", "D", "D#", "E", "F", "F#", "G", "G#", "A", "A#", "B"]


def note_to_midi(note):
    """librosa.note_to_midi for names like 'A0', 'C#4', 'Bb3'."""
    name, rest = note[0].upper(), note[1:]
    acc = 0
    while rest and rest[0] in "#b!":
        acc += 1 if rest[0] == "#" else -1
        rest = rest[1:]
    if name not in _NOTE_BASE:
        raise ParameterError("Improper note format: %r" % (note,))
    octave = int(rest) if rest else 0
    return 12 * (octave + 1) + _NOTE_BASE[name] + acc


def midi_to_hz(m):
    return 440.0 * (2.0 ** ((float(m) - 69.0) / 12.0))


def note_to_hz(note):
    return midi_to_hz(note_to_midi(note))


def midi_to_note(midi):
    n = int(round(midi))
    return "%s%d" % (_NOTE_NAMES[n % 12], int(n / 12) - 1)


def _device(device=None):
    if not torch.cuda.is_available():
        raise RuntimeError("amt_saga_b200 needs a CUDA device: there is no CPU fallback")
    d = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    if d.type != "cuda":
        raise RuntimeError("amt_saga_b200 needs a CUDA device: there is no CPU fallback")
    return d if d.index is not None else torch.device("cuda", torch.cuda.current_device())


class _Spec:
    """A [bins, frames] quantity held as frame-major storage `st` = [T, P]
    (rows = frames, P = pitch >= nb, 16-byte aligned rows, zero padding)."""
    __slots__ = ("st", "nb")

    def __init__(self, st, nb):
        self.st, self.nb = st, nb

    @property
    def view(self):
        return self.st[:, :self.nb].transpose(0, 1)

    @property
    def T(self):
        return self.st.shape[0]

    @property
    def shape(self):
        return (self.nb, self.st.shape[0])

    def rows(self, a, b):
        return _Spec(self.st[a:b], self.nb)

    def clone(self):
        return _Spec(self.st.clone(), self.nb)

    @staticmethod
    def from_view(x, dtype, dev):
        t = torch.as_tensor(np.asarray(x) if not isinstance(x, torch.Tensor) else x)
        t = t.to(device=dev, dtype=dtype)
        if t.dim() != 2:
            raise ValueError("spectral attributes are [bins, frames]")
        nb, T = t.shape
        st = torch.zeros((T, ops.frame_pitch(nb)), device=dev, dtype=dtype)
        st[:, :nb] = t.transpose(0, 1)
        return _Spec(st, nb)


class audio_complete:
    """Drop-in for util_audio.audio_complete (same signature, util_audio.py:33)."""

    def __init__(self, waveform, n_fft, hop_length=None, center=True, sample_rate=44100,
                 device=None, carrier="torch"):
        self._dev = _device(device)
        self._carrier = carrier
        self._wf = self._wave_in(waveform)
        self._F = self._mag = self._ph = self._D = None   # _Spec or None
        self._ref_mag = None      # python float once evaluated (reference: np.float32 scalar)
        self._ref_dev = None      # 1-element CUDA tensor mirror of _ref_mag (no host sync)
        self._max_hint = None     # device max of the CURRENT _mag (by-product of K1/K3) or None
        self.sr = sample_rate
        self.N = n_fft
        self.center = center
        self.hl = hop_length if hop_length is not None else int(np.floor(n_fft / 4))
        self._fft_freq = np.linspace(0, float(sample_rate) / 2, int(1 + n_fft // 2), endpoint=True)

    @classmethod
    def from_pcm16(cls, pcm, n_fft, hop_length=None, center=True, sample_rate=44100, mul=1.0, div=32768.0,
                   channel_stride=1, device=None, carrier="torch"):
        """Container over 16-bit PCM scaled on the device (K0): waveform = float32((float64(pcm) * mul) / div),
        the one scaling the reference applies to fluidsynth frames (`wf*(vel_max/128.0)**4/np.abs(wf).max()`,
        util_audio.py:776-781; pass div='peak') or to a decoded file (`/ 32768`, util_audio.py:964).
        channel_stride=2 keeps the left channel of interleaved stereo like util_audio.py:894 `[::2]`."""
        dev = _device(device)
        t = torch.as_tensor(np.asarray(pcm) if not isinstance(pcm, torch.Tensor) else pcm)
        if t.dtype != torch.int16:
            raise TypeError("from_pcm16 expects int16 samples")
        t = t.reshape(1, -1).to(dev).contiguous()
        if div == "peak":
            div = ops.pcm16_absmax(t, channel_stride)
        wave = ops.pcm16_to_wave(t, mul=mul, div=div, channel_stride=channel_stride)[0]
        return cls(wave, n_fft, hop_length, center, sample_rate, device=dev, carrier=carrier)

    # ------------------------------------------------------------------ plumbing
    def _wave_in(self, w):
        if w is None:
            return None
        t = torch.as_tensor(np.asarray(w) if not isinstance(w, torch.Tensor) else w)
        return t.to(device=self._dev, dtype=torch.float32).reshape(-1).contiguous()

    def _out(self, x):
        if isinstance(x, _Spec):
            x = x.view
        if x is None or self._carrier == "torch":
            return x
        return x.detach().cpu().numpy()

    def _plan(self):
        return ops.get_stft_plan(int(self.N), int(self.hl), bool(self.center), device=self._dev)

    def _analyse(self):
        """K1 on the waveform: fills mag and ph (and the max by-product)."""
        if self._wf is None:
            return False
        if self._wf.numel() == 0:
            raise ParameterError("Audio buffer is empty")
        if not self.center and self._wf.numel() < self.N:
            raise ParameterError("Buffer is too short (n=%d) for frame_length=%d"
                                 % (self._wf.numel(), self.N))
        plan = self._plan()
        r = ops.stft_batch(self._wf, plan, want_phase=True, want_max=True)
        self._mag = _Spec(r["mag_storage"][0], plan.n_bins)
        self._ph = _Spec(r["phase_storage"][0], plan.n_bins)
        self._max_hint = r["clip_max"]
        return True

    # ------------------------------------------------------------------ lazy fields
    @property
    def wf(self):
        return self._out(self._wave())

    def _wave(self):
        if self._wf is None:
            if self._F is not None:
                self._wf = ops.istft_batch(self._plan(), F=self._F.st.unsqueeze(0), n_bins=self._F.nb)[0]
            else:
                if not (self._mag is not None and self._ph is not None) and \
                        (self._D is not None and self._ph is not None):
                    self._mag_from_db()
                if self._mag is not None and self._ph is not None:
                    self._wf = ops.istft_batch(self._plan(), mag=self._mag.st.unsqueeze(0),
                                               phase=self._ph.st.unsqueeze(0), n_bins=self._mag.nb)[0]
        return self._wf

    @wf.setter
    def wf(self, value):
        self._D = self._mag = self._ph = self._F = None
        self._set_ref(None)
        self._max_hint = None
        self._wf = self._wave_in(value)

    def _mag_from_db(self):
        # librosa.db_to_amplitude(D, ref) = ref * 10**(D/20)   (util_audio.py:99-101)
        if not self._ref_cached():
            self._set_ref(1.0)
        st = self._ref_device() * torch.pow(10.0, 0.05 * self._D.st)
        st[:, self._D.nb:] = 0
        self._mag = _Spec(st, self._D.nb)
        self._max_hint = None

    def _spec_F(self):
        if self._F is None:
            if not (self._mag is not None and self._ph is not None) and \
                    (self._D is not None and self._ph is not None):
                self._mag_from_db()
            if self._mag is not None and self._ph is not None:
                self._F = _Spec(self._mag.st * self._ph.st, self._mag.nb)
            elif self._wave() is not None:
                plan = self._plan()
                r = ops.stft_batch(self._wf, plan, want_complex=True, want_max=False)
                self._F = _Spec(r["F_storage"][0], plan.n_bins)
        return self._F

    @property
    def F(self):
        return self._out(self._spec_F())

    @F.setter
    def F(self, value):
        self._D = self._mag = self._ph = self._wf = None
        self._set_ref(None)
        self._max_hint = None
        self._F = _Spec.from_view(value, torch.complex64, self._dev)

    def _magphase_from_F(self):
        F = self._F.st
        mag = torch.abs(F)
        ph = torch.where(mag > 0, F / mag.clamp_min(torch.finfo(torch.float32).tiny),
                         torch.ones_like(F))
        ph[:, self._F.nb:] = 0
        self._mag, self._ph = _Spec(mag, self._F.nb), _Spec(ph, self._F.nb)
        self._max_hint = None

    def _spec_mag(self):
        if self._mag is None:
            if self._D is not None and self._ph is not None:
                self._mag_from_db()
            elif self._F is not None:
                self._magphase_from_F()
            elif not self._analyse():
                raise AttributeError("audio_complete holds no data")
        return self._mag

    @property
    def mag(self):
        return self._out(self._spec_mag())

    @mag.setter
    def mag(self, val):
        self._D = self._F = self._wf = None
        self._set_ref(None)
        self._max_hint = None
        self._mag = _Spec.from_view(val, torch.float32, self._dev)
        if self._ph is not None and self._ph.shape != self._mag.shape:
            self._ph = None

    def _spec_ph(self):
        if self._ph is None:
            if self._F is not None:
                self._magphase_from_F()
            elif not self._analyse():
                raise AttributeError("audio_complete holds no data")
        return self._ph

    @property
    def ph(self):
        return self._out(self._spec_ph())

    @ph.setter
    def ph(self, val):
        self._ph = _Spec.from_view(val, torch.complex64, self._dev)
        self._F = self._wf = None

    def _set_ref(self, value, dev=None):
        self._ref_mag = None if value is None else float(value)
        self._ref_dev = dev if value is not None else None

    def _ref_cached(self):
        """Would the reference's `_ref_mag` be non-None right now?  (The value may exist on the device only.)"""
        return self._ref_mag is not None or self._ref_dev is not None

    def _ref_device(self):
        """ref_mag as a 1-element CUDA tensor, evaluated AND CACHED exactly when the reference's getter would
        (util_audio.py:170-174) -- but without a device-to-host read: the host float is fetched only when somebody
        asks for `.ref_mag` itself.  `subtract` and `D` therefore never synchronise."""
        if self._ref_dev is None:
            if self._ref_mag is not None:
                self._ref_dev = torch.tensor([self._ref_mag], device=self._dev, dtype=torch.float32)
            else:
                if self._max_hint is None:
                    m = self._spec_mag()
                    _, self._max_hint = ops.subtract_db_batch(m.st.unsqueeze(0), None, None, m.nb, want_D=False)
                self._ref_dev = self._max_hint
        return self._ref_dev

    @property
    def ref_mag(self):
        if self._ref_mag is None:
            self._ref_mag = float(self._ref_device().item())
        return np.float32(self._ref_mag)

    @property
    def D(self):
        if self._D is None:
            m = self._spec_mag()
            # the reference evaluates (and caches) ref_mag here (util_audio.py:179)
            D = ops.amplitude_to_db_batch(m.st.unsqueeze(0), m.nb, ref=self._ref_device())
            self._D = _Spec(D[0], m.nb)
        return self._out(self._D)

    @D.setter
    def D(self, val):
        self._D = _Spec.from_view(val, torch.float32, self._dev)
        if self._ph is not None and self._ph.shape != self._D.shape:
            self._ph = None
        self._mag = self._F = self._wf = None
        self._max_hint = None

    def _P(self, name):
        if name not in ("wf", "F", "mag", "ph", "D"):
            raise ValueError("Requested attribute does not exist")
        v = getattr(self, "_" + name)
        return v.view if isinstance(v, _Spec) else v

    @property
    def shape(self):
        if self._mag is not None:
            return self._mag.shape
        if self._ph is not None:
            return self._mag.shape   # (sic) util_audio.py:213-214 raises AttributeError there too
        if self._D is not None:
            return self._D.shape
        # util_audio.py:218 evaluates F here; only its shape is needed, which the
        # K1 geometry gives without running the transform
        if self._F is not None:
            return self._F.shape
        return (self.N // 2 + 1, self._plan().num_frames(self._wave().numel()))

    def _new_like(self, wav):
        ac = audio_complete(None, self.N, hop_length=self.hl, center=self.center,
                            sample_rate=self.sr, device=self._dev, carrier=self._carrier)
        ac._wf = wav
        return ac

    def clone(self):
        ac = self._new_like(None if self._wf is None else self._wf.clone())
        for k in ("_F", "_mag", "_ph", "_D"):
            v = getattr(self, k)
            setattr(ac, k, None if v is None else v.clone())
        ac._ref_mag, ac._ref_dev, ac._max_hint = self._ref_mag, self._ref_dev, self._max_hint
        return ac

    # ------------------------------------------------------------------ time <-> frames
    def _seconds_to_frames(self, time):
        # float64, exactly the reference's operation order (util_audio.py:264)
        return int(np.floor(time * self.shape[1] * self.sr / self._wave().shape[0]))

    def _frames_to_seconds(self, frames):
        return frames / self.shape[1] / self.sr * self._wave().shape[0]

    def midi_tone_to_FFT(self, tone):
        f = midi_to_hz(tone)
        ind = bisect.bisect_right(self._fft_freq, f) - 1
        return 0 if ind == 0 else ind - 1

    # ------------------------------------------------------------------ subtract (K3)
    def subtract(self, subtrahend, offset=0, attack_compensation=0,
                 normalize=True, relu=True, overkill_factor=1):
        """util_audio.py:221-259, one K3 launch (in place on this window's mag)."""
        if isinstance(subtrahend, audio_complete):
            g = subtrahend._spec_mag()
            g_ref = None
            if normalize:
                g_ref = subtrahend._ref_device()     # the reference evaluates and caches subtrahend.ref_mag too
        else:
            g = _Spec.from_view(subtrahend, torch.float32, self._dev)
            g_ref = None                    # max of the array, reduced in-kernel
        ref_init = None
        if normalize:
            ref_init = self._ref_device()   # evaluated (cached) before the update, as in :239
        off = max(self._seconds_to_frames(offset) - attack_compensation, 0)
        m = self._spec_mag()
        n_bins, T = m.shape
        if g.nb != n_bins:
            raise ValueError("operands could not be broadcast together with shapes (%d,%d) (%d,%d)"
                             % (n_bins, T, g.nb, g.T))
        if off > T:
            raise ValueError("negative dimensions are not allowed")
        ok = None
        if overkill_factor != 1:
            ok = torch.tensor([[float(overkill_factor)]], device=self._dev, dtype=torch.float32)
        _, new_max = ops.subtract_db_batch(
            m.st.unsqueeze(0), g.st.reshape(1, 1, g.T, g.st.shape[1]) if g.st.is_contiguous()
            else g.st.contiguous().reshape(1, 1, g.T, g.st.shape[1]),
            torch.full((1, 1), off, device=self._dev, dtype=torch.int32), n_bins, overkill=ok,
            guess_ref=None if g_ref is None else g_ref.reshape(1, 1),
            ref_init=None if ref_init is None else ref_init.reshape(1),
            normalize=normalize, relu=relu, want_D=False)
        # the reference's `self.mag -= ...` runs the setter: dependants are dropped
        self._D = self._F = self._wf = None
        self._set_ref(None)
        self._max_hint = new_max

    # ------------------------------------------------------------------ window mechanics
    def section(self, start, end, duration_in_frames=None):
        """util_audio.py:286-328: copy of columns [tfs:tfe], zero-padded past the end."""
        tfs = self._seconds_to_frames(start)
        tfe = self._seconds_to_frames(end) if duration_in_frames is None else tfs + duration_in_frames
        wav = None
        if self._wf is not None:
            w0 = int(np.floor(self._frames_to_seconds(tfs) * self.sr))
            w1 = int(np.floor(self._frames_to_seconds(tfe) * self.sr))
            wav = self._wf[w0:w1].clone()
            if wav.shape[0] < w1 - w0:
                # (sic) util_audio.py:306 pads with (w1 - len) zeros
                wav = torch.cat((wav, torch.zeros(w1 - wav.shape[0], device=self._dev)))
        nac = self._new_like(wav)

        def cut(f):
            if f is None:
                return None
            part = f.st[tfs:tfe]
            if f.T >= tfe:
                return _Spec(part.clone(), f.nb)
            st = torch.zeros((part.shape[0] + tfe - f.T, f.st.shape[1]), device=self._dev, dtype=f.st.dtype)
            st[:part.shape[0]] = part
            return _Spec(st, f.nb)

        nac._F, nac._mag, nac._ph, nac._D = cut(self._F), cut(self._mag), cut(self._ph), cut(self._D)
        nac._ref_mag, nac._ref_dev = self._ref_mag, self._ref_dev
        return nac

    def slice(self, start_in_frames, end_in_frames):
        """util_audio.py:351-365 (in place, views)."""
        if self._wf is not None:
            a = int(self._frames_to_seconds(start_in_frames) * self.sr)
            b = int(self._frames_to_seconds(end_in_frames) * self.sr)
            self._wf = self._wf[a:b]
        for k in ("_F", "_mag", "_ph", "_D"):
            v = getattr(self, k)
            if v is not None:
                setattr(self, k, v.rows(start_in_frames, end_in_frames))
        self._max_hint = None

    def concat(self, ac):
        """util_audio.py:374-382."""
        self._wf = None if (self._wf is None or ac._wf is None) else torch.cat((self._wf, ac._wf))
        for k in ("_F", "_mag", "_ph", "_D"):
            a, b = getattr(self, k), getattr(ac, k)
            if a is None or b is None:
                setattr(self, k, None)
            else:
                if a.nb != b.nb:
                    raise ValueError("all the input array dimensions except for the concatenation "
                                     "axis must match exactly")
                setattr(self, k, _Spec(torch.cat((a.st, b.st), dim=0), a.nb))
        self._max_hint = None

    def spectral_flatness(self):
        """util_audio.py:330-332 (librosa.feature.spectral_flatness, power 2, amin 1e-10)."""
        plan = ops.get_stft_plan(int(self.N), int(self.hl), True, device=self._dev)
        r = ops.stft_batch(self._wave(), plan, want_max=False)
        flat = ops.spectral_flatness_batch(r["mag_storage"], plan.n_bins)
        return float(flat.double().mean().item())

    def short_window_features(self, start, duration, target_frame_count, band_min, n_rows,
                              ref=None, want_phase=True):
        """Fused form of the producer loop's short-window block (training.py:337-363):
        resize(start, duration, target, ['mag','ph']) -> section_power(band_min, band_min+n_rows)
        -> x/ref, log10(1000x+1)/max, (angle(ph)+3.15)/6.3 in ONE kernel launch.  Returns a dict of
        [n_rows, target] arrays: lin, log, phase."""
        t = self._seconds_to_frames(start + duration)
        s = self._seconds_to_frames(start)
        m = self._spec_mag()
        s_c, t_c = min(max(s, 0), m.T), min(max(t, 0), m.T)
        idx = ops.resize_indices(max(t_c - s_c, 0), target_frame_count)
        src = np.where(idx >= 0, idx + s_c, -1).astype(np.int32)
        ph = self._spec_ph() if want_phase else None
        inv = 1.0 / float(ref if ref is not None else self.ref_mag)
        r = ops.short_window_features(m.st, None if ph is None else ph.st, src, band_min, n_rows, m.nb,
                                      inv_ref=inv, want_phase=want_phase)
        return {k: self._out(v) for k, v in r.items()}

    def section_power(self, name, band_min, band_max):
        """util_audio.py:334-349."""
        P = self._P(name)
        h = P.shape[0]
        part = P[band_min:band_max, :].clone()
        if band_max > h:
            part = torch.cat((part, torch.zeros((band_max - h, P.shape[1]), device=self._dev,
                                                dtype=part.dtype)), dim=0)
        return self._out(part)

    @staticmethod
    def _resize(P, target_frame_count):
        """util_audio.py:384-409, for tensors and arrays alike."""
        if isinstance(P, np.ndarray):
            return _resize_np(P, target_frame_count)
        t = P.shape[1]
        if t == 0:
            return torch.zeros((P.shape[0], target_frame_count), device=P.device, dtype=P.dtype)
        if t == target_frame_count:
            return P
        if t < 3:
            return torch.cat((P[:, :1], P[:, -1:].repeat(1, target_frame_count - 1)), dim=1)
        if t < target_frame_count:
            lim = min(1, int(np.round(t / 3)))
            reps = int(math.floor((target_frame_count - 2 * lim) / (t - 2 * lim)))
            tiled = P[:, lim:-lim].repeat(1, reps)
            tail = target_frame_count - tiled.shape[1] - lim
            return torch.cat((P[:, :lim], tiled, P[:, -tail:]), dim=1)
        return P[:, :target_frame_count]

    def slice_C(self, start, duration, target_frame_count, magnitude_only=True,
                bins_per_tone=1, filter_scale=2, highest_note="C8", lowest_note="A0", nbins=None):
        """util_audio.py:411-434; `filter_scale` is ignored there (:426) and here."""
        if nbins is None:
            nbins = int((note_to_midi(highest_note) - note_to_midi(lowest_note)) * bins_per_tone)
        wav = self._wave()
        plan = ops.get_cqt_plan(self.sr, int(self.hl), note_to_hz(lowest_note), int(nbins),
                                int(12 * bins_per_tone), 2, device=self._dev)
        plan.check_length(int(wav.numel()))
        t = self._seconds_to_frames(start + duration)
        s = self._seconds_to_frames(start)
        Tc = plan.num_frames(int(wav.numel()))
        if magnitude_only and 0 < target_frame_count <= 8 and 0 <= s < Tc:
            # whatever t - s is, `_resize` keeps columns of [s, s + target): contract only those (K2 frame window,
            # saga_cqt_frames_exec) instead of all T columns of the window
            n_avail = max(min(t, Tc) - s, 0)
            C8 = ops.cqt_frames_batch(wav.unsqueeze(0), plan, np.array([s], dtype=np.int32), target_frame_count)
            C = C8[0, :, :plan.n_bins].transpose(0, 1)[:, :min(n_avail, target_frame_count)]
            return self._out(self._resize(C, target_frame_count))
        r = ops.cqt_batch(wav, plan, want_complex=not magnitude_only)
        C = r["mag"][0] if magnitude_only else r["C"][0]
        return self._out(self._resize(C[:, s:t], target_frame_count))

    @staticmethod
    def compress_bands(spectrum, bands=80, log=True):
        """util_audio.py:436-466: mean over (log-spaced) groups of rows; CUDA tensors go through
        saga_compress_bands_exec, host arrays are averaged on the host like the reference."""
        n_rows = spectrum.shape[0]
        edges = band_edges(n_rows, bands) if log else np.arange(bands + 1) * (n_rows // bands)
        if isinstance(spectrum, np.ndarray):
            out = np.zeros((bands, spectrum.shape[1]))
            for i in range(bands):
                out[i] = np.mean(spectrum[int(edges[i]):int(edges[i + 1]), :], axis=0)
            return out
        st = spectrum.transpose(0, 1)
        if st.stride(1) != 1 or st.dtype != torch.float32 or (st.shape[0] > 1 and st.stride(0) < n_rows):
            st = _Spec.from_view(spectrum, torch.float32, spectrum.device).st
        out = ops.compress_bands_batch(st.unsqueeze(0), n_rows, edges)
        return out[0, :, :bands].transpose(0, 1)

    def resize(self, start, duration, target_frame_count, attribs=("F",)):
        """util_audio.py:469-507."""
        nac = self._new_like(None)
        if self._ref_cached():             # copied first; the F/mag setters below wipe it again,
            nac._ref_mag, nac._ref_dev = self._ref_mag, self._ref_dev   # exactly as in :489-499
        t = self._seconds_to_frames(start + duration)
        s = self._seconds_to_frames(start)
        for a in attribs:
            if a == "F":
                nac.F = self._resize(self._spec_F().view[:, s:t], target_frame_count)
            elif a == "mag":
                nac.mag = self._resize(self._spec_mag().view[:, s:t], target_frame_count)
            elif a == "ph":
                nac.ph = self._resize(self._spec_ph().view[:, s:t], target_frame_count)
            elif a == "D":
                nac.D = self._resize(self._spec_ph().view[:, s:t], target_frame_count)  # (sic) :503
            else:
                raise ValueError("Invalid attribute requested")
        return nac


def band_edges(n_rows, bands):
    """util_audio.py:451-456: int-truncated geomspace, first edge 0, strictly increasing."""
    ind = np.geomspace(1, n_rows, bands + 1).astype(int)
    ind[0] = 0
    for i in range(bands):
        sub = ind[i + 1] - ind[i]
        if sub < 1:
            ind[i + 1] += -sub + 1
    return ind


def _resize_np(P, target):
    t = P.shape[1]
    if t == 0:
        return np.zeros((P.shape[0], target))
    if t == target:
        return P
    if t < 3:
        return np.concatenate((P[:, :1], np.tile(P[:, -1:], target - 1)), axis=1)
    if t < target:
        lim = min(1, int(np.round(t / 3)))
        reps = int(np.floor((target - 2 * lim) / (t - 2 * lim)))
        tiled = np.tile(P[:, lim:-lim], reps)
        return np.concatenate((P[:, :lim], tiled, P[:, -(target - tiled.shape[1] - lim):]), axis=1)
    return P[:, :target]

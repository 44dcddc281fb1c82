"""Oracle (TEST INFRASTRUCTURE): CPU restatement of the reference's
`audio_complete` container (/root/reference/util_audio.py:32-527) on top of
the numpy restatements in oracle.spectral / oracle.cqt.

This module re-expresses the class's behaviour (lazy fields with the reference's
invalidation rules, the generative-subtractive `subtract`, the float64 time<->frame
maps, the window slide helpers, the CQT slice) so that it can travel to the GPU box,
where /root/reference does not exist.  It is asserted BIT-IDENTICAL to the reference's
own class (imported unmodified by oracle/ref_class.py) over the producer loop in
tests/test_ref_class.py.  The librosa layer underneath: see oracle/__init__.py
(CQT and dB unpinned against reference outputs).

Each method cites the reference lines it follows.
"""
import bisect
import copy

import numpy as np

from . import cqt as _cqt
from . import spectral as _sp

_FIELDS = ("wf", "F", "mag", "ph", "D")


class AudioOracle:
    """Same constructor and public surface as util_audio.audio_complete
    (util_audio.py:33)."""

    def __init__(self, waveform, n_fft, hop_length=None, center=True, sample_rate=44100):
        self._v = {k: None for k in _FIELDS}
        self._v["wf"] = waveform
        self._ref = None
        self.sr = sample_rate
        self.N = n_fft
        self.center = center
        # util_audio.py:65
        self.hl = hop_length if hop_length is not None else int(np.floor(n_fft / 4))
        self._fft_freq = _sp.fft_frequencies(sample_rate, n_fft)  # :67

    # -- raw access ---------------------------------------------------------
    def _P(self, name):  # util_audio.py:192-207
        if name not in _FIELDS:
            raise ValueError("Requested attribute does not exist")
        return self._v[name]

    def _drop(self, *names):
        for n in names:
            if n == "ref":
                self._ref = None
            else:
                self._v[n] = None

    def _mag_from_db(self):  # util_audio.py:99-101 / :122-124 / :143-145
        if self._ref is None:
            self._ref = 1.0
        self._v["mag"] = _sp.db_to_amplitude(self._v["D"], ref=self._ref)

    def _istft(self):
        return _sp.istft(self._v["F"], hop_length=self.hl, center=self.center)

    # -- lazy properties (util_audio.py:88-190) -------------------------------
    @property
    def wf(self):
        v = self._v
        if v["wf"] is None:
            if v["F"] is not None:
                v["wf"] = self._istft()
            elif v["mag"] is not None and v["ph"] is not None:
                v["F"] = v["mag"] * v["ph"]
                v["wf"] = self._istft()
            elif v["D"] is not None and v["ph"] is not None:
                self._mag_from_db()
                v["F"] = v["mag"] * v["ph"]
                v["wf"] = self._istft()
        return v["wf"]

    @wf.setter
    def wf(self, value):
        self._drop("D", "ref", "mag", "ph", "F")
        self._v["wf"] = value

    @property
    def F(self):
        v = self._v
        if v["F"] is None:
            if v["mag"] is not None and v["ph"] is not None:
                v["F"] = v["mag"] * v["ph"]
            elif v["D"] is not None and v["ph"] is not None:
                self._mag_from_db()
                v["F"] = v["mag"] * v["ph"]
            elif self.wf is not None:
                v["F"] = _sp.stft(self.wf, n_fft=self.N, hop_length=self.hl, center=self.center)
        return v["F"]

    @F.setter
    def F(self, value):
        self._drop("D", "ref", "mag", "ph", "wf")
        self._v["F"] = value

    @property
    def mag(self):
        v = self._v
        if v["mag"] is None:
            if v["D"] is not None and v["ph"] is not None:
                self._mag_from_db()
            else:
                v["mag"], v["ph"] = _sp.magphase(self.F)
        return v["mag"]

    @mag.setter
    def mag(self, val):  # util_audio.py:149-157
        self._drop("D", "ref", "F", "wf")
        self._v["mag"] = val
        if self._v["ph"] is not None and self._v["ph"].shape != val.shape:
            self._v["ph"] = None

    @property
    def ph(self):
        if self._v["ph"] is None:
            self._v["mag"], self._v["ph"] = _sp.magphase(self.F)
        return self._v["ph"]

    @ph.setter
    def ph(self, val):
        self._v["ph"] = val
        self._drop("F", "wf")

    @property
    def ref_mag(self):  # util_audio.py:170-174
        if self._ref is None:
            self._ref = np.max(self.mag)
        return self._ref

    @property
    def D(self):  # util_audio.py:176-180
        if self._v["D"] is None:
            self._v["D"] = _sp.amplitude_to_db(self.mag, ref=self.ref_mag)
        return self._v["D"]

    @D.setter
    def D(self, val):  # util_audio.py:181-190 (ref_mag deliberately kept)
        self._v["D"] = val
        if self._v["ph"] is not None and self._v["ph"].shape != val.shape:
            self._v["ph"] = None
        self._drop("mag", "F", "wf")

    @property
    def shape(self):  # util_audio.py:209-218
        for k in ("mag", "ph", "D"):
            if self._v[k] is not None:
                # the reference returns _mag.shape in the `ph` branch, which
                # raises when only ph is set; same shape whenever both exist
                return self._v[k].shape if k != "ph" else self._v["mag"].shape
        return self.F.shape

    def clone(self):  # util_audio.py:69-87
        ac = AudioOracle(copy.deepcopy(self._v["wf"]), self.N, hop_length=self.hl,
                         center=self.center, sample_rate=self.sr)
        for k in ("F", "mag", "ph", "D"):
            ac._v[k] = copy.deepcopy(self._v[k])
        ac._ref = self._ref
        return ac

    # -- time <-> frame maps (util_audio.py:261-272): float64, this order ----
    def _seconds_to_frames(self, time):
        return int(np.floor(time * self.shape[1] * self.sr / self.wf.shape[0]))

    def _frames_to_seconds(self, frames):
        return frames / self.shape[1] / self.sr * self.wf.shape[0]

    def midi_tone_to_FFT(self, tone):  # util_audio.py:278-284
        f = _sp.midi_to_hz(tone)
        ind = bisect.bisect_right(self._fft_freq, f) - 1
        return 0 if ind == 0 else ind - 1

    # -- generative-subtractive step (util_audio.py:221-259) ------------------
    def subtract(self, subtrahend, offset=0, attack_compensation=0,
                 normalize=True, relu=True, overkill_factor=1):
        if isinstance(subtrahend, type(self)):
            mag_sub = copy.deepcopy(subtrahend.mag)
            if normalize:
                mag_sub *= self.ref_mag / subtrahend.ref_mag
        else:
            mag_sub = copy.deepcopy(subtrahend)
            peak = np.max(mag_sub)
            if normalize:
                mag_sub *= self.ref_mag / peak
        mag_sub *= overkill_factor
        offset = max(self._seconds_to_frames(offset) - attack_compensation, 0)
        n_bins, n_frames = self.mag.shape
        if mag_sub.shape[1] + offset > n_frames:
            mag_sub = mag_sub[:, : (n_frames - offset)]
        padded = np.concatenate(
            (np.zeros((n_bins, offset)), mag_sub,
             np.zeros((n_bins, n_frames - offset - mag_sub.shape[1]))), axis=1)
        m = self.mag
        m -= padded          # in place, float64 operand cast back to mag's dtype
        self.mag = m         # setter runs: ref_mag / D / F / wf invalidated
        if relu:
            self.mag = np.maximum(self.mag, 0, self.mag)

    # -- window mechanics ----------------------------------------------------
    def section(self, start, end, duration_in_frames=None):  # util_audio.py:286-328
        tfs = self._seconds_to_frames(start)
        tfe = self._seconds_to_frames(end) if duration_in_frames is None else tfs + duration_in_frames
        if self._v["wf"] is not None:
            w0 = int(np.floor(self._frames_to_seconds(tfs) * self.sr))
            w1 = int(np.floor(self._frames_to_seconds(tfe) * self.sr))
            wav = copy.deepcopy(self.wf[w0:w1])
            if wav.shape[0] < w1 - w0:
                # (sic) the reference pads by w1 - len, not by the shortfall
                wav = np.concatenate((wav, np.zeros(w1 - wav.shape[0])))
        else:
            wav = None
        nac = AudioOracle(wav, self.N, hop_length=self.hl, center=self.center,
                          sample_rate=self.sr)

        def cut(f):
            if f is None:
                return None
            part = copy.deepcopy(f[:, tfs:tfe])
            if f.shape[1] >= tfe:
                return part
            return np.concatenate((part, np.zeros((f.shape[0], tfe - f.shape[1]))), axis=1)

        for k in ("F", "mag", "ph", "D"):
            nac._v[k] = cut(self._v[k])
        nac._ref = self._ref
        return nac

    def spectral_flatness(self):  # util_audio.py:330-332
        return np.mean(_sp.spectral_flatness(self.wf, n_fft=self.N, hop_length=self.hl))

    def section_power(self, name, band_min, band_max):  # util_audio.py:334-349
        P = self._P(name)
        h = P.shape[0]
        part = copy.deepcopy(P[band_min:band_max, :])
        if band_max > h:
            part = np.concatenate((part, np.zeros((band_max - h, P.shape[1]))), axis=0)
        return part

    def slice(self, start_in_frames, end_in_frames):  # util_audio.py:351-365
        v = self._v
        if v["wf"] is not None:
            a = int(self._frames_to_seconds(start_in_frames) * self.sr)
            b = int(self._frames_to_seconds(end_in_frames) * self.sr)
            v["wf"] = v["wf"][a:b]
        for k in ("F", "mag", "ph", "D"):
            if v[k] is not None:
                v[k] = v[k][:, start_in_frames:end_in_frames]

    def concat(self, ac):  # util_audio.py:368-382
        def join(dst, src, axis):
            return None if (src is None or dst is None) else np.concatenate((dst, src), axis=axis)
        self._v["wf"] = join(self._v["wf"], ac._v["wf"], 0)
        for k in ("F", "mag", "ph", "D"):
            self._v[k] = join(self._v[k], ac._v[k], 1)

    @staticmethod
    def _resize(P, target_frame_count):  # util_audio.py:384-409
        t = P.shape[1]
        if t == 0:
            return np.zeros((P.shape[0], target_frame_count))
        if t == target_frame_count:
            return P
        if t < 3:
            return np.concatenate((P[:, :1], np.tile(P[:, -1:], target_frame_count - 1)), axis=1)
        if t < target_frame_count:
            lim = np.min((1, int(np.round(t / 3))))
            reps = int(np.floor((target_frame_count - 2 * lim) / (t - 2 * lim)))
            tiled = np.tile(P[:, lim:-lim], reps)
            tail = target_frame_count - tiled.shape[1] - lim
            return np.concatenate((P[:, :lim], tiled, P[:, -tail:]), axis=1)
        return P[:, :target_frame_count]

    def slice_C(self, start, duration, target_frame_count, magnitude_only=True,
                bins_per_tone=1, filter_scale=2, highest_note="C8", lowest_note="A0",
                nbins=None):  # util_audio.py:411-434 (filter_scale argument ignored: :426)
        if nbins is None:
            nbins = int((_sp.note_to_midi(highest_note) - _sp.note_to_midi(lowest_note))
                        * bins_per_tone)
        C = _cqt.cqt(self.wf, sr=self.sr, fmin=_sp.note_to_hz(lowest_note), n_bins=nbins,
                     bins_per_octave=int(12 * bins_per_tone), filter_scale=2,
                     hop_length=self.hl)
        if magnitude_only:
            C = np.abs(C)
        t = self._seconds_to_frames(start + duration)
        s = self._seconds_to_frames(start)
        return self._resize(C[:, s:t], target_frame_count)

    @staticmethod
    def compress_bands(spectrum, bands=80, log=True):  # util_audio.py:436-466
        out = np.zeros((bands, spectrum.shape[1]))
        # the reference averages one 1-D column slice at a time (util_audio.py:458-465); numpy's
        # float32 summation order depends on that, so the oracle keeps the same granularity
        if log:
            ind = band_edges(spectrum.shape[0], bands)
            lo, hi = [int(v) for v in ind[:-1]], [int(v) for v in ind[1:]]
        else:
            r = spectrum.shape[0] // bands
            lo, hi = [r * i for i in range(bands)], [r * (i + 1) for i in range(bands)]
        for j in range(spectrum.shape[1]):
            col = spectrum[:, j]
            for i in range(bands):
                out[i, j] = np.mean(col[lo[i]:hi[i]])
        return out

    def resize(self, start, duration, target_frame_count, attribs=("F",)):  # :469-507
        nac = AudioOracle(None, self.N, hop_length=self.hl, center=self.center,
                          sample_rate=self.sr)
        if self._ref is not None:
            nac._ref = self._ref
        t = self._seconds_to_frames(start + duration)
        s = self._seconds_to_frames(start)
        for a in attribs:
            if a == "F":
                nac.F = self._resize(self.F[:, s:t], target_frame_count)
            elif a == "mag":
                nac.mag = self._resize(self.mag[:, s:t], target_frame_count)
            elif a == "ph":
                nac.ph = self._resize(self.ph[:, s:t], target_frame_count)
            elif a == "D":
                # (sic) util_audio.py:503 resizes `ph` here
                nac.D = self._resize(self.ph[:, s:t], target_frame_count)
            else:
                raise ValueError("Invalid attribute requested")
        return nac


def band_edges(n_rows, bands):
    """Band edges of compress_bands(log=True) (util_audio.py:451-456):
    integer-truncated geomspace(1, n_rows, bands+1), first edge forced to 0,
    then forced strictly increasing left to right."""
    ind = np.geomspace(1, n_rows, bands + 1).astype(int)
    ind[0] = 0
    for i in range(bands):
        sub = ind[i + 1] - ind[i]
        if sub < 1:
            ind[i + 1] += -sub + 1
    return ind

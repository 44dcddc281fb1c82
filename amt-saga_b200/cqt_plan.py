"""Host-side constant-Q plan builder (float64 numpy, runs once per parameter
set).  Produces the descriptor `saga_cqt_plan_create` consumes.

What librosa 0.6.3 `cqt()` does per call with audio-independent data
(/root/reference/util_audio.py:424-426 is the only call site) is hoisted here:

  * octave geometry, early-downsample factor (librosa.core.constantq:
    cqt / __early_downsample_count),
  * the top-octave filter bank: complex exponentials x periodic Hann,
    L1-normalised, centre-padded to a power of two, complex64, FFT'd, each row
    sparsified at the `sparsity` L1 quantile (filters.constant_q,
    __cqt_filter_fft, util.sparsify_rows),
  * the kaiser_fast decimator of resampy, which for the integer ratios used
    here is a fixed symmetric FIR.

The device does not run "rect-window FFT, then sparse basis": that product is
linear in the frame, so each octave's (sparsified) FFT-domain basis B[f, b] is
folded with the DFT into a dense time-domain bank

    g_f[n] = sum_{b=0}^{n_fft/2} B[f, b] * exp(-2*pi*i*b*n/n_fft)

(all gains -- sqrt(2) per decimation level, 1/sqrt(filter length) -- folded in),
stored as G[n, 2f] = Re g_f[n], G[n, 2f+1] = Im g_f[n].  C[f, t] = sum_n
frame_t[n] * g_f[n] is then exactly librosa's number, evaluated as a GEMM.
"""
import ctypes as C
import math

import numpy as np

from . import _lib

HANN_BANDWIDTH = 1.50018310546875
KAISER_FAST = dict(zeros=16, table=512, beta=8.555504641634386, rolloff=0.85)


class ParameterError(ValueError):
    """Same role as librosa.util.exceptions.ParameterError (a ValueError)."""


def _hann_periodic(n):
    if n == 1:
        return np.ones(1)
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n)


def kaiser_fast_taps(factor):
    """Symmetric FIR equivalent of resampy.resample(x, factor, 1, 'kaiser_fast')
    followed by librosa's 1/sqrt(ratio) rescale.  Returns taps[0..S], centre first:
    y[t] = sum_{|m|<=S} taps[|m|] x[factor*t + m]."""
    k = KAISER_FAST
    if k["table"] % factor:
        raise ParameterError("decimation factor %d unsupported" % factor)
    n = k["table"] * k["zeros"]
    grid = np.linspace(0, k["zeros"], num=n + 1, endpoint=True)
    half = np.kaiser(2 * n + 1, k["beta"])[n:] * (k["rolloff"] * np.sinc(k["rolloff"] * grid))
    step = k["table"] // factor
    count = half.shape[0] // step
    return half[0:count * step:step] / factor * math.sqrt(factor)


def _two_factors(x):
    n = 0
    while x > 0 and x % 2 == 0:
        n += 1
        x //= 2
    return n


def _filter_lengths(sr, fmin, n_bins, bpo, filter_scale):
    if fmin <= 0 or bpo <= 0 or filter_scale <= 0 or n_bins <= 0:
        raise ParameterError("fmin, bins_per_octave, filter_scale, n_bins must be positive")
    Q = float(filter_scale) / (2.0 ** (1.0 / bpo) - 1)
    freq = fmin * (2.0 ** (np.arange(n_bins, dtype=float) / bpo))
    if freq[-1] * (1 + 0.5 * HANN_BANDWIDTH / Q) > sr / 2.0:
        raise ParameterError("Filter pass-band lies beyond Nyquist")
    return Q * sr / freq, Q


def _fft_basis(sr, fmin, n_filters, bpo, filter_scale, norm, sparsity):
    """Sparsified FFT-domain bank of one octave: complex [n_filters, n_fft/2+1]."""
    lengths, Q = _filter_lengths(sr, fmin, n_filters, bpo, filter_scale)
    n_fft = int(2.0 ** math.ceil(math.log2(lengths.max())))
    basis = np.zeros((n_filters, n_fft), dtype=np.complex64)
    for row, ilen in enumerate(lengths):
        freq = Q * sr / ilen
        t = np.arange(-ilen // 2, ilen // 2, dtype=float)
        sig = np.exp(t * 1j * 2 * np.pi * freq / sr) * _hann_periodic(len(t))
        if norm == 1:
            sig = sig / max(np.sum(np.abs(sig)), np.finfo(float).tiny)
        elif norm == 2:
            sig = sig / max(math.sqrt(np.sum(np.abs(sig) ** 2)), np.finfo(float).tiny)
        elif norm is not None:
            raise ParameterError("Unsupported norm: %r" % (norm,))
        lpad = (n_fft - len(sig)) // 2
        basis[row, lpad:lpad + len(sig)] = sig
    basis = (basis.astype(np.complex128) * (lengths[:, None] / float(n_fft))).astype(np.complex64)
    spec = np.fft.fft(basis, n=n_fft, axis=1)[:, :n_fft // 2 + 1]
    # util.sparsify_rows
    if not 0.0 <= sparsity < 1:
        raise ParameterError("Invalid quantile %r" % (sparsity,))
    mags = np.abs(spec)
    order = np.sort(mags, axis=1)
    share = np.cumsum(order / np.sum(mags, axis=1, keepdims=True), axis=1)
    cut = order[np.arange(n_filters), np.argmin(share < sparsity, axis=1)]
    spec = np.where(mags >= cut[:, None], spec, 0)
    return spec, n_fft


class CqtPlan:
    """Owns the device-side plan handle.  Geometry attributes mirror the
    descriptor (early_factor, octaves[i] = dict(level, hop, n_fft, n_filters,
    first_bin, bank))."""

    def __init__(self, sr, hop_length, fmin, n_bins, bins_per_octave=12, filter_scale=1,
                 norm=1, sparsity=0.01, scale=True, tuning=0.0, create_device_plan=True):
        if tuning != 0.0:
            raise _lib.SagaUnsupported("tuning != 0 is not on the reference's path")
        self.sr, self.hop, self.fmin, self.n_bins = sr, int(hop_length), float(fmin), int(n_bins)
        bpo = int(bins_per_octave)
        n_oct = int(math.ceil(float(n_bins) / bpo))
        n_filt = min(bpo, n_bins)
        freqs = (fmin * 2.0 ** (np.arange(n_bins, dtype=float) / bpo))[-bpo:]
        fmin_t, fmax_t = freqs.min(), freqs.max()
        Q = float(filter_scale) / (2.0 ** (1.0 / bpo) - 1)
        cutoff = fmax_t * (1 + 0.5 * HANN_BANDWIDTH / Q)
        nyq = sr / 2.0
        fast = cutoff < KAISER_FAST["rolloff"] * nyq
        early = 0
        if fast:
            c1 = max(0, int(math.ceil(math.log2(KAISER_FAST["rolloff"] * nyq / cutoff)) - 1) - 1)
            c2 = max(0, _two_factors(self.hop) - n_oct + 1)
            early = min(c1, c2)
        self.early_factor = 2 ** early
        sr_e = sr / float(self.early_factor) if early else sr
        hop_e = self.hop // self.early_factor
        jobs = []  # (level, hop, fft-domain basis, n_fft), top octave first
        rem = n_oct
        if not fast:
            fb, n_fft = _fft_basis(sr_e, fmin_t, n_filt, bpo, filter_scale, norm, sparsity)
            jobs.append((0, hop_e, fb, n_fft))
            fmin_t /= 2
            rem -= 1
        if _two_factors(hop_e) < rem - 1:
            raise ParameterError("hop_length must be a positive integer multiple of 2^%d for "
                                 "%d-octave CQT" % (rem - 1, rem))
        if rem > 0:
            fb, n_fft = _fft_basis(sr_e, fmin_t, n_filt, bpo, filter_scale, norm, sparsity)
            for i in range(rem):
                jobs.append((i, hop_e >> i, fb * (math.sqrt(2.0) ** i), n_fft))
        lengths, _ = _filter_lengths(sr_e, fmin, n_bins, bpo, filter_scale)
        gain = math.sqrt(self.early_factor) if (early and not scale) else 1.0
        self.octaves = []
        for idx, (level, hop_o, fb, n_fft) in enumerate(jobs):
            first_bin = n_bins - (idx + 1) * n_filt
            padded = np.zeros((n_filt, n_fft), dtype=np.complex128)
            padded[:, :n_fft // 2 + 1] = fb
            g = np.fft.fft(padded, axis=1) * gain        # g[f, n]
            if scale:
                rows = np.arange(n_filt) + first_bin
                inv = np.where(rows >= 0, 1.0 / np.sqrt(lengths[np.clip(rows, 0, None)]), 0.0)
                g = g * inv[:, None]
            bank = np.empty((n_fft, 2 * n_filt), dtype=np.float32)
            bank[:, 0::2] = g.real.T
            bank[:, 1::2] = g.imag.T
            self.octaves.append(dict(level=level, hop=hop_o, n_fft=n_fft, n_filters=n_filt,
                                     first_bin=first_bin, bank=np.ascontiguousarray(bank)))
        self.early_taps = (kaiser_fast_taps(self.early_factor).astype(np.float32)
                           if self.early_factor > 1 else np.zeros(0, np.float32))
        self.half_taps = kaiser_fast_taps(2).astype(np.float32)
        self.max_level = max(o["level"] for o in self.octaves)
        self._h = None
        if create_device_plan:
            self._create()

    # -- frames librosa.cqt returns for a clip of `n` samples --------------------
    def level_len(self, n, level):
        if self.early_factor > 1:
            n = -(-n // self.early_factor)
        for _ in range(level):
            n = (n + 1) // 2
        return n

    def num_frames(self, n):
        if n <= 0:
            return 0
        return min(1 + self.level_len(n, o["level"]) // o["hop"] for o in self.octaves)

    def geometry(self):
        """What the decimation cascade and the level buffers depend on (not the bank): plans of equal geometry can share
        one cascade (saga_cqt_frames_shared_exec)."""
        return (self.sr, self.hop, self.early_factor, self.max_level,
                tuple((o["level"], o["hop"], o["n_fft"], o["n_filters"]) for o in self.octaves))

    def check_length(self, n):
        """librosa raises when the signal is too short to decimate."""
        if n < self.early_factor or self.level_len(n, max(self.max_level - 1, 0)) < 2 and self.max_level > 0:
            raise ParameterError("Input signal length=%d is too short for %d-octave CQT"
                                 % (n, len(self.octaves)))

    def _create(self):
        octs = (_lib.CqtOctave * len(self.octaves))()
        for i, o in enumerate(self.octaves):
            octs[i].level, octs[i].hop, octs[i].n_fft = o["level"], o["hop"], o["n_fft"]
            octs[i].n_filters, octs[i].first_bin = o["n_filters"], o["first_bin"]
            octs[i].bank_host = o["bank"].ctypes.data_as(C.POINTER(C.c_float))
        d = _lib.CqtDesc()
        d.n_bins, d.hop, d.early_factor = self.n_bins, self.hop, self.early_factor
        d.n_early_taps = len(self.early_taps)
        d.early_taps_host = self.early_taps.ctypes.data_as(C.POINTER(C.c_float))
        d.n_half_taps = len(self.half_taps)
        d.half_taps_host = self.half_taps.ctypes.data_as(C.POINTER(C.c_float))
        d.n_octaves, d.octaves = len(self.octaves), octs
        h = C.c_void_p()
        _lib.check(_lib.lib().saga_cqt_plan_create(C.byref(h), C.byref(d)), ParameterError)
        self._h = h

    @property
    def handle(self):
        if self._h is None:
            self._create()
        return self._h

    def close(self):
        if self._h is not None:
            _lib.lib().saga_cqt_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

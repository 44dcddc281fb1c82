"""Oracle (TEST INFRASTRUCTURE): restatement of `librosa.core.constantq.cqt`
as of librosa 0.6.3 (multi-rate, per-octave rectangular-window STFT times a
1%-sparsified FFT-domain filter bank), the routine the reference reaches at
/root/reference/util_audio.py:424-426 (`slice_C`) and training.py:271-388.

librosa is an unpinned third-party dependency that is absent here (no
network): PARITY UNPINNED.  Algorithm per SURVEY.md Appendix A.6; anchored by
the analytic bin-centre response and the frame-count rules in
tests/test_oracle_cqt.py.

Also provides `cqt_plan`: the per-octave geometry + FFT-domain basis, which the
CUDA plan builder (amt-saga_b200/cqt_plan.py) must reproduce; the tests
compare the two builders entry by entry.
"""
import numpy as np

from . import resample as _rs
from .spectral import get_window, pad_center, stft

HANN_BANDWIDTH = 1.50018310546875  # librosa.filters.WINDOW_BANDWIDTHS['hann']


class ParameterError(ValueError):
    """Stands in for librosa.util.exceptions.ParameterError."""


def cqt_frequencies(n_bins, fmin, bins_per_octave=12, tuning=0.0):
    correction = 2.0 ** (float(tuning) / bins_per_octave)
    return correction * fmin * 2.0 ** (np.arange(0, n_bins, dtype=float) / bins_per_octave)


def constant_q_lengths(sr, fmin, n_bins, bins_per_octave, tuning=0.0, filter_scale=1):
    if fmin <= 0:
        raise ParameterError("fmin must be positive")
    if bins_per_octave <= 0:
        raise ParameterError("bins_per_octave must be positive")
    if filter_scale <= 0:
        raise ParameterError("filter_scale must be positive")
    if n_bins <= 0 or not isinstance(n_bins, (int, np.integer)):
        raise ParameterError("n_bins must be a positive integer")
    fmin = 2.0 ** (float(tuning) / bins_per_octave) * fmin
    Q = float(filter_scale) / (2.0 ** (1.0 / bins_per_octave) - 1)
    freq = fmin * (2.0 ** (np.arange(n_bins, dtype=float) / bins_per_octave))
    if freq[-1] * (1 + 0.5 * HANN_BANDWIDTH / Q) > sr / 2.0:
        raise ParameterError("Filter pass-band lies beyond Nyquist")
    return Q * sr / freq


def constant_q(sr, fmin, n_bins, bins_per_octave, tuning=0.0, filter_scale=1,
               norm=1, dtype=np.complex64):
    """librosa.filters.constant_q(pad_fft=True, window='hann'): time-domain
    complex filters, L1-normalised, centre-padded to a power of two, stored
    as complex64 (librosa's default dtype)."""
    lengths = constant_q_lengths(sr, fmin, n_bins, bins_per_octave, tuning, filter_scale)
    Q = float(filter_scale) / (2.0 ** (1.0 / bins_per_octave) - 1)
    freqs = Q * sr / lengths
    filters = []
    for ilen, freq in zip(lengths, freqs):
        # np.arange(-ilen//2, ilen//2): float floor-division => ceil(ilen)-ish samples
        sig = np.exp(np.arange(-ilen // 2, ilen // 2, dtype=float) * 1j * 2 * np.pi * freq / sr)
        sig = sig * get_window("hann", len(sig))
        if norm == 1:
            length = np.sum(np.abs(sig))
        elif norm == 2:
            length = np.sqrt(np.sum(np.abs(sig) ** 2))
        elif norm is None:
            length = 1.0
        else:
            raise ParameterError("Unsupported norm")
        if length < np.finfo(np.float64).tiny:
            length = 1.0
        filters.append(sig / length)
    max_len = int(2.0 ** (np.ceil(np.log2(max(lengths)))))
    basis = np.asarray([pad_center(f, max_len) for f in filters], dtype=dtype)
    return basis, np.asarray(lengths)


def sparsify_rows(x, quantile=0.01):
    """librosa.util.sparsify_rows as a dense array with zeros: per row zero
    every entry smaller than the first sorted magnitude whose cumulative L1
    share reaches `quantile`."""
    if not 0.0 <= quantile < 1:
        raise ParameterError("Invalid quantile")
    mags = np.abs(x)
    norms = np.sum(mags, axis=1, keepdims=True)
    mag_sort = np.sort(mags, axis=1)
    cumulative = np.cumsum(mag_sort / norms, axis=1)
    thr_idx = np.argmin(cumulative < quantile, axis=1)
    out = np.zeros_like(x)
    for i, j in enumerate(thr_idx):
        keep = mags[i] >= mag_sort[i, j]
        out[i, keep] = x[i, keep]
    return out


def _basis_fft(basis, n_fft):
    try:
        from scipy import fftpack
        return fftpack.fft(basis, n=n_fft, axis=1)
    except ImportError:        # numpy >= 2.0 keeps complex64 as well
        return np.fft.fft(basis, n=n_fft, axis=1)


def cqt_filter_fft(sr, fmin, n_bins, bins_per_octave, tuning, filter_scale,
                   norm, sparsity):
    basis, lengths = constant_q(sr, fmin, n_bins, bins_per_octave, tuning,
                                filter_scale, norm)
    n_fft = basis.shape[1]
    # in-place `basis *= float64` on a complex64 array: computed wide, stored complex64
    basis = (basis.astype(np.complex128) * (lengths[:, np.newaxis] / float(n_fft))).astype(np.complex64)
    # librosa 0.6.3 (core/constantq.py, `import scipy.fftpack as fft`) transforms the complex64 basis with
    # scipy.fftpack.fft, which computes and returns SINGLE precision for complex64 input; so does numpy >= 2.0's
    # np.fft.fft (the two agree to 1e-7 of the basis peak and to the same sparsity pattern at every shape the
    # reference uses: tests/test_oracle_pins.py::test_basis_fft_precision).  Use the reference's own function
    # where it is importable.
    fft_basis = _basis_fft(basis, n_fft)[:, : (n_fft // 2) + 1]
    fft_basis = sparsify_rows(fft_basis, quantile=sparsity)
    return fft_basis, n_fft, lengths


def num_two_factors(x):
    if x <= 0:
        return 0
    n = 0
    while x % 2 == 0:
        n += 1
        x //= 2
    return n


def early_downsample_count(nyquist, filter_cutoff, hop_length, n_octaves):
    c1 = max(0, int(np.ceil(np.log2(_rs.BW_FASTEST * nyquist / filter_cutoff)) - 1) - 1)
    c2 = max(0, num_two_factors(hop_length) - n_octaves + 1)
    return min(c1, c2)


def cqt_plan(sr, hop_length, fmin, n_bins, bins_per_octave, tuning=0.0,
             filter_scale=1, norm=1, sparsity=0.01):
    """Everything about librosa.cqt that does not depend on the audio: the
    early-downsample factor and, per octave job, (decimation level relative to
    the early-downsampled signal, hop, sqrt(2)**level-scaled FFT-domain basis).
    Jobs are listed top octave first, as librosa computes them."""
    n_octaves = int(np.ceil(float(n_bins) / bins_per_octave))
    n_filters = min(bins_per_octave, n_bins)
    freqs = cqt_frequencies(n_bins, fmin, bins_per_octave, tuning)[-bins_per_octave:]
    fmin_t, fmax_t = np.min(freqs), np.max(freqs)
    Q = float(filter_scale) / (2.0 ** (1.0 / bins_per_octave) - 1)
    filter_cutoff = fmax_t * (1 + 0.5 * HANN_BANDWIDTH / Q)
    nyquist = sr / 2.0
    fast = filter_cutoff < _rs.BW_FASTEST * nyquist
    early = early_downsample_count(nyquist, filter_cutoff, hop_length, n_octaves) if fast else 0
    plan = {"early_factor": 2 ** early, "jobs": [], "n_octaves": n_octaves,
            "n_filters": n_filters, "n_bins": n_bins}
    sr_e = sr / float(2 ** early) if early > 0 else sr
    hop_e = hop_length // (2 ** early)
    plan["sr_early"], plan["hop_early"] = sr_e, hop_e
    n_oct = n_octaves
    if not fast:
        fb, n_fft, _ = cqt_filter_fft(sr_e, fmin_t, n_filters, bins_per_octave, tuning,
                                      filter_scale, norm, sparsity)
        plan["jobs"].append({"level": 0, "hop": hop_e, "n_fft": n_fft, "fft_basis": fb})
        fmin_t /= 2
        fmax_t /= 2
        n_oct -= 1
    if num_two_factors(hop_e) < n_oct - 1:
        raise ParameterError("hop_length must be a positive integer multiple of 2^%d for "
                             "%d-octave CQT" % (n_oct - 1, n_oct))
    fb, n_fft, _ = cqt_filter_fft(sr_e, fmin_t, n_filters, bins_per_octave, tuning,
                                  filter_scale, norm, sparsity)
    hop_i = hop_e
    for i in range(n_oct):
        if i > 0:
            fb = fb * np.sqrt(2)
            hop_i //= 2
        plan["jobs"].append({"level": i, "hop": hop_i, "n_fft": n_fft, "fft_basis": fb.copy()})
    plan["lengths"] = constant_q_lengths(sr_e, fmin, n_bins, bins_per_octave, tuning,
                                         filter_scale)
    return plan


def cqt(y, sr=22050, hop_length=512, fmin=None, n_bins=84, bins_per_octave=12,
        tuning=0.0, filter_scale=1, norm=1, sparsity=0.01, scale=True):
    """librosa.cqt (0.6.3): complex [n_bins, T]."""
    y = np.asarray(y)
    if fmin is None:
        fmin = 32.70319566257483  # note_to_hz('C1')
    plan = cqt_plan(sr, hop_length, fmin, n_bins, bins_per_octave, tuning,
                    filter_scale, norm, sparsity)
    len_orig = len(y)
    if plan["early_factor"] > 1:
        if len(y) < plan["early_factor"]:
            raise ParameterError("Input signal length=%d is too short for %d-octave CQT"
                                 % (len_orig, plan["n_octaves"]))
        y = _rs.librosa_resample(y, sr, plan["sr_early"], scale=True)
        if not scale:
            y = y * np.sqrt(plan["early_factor"])
    resp = []
    my_y, level = y, 0
    for job in plan["jobs"]:
        while level < job["level"]:
            if len(my_y) < 2:
                raise ParameterError("Input signal length=%d is too short for %d-octave CQT"
                                     % (len_orig, plan["n_octaves"]))
            my_y = _rs.librosa_resample(my_y, 2, 1, scale=True)
            level += 1
        D = stft(my_y, n_fft=job["n_fft"], hop_length=job["hop"], window="ones")
        resp.append(job["fft_basis"].dot(D))
    max_col = min(x.shape[1] for x in resp)
    C = np.vstack([x[:, :max_col] for x in resp][::-1])[-n_bins:]
    if scale:
        C = C / np.sqrt(plan["lengths"][:, np.newaxis])
    return C

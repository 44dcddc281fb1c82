"""Oracle (TEST INFRASTRUCTURE): restatement of resampy 0.2.x band-limited
sinc interpolation with the `kaiser_fast` / `kaiser_best` filters, and of
`librosa.core.audio.resample(..., fix=True, scale=True)` built on it.

resampy is a third-party dependency of librosa (itself an unpinned dependency
of /root/reference: no requirements file; era fixed to librosa 0.6.3 / resampy
0.2.x by API usage).  It is absent here, so this is a
restatement of its published algorithm (Smith, "Digital Audio Resampling Home
Page"; resampy/interpn.py and resampy/filters.py) -- PARITY UNPINNED.

Reached from the reference through librosa.cqt at util_audio.py:424-426.
"""
import numpy as np

# resampy's two stock filters: (zero crossings, table bits, kaiser beta, rolloff)
_FILTERS = {
    "kaiser_fast": (16, 9, 8.555504641634386, 0.85),
    "kaiser_best": (64, 9, 14.769656459379492, 0.9475937167399596),
}
BW_FASTEST = _FILTERS["kaiser_fast"][3]
BW_BEST = _FILTERS["kaiser_best"][3]


def sinc_window(num_zeros, precision, beta, rolloff):
    """resampy.filters.sinc_window with a Kaiser taper: right wing of a
    windowed sinc sampled 2**precision times per zero crossing."""
    num_bits = 2 ** precision
    n = num_bits * num_zeros
    sinc_win = rolloff * np.sinc(rolloff * np.linspace(0, num_zeros, num=n + 1, endpoint=True))
    taper = np.kaiser(2 * n + 1, beta)[n:]
    return taper * sinc_win, num_bits, rolloff


_CACHE = {}


def get_filter(name):
    if name not in _CACHE:
        zeros, prec, beta, rolloff = _FILTERS[name]
        _CACHE[name] = sinc_window(zeros, prec, beta, rolloff)
    half, bits, rolloff = _CACHE[name]
    return half.copy(), bits, rolloff


def resample_f_literal(x, n_out, sample_ratio, interp_win, interp_delta, num_table):
    """The interpolation loop of resampy.interpn.resample_f, statement for
    statement (1-D).  Slow; used for small inputs and to validate
    `decimate_fir`."""
    x = np.asarray(x)
    y = np.zeros(n_out, dtype=x.dtype)
    scale = min(1.0, sample_ratio)
    time_increment = 1.0 / sample_ratio
    index_step = int(scale * num_table)
    time_register = 0.0
    nwin = interp_win.shape[0]
    n_orig = x.shape[0]
    for t in range(n_out):
        n = int(time_register)
        frac = scale * (time_register - n)
        index_frac = frac * num_table
        offset = int(index_frac)
        eta = index_frac - offset
        i_max = min(n + 1, (nwin - offset) // index_step)
        for i in range(i_max):
            w = interp_win[offset + i * index_step] + eta * interp_delta[offset + i * index_step]
            y[t] += w * x[n - i]
        frac = scale - frac
        index_frac = frac * num_table
        offset = int(index_frac)
        eta = index_frac - offset
        k_max = min(n_orig - n - 1, (nwin - offset) // index_step)
        for k in range(k_max):
            w = interp_win[offset + k * index_step] + eta * interp_delta[offset + k * index_step]
            y[t] += w * x[n + k + 1]
        time_register += time_increment
    return y


def decimation_taps(factor, filt="kaiser_fast"):
    """For an integer decimation factor D the time register of resample_f is
    integral (frac == 0), so the loop degenerates to a fixed symmetric FIR:
        y[t] = sum_{m=-S}^{S} taps[|m|] * x[D*t + m]      (zeros outside x)
    with taps[|m|] = h[|m| * (table/D)] / D and S = len(h)//(table/D) - 1
    (left wing: i < len(h)//step incl. the centre; right wing starts one step
    in and has (len(h)-step)//step = S entries).  kaiser_fast, D=2: S=31, i.e.
    63 taps.  Returns taps[0..S], centre first.  Needs table % D == 0."""
    half, table, _ = get_filter(filt)
    if table % factor != 0:
        raise ValueError("decimation factor must divide the filter table resolution")
    step = table // factor
    n_left = half.shape[0] // step
    return half[0 : n_left * step : step] / factor


def decimate_fir(x, factor, filt="kaiser_fast"):
    """Vectorised equivalent of resampy.resample(x, D, 1) for integer D
    (validated against resample_f_literal in tests/test_oracle_cqt.py)."""
    x = np.asarray(x, dtype=np.float64)
    n_out = int(x.shape[0] * (1.0 / factor))
    taps = decimation_taps(factor, filt)
    side = taps.shape[0] - 1
    h = np.concatenate([taps[:0:-1], taps])  # m = -side .. side
    xp = np.concatenate([np.zeros(side), x, np.zeros(side + factor)])
    out = np.empty(n_out, dtype=np.float64)
    blk = 1 << 16
    for s in range(0, n_out, blk):
        e = min(n_out, s + blk)
        idx = factor * np.arange(s, e)[:, None] + np.arange(h.shape[0])[None, :]
        out[s:e] = xp[idx] @ h
    return out


def resampy_resample(x, sr_orig, sr_new, filt="kaiser_fast"):
    """resampy.resample for 1-D input."""
    x = np.asarray(x)
    sample_ratio = float(sr_new) / sr_orig
    n_out = int(x.shape[0] * sample_ratio)
    if n_out < 1:
        raise ValueError("Input signal length=%d is too small to resample" % x.shape[0])
    ratio_inv = float(sr_orig) / sr_new
    if ratio_inv == int(ratio_inv) and get_filter(filt)[1] % int(ratio_inv) == 0:
        return decimate_fir(x, int(ratio_inv), filt).astype(x.dtype)
    interp_win, precision, _ = get_filter(filt)
    if sample_ratio < 1:
        interp_win *= sample_ratio
    interp_delta = np.zeros_like(interp_win)
    interp_delta[:-1] = np.diff(interp_win)
    return resample_f_literal(x, n_out, sample_ratio, interp_win, interp_delta, precision)


def librosa_resample(y, orig_sr, target_sr, res_type="kaiser_fast", scale=True):
    """librosa.core.audio.resample(..., fix=True, scale=scale)."""
    if orig_sr == target_sr:
        return y
    ratio = float(target_sr) / orig_sr
    n_samples = int(np.ceil(y.shape[-1] * ratio))
    y_hat = resampy_resample(y, orig_sr, target_sr, filt=res_type)
    if y_hat.shape[0] > n_samples:
        y_hat = y_hat[:n_samples]
    elif y_hat.shape[0] < n_samples:
        y_hat = np.concatenate([y_hat, np.zeros(n_samples - y_hat.shape[0], dtype=y_hat.dtype)])
    if scale:
        y_hat = y_hat / np.sqrt(ratio)
    return np.ascontiguousarray(y_hat, dtype=y.dtype)

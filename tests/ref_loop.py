"""The reference's producer loop (`/root/reference/training.py:265-449`) as a
test driver, parameterised over the class that plays `audio_complete`:

* `oracle.ref_class.audio_complete`  -- the reference's own class (this container),
* `oracle.audio_oracle.AudioOracle`  -- the numpy restatement,
* `amt_saga_b200.util_audio.audio_complete` -- the CUDA class (GPU box).

Same call order, same arguments, same normalisers as the reference; the MIDI
side (fluidsynth render, `relevant_notes`) is replaced by a fixed note list with
pre-rendered single-note clips, which is the input boundary of the hot path.
Returns every intermediate as a host numpy array, keyed by step.
"""
import numpy as np

from tests.synth import piano_clip

MIDI_NAMES = ["C", "C#", "D", "D#", "E", "F", "F#", "G", "G#", "A", "A#", "B"]


def midi_to_note(m):  # librosa.midi_to_note, sharps, octave numbers (training.py:17)
    return "%s%d" % (MIDI_NAMES[m % 12], m // 12 - 1)


class Params:  # util_train_test.py:15-79 with main.py's defaults (bins_per_tone 4, N 4096)
    N = 4096
    sr = 44100
    H = 1024
    window_size_note_time = 6
    timing_frames = int(6 * 44100 / 1024)            # 258
    timing_bands = 20
    pitch_frames = 8
    instrument_frames = 8
    pitch_bins_per_tone = 2
    instrument_bins_per_tone = 4
    instrument_bands = 348
    bins_velocity = 36


def small_params(timing_frames=64):
    """Same loop, shorter window (CPU test budget)."""
    p = Params()
    p.timing_frames = timing_frames
    p.window_size_note_time = timing_frames * 1024 / 44100.0
    return p


def make_inputs(seed, song_seconds, notes, dtype=np.float64, sr=44100):
    """Synthetic song + one single-note clip per note (the fluidsynth outputs are
    float64, util_audio.py:776-781)."""
    song = piano_clip(seed, int(sr * song_seconds), n_notes=int(4 * song_seconds)).astype(dtype)
    clips = []
    for i, (onset, dur, pitch) in enumerate(notes):
        clips.append(piano_clip(seed + 100 + i, int(sr * (dur + 0.3)), n_notes=1,
                                pitch_range=(pitch, pitch)).astype(dtype))
    return song, clips


def host(x):
    if hasattr(x, "detach"):
        x = x.detach().cpu().numpy()
    return np.array(x)


def run_loop(AC, song, notes, clips, params=Params, foc_bpt_scale=4, slide_after=None, out=None):
    """training.py:265-449.  `notes` = [(onset_s, duration_s, pitch)], times relative to the
    current window; `slide_after` = index of the note before which the window slides half a window
    (training.py:317-328).  `foc_bpt_scale` = the `*4` of training.py:281/:368 (1 keeps the CPU
    oracle's 8192-sample banks out of the quick tests)."""
    p = params
    r = {} if out is None else out
    halfwindow_frames = p.timing_frames // 2
    halfwindow_time = p.window_size_note_time / 2

    mid_wf = AC(song, p.N)                                                 # :265
    r["flatness"] = np.float64(host(mid_wf.spectral_flatness()))           # :266
    mid_wf.mag                                                             # :269
    dur = mid_wf._frames_to_seconds(mid_wf.shape[1])
    r["song_dur"] = np.float64(dur)
    T = mid_wf.shape[1]
    ref_C_1 = np.max(host(mid_wf.slice_C(0, dur, T, p.pitch_frames, bins_per_tone=1)))        # :271
    ref_C_inst = np.max(host(mid_wf.slice_C(0, dur, T, p.pitch_frames,
                                            bins_per_tone=p.instrument_bins_per_tone)))      # :275
    ref_C_foc = np.max(host(mid_wf.slice_C(0, dur, T, p.pitch_frames,
                                           bins_per_tone=p.instrument_bins_per_tone * foc_bpt_scale)))  # :279
    r["ref_C"] = np.array([ref_C_1, ref_C_inst, ref_C_foc], dtype=np.float64)
    offset = 0
    audio_w = mid_wf.section(offset, None, p.timing_frames)               # :284
    fft_bin_min_const = mid_wf.midi_tone_to_FFT(60)                        # :289
    fft_bin_max_const = fft_bin_min_const + p.instrument_bands
    r["fft_bin_min_const"] = np.int64(fft_bin_min_const)
    # the song's ref_mag is first evaluated HERE (training.py:336), after `section` copied a still-empty `_ref_mag`:
    # the window's first subtraction therefore scales by the window's own maximum, not by the song's
    song_ref = host(mid_wf.ref_mag)
    r["song_ref_mag"] = np.float64(song_ref)

    for i, (onset, duration, pitch) in enumerate(notes):
        k = "n%d_" % i
        if slide_after is not None and i == slide_after:                   # :317-323
            offset += halfwindow_time
            new = mid_wf.section(offset + halfwindow_time, None, halfwindow_frames)
            audio_w.slice(halfwindow_frames, 2 * halfwindow_frames)
            audio_w.concat(new)
            r[k + "slid_shape"] = np.array(audio_w.shape)
        C_timing = AC.compress_bands(audio_w.mag, bands=p.timing_bands)                     # :333
        r[k + "C_timing"] = host(AC._resize(C_timing, p.timing_frames)) / song_ref          # :335
        audio_sw = audio_w.resize(onset, duration, p.pitch_frames, attribs=["mag", "ph"])   # :337
        r[k + "sw_mag"] = host(audio_sw.mag)
        r[k + "C_sw_pitch"] = host(audio_w.slice_C(onset, duration, p.pitch_frames,
                                                   bins_per_tone=p.pitch_bins_per_tone)) / ref_C_1   # :340
        r[k + "C_sw_inst"] = host(audio_w.slice_C(onset, duration, p.instrument_frames,
                                                  bins_per_tone=p.instrument_bins_per_tone)) / ref_C_inst  # :343
        F_const = host(audio_sw.section_power("mag", fft_bin_min_const, fft_bin_max_const))  # :347
        F_const_log10 = np.log10(F_const * 1000 + 1)
        F_const_log10 /= np.max(F_const_log10)
        r[k + "F_const_log10"] = F_const_log10
        r[k + "F_const"] = F_const / song_ref
        fft_bin_min = audio_w.midi_tone_to_FFT(pitch)                                       # :354
        fft_bin_max = fft_bin_min + p.instrument_bands
        r[k + "fft_bin_min"] = np.int64(fft_bin_min)
        F_foc = host(audio_sw.section_power("mag", fft_bin_min, fft_bin_max))               # :356
        F_foc_log10 = np.log10(F_foc * 1000 + 1)
        F_foc_log10 /= np.max(F_foc_log10)
        r[k + "F_foc_log10"] = F_foc_log10
        r[k + "F_foc"] = F_foc / song_ref
        ph = host(audio_sw.section_power("ph", fft_bin_min, fft_bin_max))                   # :362
        r[k + "ph"] = (np.angle(ph) + 3.15) / 6.3
        r[k + "C_foc"] = host(audio_w.slice_C(onset, duration, p.instrument_frames,
                                              bins_per_tone=p.instrument_bins_per_tone * foc_bpt_scale,
                                              highest_note=None, nbins=p.instrument_bands,
                                              lowest_note=midi_to_note(pitch))) / ref_C_foc  # :365
        r[k + "C_foc_const"] = host(audio_w.slice_C(onset, duration, p.instrument_frames,
                                                    bins_per_tone=p.instrument_bins_per_tone * foc_bpt_scale,
                                                    highest_note=None, nbins=p.instrument_bands,
                                                    lowest_note=midi_to_note(60))) / ref_C_foc  # :373
        r[k + "C_velocity"] = host(audio_w.slice_C(onset, duration, p.instrument_frames, bins_per_tone=2,
                                                   highest_note=None, nbins=p.bins_velocity,
                                                   lowest_note=midi_to_note(pitch - 10))) / ref_C_foc  # :382
        guess = AC(clips[i], p.N)                                                           # :426
        if i == 0:                                                                          # :445-447
            sub = audio_w.clone()
            sub.subtract(guess, offset=onset)
            r[k + "clone_after_subtr_ref"] = np.float64(host(sub.ref_mag))
        r[k + "off_frames"] = np.int64(audio_w._seconds_to_frames(onset))
        audio_w.subtract(guess, offset=onset)                                               # :449
        r[k + "mag_after"] = host(audio_w.mag)
        r[k + "ref_after"] = np.float64(host(audio_w.ref_mag))
        r[k + "wf_len_after"] = np.int64(audio_w.wf.shape[0])
    r["D_final"] = host(audio_w.D)
    r["wf_final"] = host(audio_w.wf)
    return r


def run_setters(AC, song, N=2048):
    """Property setters and their invalidation rules (util_audio.py:107-190)."""
    r = {}
    a = AC(song, N)
    mag0, ph0 = host(a.mag).copy(), host(a.ph).copy()
    r["ref0"] = np.float64(host(a.ref_mag))
    D0 = host(a.D).copy()
    r["D0"] = D0
    b = a.clone()
    b.mag = a.mag * 0.5                                   # setter: ref, D, F, wf dropped
    r["ref_half"] = np.float64(host(b.ref_mag))
    r["wf_half"] = host(b.wf)
    c = a.clone()
    c.D = a.D - 6.0                                       # D setter keeps ref_mag
    r["mag_from_D"] = host(c.mag)
    r["ref_after_D"] = np.float64(host(c.ref_mag))
    d = AC(None, N)
    d.F = a.F
    r["wf_from_F"] = host(d.wf)
    r["shape"] = np.array(d.shape)
    return r

"""Pins for the CPU oracle (runs without a GPU).

The reference has no asserting tests and cannot be imported here, so the oracle
is anchored on (1) independent implementations available in this image
(torch.stft/istft, transformers.audio_utils), (2) analytic responses, (3) the
frame-count arithmetic implied by the reference's own FLAC fixtures
(SURVEY.md Appendix C), and (4) the committed golden vectors."""
import os

import numpy as np
import pytest
import torch

from oracle import cqt as ocqt
from oracle import resample as ors
from oracle import spectral as osp
from oracle.audio_oracle import AudioOracle, band_edges
from tests.synth import piano_clip

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("n_fft,hop", [(2048, 512), (4096, 1024)])
def test_stft_matches_torch(cfg1, n_fft, hop):
    y, _ = cfg1
    F = osp.stft(y, n_fft, hop)
    win = torch.hann_window(n_fft, periodic=True, dtype=torch.float64)
    Ft = torch.stft(torch.from_numpy(y).double(), n_fft, hop, window=win, center=True, pad_mode="reflect",
                    return_complex=True).numpy()
    assert F.shape == Ft.shape == (n_fft // 2 + 1, 1 + len(y) // hop)
    assert F.dtype == np.complex64 and np.isfortran(F)
    assert np.abs(F - Ft).max() <= 2e-7 * np.abs(Ft).max()


def test_reflect_index_equals_numpy_pad():
    for L in (1, 2, 3, 5, 17):
        for pad in (1, 2, 7, 40):
            x = np.arange(L, dtype=float) + 1
            assert np.array_equal(np.pad(x, pad, mode="reflect"), x[osp.reflect_index(np.arange(-pad, L + pad), L)])


def test_istft_matches_torch_and_length_rule():
    y = piano_clip(5, 30000)
    F = osp.stft(y, 2048, 512)
    w = osp.istft(F, 512)
    assert w.shape[0] == 512 * (F.shape[1] - 1) and w.dtype == np.float32
    wt = torch.istft(torch.from_numpy(F), 2048, 512, window=torch.hann_window(2048), center=True,
                     length=w.shape[0]).numpy()
    assert np.abs(w[2048:-2048] - wt[2048:-2048]).max() < 1e-5


def test_amplitude_to_db_matches_transformers():
    au = pytest.importorskip("transformers.audio_utils")
    rng = np.random.default_rng(0)
    S = (rng.random((257, 40)) ** 6).astype(np.float32)
    ours = osp.amplitude_to_db(S, ref=S.max())
    theirs = au.amplitude_to_db(S.astype(np.float64), reference=float(S.max()), min_value=1e-5, db_range=80.0)
    assert np.abs(ours - theirs).max() < 1e-4
    assert ours.max() == 0.0 and ours.min() == -80.0
    assert np.allclose(osp.db_to_amplitude(ours, ref=S.max())[ours > -79], S[ours > -79], rtol=1e-5)


def test_magphase_zero_bin_convention():
    mag, ph = osp.magphase(np.zeros((3, 2), dtype=np.complex64))
    assert np.all(mag == 0) and np.all(ph == 1 + 0j)


def test_fixture_frame_arithmetic():
    """lengths of the reference's own FLAC dumps."""
    # *_test.flac 132300 samples -> T = 130; *_test_sub.flac = istft -> 1024 * 129 samples
    T = 1 + 132300 // 1024
    assert T == 130 and 1024 * (T - 1) == 132096
    F = osp.stft(np.zeros(132300, dtype=np.float32) + 1e-3, 4096, 1024)
    assert F.shape == (2049, 130) and osp.istft(F, 1024).shape[0] == 132096
    # window dumps: 263168 = 1024 * 257 -> timing_frames 258 = int(6 * 44100 / 1024)
    assert int(6 * 44100 / 1024) == 258 and 1024 * 257 == 263168
    # short_window_demo/{j}: _resize returns exactly j frames -> 1024*(j-1) samples
    for j in (6, 8, 10, 15, 20):
        P = np.arange(3 * 11, dtype=float).reshape(3, 11)
        for t in (0, 1, 2, 3, 5, 11):
            assert AudioOracle._resize(P[:, :t], j).shape == (3, j)


def test_resize_rules():
    P = np.arange(2 * 5, dtype=float).reshape(2, 5)
    r = AudioOracle._resize(P, 8)                      # t < target: first col, tiled middle, tail
    assert r.shape == (2, 8) and np.array_equal(r[0], [0, 1, 2, 3, 1, 2, 3, 4])
    assert np.array_equal(AudioOracle._resize(P[:, :2], 4)[0], [0, 1, 1, 1])
    assert np.array_equal(AudioOracle._resize(P, 3), P[:, :3])
    assert AudioOracle._resize(P[:, :0], 4).sum() == 0


def test_band_edges_and_bins():
    assert list(band_edges(2049, 20)) == [0, 1, 2, 3, 4, 6, 9, 14, 21, 30, 45, 66, 97, 142, 208, 304, 445,
                                          652, 955, 1399, 2049]
    assert AudioOracle(np.zeros(8), 4096).midi_tone_to_FFT(60) == 23
    assert osp.note_to_midi("A0") == 21 and osp.note_to_midi("C1") == 24 and osp.note_to_midi("C8") == 108
    assert osp.midi_to_note(60) == "C4" and osp.midi_to_note(61) == "C#4"
    assert abs(osp.note_to_hz("A4") - 440.0) < 1e-9


def test_resampler_fir_equals_literal_loop_and_has_unit_dc_gain():
    rng = np.random.default_rng(0)
    for D in (2, 4, 8):
        x = rng.standard_normal(333)
        win, tab, _ = ors.get_filter("kaiser_fast")
        win = win / D
        dl = np.zeros_like(win)
        dl[:-1] = np.diff(win)
        lit = ors.resample_f_literal(x, int(333 / D), 1.0 / D, win, dl, tab)
        assert np.allclose(lit, ors.decimate_fir(x, D), atol=1e-13)
    taps = ors.decimation_taps(2)
    assert taps.shape == (32,)                              # 63-tap symmetric FIR
    assert abs(taps[0] + 2 * taps[1:].sum() - 1.0) < 2e-4   # DC gain (passband error ~1e-4)
    y = ors.librosa_resample(np.ones(1001), 2, 1)
    assert y.shape[0] == 501 and abs(y[250] - np.sqrt(2)) < 1e-3   # fix_length + 1/sqrt(ratio) scale


def test_cqt_analytic_response_and_geometry():
    sr = 16000
    t = np.arange(3 * sr)
    fmin = osp.note_to_hz("C1")
    L = ocqt.constant_q_lengths(sr, fmin, 84, 12, 0.0, 2)
    for k in (30, 45, 60):
        y = np.sin(2 * np.pi * fmin * 2 ** (k / 12) * t / sr)
        C = np.abs(ocqt.cqt(y, sr=sr, hop_length=512, fmin=fmin, n_bins=84, filter_scale=2))
        assert C.shape == (84, 1 + len(y) // 512)
        assert np.argmax(C[:, 40]) == k
        assert abs(C[k, 40] / (0.5 * np.sqrt(L[k])) - 1) < 3e-3      # (A/2) sqrt(L_k), up to the 1% sparsification
    plan = ocqt.cqt_plan(44100, 512, fmin, 84, 12, 0.0, 2)
    assert plan["early_factor"] == 2 and [j["hop"] for j in plan["jobs"]] == [256, 128, 64, 32, 16, 8, 4]
    nnz = [(np.abs(j["fft_basis"]) > 0).sum(axis=1) for j in plan["jobs"]]
    assert min(n.min() for n in nnz) == 9 and max(n.max() for n in nnz) == 17   # SURVEY probe
    with pytest.raises(ocqt.ParameterError):
        ocqt.cqt(np.zeros(4096), sr=44100, hop_length=100, fmin=fmin, n_bins=84, filter_scale=2)


def test_subtract_semantics_of_the_container():
    """ref_mag staleness and setter invalidation (util_audio.py:149-157, :323)."""
    sr, N = 44100, 2048
    song = piano_clip(1, sr * 2)
    note = piano_clip(2, sr // 2, n_notes=1)
    a = AudioOracle(song, N, 512)
    a.mag
    song_ref = a.ref_mag
    w = a.section(0, None, 100)
    assert w._ref == song_ref                       # section copies the song-level ref
    before = w.mag.copy()
    g = AudioOracle(note, N, 512)
    w.subtract(g, offset=0.3)
    assert w._ref is None and w._v["D"] is None and w._v["F"] is None
    assert w.ref_mag == np.max(w.mag) and w.mag.min() >= 0
    changed = np.nonzero(np.any(w.mag != before, axis=0))[0]
    assert changed.min() >= 25 and changed.max() < 100    # 0.3 s at hop 512 -> frame 25/26
    assert w.wf.shape[0] == 512 * 99                       # wf rebuilt by iSTFT: hop*(T-1)


def test_golden_vectors_still_reproduced(cfg1):
    y, sr = cfg1
    g = np.load(os.path.join(GOLD, "cfg1.npz"))
    a = AudioOracle(y, 2048, 512, sample_rate=sr)
    assert tuple(g["mag_shape"]) == a.mag.shape == (1025, 313)
    assert np.array_equal(a.mag[:, g["cols"]], g["mag_cols"])
    assert np.allclose(a.D[:, g["cols"]], g["D_cols"], atol=1e-5)
    C = a.slice_C(0, 10.0, 313, bins_per_tone=1, lowest_note="C1", nbins=84)
    assert tuple(g["cqt_shape"]) == C.shape == (84, 313)
    assert np.allclose(C[:, g["cols"]], g["cqt_cols"], rtol=1e-9, atol=1e-12)
    s = np.load(os.path.join(GOLD, "subtract_chain.npz"))
    w = AudioOracle(None, 256, 64)
    w.mag = s["win"].copy()
    for j in range(3):
        # the mag setter drops wf (util_audio.py:157); with no phase there is nothing to rebuild it
        # from, so hand the container a waveform of the right length (only len() enters the frame map)
        w._v["wf"] = np.zeros(64 * 39, dtype=np.float32)
        w.subtract(s["guesses"][j], offset=w._frames_to_seconds(int(s["offsets"][j])) + 1e-9)
    assert np.array_equal(w.mag, s["result"])
    assert np.allclose(w.D, s["D"], atol=1e-5)


def test_basis_fft_precision():
    """librosa 0.6.3 transforms the complex64 constant-Q basis with scipy.fftpack.fft, i.e. in SINGLE precision; the
    oracle calls the same function (numpy >= 2.0's np.fft.fft keeps complex64 as well).  The precision of that
    transform is not a parity risk: single (either library) and double precision agree to 2e-7 of the basis peak and
    select the SAME 1 % sparsity pattern at every (octave rate, bins per octave) the reference's calls produce."""
    from scipy import fftpack
    assert ocqt._basis_fft(np.zeros((2, 8), dtype=np.complex64), 8).dtype == np.complex64
    for sr, fmin, nf, bpo in ((22050, 1046.5, 12, 12), (22050, 1046.5, 24, 24), (22050, 1046.5, 48, 48), (11025, 523.25, 192, 192)):
        basis, lengths = ocqt.constant_q(sr, fmin, nf, bpo, 0.0, 2, 1)
        n_fft = basis.shape[1]
        basis = (basis.astype(np.complex128) * (lengths[:, None] / float(n_fft))).astype(np.complex64)
        half = slice(0, n_fft // 2 + 1)
        single_sp = fftpack.fft(basis, n=n_fft, axis=1)[:, half]
        single_np = np.fft.fft(basis, n=n_fft, axis=1)[:, half]
        double = np.fft.fft(basis.astype(np.complex128), n=n_fft, axis=1)[:, half]
        peak = np.abs(double).max()
        assert np.abs(single_sp - double).max() <= 2e-7 * peak and np.abs(single_np - double).max() <= 2e-7 * peak
        keep = [ocqt.sparsify_rows(x, quantile=0.01) != 0 for x in (single_sp, single_np, double)]
        assert np.array_equal(keep[0], keep[1]) and np.array_equal(keep[0], keep[2])

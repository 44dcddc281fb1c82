"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the
same inputs.  Tolerances are BASELINE.json's: frame indexing / bin layout exact,
magnitudes within 1e-4 of the per-clip peak, dB within 0.01 dB above the floor.
"""
import numpy as np
import pytest
import torch

from oracle import cqt as ocqt
from oracle import spectral as osp
from oracle.audio_oracle import AudioOracle
from tests.synth import piano_clip

pytestmark = pytest.mark.gpu

MAG_TOL = 1e-4      # relative to the per-clip peak (north_star)
DB_TOL = 0.01       # dB, where the oracle is above its floor


@pytest.fixture(scope="module")
def saga():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import amt_saga_b200  # noqa: F401
    from amt_saga_b200 import ops, util_audio
    return ops, util_audio


def dev(x):
    return torch.as_tensor(x, device="cuda")


def check_mag(got, ref, tol=MAG_TOL):
    assert got.shape == ref.shape
    peak = max(float(np.abs(ref).max()), 1e-30)
    wide = np.complex128 if (np.iscomplexobj(got) or np.iscomplexobj(ref)) else np.float64
    err = float(np.abs(got.astype(wide) - ref.astype(wide)).max()) / peak
    assert err <= tol, "max |err| / peak = %.3e" % err
    return err


def check_db(got, ref, floor_margin=1e-3):
    above = ref > (ref.max() - 80.0 + floor_margin)
    err = float(np.abs(got[above].astype(np.float64) - ref[above]).max()) if above.any() else 0.0
    assert err <= DB_TOL, "max dB err above the floor = %.4f" % err
    # and the floor itself is reproduced
    assert np.all(got >= ref.max() - 80.0 - 1e-3)


# --------------------------------------------------------------------------- K1
@pytest.mark.parametrize("n_fft,hop", [(2048, 512), (4096, 1024), (1024, 256), (512, 128), (256, 64), (8192, 2048)])
def test_stft_mag_phase_cfg1(saga, cfg1, n_fft, hop):
    ops, _ = saga
    y, sr = cfg1
    plan = ops.StftPlan(n_fft, hop, True)
    r = ops.stft_batch(dev(y), plan, want_phase=True, want_complex=True)
    F = osp.stft(y, n_fft, hop)
    mag, ph = osp.magphase(F)
    assert tuple(r["mag"].shape) == (1,) + mag.shape == (1, n_fft // 2 + 1, 1 + len(y) // hop)
    check_mag(r["mag"][0].cpu().numpy(), mag)
    check_mag(r["F"][0].cpu().numpy(), F)
    # phase only where the magnitude is meaningfully above rounding noise
    g = r["phase"][0].cpu().numpy()
    big = mag > 1e-3 * mag.max()
    assert np.abs(g[big] - ph[big]).max() < 2e-3
    assert abs(float(r["clip_max"][0]) - float(mag.max())) <= 1e-6 * mag.max()
    assert np.allclose(r["frame_max"][0].cpu().numpy(), mag.max(axis=0), rtol=1e-5, atol=1e-6 * mag.max())


def test_stft_db_within_tolerance(saga, cfg1):
    ops, _ = saga
    y, _ = cfg1
    for n_fft, hop in [(2048, 512), (4096, 1024)]:
        plan = ops.StftPlan(n_fft, hop, True)
        r = ops.stft_batch(dev(y), plan)
        D = ops.amplitude_to_db_batch(r["mag_storage"], plan.n_bins)
        ref_mag = np.abs(osp.stft(y, n_fft, hop))
        ref = osp.amplitude_to_db(ref_mag, ref=ref_mag.max())
        check_db(D[0, :, :plan.n_bins].T.cpu().numpy(), ref)


@pytest.mark.parametrize("n_fft,hop", [(2048, 512), (4096, 1024), (8192, 2048)])
@pytest.mark.parametrize("center", [True, False])
def test_stft_ragged_batch_and_edges(saga, center, n_fft, hop):
    ops, _ = saga
    rng = np.random.default_rng(3)
    lens = [5000, 2048, 2049, 1, 300, 44100, 1023, 4095, 0 + 2048 * 3 + 7, 8192, 8193, 4096 * 5 + 1]
    if not center:
        lens = [x for x in lens if x >= n_fft]
    width = max(lens)
    wav = np.zeros((len(lens), width), dtype=np.float32)
    for i, n in enumerate(lens):
        wav[i, :n] = rng.standard_normal(n).astype(np.float32) * (0.1 + i)
    plan = ops.StftPlan(n_fft, hop, center)
    r = ops.stft_batch(dev(wav), plan, lens=lens)
    for i, n in enumerate(lens):
        ref = np.abs(osp.stft(wav[i, :n], n_fft, hop, center=center))
        T = ref.shape[1]
        assert T == plan.num_frames(n)
        got = r["mag"][i].cpu().numpy()
        check_mag(got[:, :T], ref)
        assert np.all(got[:, T:] == 0)


@pytest.mark.parametrize("n_fft,hop", [(2048, 512), (4096, 1024)])
def test_stft_all_zero_clip(saga, n_fft, hop):
    ops, _ = saga
    plan = ops.StftPlan(n_fft, hop, True)
    r = ops.stft_batch(torch.zeros(1, 8000, device="cuda"), plan, want_phase=True)
    assert float(r["mag"].abs().max()) == 0.0
    ph = r["phase"][0].cpu().numpy()
    assert np.all(ph == 1.0 + 0.0j)      # angle(0) = 0 -> phase 1+0j (magphase convention)
    D = ops.amplitude_to_db_batch(r["mag_storage"], plan.n_bins)
    assert np.all(D[0, :, :plan.n_bins].cpu().numpy() == 0.0)   # max(amin, 0) path: D == 0 everywhere


@pytest.mark.parametrize("n_fft,hop", [(2048, 512), (4096, 1024)])
def test_stft_piano_noise_floor_db(saga, n_fft, hop):
    ops, _ = saga
    y = piano_clip(11, 44100 * 2)
    plan = ops.StftPlan(n_fft, hop, True)
    r = ops.stft_batch(dev(y), plan)
    ref = np.abs(osp.stft(y, n_fft, hop))
    check_mag(r["mag"][0].cpu().numpy(), ref)
    D = ops.amplitude_to_db_batch(r["mag_storage"], plan.n_bins)
    check_db(D[0, :, :plan.n_bins].T.cpu().numpy(), osp.amplitude_to_db(ref, ref=ref.max()))


def test_stft_4096_split_kernel_against_three_pass_form(saga):
    """n_fft 4096 (the reference's default N) runs as two 1024-point half transforms + one combine stage; the
    three-pass radix-16/16/8 kernel stays available (SAGA_STFT_NO_EO) and both must agree to fp32 rounding on
    magnitude, phase and complex outputs, including the clip-edge tiles and the per-clip / per-frame maxima."""
    import os
    ops, _ = saga
    y = np.stack([piano_clip(70 + i, 44100) for i in range(3)])
    plan = ops.StftPlan(4096, 1024, True)
    out = {}
    for mode in ("split", "three_pass"):
        with ops.options(SAGA_STFT_NO_EO="1" if mode == "three_pass" else None):
            r = ops.stft_batch(dev(y), plan, want_phase=True, want_complex=True, want_max=True)
            out[mode] = {k: r[k].cpu().numpy() for k in ("mag", "phase", "F", "clip_max", "frame_max") if k in r}
    a, b = out["split"], out["three_pass"]
    peak = float(b["mag"].max())
    assert np.abs(a["mag"] - b["mag"]).max() <= 2e-6 * peak
    assert np.abs(a["F"] - b["F"]).max() <= 2e-6 * peak
    assert np.allclose(a["clip_max"], b["clip_max"], rtol=2e-6)
    ref = np.stack([osp.stft(y[i], 4096, 1024) for i in range(3)])
    check_mag(a["F"], ref, tol=2e-6)
    strong = np.abs(ref) > 1e-3 * peak                     # phase is only meaningful away from the noise floor
    assert np.abs(a["phase"][strong] - ref[strong] / np.abs(ref[strong])).max() <= 1e-3


# --------------------------------------------------------------------------- K4
@pytest.mark.parametrize("n_fft,hop", [(2048, 512), (4096, 1024), (1024, 256)])
def test_istft_matches_oracle(saga, n_fft, hop):
    ops, _ = saga
    y = piano_clip(5, 30000)
    F = osp.stft(y, n_fft, hop)
    ref = osp.istft(F, hop)
    plan = ops.StftPlan(n_fft, hop, True)
    r = ops.stft_batch(dev(y), plan, want_phase=True, want_complex=True)
    w1 = ops.istft_batch(plan, F=r["F_storage"])
    w2 = ops.istft_batch(plan, mag=r["mag_storage"], phase=r["phase_storage"])
    assert w1.shape[1] == ref.shape[0] == hop * (F.shape[1] - 1)
    for w in (w1, w2):
        assert np.abs(w[0].cpu().numpy() - ref).max() <= 2e-5 * np.abs(ref).max()
    # round trip reproduces the interior of the signal
    n = min(len(y), ref.shape[0])
    assert np.abs(w1[0, n_fft:n - n_fft].cpu().numpy() - y[n_fft:n - n_fft]).max() < 1e-4


# --------------------------------------------------------------------------- K3
def _numpy_chain(win, guesses, offs, overkill=None):
    """numpy float32 replay of util_audio.py:236-259 applied S times."""
    win = win.copy()
    for j in range(guesses.shape[0]):
        ref, gref = np.max(win), np.max(guesses[j])
        g = guesses[j].copy()
        g *= ref / gref
        if overkill is not None:
            g *= np.float32(overkill[j])
        off = int(offs[j])
        T = win.shape[1]
        if off >= T:
            continue
        g = g[:, : T - off]
        pad = np.concatenate((np.zeros((win.shape[0], off)), g,
                              np.zeros((win.shape[0], T - off - g.shape[1]))), axis=1)
        win -= pad
        win = np.maximum(win, 0, win)
    return win


@pytest.mark.parametrize("B,T,Tg,S", [(1025, 516, 128, 16), (2049, 258, 54, 3), (1025, 40, 128, 2), (129, 33, 5, 4)])
def test_subtract_chain_bit_exact(saga, B, T, Tg, S):
    ops, _ = saga
    rng = np.random.default_rng(B + T)
    W = 5
    P = ops.frame_pitch(B)
    win = rng.random((W, B, T), dtype=np.float32) ** 4
    gs = rng.random((W, S, B, Tg), dtype=np.float32) ** 3
    offs = rng.integers(0, T, size=(W, S)).astype(np.int32)
    offs[0, 0] = 0
    offs[1, 0] = T - 1
    ok = (1.0 + rng.random((W, S))).astype(np.float32)
    for overkill in (None, ok):
        st = torch.zeros((W, T, P), device="cuda")
        st[:, :, :B] = dev(win).transpose(1, 2)
        g = torch.zeros((W, S, Tg, P), device="cuda")
        g[:, :, :, :B] = dev(gs).transpose(2, 3)
        D, ref = ops.subtract_db_batch(st, g, dev(offs), B, overkill=None if overkill is None else dev(overkill))
        got = st[:, :, :B].transpose(1, 2).cpu().numpy()
        for w in range(W):
            exp = _numpy_chain(win[w], gs[w], offs[w], None if overkill is None else overkill[w])
            assert np.array_equal(got[w], exp), "window %d differs (max %.3e)" % (w, np.abs(got[w] - exp).max())
            assert float(ref[w]) == float(exp.max())
            check_db(D[w, :, :B].T.cpu().numpy(), osp.amplitude_to_db(exp, ref=exp.max()))
        assert float(st[:, :, B:].abs().max()) == 0.0


@pytest.mark.parametrize("B,T,Tg", [(1025, 516, 128), (2049, 258, 54), (129, 33, 40), (1025, 7, 3)])
def test_subtract_single_step_flat_kernel_bit_exact(saga, B, T, Tg):
    """One guessed note per window with K1's by-products at hand (the producer loop's case) runs on the flat
    single-step kernel: bit-identical to the numpy replay, to the one-CTA-per-window chain kernel and to both dB
    kernels, for offsets outside / at the edge of the window, ragged guess lengths, optional ref_init / overkill,
    and garbage in the window's padding columns (neither changed nor counted)."""
    import os
    ops, _ = saga
    rng = np.random.default_rng(B + T + Tg)
    W = 11
    P = ops.frame_pitch(B)
    win = rng.random((W, B, T), dtype=np.float32) ** 4
    gs = rng.random((W, 1, B, Tg), dtype=np.float32) ** 3
    offs = rng.integers(0, T, size=(W, 1)).astype(np.int32)
    offs[0, 0], offs[1, 0], offs[2, 0], offs[3, 0] = 0, T - 1, T + 5, -3
    gframes = rng.integers(1, Tg + 1, size=(W, 1)).astype(np.int32)
    ok = (1.0 + rng.random((W, 1))).astype(np.float32)
    ref_init = np.where(rng.random(W) < 0.5, -1.0, 2.5).astype(np.float32)
    st0 = torch.zeros((W, T, P), device="cuda")
    st0[:, :, :B] = dev(win).transpose(1, 2)
    st0[:, :, B:] = 7.0                                   # garbage where the layout says padding
    g = torch.zeros((W, 1, Tg, P), device="cuda")
    g[:, :, :, :B] = dev(gs).transpose(2, 3)
    fmax = st0[:, :, :B].amax(dim=2).contiguous()
    out = {}
    for mode in ("flat", "chain", "shallow_db"):
        opt = {"chain": dict(SAGA_SUB_NO_FLAT="1"), "shallow_db": dict(SAGA_DB_LEAN="0")}.get(mode, {})
        with ops.options(**opt):
            res = []
            for kw in (dict(), dict(overkill=dev(ok)), dict(guess_frames=dev(gframes), ref_init=dev(ref_init))):
                st = st0.clone()
                gref = torch.stack([g[w, 0, :int(gframes[w, 0]) if "guess_frames" in kw else Tg].amax() for w in range(W)])
                D, ref = ops.subtract_db_batch(st, g, dev(offs), B, frame_max=fmax, guess_ref=gref.reshape(W, 1), **kw)
                res.append((st.cpu().numpy(), D[:, :, :B].cpu().numpy(), ref.cpu().numpy()))
            out[mode] = res
    for mode in ("chain", "shallow_db"):
        for a, b in zip(out["flat"], out[mode]):
            for x, y in zip(a, b):
                assert np.array_equal(x, y), mode
    assert np.all(out["flat"][0][0][:, :, B:] == 7.0)
    got = out["flat"][1][0][:, :, :B].transpose(0, 2, 1)
    for w in range(W):
        if 0 <= offs[w, 0] < T:
            exp = _numpy_chain(win[w], gs[w], offs[w], ok[w])
        else:
            exp = win[w]
        assert np.array_equal(got[w], exp)
        assert float(out["flat"][1][2][w]) == float(exp.max())


@pytest.mark.parametrize("normalize,relu", [(True, True), (False, True), (True, False)])
def test_subtract_cluster_kernel_equals_single_cta_kernel(saga, normalize, relu):
    """Multi-step chains run on clusters of 4 CTAs per window; every optional input (ragged guess lengths,
    caller-provided guess / initial reference levels, K1 frame maxima) must give bit-identical results to
    the one-CTA-per-window kernel."""
    import os
    ops, _ = saga
    rng = np.random.default_rng(5)
    W, B, T, Tg, S = 9, 1025, 131, 40, 5
    P = ops.frame_pitch(B)
    win = torch.zeros((W, T, P), device="cuda")
    win[:, :, :B] = dev(rng.random((W, T, B), dtype=np.float32) ** 3)
    g = torch.zeros((W, S, Tg, P), device="cuda")
    g[:, :, :, :B] = dev(rng.random((W, S, Tg, B), dtype=np.float32) ** 2)
    offs = dev(rng.integers(-2, T + 3, size=(W, S)).astype(np.int32))
    gframes = dev(rng.integers(1, Tg + 1, size=(W, S)).astype(np.int32))
    gref = g.amax(dim=(2, 3)) * 1.25
    ref_init = dev(np.where(rng.random(W) < 0.5, -1.0, 3.0).astype(np.float32))
    fmax = win.amax(dim=2).contiguous()
    out = {}
    for mode in ("cluster", "single"):
        with ops.options(SAGA_SUB_NO_CLUSTER="1" if mode == "single" else None):
            res = []
            for kw in (dict(), dict(guess_ref=gref), dict(guess_frames=gframes, ref_init=ref_init),
                       dict(frame_max=fmax, guess_ref=gref, guess_frames=gframes)):
                st = win.clone()
                D, ref = ops.subtract_db_batch(st, g, offs, B, normalize=normalize, relu=relu, **kw)
                res.append((st.cpu().numpy(), D.cpu().numpy(), ref.cpu().numpy()))
            out[mode] = res
    for a, b in zip(out["cluster"], out["single"]):
        for x, y in zip(a, b):
            assert np.array_equal(x, y)


def test_subtract_matches_oracle_class(saga):
    """audio_complete.subtract vs the oracle container, incl. the stale song-level
    ref_mag a `section` hands to its first subtraction (util_audio.py:323)."""
    _, ua = saga
    sr, N = 44100, 4096
    song = piano_clip(21, sr * 8)
    note = piano_clip(22, sr * 1, n_notes=1)
    a, o = ua.audio_complete(song, N), AudioOracle(song, N)
    a.mag, o.mag
    aw, ow = a.section(0, None, 258), o.section(0, None, 258)
    for onset in (0.5, 2.25, 5.9):
        aw.subtract(ua.audio_complete(note, N), offset=onset)
        ow.subtract(AudioOracle(note, N), offset=onset)
        check_mag(aw.mag.cpu().numpy(), ow.mag, tol=2e-5)
        assert abs(float(aw.ref_mag) - float(ow.ref_mag)) <= 2e-5 * float(ow.ref_mag)
    check_db(aw.D.cpu().numpy(), ow.D)


# --------------------------------------------------------------------------- K2
CQT_CASES = [
    (16000, 512, "C1", 84, 12, 160000),
    (44100, 512, "C1", 84, 12, 66150),
    (44100, 1024, "A0", 87, 12, 88200),
    (44100, 1024, "A0", 174, 24, 66150),
    (44100, 1024, "D3", 36, 24, 66150),
    (44100, 512, "C1", 84, 12, 512 * 133 + 100),    # 134 frames: the 6-frame partial tile goes to cqt_tail_kernel
    (44100, 1024, "A0", 348, 48, 66150),            # ref_C_4 (training.py:277): 96 columns per octave, fp32 contraction
    (44100, 1024, "C4", 348, 192, 132300),          # bins_per_tone=16 from the note (training.py:366-381): 3 column blocks
    (44100, 1024, "F#1", 36, 24, 264192),           # C_velocity below a low note (training.py:382): early decimation by 64
]


# impl: 1 = fp32 CUDA-core contraction, 2 = tcgen05 3xTF32 (fp32-grade), 3 = tcgen05 single TF32
@pytest.mark.parametrize("impl", [1, 2, 3, 0])
@pytest.mark.parametrize("sr,hop,low,n_bins,bpo,n", CQT_CASES)
def test_cqt_matches_oracle(saga, cfg1, sr, hop, low, n_bins, bpo, n, impl):
    ops, _ = saga
    from amt_saga_b200._lib import SagaUnsupported
    y = cfg1[0][:n] if sr == 16000 else piano_clip(31, n, sr=sr)
    fmin = osp.note_to_hz(low)
    ref = ocqt.cqt(y, sr=sr, hop_length=hop, fmin=fmin, n_bins=n_bins, bins_per_octave=bpo, filter_scale=2)
    plan = ops.CqtPlan(sr, hop, fmin, n_bins, bpo, filter_scale=2)
    try:
        r = ops.cqt_batch(dev(y), plan, want_complex=True, impl=impl, fill=float("nan"))
    except SagaUnsupported:
        pytest.skip("bank does not fit the resident-B tensor path (falls back to fp32 under impl=0)")
    assert tuple(r["mag"].shape) == (1,) + ref.shape
    # fp32 path ~1e-6; 3xTF32 ~3e-6 (tensor-core accumulation); both far inside the 1e-4 bar.
    # impl=3 (single TF32 pass) measures ~1.1e-4: NOT parity-grade, opt-in only, checked loosely.
    # 192 bins per octave = 8192-sample kernels: longer fp32 sums, and on the streamed tensor path (impl 0) twice
    # the accumulation steps per TMEM partial that the 2048-sample kernels see (cqt_umma_stream.cu): 2e-5
    tol = {1: 2e-6 if bpo <= 48 else 5e-6,
           2: 1e-5 if bpo <= 48 else 2e-5, 0: 1e-5 if bpo <= 48 else 2e-5, 3: 3e-4}[impl]
    check_mag(r["mag"][0].cpu().numpy(), np.abs(ref), tol=tol)
    check_mag(r["C"][0].cpu().numpy(), ref, tol=tol)


@pytest.mark.parametrize("sr,hop,low,n_bins,bpo,n", [c for c in CQT_CASES if c[0] == 44100 and c[1] == 1024])
def test_cqt_frame_window_equals_columns_of_the_full_transform(saga, sr, hop, low, n_bins, bpo, n):
    """saga_cqt_frames_exec (the per-note form: only the <= 8 columns slice_C + _resize keep) against the same
    columns of the full transform and of the oracle, at the clip start, in the interior, across the clip end and
    past it (zeros), on a ragged batch."""
    ops, _ = saga
    fmin = osp.note_to_hz(low)
    plan = ops.CqtPlan(sr, hop, fmin, n_bins, bpo, filter_scale=2)
    lens = [n, n - 5000, n - 1]
    wav = np.zeros((3, n), dtype=np.float32)
    for i, m in enumerate(lens):
        wav[i, :m] = piano_clip(50 + i, m, sr=sr)
    full = ops.cqt_batch(dev(wav), plan, lens=lens, impl=1)["mag"].cpu().numpy()          # [3, bins, T]
    for first in ([0, 0, 0], [3, 11, 20], [n // hop - 2, (n - 5000) // hop - 6, n // hop + 5]):
        got = ops.cqt_frames_batch(dev(wav), plan, np.array(first, dtype=np.int32), 8, lens=lens)
        got = got[:, :, :n_bins].transpose(1, 2).cpu().numpy()                           # [3, bins, 8]
        for i, m in enumerate(lens):
            T = plan.num_frames(m)
            ref = np.abs(ocqt.cqt(wav[i, :m], sr=sr, hop_length=hop, fmin=fmin, n_bins=n_bins, bins_per_octave=bpo,
                                  filter_scale=2)) if first[0] == 3 else None
            for j in range(8):
                t = first[i] + j
                if t >= T:
                    assert np.all(got[i, :, j] == 0)
                    continue
                tol = 1e-5 if bpo <= 48 else 2e-5       # streamed tensor-core contraction (see test_cqt_matches_oracle)
                assert np.abs(got[i, :, j] - full[i, :, t]).max() <= tol * full[i].max()
                if ref is not None:
                    assert np.abs(got[i, :, j] - ref[:, t]).max() <= tol * ref.max()


@pytest.mark.parametrize("rows_in_smem", [False, True], ids=["rows_in_tmem", "rows_in_smem"])
@pytest.mark.parametrize("low,n_bins,bpo", [("A0", 174, 24), ("A0", 348, 48), ("C4", 348, 192), ("C1", 84, 12)])
def test_cqt_streamed_bank_kernel_against_its_fp32_twin(saga, low, n_bins, bpo, rows_in_smem):
    """cqt_umma_stream_kernel (gathered rows x streamed bank on tcgen05: the transforms whose bank does not fit the
    resident kernel, and every frame window) against the fp32 CUDA-core kernels (SAGA_CQT_STREAM=0) on a ragged
    batch whose 128-row tiles straddle clips: whole transform (magnitude and complex), an empty-frame clip tail,
    and 8-column frame windows before the start, inside, across and past the end of the clips.  Both operand
    placements: rows (A) in tensor memory (default where the accumulators leave room for a deep enough ring: 12 / 24
    per octave) and in shared memory (SAGA_CQT_STREAM_SS=1; what 48 and 192 per octave use anyway)."""
    ops, _ = saga
    sr, hop = 44100, 1024
    plan = ops.CqtPlan(sr, hop, osp.note_to_hz(low), n_bins, bpo, filter_scale=2)
    n = 66150 if bpo < 192 else 99000
    lens = [n, n - 7001, n - 1, 30000 if bpo < 192 else 70000, n, n - 12345, n - 2048]
    wav = np.zeros((len(lens), n), dtype=np.float32)
    for i, m in enumerate(lens):
        wav[i, :m] = piano_clip(70 + i, m, sr=sr)
    x = dev(wav)
    tol = 1e-5 if bpo <= 48 else 2e-5
    firsts = [np.array([-3, 0, 5, 11, 20, 40, 60], dtype=np.int32),
              np.array([plan.num_frames(m) - 4 for m in lens], dtype=np.int32),
              np.array([plan.num_frames(m) + 2 for m in lens], dtype=np.int32)]
    with ops.options(SAGA_CQT_STREAM_SS="1" if rows_in_smem else None):
        got = ops.cqt_batch(x, plan, lens=lens, want_complex=True, fill=float("nan"))
        fr = [ops.cqt_frames_batch(x, plan, f, 8, lens=lens).clone() for f in firsts]
        fr5 = ops.cqt_frames_batch(x, plan, firsts[0], 5, lens=lens).clone()
        # same call again: bit-identical (fixed MMA order per tile, split-K slices added in slice order)
        again = ops.cqt_batch(x, plan, lens=lens, want_complex=True, fill=float("nan"))
        for i, m in enumerate(lens):
            T = plan.num_frames(m)
            assert torch.equal(again["mag"][i, :, :T], got["mag"][i, :, :T])
        assert torch.equal(ops.cqt_frames_batch(x, plan, firsts[0], 8, lens=lens), fr[0])
    with ops.options(SAGA_CQT_STREAM="0"):
        ref = ops.cqt_batch(x, plan, lens=lens, want_complex=True, impl=1, fill=float("nan"))
        fr_ref = [ops.cqt_frames_batch(x, plan, f, 8, lens=lens).clone() for f in firsts]
    peak = max(float(ref["mag"][i, :, :plan.num_frames(m)].max()) for i, m in enumerate(lens))   # columns >= T keep the fill
    for i, m in enumerate(lens):
        T = plan.num_frames(m)
        a, b = got["mag"][i, :, :T], ref["mag"][i, :, :T]
        assert torch.isfinite(a).all()
        assert float((a - b).abs().max()) <= tol * peak
        assert float((got["C"][i, :, :T] - ref["C"][i, :, :T]).abs().max()) <= tol * peak
    for g, r in zip(fr, fr_ref):
        assert torch.isfinite(g).all()
        assert float((g - r).abs().max()) <= tol * peak
        assert bool((g[r == 0] == 0).all())                     # dead columns and pitch padding are exact zeros
    assert float((fr5[:, :5] - fr_ref[0][:, :5]).abs().max()) <= tol * peak
    assert float(fr[2].abs().max()) == 0.0                     # windows past the clip end are all zero


@pytest.mark.parametrize("midi,factor", [(69, 4), (57, 8), (45, 16), (33, 32), (21, 64), (12, 128)])
def test_cqt_polyphase_early_decimator_against_the_generic_one(saga, midi, factor):
    """decimate_phase_kernel (polyphase form of the early resampy stage for factors >= 4: the note-relative transforms
    of training.py:366-388 start below the note and decimate by up to 128 first) against the generic tile decimator
    (SAGA_DEC_NO_PHASE=1), through the fp32 contraction, on a ragged batch with lengths that are not multiples of the
    factor (librosa's zero-padded last sample), and against the oracle."""
    ops, _ = saga
    sr, hop = 44100, 1024
    fmin = 440.0 * 2.0 ** ((midi - 69) / 12.0)
    plan = ops.CqtPlan(sr, hop, fmin, 36, 24, filter_scale=2)
    assert plan.early_factor == factor
    n = 264600
    lens = [n, n - 12345, n - 1, 200001]
    wav = np.zeros((len(lens), n), dtype=np.float32)
    for i, m in enumerate(lens):
        wav[i, :m] = piano_clip(80 + i, m, sr=sr)
    x = dev(wav)
    got = ops.cqt_batch(x, plan, lens=lens, impl=1, fill=float("nan"))["mag"]
    with ops.options(SAGA_DEC_NO_PHASE="1"):
        ref = ops.cqt_batch(x, plan, lens=lens, impl=1, fill=float("nan"))["mag"]
    for i, m in enumerate(lens):
        T = plan.num_frames(m)
        a, b = got[i, :, :T], ref[i, :, :T]
        assert torch.isfinite(a).all()
        assert float((a - b).abs().max()) <= 2e-6 * float(b.max())
    o = np.abs(ocqt.cqt(wav[3, :lens[3]], sr=sr, hop_length=hop, fmin=fmin, n_bins=36, bins_per_octave=24, filter_scale=2))
    check_mag(got[3, :, :o.shape[1]].cpu().numpy(), o, tol=5e-6)


def test_cqt_streamed_bank_kernel_small_and_odd_batches(saga):
    """Edges of the streamed tensor-core contraction: one clip, 129 clips of 3 frames (rows of 43 clips per tile, last
    tile partly empty), every frame_count 1..8, windows entirely before / after the clip, against the fp32 twin."""
    ops, _ = saga
    sr, hop = 44100, 1024
    plan = ops.CqtPlan(sr, hop, osp.note_to_hz("A0"), 174, 24, filter_scale=2)
    n_min = 2 * hop + 1
    while True:
        try:
            plan.check_length(n_min)
            break
        except Exception:
            n_min += hop
    for n_clips, n in ((1, n_min), (1, 50000), (129, n_min + 2 * hop)):
        wav = np.stack([piano_clip(200 + (i % 7), n, sr=sr) * (1.0 + 0.01 * i) for i in range(n_clips)]).astype(np.float32)
        x = dev(wav)
        T = plan.num_frames(n)
        first = ((np.arange(n_clips) * 5) % (T + 6) - 3).astype(np.int32)
        res = {}
        for tag, opt in (("tc", None), ("fp32", "0")):
            with ops.options(SAGA_CQT_STREAM=opt):
                res[tag] = (ops.cqt_batch(x, plan, impl=1 if opt else 0)["mag"].clone(),
                            [ops.cqt_frames_batch(x, plan, first, fc).clone() for fc in range(1, 9)])
        peak = float(res["fp32"][0].max())
        assert float((res["tc"][0] - res["fp32"][0]).abs().max()) <= 1e-5 * peak
        for a, b in zip(res["tc"][1], res["fp32"][1]):
            assert a.shape == b.shape and torch.isfinite(a).all()
            assert float((a - b).abs().max()) <= 1e-5 * peak
            assert bool((a[b == 0] == 0).all())


def test_cqt_ragged_batch(saga):
    ops, _ = saga
    sr, hop = 44100, 512
    fmin = osp.note_to_hz("C1")
    plan = ops.CqtPlan(sr, hop, fmin, 84, 12, filter_scale=2)
    lens = [30000, 44100, 12345, 700]
    wav = np.zeros((len(lens), max(lens)), dtype=np.float32)
    for i, n in enumerate(lens):
        wav[i, :n] = piano_clip(40 + i, n)
    r = ops.cqt_batch(dev(wav), plan, lens=lens, fill=float("nan"))
    for i, n in enumerate(lens):
        ref = np.abs(ocqt.cqt(wav[i, :n], sr=sr, hop_length=hop, fmin=fmin, n_bins=84, filter_scale=2))
        got = r["mag"][i].cpu().numpy()
        assert plan.num_frames(n) == ref.shape[1]
        check_mag(got[:, :ref.shape[1]], ref)


@pytest.mark.parametrize("n", [66150, 512 * 300 + 7])
def test_cqt_tensor_path_equals_fp32_path_on_a_batch(saga, n):
    """Many tiles per persistent CTA (tile walk, stage ring wrap-around, octave interleaving, tail kernel):
    the tcgen05 path against the CUDA-core path on 40 clips, outputs poisoned with NaN beforehand."""
    ops, _ = saga
    plan = ops.CqtPlan(44100, 512, osp.note_to_hz("C1"), 84, 12, filter_scale=2)
    wav = dev(np.stack([piano_clip(500 + i, n) for i in range(40)]))
    ref = ops.cqt_batch(wav, plan, impl=1, fill=float("nan"))["mag"]
    got = ops.cqt_batch(wav, plan, impl=2, fill=float("nan"))["mag"]
    assert not torch.isnan(got).any() and not torch.isnan(ref).any()
    err = float((got - ref).abs().max() / ref.max())
    assert err <= 1e-5, err


@pytest.mark.parametrize("sr,hop,low,n_bins,bpo", [(44100, 512, "C1", 84, 12), (16000, 512, "C1", 84, 12),
                                                  (44100, 1024, "A0", 87, 12), (44100, 256, "C2", 60, 12)])
def test_cqt_fused_cascade_is_bit_identical_to_level_by_level(saga, sr, hop, low, n_bins, bpo):
    """decimate2x2_kernel (two cascade levels per launch, the intermediate level kept in shared memory) must
    leave exactly the numbers of the level-by-level cascade: ragged lengths (odd, shorter than one tile,
    longer than several), early factor 1 and 2, odd and even numbers of levels."""
    import os
    ops, _ = saga
    plan = ops.CqtPlan(sr, hop, osp.note_to_hz(low), n_bins, bpo, filter_scale=2)
    lens = [264600, 264599, 7937, 7938, 7939, 3968, 1985, 701, 133001, 100003]
    wav = np.zeros((len(lens), max(lens)), dtype=np.float32)
    for i, n in enumerate(lens):
        wav[i, :n] = piano_clip(60 + i, n, sr=sr)
    out = {}
    for mode in ("fused", "default", "levels"):
        # "fused": every pair, also the short deep levels the default leaves alone
        with ops.options(SAGA_DEC_NO_FUSE="1" if mode == "levels" else None,
                         SAGA_DEC_FUSE_MASK="0xffff" if mode == "fused" else None):
            res = []
            for kw in (dict(lens=lens), dict()):                 # ragged and equal-length (clip_lens NULL) batches
                for impl in (1, 0):
                    r = ops.cqt_batch(dev(wav), plan, want_complex=True, impl=impl, fill=float("nan"), **kw)
                    res.append((r["mag"].cpu().numpy(), r["C"].cpu().numpy()))
            out[mode] = res
    for mode in ("fused", "default"):
        for a, b in zip(out[mode], out["levels"]):
            for x, y in zip(a, b):
                assert np.array_equal(x, y, equal_nan=True)
    ref = np.abs(ocqt.cqt(wav[2, :lens[2]], sr=sr, hop_length=hop, fmin=osp.note_to_hz(low), n_bins=n_bins,
                          bins_per_octave=bpo, filter_scale=2))
    check_mag(out["fused"][0][0][2][:, :ref.shape[1]], ref)


def test_cqt_tensor_path_without_shared_bank(saga):
    """Plans whose octave banks are not multiples of one another keep octave-major tiles and swap the
    resident bank with a bulk copy per octave; force that path (the plan reads the switch at creation)."""
    import subprocess, sys, os
    code = r"""
import sys, numpy as np, torch
sys.path.insert(0, %r)
import amt_saga_b200
from amt_saga_b200 import ops
from amt_saga_b200.util_audio import note_to_hz
from tests.synth import piano_clip
plan = ops.CqtPlan(44100, 512, note_to_hz("C1"), 84, 12, filter_scale=2)
wav = torch.as_tensor(np.stack([piano_clip(700 + i, 512 * 140 + 3) for i in range(12)]), device="cuda")
ref = ops.cqt_batch(wav, plan, impl=1, fill=float("nan"))["mag"]
got = ops.cqt_batch(wav, plan, impl=2, fill=float("nan"))["mag"]
assert not torch.isnan(got).any()
print("ERR", float((got - ref).abs().max() / ref.max()))
""" % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, SAGA_UMMA_NO_SHARED_BANK="1")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    err = float(r.stdout.strip().split("ERR")[-1])
    assert err <= 1e-5, err


# --------------------------------------------------------------------------- class flow
def test_audio_complete_loop_matches_oracle(saga):
    """The producer loop's call sequence (training.py:265-449) on both containers."""
    _, ua = saga
    sr, N = 44100, 4096
    song = piano_clip(77, sr * 10)
    note = piano_clip(78, int(sr * 1.2), n_notes=1)
    a, o = ua.audio_complete(song, N), AudioOracle(song, N)
    assert a.shape == o.shape
    check_mag(a.mag.cpu().numpy(), o.mag)
    assert abs(a.spectral_flatness() - o.spectral_flatness()) < 1e-4
    dur = o._frames_to_seconds(o.shape[1])
    assert a._frames_to_seconds(a.shape[1]) == dur
    ca = a.slice_C(0, dur, a.shape[1], 8, bins_per_tone=1)
    co = o.slice_C(0, dur, o.shape[1], 8, bins_per_tone=1)
    check_mag(ca.cpu().numpy(), co)
    aw, ow = a.section(0, None, 258), o.section(0, None, 258)
    # slide half a window (training.py:318-323)
    an, on = a.section(6, None, 129), o.section(6, None, 129)
    aw.slice(129, 258); ow.slice(129, 258)
    aw.concat(an); ow.concat(on)
    assert aw.shape == ow.shape == (2049, 258)
    check_mag(aw.mag.cpu().numpy(), ow.mag)
    assert aw.wf.shape[0] == ow.wf.shape[0]
    onset, d = 1.3, 0.8
    for _ in range(2):
        assert aw._seconds_to_frames(onset) == ow._seconds_to_frames(onset)
        sa = aw.resize(onset, d, 8, attribs=["mag", "ph"])
        so = ow.resize(onset, d, 8, attribs=["mag", "ph"])
        check_mag(sa.mag.cpu().numpy(), so.mag)
        ct = ua.audio_complete._resize(ua.audio_complete.compress_bands(aw.mag, bands=20), 258)
        cto = AudioOracle._resize(AudioOracle.compress_bands(ow.mag, bands=20), 258)
        check_mag(ct.cpu().numpy(), cto, tol=1e-5)
        cp = aw.slice_C(onset, d, 8, bins_per_tone=2)
        cpo = ow.slice_C(onset, d, 8, bins_per_tone=2)
        assert tuple(cp.shape) == cpo.shape == (174, 8)
        check_mag(cp.cpu().numpy(), cpo)
        b0 = aw.midi_tone_to_FFT(60)
        assert b0 == ow.midi_tone_to_FFT(60)
        check_mag(sa.section_power("mag", b0, b0 + 348).cpu().numpy(), so.section_power("mag", b0, b0 + 348))
        aw.subtract(ua.audio_complete(note, N), offset=onset)
        ow.subtract(AudioOracle(note, N), offset=onset)
        check_mag(aw.mag.cpu().numpy(), ow.mag, tol=2e-5)
        # after the subtraction wf is rebuilt by an iSTFT of mag*ph: its length changes
        assert aw.wf.shape[0] == ow.wf.shape[0] == 1024 * 257
        check_mag(aw.wf.cpu().numpy()[None], ow.wf[None], tol=1e-4)
        onset, d = 3.7, 1.1


def test_numpy_carrier_and_errors(saga):
    _, ua = saga
    y = piano_clip(3, 20000)
    a = ua.audio_complete(y, 2048, hop_length=512, carrier="numpy")
    assert isinstance(a.mag, np.ndarray) and a.mag.shape == (1025, 40)
    with pytest.raises(ValueError):
        a._P("nope")
    with pytest.raises(ValueError):
        a.resize(0, 0.1, 8, attribs=["zzz"])
    with pytest.raises(ValueError):       # librosa ParameterError is a ValueError
        ua.audio_complete(y, 2048, hop_length=100, carrier="numpy").slice_C(0, 0.2, 8)


# --------------------------------------------------------------------------- pipeline (bench unit)
def test_window_pipeline_matches_oracle(saga):
    """The bench's step on 3 full-size 6 s windows vs the oracle, incl. the
    song-level ref handed to the first subtraction and K1's frame maxima reuse."""
    from amt_saga_b200.pipeline import WindowFeaturePipeline
    sr, n_fft, hop, ns, ng = 44100, 2048, 512, 264600, 65024
    W = 3
    pipe = WindowFeaturePipeline(W, ns, ng, sr, n_fft, hop)
    wav = np.stack([piano_clip(500 + i, ns) for i in range(W)])
    gue = np.stack([piano_clip(600 + i, ng, n_notes=1) for i in range(W)])
    offs = np.array([[0], [200], [500]], dtype=np.int32)
    pipe.run(dev(wav), dev(gue), dev(offs))
    torch.cuda.synchronize()
    assert (pipe.T, pipe.T_clip, pipe.Tg, pipe.Tc) == (516, 517, 128, 517)
    for w in range(W):
        full = np.abs(osp.stft(wav[w], n_fft, hop))
        mag = full[:, :516].copy()
        g = np.abs(osp.stft(gue[w], n_fft, hop))
        g *= mag.max() / g.max()          # the window's own maximum (training.py:284 precedes the song's ref_mag, :336)
        g = g[:, :516 - offs[w, 0]]
        mag[:, offs[w, 0]:offs[w, 0] + g.shape[1]] -= g
        np.maximum(mag, 0, mag)
        check_mag(pipe.mag[w, :516, :1025].T.cpu().numpy(), mag, tol=2e-5)
        check_db(pipe.D[w, :516, :1025].T.cpu().numpy(), osp.amplitude_to_db(mag, ref=mag.max()))
        assert abs(float(pipe.ref[w]) - mag.max()) <= 2e-5 * mag.max()
        C = np.abs(ocqt.cqt(wav[w], sr=sr, hop_length=hop, fmin=osp.note_to_hz("C1"), n_bins=84, filter_scale=2))
        check_mag(pipe.C[w, :, :84].T.cpu().numpy(), C)


# --------------------------------------------------------------------------- committed golden vectors
GOLD = __import__("os").path.join(__import__("os").path.dirname(__file__), "golden")


def test_golden_cfg1(saga, cfg1):
    _, ua = saga
    y, sr = cfg1
    g = np.load(GOLD + "/cfg1.npz")
    a = ua.audio_complete(y, 2048, hop_length=512, sample_rate=sr)
    assert a.shape == tuple(g["mag_shape"]) and a.mag.stride() == (1, 1028)   # [bins, frames], frame-major
    mag = a.mag.cpu().numpy()
    assert np.abs(mag[:, g["cols"]] - g["mag_cols"]).max() <= MAG_TOL * float(g["ref_mag"])
    assert np.abs(mag.sum(axis=0) - g["mag_colsum"]).max() <= 1e-5 * g["mag_colsum"].max()
    assert abs(float(a.ref_mag) - float(g["ref_mag"])) <= 1e-6 * float(g["ref_mag"])
    D = a.D.cpu().numpy()[:, g["cols"]]
    above = g["D_cols"] > -79.9
    assert np.abs(D[above] - g["D_cols"][above]).max() <= DB_TOL
    C = a.slice_C(0, 10.0, 313, bins_per_tone=1, lowest_note="C1", nbins=84).cpu().numpy()
    assert C.shape == tuple(g["cqt_shape"])
    assert np.abs(C[:, g["cols"]] - g["cqt_cols"]).max() <= MAG_TOL * g["cqt_cols"].max()
    assert np.abs(C.sum(axis=1) - g["cqt_rowsum"]).max() <= 1e-4 * g["cqt_rowsum"].max()


def test_golden_subtract_istft_cqt87(saga):
    ops, _ = saga
    s = np.load(GOLD + "/subtract_chain.npz")
    B, T = s["win"].shape
    P = ops.frame_pitch(B)
    st = torch.zeros((1, T, P), device="cuda")
    st[0, :, :B] = dev(s["win"]).T
    g = torch.zeros((1, 3, 9, P), device="cuda")
    g[0, :, :, :B] = dev(s["guesses"]).transpose(1, 2)
    D, ref = ops.subtract_db_batch(st, g, dev(s["offsets"]).reshape(1, 3), B)
    assert np.array_equal(st[0, :, :B].T.cpu().numpy(), s["result"])          # bit exact
    check_db(D[0, :, :B].T.cpu().numpy(), s["D"])
    i = np.load(GOLD + "/istft.npz")
    plan = ops.StftPlan(1024, 256, True)
    r = ops.stft_batch(dev(i["wav"]), plan, want_complex=True)
    w = ops.istft_batch(plan, F=r["F_storage"])[0].cpu().numpy()
    assert w.shape == i["istft"].shape and np.abs(w - i["istft"]).max() <= 2e-5 * np.abs(i["istft"]).max()
    c = np.load(GOLD + "/cqt87.npz")
    cp = ops.CqtPlan(44100, 1024, osp.note_to_hz("A0"), 87, 12, filter_scale=2)
    got = ops.cqt_batch(dev(c["wav"]), cp)["mag"][0].cpu().numpy()
    check_mag(got, c["cqt"], tol=1e-5)


# --------------------------------------------------------------------------- full-size properties
def test_full_size_properties_one_hour(saga):
    """BASELINE cfg2/cfg3 size (1 h = 600 x 6 s windows): properties that do not
    need the oracle at that size."""
    ops, _ = saga
    from amt_saga_b200 import synth
    from amt_saga_b200.pipeline import WindowFeaturePipeline
    W, ns, ng = 600, 264600, 65024
    wav = synth.piano_batch(range(W), ns, seed_base=50000)
    guess = synth.piano_batch(range(W), ng, n_notes=1, seed_base=90000)
    plan = ops.get_stft_plan(2048, 512, True)
    r = ops.stft_batch(wav, plan, want_complex=True)
    assert tuple(r["mag"].shape) == (W, 1025, 517)
    # (1) |F| == mag, per-frame / per-clip maxima consistent (a checksum of checksums)
    assert float((r["F"].abs() - r["mag"]).abs().max()) <= 1e-5 * float(r["mag"].max())
    assert torch.equal(r["frame_max"].amax(dim=1), r["clip_max"])
    assert torch.allclose(r["mag"].amax(dim=1), r["frame_max"], rtol=0, atol=0)
    # (2) linearity: STFT(2.5 x) == 2.5 STFT(x)
    r2 = ops.stft_batch(wav * 2.5, plan)
    assert float((r2["mag"] - 2.5 * r["mag"]).abs().max()) <= 2e-6 * float(r2["mag"].max())
    # (3) analysis -> synthesis round trip reproduces the interior samples
    back = ops.istft_batch(plan, F=r["F_storage"])
    assert back.shape == (W, 512 * 516)
    assert float((back[:, 2048:-2048] - wav[:, 2048:512 * 516 - 2048]).abs().max()) <= 2e-5
    del r2, back
    # (4) Parseval on the CQT side is not available (non-unitary bank); use shift invariance:
    #     interior CQT frames of a clip delayed by 8 hops equal the original's frames 8 later.
    cq = ops.get_cqt_plan(44100, 512, osp.note_to_hz("C1"), 84, 12, 2)
    c0 = ops.cqt_batch(wav[:8], cq)["mag"]
    c1 = ops.cqt_batch(torch.roll(wav[:8], 8 * 512, dims=1), cq)["mag"]
    assert float((c1[:, :, 120:400] - c0[:, :, 112:392]).abs().max()) <= 2e-5 * float(c0.max())
    # (5) a guess placed at the window's end is clipped away (util_audio.py:250-251): the pipeline is
    #     then the identity on the window and D = dB(mag)
    pipe = WindowFeaturePipeline(W, ns, ng)
    offs = torch.zeros((W, 1), device="cuda", dtype=torch.int32)
    pipe.run(wav, guess, offs + 516)
    torch.cuda.synchronize()
    assert torch.equal(pipe.mag[:, :516, :1025], r["mag_storage"][:, :516, :1025])
    assert float(pipe.D.max()) <= 0.0 and float(pipe.D[:, :516, :1025].min()) >= -80.0
    assert torch.equal(pipe.ref, pipe.mag[:, :516, :1025].amax(dim=(1, 2)))
    # (6) a real guess only lowers the window, never below zero, and only inside its column range
    pipe.run(wav, guess, offs + 100)
    torch.cuda.synchronize()
    assert float(pipe.mag.min()) >= 0.0
    d = r["mag_storage"][:, :516, :1025] - pipe.mag[:, :516, :1025]
    assert float(d.min()) >= 0.0 and float(d[:, :100].abs().max()) == 0.0 and float(d[:, 228:].abs().max()) == 0.0
    assert float(d[:, 100:228].max()) > 0.0


def test_pipeline_schedules_and_phase_flags_give_identical_results(saga):
    """The step's kernels can be enqueued on one stream ("serial", default), with the CQT chain forked onto a second
    stream ("chains"), or with the contraction held back beside the dB pass ("db_with_contraction", which uses the
    SAGA_SUB_SKIP_DB / SAGA_SUB_ONLY_DB phase flags of K3): all bit-identical, also when run chunk by chunk."""
    from amt_saga_b200.pipeline import WindowFeaturePipeline
    W, ns, ng = 9, 44100, 16384
    pipe = WindowFeaturePipeline(W, ns, ng)
    wav = dev(np.stack([piano_clip(300 + i, ns, n_notes=5) for i in range(W)]))
    gue = dev(np.stack([piano_clip(400 + i, ng, n_notes=1) for i in range(W)]))
    offs = dev((np.arange(W, dtype=np.int32) * 11 % 80).reshape(-1, 1))
    ref = None
    for sched in ("serial", "chains", "db_with_contraction", "serial-chunks"):
        pipe.schedule = sched.split("-")[0]
        for t in (pipe.mag, pipe.D, pipe.C, pipe.ref):
            t.fill_(float("nan"))
        if sched.endswith("chunks"):
            pipe.run(wav, gue, offs, parts=("cqt",))
            for a in range(0, W, 4):
                pipe.run(wav, gue, offs, w0=a, w1=min(W, a + 4), parts=("stft",))
        else:
            pipe.run(wav, gue, offs)
        torch.cuda.synchronize()
        out = [pipe.mag.clone(), pipe.D[:, :pipe.T].clone(), pipe.C.clone(), pipe.ref.clone()]
        assert not any(torch.isnan(t).any() for t in out)
        if ref is None:
            ref = out
        for x, y in zip(out, ref):
            assert torch.equal(x, y), sched


def test_run_host_chunked_equals_resident_run(saga):
    """The overlapped host path (chunks on three streams) returns exactly what the
    resident single-shot pass computes."""
    from amt_saga_b200.pipeline import WindowFeaturePipeline
    W, ns, ng = 7, 44100, 16384
    pipe = WindowFeaturePipeline(W, ns, ng)
    wav = np.stack([piano_clip(700 + i, ns, n_notes=5) for i in range(W)])
    gue = np.stack([piano_clip(800 + i, ng, n_notes=1) for i in range(W)])
    offs = (np.arange(W, dtype=np.int32) * 9 % 70).reshape(-1, 1)
    pipe.run(dev(wav), dev(gue), dev(offs))
    torch.cuda.synchronize()
    C0, ref0, mag0 = pipe.C.clone(), pipe.ref.clone(), pipe.mag.clone()
    h = pipe.host_buffers()
    h["wav"].copy_(torch.from_numpy(wav)); h["guess"].copy_(torch.from_numpy(gue)); h["offs"].copy_(torch.from_numpy(offs))
    for chunks in (1, 3, 7):
        h["C"].zero_(); h["ref"].zero_()
        pipe.run_host(chunks)
        torch.cuda.synchronize()
        assert torch.equal(h["C"], C0.cpu()) and torch.equal(h["ref"], ref0.cpu())
        assert torch.equal(pipe.mag, mag0)


# --------------------------------------------------------------------------- K5 feature gather
def test_feature_gather_matches_oracle_sequence(saga):
    """compress_bands, the fused short-window block and spectral flatness vs the
    oracle doing the reference's call sequence (training.py:333-363)."""
    ops, ua = saga
    sr, N = 44100, 4096
    song = piano_clip(91, sr * 7)
    a, o = ua.audio_complete(song, N), AudioOracle(song, N)
    a.mag, o.mag
    aw, ow = a.section(0, None, 258), o.section(0, None, 258)
    # C_timing (training.py:333-336)
    ct = ua.audio_complete._resize(ua.audio_complete.compress_bands(aw.mag, bands=20), 258) / float(a.ref_mag)
    cto = AudioOracle._resize(AudioOracle.compress_bands(ow.mag, bands=20), 258) / o.ref_mag
    check_mag(ct.cpu().numpy(), cto, tol=1e-5)
    assert abs(a.spectral_flatness() - o.spectral_flatness()) <= 1e-5
    for onset, dur, pitch in ((0.7, 0.9, 60), (2.0, 0.02, 40), (4.1, 3.0, 88), (5.9, 0.5, 100)):
        b0 = aw.midi_tone_to_FFT(pitch)
        f = aw.short_window_features(onset, dur, 8, b0, 348, ref=a.ref_mag)
        so = ow.resize(onset, dur, 8, attribs=["mag", "ph"])
        lin = so.section_power("mag", b0, b0 + 348)
        log = np.log10(lin * 1000 + 1)
        log /= np.max(log)
        pha = (np.angle(so.section_power("ph", b0, b0 + 348)) + 3.15) / 6.3
        assert tuple(f["lin"].shape) == lin.shape == (348, 8)
        check_mag(f["lin"].cpu().numpy(), lin / o.ref_mag, tol=2e-5)
        assert np.abs(f["log"].cpu().numpy() - log).max() <= 2e-4
        big = lin > 1e-3 * lin.max()            # phase is noise where the magnitude is
        assert np.abs(f["phase"].cpu().numpy() - pha)[big].max() <= 2e-3
    for t in (0, 1, 2, 3, 5, 8, 11, 300):
        for target in (8, 258):
            idx = ops.resize_indices(t, target)
            P = np.arange(t, dtype=float)[None, :]
            exp = AudioOracle._resize(P, target)[0]
            got = np.where(idx >= 0, P[0][np.clip(idx, 0, max(t - 1, 0))] if t else 0.0, 0.0)
            assert np.array_equal(got, exp)


def test_run_host_feature_return_matches_the_full_images(saga):
    """run_host(returns="features"): the reduced outputs the e2e path ships (K5 on the device) are exactly the
    columns / band means of the full dB, CQT and magnitude images the resident pass leaves on the device."""
    from amt_saga_b200.pipeline import WindowFeaturePipeline
    from amt_saga_b200.util_audio import band_edges
    W, ns, ng = 5, 44100, 16384
    pipe = WindowFeaturePipeline(W, ns, ng)
    wav = np.stack([piano_clip(710 + i, ns, n_notes=5) for i in range(W)])
    gue = np.stack([piano_clip(810 + i, ng, n_notes=1) for i in range(W)])
    offs = np.array([[0], [3], [40], [80], [85]], dtype=np.int32)
    h = pipe.host_buffers()
    h["wav"].copy_(torch.as_tensor(wav)); h["guess"].copy_(torch.as_tensor(gue)); h["offs"].copy_(torch.as_tensor(offs))
    pipe.run_host(2, returns="features")
    torch.cuda.synchronize()
    D, C, mag = pipe.D.cpu().numpy(), pipe.C.cpu().numpy(), pipe.mag.cpu().numpy()
    edges = band_edges(pipe.nb, 20)
    for w in range(W):
        for j in range(8):
            t = offs[w, 0] + j
            want_d = D[w, t, :pipe.nb] if t < pipe.T else np.zeros(pipe.nb, np.float32)
            want_c = C[w, t, :84] if t < pipe.Tc else np.zeros(84, np.float32)
            assert np.array_equal(h["D8"][w, j, :pipe.nb].numpy(), want_d)
            assert np.array_equal(h["C8"][w, j, :84].numpy(), want_c)
        inv = np.float32(1.0) / pipe.clip_max[w].cpu().numpy()
        for b in range(20):
            want = mag[w, :pipe.T, edges[b]:edges[b + 1]].mean(axis=1) * inv
            assert np.allclose(h["timing"][w, :, b].numpy(), want, rtol=2e-6, atol=1e-9)
        assert h["ref"][w] == pipe.ref[w].cpu()

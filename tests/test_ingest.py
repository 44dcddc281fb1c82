"""K0 PCM ingest (saga_pcm16_absmax_exec / saga_pcm16_ingest_exec): the reference's one scaling of its
16-bit PCM, util_audio.py:776-781 / :894 / :964.  Integer / IEEE work: the bar is bit-exact."""
import numpy as np
import pytest
import torch

from oracle import ingest as oin
from oracle import spectral as osp

from tests.synth import piano_clip


# ----------------------------------------------------------------------------- oracle, no GPU
def test_oracle_render_scale_follows_the_reference_lines():
    pcm = np.array([0, 100, -3000, 2999, 7], dtype=np.int16)
    # several notes: vel_max as is; a single note: max(1, vel - 12)   (util_audio.py:778-780)
    mul, div = oin.render_scale(pcm, [40, 90, 64])
    assert mul == (90 / 128.0) ** 4 and div == 3000.0
    mul1, _ = oin.render_scale(pcm, [90])
    assert mul1 == (78 / 128.0) ** 4
    assert oin.render_scale(pcm, [5])[0] == (1 / 128.0) ** 4
    w = oin.render(pcm, [90])
    # operation order of `wf*(vel_max/128.0)**4/np.abs(wf).max()`: multiply, then divide, in float64
    assert w.dtype == np.float64 and np.array_equal(w, (pcm.astype(np.float64) * mul1) / 3000.0)
    assert abs(w).max() == mul1
    # soundfile convention: exact in float32
    f = oin.pcm_to_wave(pcm)
    assert np.array_equal(f.astype(np.float32).astype(np.float64), f)
    st = np.arange(12, dtype=np.int16)
    assert np.array_equal(oin.left_channel(st), st[::2])


def test_oracle_per_clip_scales_broadcast():
    rng = np.random.default_rng(0)
    pcm = rng.integers(-32768, 32768, size=(3, 50)).astype(np.int16)
    mul = np.array([1.0, 0.3, 0.07])
    div = np.abs(pcm.astype(np.int32)).max(axis=1).astype(np.float64)
    w = oin.pcm_to_wave(pcm, mul, div)
    for c in range(3):
        assert np.array_equal(w[c], (pcm[c].astype(np.float64) * mul[c]) / div[c])


# ----------------------------------------------------------------------------- CUDA path
@pytest.fixture(scope="module")
def saga():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import amt_saga_b200  # noqa: F401
    from amt_saga_b200 import ops, util_audio
    return ops, util_audio


def _pcm(seed, shape):
    rng = np.random.default_rng(seed)
    x = rng.integers(-32768, 32768, size=shape).astype(np.int16)
    x.flat[:4] = [-32768, 32767, 0, -1]
    return x


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 7, 8, 2047, 2048, 2049, 264600, 65024 + 3])
def test_ingest_bit_exact_generic_scales(saga, n):
    ops, _ = saga
    pcm = _pcm(n, (5, n))
    pcm[3] = 0                                 # silent render: 0 * mul / 0 = nan, as numpy gives the reference
    mul = np.array([(v / 128.0) ** 4 for v in (90, 31, 120, 64, 1)])
    d = torch.as_tensor(pcm, device="cuda")
    peak = ops.pcm16_absmax(d)
    ref_peak = np.abs(pcm.astype(np.int32)).max(axis=1)
    assert np.array_equal(peak.cpu().numpy(), ref_peak)
    got = ops.pcm16_to_wave(d, mul=torch.as_tensor(mul, device="cuda"), div=peak).cpu().numpy()
    want = oin.pcm_to_wave(pcm, mul, ref_peak.astype(np.float64)).astype(np.float32)
    assert got.dtype == np.float32 and got.shape == want.shape
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))          # bit for bit, nan rows included
    assert np.isnan(got[3]).all()


@pytest.mark.gpu
def test_ingest_file_convention_and_scalar_scales(saga):
    ops, _ = saga
    pcm = _pcm(3, (2, 4099))
    d = torch.as_tensor(pcm, device="cuda")
    got = ops.pcm16_to_wave(d).cpu().numpy()                                   # / 32768 (soundfile)
    assert np.array_equal(got, (pcm.astype(np.float64) / 32768.0).astype(np.float32))
    got = ops.pcm16_to_wave(d, mul=0.2373046875 ** 4, div=29999.0).cpu().numpy()
    want = oin.pcm_to_wave(pcm, 0.2373046875 ** 4, 29999.0).astype(np.float32)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    # per-clip float64 divisors
    div = np.array([3.0, 32768.0])
    got = ops.pcm16_to_wave(d, div=torch.as_tensor(div, device="cuda")).cpu().numpy()
    want = oin.pcm_to_wave(pcm, 1.0, div).astype(np.float32)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


@pytest.mark.gpu
def test_ingest_left_channel_of_interleaved_stereo_and_unaligned_rows(saga):
    ops, _ = saga
    n = 3001
    st = _pcm(9, (3, 2 * n + 1))[:, 1:]                      # odd row pitch: rows are not 16-byte aligned
    d = torch.as_tensor(np.ascontiguousarray(_pcm(9, (3, 2 * n + 1))), device="cuda")[:, 1:]
    peak = ops.pcm16_absmax(d, channel_stride=2)
    left = oin.left_channel(st)
    assert left.shape == (3, n)
    assert np.array_equal(peak.cpu().numpy(), np.abs(left.astype(np.int32)).max(axis=1))
    got = ops.pcm16_to_wave(d, mul=0.5, div=peak, channel_stride=2).cpu().numpy()
    want = oin.pcm_to_wave(left, 0.5, np.abs(left.astype(np.int32)).max(axis=1).astype(np.float64)).astype(np.float32)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


@pytest.mark.gpu
def test_ingest_argument_errors(saga):
    ops, _ = saga
    with pytest.raises(TypeError):
        ops.pcm16_to_wave(torch.zeros((1, 8), device="cuda"))                  # not int16
    with pytest.raises(TypeError):
        ops.pcm16_to_wave(torch.zeros((1, 8), dtype=torch.int16))              # not on the device
    with pytest.raises(ValueError):
        ops.pcm16_to_wave(torch.zeros((2, 8), dtype=torch.int16, device="cuda"), div=0.0)
    with pytest.raises(ValueError):
        ops.pcm16_to_wave(torch.zeros((2, 8), dtype=torch.int16, device="cuda"),
                          mul=torch.ones(3, dtype=torch.float64, device="cuda"))
    assert ops.pcm16_to_wave(torch.zeros((0, 8), dtype=torch.int16, device="cuda")).shape == (0, 8)


@pytest.mark.gpu
def test_audio_complete_from_pcm16_equals_float_construction(saga):
    """A rendered note handed over as PCM gives the same container, bit for bit, as the reference's float64
    waveform handed to the ordinary constructor (util_audio.py:776-781 then :33)."""
    _, ua = saga
    y = piano_clip(5, 30000, n_notes=1)
    pcm = np.round(y / np.abs(y).max() * 20000).astype(np.int16)
    wf = oin.render(pcm, [77])
    a = ua.audio_complete.from_pcm16(pcm, 2048, mul=oin.render_scale(pcm, [77])[0], div="peak")
    b = ua.audio_complete(wf, 2048)
    assert torch.equal(a.wf, b.wf) and torch.equal(a.mag, b.mag)
    ref = np.abs(osp.stft(wf, 2048, 512))
    assert np.abs(a.mag.cpu().numpy() - ref).max() <= 1e-4 * ref.max()


@pytest.mark.gpu
def test_run_host_pcm16_equals_float_path(saga):
    """The PCM-ingest variant of the host path returns exactly what the float32 host path returns for the
    waveform the oracle builds from the same PCM."""
    from amt_saga_b200.pipeline import WindowFeaturePipeline
    W, ns, ng = 5, 44100, 16384
    pipe = WindowFeaturePipeline(W, ns, ng)
    rng = np.random.default_rng(1)
    wav = np.stack([piano_clip(900 + i, ns, n_notes=5) for i in range(W)])
    gue = np.stack([piano_clip(950 + i, ng, n_notes=1) for i in range(W)])
    wav_pcm = np.round(wav / np.abs(wav).max(axis=1, keepdims=True) * 30000).astype(np.int16)
    gue_pcm = np.round(gue / np.abs(gue).max(axis=1, keepdims=True) * 12000).astype(np.int16)
    mul = (rng.integers(30, 121, size=(2, W)) / 128.0) ** 4
    offs = (np.arange(W, dtype=np.int32) * 13 % 60).reshape(-1, 1)
    wf = oin.pcm_to_wave(wav_pcm, mul[0], np.abs(wav_pcm.astype(np.int32)).max(axis=1)).astype(np.float32)
    gf = oin.pcm_to_wave(gue_pcm, mul[1], np.abs(gue_pcm.astype(np.int32)).max(axis=1)).astype(np.float32)
    h = pipe.host_pcm_buffers()
    h["wav"].copy_(torch.from_numpy(wf)); h["guess"].copy_(torch.from_numpy(gf)); h["offs"].copy_(torch.from_numpy(offs))
    pipe.run_host(2)
    torch.cuda.synchronize()
    C0, ref0, mag0, D0 = h["C"].clone(), h["ref"].clone(), pipe.mag.clone(), pipe.D.clone()
    h["wav_pcm"].copy_(torch.from_numpy(wav_pcm)); h["guess_pcm"].copy_(torch.from_numpy(gue_pcm))
    h["mul"].copy_(torch.from_numpy(mul))
    h["d_wav"].zero_(); h["d_guess"].zero_()
    for chunks in (1, 3):
        h["C"].zero_(); h["ref"].zero_()
        pipe.run_host(chunks, pcm16=True)
        torch.cuda.synchronize()
        assert torch.equal(h["d_wav"].cpu(), torch.from_numpy(wf))
        assert torch.equal(h["C"], C0) and torch.equal(h["ref"], ref0)
        assert torch.equal(pipe.mag, mag0) and torch.equal(pipe.D[:, :pipe.T], D0[:, :pipe.T])   # row T of D is never written
    # windows cut from a longer render: the divisor is the SONG's peak (util_audio.py:781 runs before `section`), given
    # explicitly per window; the guesses keep their own peak
    song_peak = np.abs(wav_pcm.astype(np.int32)).max(axis=1) + np.arange(W) * 7 + 100
    wf2 = oin.pcm_to_wave(wav_pcm, mul[0], song_peak).astype(np.float32)
    h["wav"].copy_(torch.from_numpy(wf2))
    pipe.run_host(2)
    torch.cuda.synchronize()
    C1, ref1 = h["C"].clone(), h["ref"].clone()
    h["wav_div"].copy_(torch.from_numpy(song_peak.astype(np.float64)))
    h["C"].zero_(); h["ref"].zero_()
    pipe.run_host(3, pcm16=True)
    torch.cuda.synchronize()
    assert torch.equal(h["d_wav"].cpu(), torch.from_numpy(wf2))
    assert torch.equal(h["C"], C1) and torch.equal(h["ref"], ref1)
    assert not torch.equal(C1, C0)

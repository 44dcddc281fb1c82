"""Property tests (hypothesis) of the host-side logic against the oracle's restatement of the reference:
`_resize` index map (util_audio.py:384-409), the float64 time<->frame maps (:261-272), band edges (:451-456),
the shard partition, the PCM scaling.  No GPU."""
import numpy as np
from hypothesis import given, settings, strategies as st

import amt_saga_b200  # noqa: F401
from amt_saga_b200 import ops, shard, util_audio
from amt_saga_b200.pipeline import seconds_to_frames
from oracle import ingest as oin
from oracle.audio_oracle import AudioOracle, band_edges


@settings(max_examples=300, deadline=None)
@given(t=st.integers(0, 300), target=st.integers(3, 300), bands=st.integers(1, 5))
def test_resize_index_map_equals_reference_resize(t, target, bands):
    """ops.resize_indices (what the fused K5 gather consumes) reproduces util_audio._resize column for column."""
    P = np.arange(bands * max(t, 1), dtype=np.float64).reshape(bands, max(t, 1))[:, :t] + 1.0
    ref = AudioOracle._resize(P, target)
    idx = ops.resize_indices(t, target)
    assert ref.shape == (bands, target) and idx.shape == (target,)
    got = np.where(idx[None, :] >= 0, P[:, np.clip(idx, 0, t - 1)], 0.0) if t else np.zeros((bands, target))
    if t == 0:
        assert np.all(idx == -1)
    assert np.array_equal(got, ref)
    assert np.array_equal(util_audio.audio_complete._resize(P, target), ref)


@settings(max_examples=300, deadline=None)
@given(time=st.floats(0.0, 30.0, allow_nan=False), T=st.integers(1, 4000), hop=st.sampled_from([256, 512, 1024]),
       sr=st.sampled_from([16000, 22050, 44100]))
def test_seconds_to_frames_follows_the_reference_operation_order(time, T, hop, sr):
    n = hop * max(T - 1, 1)
    want = int(np.floor(time * T * sr / n))            # util_audio.py:264, float64, this operation order
    assert seconds_to_frames(time, T, sr, n) == want


@settings(max_examples=200, deadline=None)
@given(n_items=st.integers(0, 100000), world=st.integers(1, 16))
def test_shard_ranges_partition_the_items(n_items, world):
    r = [shard.shard_range(n_items, k, world) for k in range(world)]
    assert r[0][0] == 0 and r[-1][1] == n_items
    assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
    sizes = [b - a for a, b in r]
    assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)


@settings(max_examples=100, deadline=None)
@given(n_bins=st.integers(30, 5000), bands=st.integers(1, 60))
def test_band_edges_equal_the_oracle(n_bins, bands):
    e = util_audio.band_edges(n_bins, bands)
    assert np.array_equal(e, band_edges(n_bins, bands))
    assert e[0] == 0 and np.all(np.diff(e) >= 1)       # forced strictly increasing (util_audio.py:451-456)


@settings(max_examples=100, deadline=None)
@given(vel=st.lists(st.integers(1, 127), min_size=1, max_size=6), seed=st.integers(0, 2 ** 31 - 1))
def test_render_scale_is_the_reference_expression(vel, seed):
    pcm = np.random.default_rng(seed).integers(-32768, 32768, size=257).astype(np.int16)
    pcm[0] = 12345
    wf = pcm.astype(np.float64)
    vel_max = max(vel)
    if len(vel) == 1:
        vel_max = max(1, vel_max - 12)
    want = wf * (vel_max / 128.0) ** 4 / np.abs(wf).max()          # util_audio.py:778-781, literally
    assert np.array_equal(oin.render(pcm, vel), want)

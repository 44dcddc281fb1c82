"""Classifier-input assembly: host logic on CPU, device batches on the GPU."""
import numpy as np
import pytest

from oracle import batches as ob


def _sample(bt, seed):
    rng = np.random.default_rng(seed)
    a = lambda b, f: rng.standard_normal((b, f)).astype(np.float32)
    return bt.note_sample("f.mid", a(20, 258), a(174, 8), a(348, 8), a(348, 8), a(348, 8), a(348, 8), a(348, 8),
                          a(348, 8), a(348, 8), a(36, 8), a(348, 8), 60 + seed, seed % 5, 0.1 * seed, 0.1 * seed + 0.5,
                          80 + seed)


def test_check_shape_and_model_order():
    import amt_saga_b200  # noqa: F401
    from amt_saga_b200 import batches as bt
    s = _sample(bt, 1)
    bt.check_shape(s.C_timing, 20, 258)
    bt.check_shape([s.C_sw_inst, s.C_sw_inst], 348, 8)
    bt.check_shape([[s.C_sw_inst, s.F_sw_inst_foc]], 348, 8)
    with pytest.raises(ValueError, match=r"Invalid Input shape\. Expected: \(20, 8\) \. Got: \(20, 258\)"):
        bt.check_shape(s.C_timing, 20, 8)
    all_x, all_y = bt.model_inputs(s)
    assert len(all_x) == len(all_y) == 14                       # training.py:509-536
    assert all_x[0] is s.C_timing and all_x[1] is s.C_timing and all_x[13] is s.C_velocity
    assert all_x[8][0] is s.C_sw_inst and all_x[8][1] is s.F_sw_inst_foc and all_x[9][1] is s.ph
    assert all_y[0] == s.time_start and all_y[1] == s.time_end and all_y[2] == s.pitch
    assert all(y == s.instrument for y in all_y[3:13]) and all_y[13] == s.velocity


def test_oracle_list_to_nd_array_shapes():
    x, y = ob.list_to_nd_array([np.ones((3, 4)), np.zeros((3, 4))], [1, 2])
    assert x.shape == (2, 3, 4, 1) and y.shape == (2, 1) and x.dtype == np.float64
    # multi-input models: every channel buffer is created with the FIRST channel's shape (util_train_test.py:119),
    # so the reference only supports equal-shaped channels -- which is what it feeds ([348, 8] pairs)
    xs, y = ob.list_to_nd_array([[np.ones((3, 4)), 2 * np.ones((3, 4))]] * 3, [7, 8, 9])
    assert [t.shape for t in xs] == [(3, 3, 4, 1), (3, 3, 4, 1)] and y[:, 0].tolist() == [7, 8, 9]
    with pytest.raises(ValueError):
        ob.list_to_nd_array([[np.ones((3, 4)), np.ones((5, 4))]], [1])
    x, y = ob.list_to_nd_array(np.ones((3, 4)), [5])
    assert x.shape == (1, 3, 4, 1) and y.shape == (1, 1)


@pytest.mark.gpu
def test_device_batches_match_the_reference_layout():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import amt_saga_b200  # noqa: F401
    from amt_saga_b200 import batches as bt
    samples = [_sample(bt, i) for i in range(5)]
    models = [0, 2, 8, 9, 13]
    batcher = bt.SampleBatcher(models, batch_size=5)
    out = None
    for s in samples:
        dev_s = bt.note_sample(*[torch.as_tensor(v, device="cuda") if isinstance(v, np.ndarray) else v
                                 for v in (s.filename, s.C_timing, s.C_sw_pitch, s.C_sw_inst, s.F_sw_inst_foc,
                                           s.F_sw_inst_foc_log10, s.F_sw_inst_foc_const, s.F_sw_inst_foc_const_log10,
                                           s.C_sw_inst_foc, s.C_sw_inst_foc_const, s.C_velocity, s.ph, s.pitch,
                                           s.instrument, s.time_start, s.time_end, s.velocity)])
        out = batcher.add(dev_s)
    assert out is not None and len(out) == len(models) and len(batcher) == 0
    # the reference: per model, lists over samples -> list_to_nd_array
    per = [bt.model_inputs(s) for s in samples]
    for (x, y), mi in zip(out, models):
        rx, ry = ob.list_to_nd_array([p[0][mi] for p in per], [p[1][mi] for p in per])
        assert np.allclose(y.cpu().numpy(), ry)
        if isinstance(rx, list):
            assert isinstance(x, list) and len(x) == len(rx)
            for a, b in zip(x, rx):
                assert a.is_cuda and tuple(a.shape) == b.shape and np.array_equal(a.cpu().numpy(), b.astype(np.float32))
        else:
            assert x.is_cuda and tuple(x.shape) == rx.shape and np.array_equal(x.cpu().numpy(), rx.astype(np.float32))
    caps = bt.to_dlpack(out[0])
    back = torch.utils.dlpack.from_dlpack(caps[0])
    assert back.data_ptr() == out[0][0].data_ptr()          # zero copy

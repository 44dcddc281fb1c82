"""Classifier-input assembly: the hand-over between the feature hot path and
the reference's classifiers, kept on the device.

Mirrors, name for name and shape for shape,
  * `note_sample`        /root/reference/util_train_test.py:177-209   one note's 11 feature arrays + labels
  * `check_shape`        /root/reference/util_train_test.py:92-112    same ValueError text
  * `list_to_nd_array`   /root/reference/util_train_test.py:114-146   [bands, frames] -> [B, bands, frames, 1]
  * the per-model input selection of `thread_training` /root/reference/training.py:509-555
    (`all_x` / `all_y` tuples indexed by `models_to_train`, then transposed into one (x, y) per model).

The reference concatenates float64 numpy arrays one sample at a time (`np.concatenate` in a loop, then
pickles them through a multiprocessing.Queue).  Here a batch is ONE stacked CUDA tensor per model input;
nothing is copied to the host, and `to_dlpack()` hands the tensors to whatever framework hosts the
classifiers.  This module moves data only -- there is no arithmetic on the path.
"""
import numpy as np
import torch


class note_sample:
    """Same constructor and attributes as util_train_test.note_sample (util_train_test.py:177-209)."""

    def __init__(self, filename, C_timing, C_sw_pitch, C_sw_inst, F_sw_inst_foc, F_sw_inst_foc_log10,
                 F_sw_inst_foc_const, F_sw_inst_foc_const_log10, C_sw_inst_foc, C_sw_inst_foc_const, C_velocity,
                 ph, pitch, instrument, time_start, time_end, velocity):
        self.filename = filename
        self.C_timing = C_timing
        self.C_sw_pitch = C_sw_pitch
        self.C_sw_inst = C_sw_inst
        self.F_sw_inst_foc = F_sw_inst_foc
        self.F_sw_inst_foc_log10 = F_sw_inst_foc_log10
        self.F_sw_inst_foc_const = F_sw_inst_foc_const
        self.F_sw_inst_foc_const_log10 = F_sw_inst_foc_const_log10
        self.C_sw_inst_foc = C_sw_inst_foc
        self.C_sw_inst_foc_const = C_sw_inst_foc_const
        self.C_velocity = C_velocity
        self.ph = ph
        self.pitch = pitch
        self.instrument = instrument
        self.time_start = time_start
        self.time_end = time_end
        self.velocity = velocity


def _shape_of(spec):
    if isinstance(spec, (list, tuple)):
        first = spec[0]
        if isinstance(first, (list, tuple)):
            return tuple(first[0].shape)
        return tuple(first.shape)
    return tuple(spec.shape)


def check_shape(spec, bands, frames):
    """util_train_test.check_shape: raises ValueError with the reference's message on a mismatch."""
    spec_shape = _shape_of(spec)
    if spec_shape != (bands, frames):
        raise ValueError('Invalid Input shape. Expected: {} . Got: {}'.format((bands, frames), spec_shape))


def _as_device(x, device, dtype):
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=dtype)
    return torch.as_tensor(np.asarray(x), device=device).to(dtype)


def _stack(specs, device, dtype):
    """list of [bands, frames] -> [B, bands, frames, 1] (one stacked tensor, not B concatenations)."""
    return torch.stack([_as_device(s, device, dtype) for s in specs], dim=0).unsqueeze(-1)


def list_to_nd_array(spec, label, device=None, dtype=torch.float32):
    """util_train_test.list_to_nd_array on the device.

    spec: one [bands, frames] array, a list of them (one per sample), or a list of lists (per sample,
    one array per model input channel).  Returns (x, y) with x `[B, bands, frames, 1]` (a list of such
    tensors in the multi-input case) and y `[B, 1]`, exactly the reference's shapes.  dtype defaults to
    float32 (what the Keras models compute in); pass torch.float64 for the reference's storage type."""
    if device is None:
        if not torch.cuda.is_available():
            raise RuntimeError("amt_saga_b200 needs a CUDA device: there is no CPU fallback")
        device = torch.device("cuda", torch.cuda.current_device())
    if isinstance(spec, (list, tuple)):
        y = torch.as_tensor(np.asarray(label, dtype=np.float64).reshape(len(spec), 1), device=device).to(dtype)
        if isinstance(spec[0], (list, tuple)):
            n_in = len(spec[0])
            x = [_stack([sp[i] for sp in spec], device, dtype) for i in range(n_in)]
        else:
            x = _stack(spec, device, dtype)
        return x, y
    x = _as_device(spec, device, dtype).unsqueeze(0).unsqueeze(-1)
    y = torch.as_tensor(np.expand_dims(np.asarray(label, dtype=np.float64), axis=0), device=device).to(dtype)
    return x, y


def model_inputs(sample):
    """The 14 (x, y) pairs of training.py:509-536, in the reference's order (index = model id)."""
    s = sample
    all_x = (s.C_timing, s.C_timing, s.C_sw_pitch, s.C_sw_inst, s.F_sw_inst_foc, s.F_sw_inst_foc_log10,
             s.F_sw_inst_foc_const, s.F_sw_inst_foc_const_log10, [s.C_sw_inst, s.F_sw_inst_foc],
             [s.F_sw_inst_foc, s.ph], s.C_sw_inst_foc, s.C_sw_inst_foc_const, [s.C_sw_inst, s.C_sw_inst_foc],
             s.C_velocity)
    all_y = (s.time_start, s.time_end, s.pitch) + (s.instrument,) * 10 + (s.velocity,)
    return all_x, all_y


class SampleBatcher:
    """Collects note samples and emits, per trained model, one device batch -- the role of the batching
    loop in training.py:486-555 (which pickles per-sample numpy lists into one Queue per model)."""

    def __init__(self, models_to_train, batch_size, device=None, dtype=torch.float32):
        self.models = list(models_to_train)
        self.batch_size = int(batch_size)
        self.device, self.dtype = device, dtype
        self._x, self._y = [], []

    def __len__(self):
        return len(self._x)

    def add(self, sample):
        """Returns the finished batch (see `flush`) when `batch_size` samples are in, else None."""
        all_x, all_y = model_inputs(sample)
        self._x.append([all_x[mi] for mi in self.models])
        self._y.append([all_y[mi] for mi in self.models])
        return self.flush() if len(self._x) >= self.batch_size else None

    def flush(self):
        """[(x, y) per model]: x `[B, bands, frames, 1]` or a list of them, y `[B, 1]`."""
        if not self._x:
            return []
        per_model_x = list(map(list, zip(*self._x)))     # training.py:549-550
        per_model_y = list(map(list, zip(*self._y)))
        self._x, self._y = [], []
        return [list_to_nd_array(x, y, self.device, self.dtype) for x, y in zip(per_model_x, per_model_y)]


def to_dlpack(batch):
    """DLPack capsules for a batch returned by SampleBatcher.flush / list_to_nd_array (zero copy)."""
    from torch.utils.dlpack import to_dlpack as _to

    def conv(t):
        return [conv(u) for u in t] if isinstance(t, (list, tuple)) else _to(t.contiguous())
    return conv(batch)

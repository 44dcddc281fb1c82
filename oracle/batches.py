"""Oracle (TEST INFRASTRUCTURE): numpy restatement of the reference's classifier-input helpers,
/root/reference/util_train_test.py:114-146 (`list_to_nd_array`) -- float64 arrays grown by
`np.concatenate` one sample at a time, exactly as the reference does."""
import numpy as np


def list_to_nd_array(spec, label):
    if isinstance(spec, (list, tuple)):
        if isinstance(spec[0], (list, tuple)):                       # util_train_test.py:118-129
            cb_xa = [np.zeros([0] + list(spec[0][0].shape) + [1]) for _ in range(len(spec[0]))]
            cb_y = np.zeros([0, 1])
            for sp, lab in zip(spec, label):
                cb_y = np.concatenate((cb_y, np.expand_dims([lab], axis=0)))
                for ind, chan in enumerate(sp):
                    cb_xa[ind] = np.concatenate((cb_xa[ind], chan[np.newaxis, :, :, np.newaxis]))
            return cb_xa, cb_y
        cb_x = np.zeros([0] + list(spec[0].shape) + [1])              # :131-140
        cb_y = np.zeros([0, 1])
        for specs, lab in zip(spec, label):
            cb_x = np.concatenate((cb_x, specs[np.newaxis, :, :, np.newaxis]))
            cb_y = np.concatenate((cb_y, np.expand_dims([lab], axis=0)))
        return cb_x, cb_y
    return spec[np.newaxis, :, :, np.newaxis], np.expand_dims(label, axis=0)   # :142-144

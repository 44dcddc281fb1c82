"""No-GPU checks of the boundary: the C-ABI library loads, exports every symbol
include/saga_b200.h declares, fails loudly without a device; host helpers of the
audio_complete mirror agree with the oracle's restatement."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import amt_saga_b200  # noqa: F401
from amt_saga_b200 import _lib, util_audio
from oracle import spectral as osp
from oracle.audio_oracle import AudioOracle, band_edges

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__
    __graft_entry__.build()
    return _lib.lib()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "saga_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(saga_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(lib):
    names = declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), "libsaga_b200.so does not export %s" % n
        assert n in _lib.SIGNATURES, "ctypes binding missing for %s" % n
    assert sorted(_lib.SIGNATURES) == names
    assert lib.saga_abi_version() == 3


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    rc = lib.saga_stft_plan_create(C.byref(h), 2048, 512, 1, None)
    assert rc == _lib.SAGA_ERR_CUDA and lib.saga_last_error_string()
    with pytest.raises(RuntimeError):
        util_audio.audio_complete(np.zeros(10, dtype=np.float32), 2048)


def test_argument_errors_do_not_need_a_device(lib):
    h = C.c_void_p()
    assert lib.saga_stft_plan_create(C.byref(h), 1000, 512, 1, None) == _lib.SAGA_ERR_UNSUPPORTED
    assert lib.saga_stft_plan_create(C.byref(h), 2048, 0, 1, None) == _lib.SAGA_ERR_INVALID
    assert b"hop_length" in lib.saga_last_error_string()
    assert lib.saga_subtract_db_exec(None, None, 0, None, None, 0, None, 0, None, None, None, None, None, 0, 0,
                                     None, None, 1, 0, 4, 4, 4, 1e-5, 80.0, None) == _lib.SAGA_ERR_INVALID
    with pytest.raises(ValueError):
        _lib.check(_lib.SAGA_ERR_INVALID)
    with pytest.raises(_lib.SagaUnsupported):
        _lib.check(_lib.SAGA_ERR_UNSUPPORTED)


def test_host_helpers_match_oracle():
    for note in ("A0", "C1", "C8", "C#4", "Bb3", "D3"):
        assert util_audio.note_to_midi(note) == osp.note_to_midi(note)
        assert abs(util_audio.note_to_hz(note) - osp.note_to_hz(note)) < 1e-9
    for m in (21, 50, 59, 60, 61, 108):
        assert util_audio.midi_to_note(m) == osp.midi_to_note(m)
    for n, b in ((2049, 20), (1025, 20), (1025, 80)):
        assert np.array_equal(util_audio.band_edges(n, b), band_edges(n, b))
    P = np.arange(3 * 11, dtype=float).reshape(3, 11)
    for t in (0, 1, 2, 3, 5, 8, 11):
        for target in (6, 8, 258):
            assert np.array_equal(util_audio.audio_complete._resize(P[:, :t], target),
                                  AudioOracle._resize(P[:, :t], target))
    S = np.random.default_rng(0).random((2049, 7))
    assert np.allclose(util_audio.audio_complete.compress_bands(S, bands=20), AudioOracle.compress_bands(S, bands=20))


def test_frame_arithmetic_of_the_pipeline_without_gpu():
    from amt_saga_b200.pipeline import seconds_to_frames
    o = AudioOracle(np.zeros(263680, dtype=np.float32), 2048, 512)
    o.mag = np.zeros((1025, 516), dtype=np.float32)
    o._v["wf"] = np.zeros(263680, dtype=np.float32)
    for t in (0.0, 0.5, 1.999, 3.0, 5.99):
        assert seconds_to_frames(t, 516, 44100, 263680) == o._seconds_to_frames(t)


def test_options_table_replaces_the_environment_on_the_hot_path():
    """saga_set_option / saga_get_option: unknown names are refused, values round-trip, None unsets."""
    from amt_saga_b200 import _lib
    lib = _lib.lib()
    assert lib.saga_set_option(b"SAGA_NOT_AN_OPTION", b"1") == _lib.SAGA_ERR_INVALID
    before = lib.saga_get_option(b"SAGA_DB_CHUNKS")
    assert lib.saga_set_option(b"SAGA_DB_CHUNKS", b"17") == 0
    assert lib.saga_get_option(b"SAGA_DB_CHUNKS") == b"17"
    assert lib.saga_set_option(b"SAGA_DB_CHUNKS", before) == 0
    assert lib.saga_get_option(b"SAGA_DB_CHUNKS") == before


def test_every_switch_the_docs_and_sources_name_is_in_the_options_table():
    """DESIGN.md lists the tuning / A-B switches and the CUDA sources read them with SAGA_OPT("..."): every such name
    must be known to saga_set_option (a name missing from the table would silently read as unset)."""
    import re
    from amt_saga_b200 import _lib
    lib = _lib.lib()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    names = set()
    csrc = os.path.join(root, "amt-saga_b200", "csrc")
    for f in os.listdir(csrc):
        names |= set(re.findall(r'SAGA_OPT\("(SAGA_[A-Z0-9_]+)"\)', open(os.path.join(csrc, f)).read()))
    doc = open(os.path.join(root, "DESIGN.md")).read()
    table = doc[doc.index("Tuning / A-B switches"):doc.index("No launcher calls `getenv`")]
    names |= set(re.findall(r"`(SAGA_[A-Z0-9_]+)`", table))
    names.discard("SAGA_X")            # the placeholder in the macro's own comment (saga_common.cuh)
    assert len(names) >= 20
    for n in sorted(names):
        cur = lib.saga_get_option(n.encode())
        assert lib.saga_set_option(n.encode(), cur) == 0, n


def test_plan_cache_drops_handles_of_another_process():
    """The reference forks its producers (training.py:623-630).  Plans are cached per (pid, device); entries that
    belong to another process are dropped WITHOUT calling *_destroy on them (their handles live in the parent's
    CUDA context)."""
    import os
    from amt_saga_b200 import ops

    class FakePlan:
        def __init__(self):
            self._h = object()
    foreign = FakePlan()
    key = (os.getpid() + 1, 0, "stft", 2048, 512, True)
    ops._plans[key] = foreign
    try:
        ops._forget_foreign(ops._plans)
        assert key not in ops._plans and foreign._h is None
    finally:
        ops._plans.pop(key, None)


@pytest.mark.gpu
def test_forked_workers_build_their_own_plans():
    """fork-after-import (the reference's Pool): the parent imports the package but never touches CUDA, every forked
    worker initialises its own context lazily and builds its own plans; results agree with the oracle."""
    import subprocess
    import sys
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r"""
import sys, multiprocessing as mp
sys.path.insert(0, %r)
import numpy as np
import amt_saga_b200                      # imported BEFORE the fork, like `import util_audio` in training.py
from amt_saga_b200 import util_audio as ua
from tests.synth import piano_clip

def work(seed):
    import os
    y = piano_clip(seed, 30000)
    a = ua.audio_complete(y, 2048, hop_length=512, carrier="numpy")
    C = a.slice_C(0, 0.5, 8, bins_per_tone=1)
    return os.getpid(), float(a.mag.sum()), float(a.ref_mag), float(C.sum())

if __name__ == "__main__":
    with mp.get_context("fork").Pool(2) as pool:
        res = pool.map(work, [5, 5, 6, 6])
    from oracle.audio_oracle import AudioOracle
    for seed, (pid, s, r, c) in zip([5, 5, 6, 6], res):
        o = AudioOracle(piano_clip(seed, 30000), 2048, hop_length=512)
        assert abs(s - float(o.mag.sum())) <= 1e-4 * float(o.mag.sum()), (s, float(o.mag.sum()))
        assert abs(r - float(o.ref_mag)) <= 1e-5 * float(o.ref_mag)
        assert abs(c - float(o.slice_C(0, 0.5, 8, bins_per_tone=1).sum())) <= 1e-3 * c
    assert res[0][1:] == res[1][1:] and res[2][1:] == res[3][1:]
    assert len({r[0] for r in res}) >= 1
    print("FORK_OK", len({r[0] for r in res}))
""" % root
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and "FORK_OK" in p.stdout, p.stdout[-1500:] + p.stderr[-3000:]

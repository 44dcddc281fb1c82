"""CPU oracle for the AMT-SAGA feature hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, in plain numpy (float64 where the reference's stack is
float64), the algorithm that the reference executes for the hot path named in
BASELINE.json: `util_audio.audio_complete` (/root/reference/util_audio.py:32-527)
and the third-party routines it calls (librosa 0.6.3-era `stft`, `istft`,
`magphase`, `amplitude_to_db`, `cqt`; resampy 0.2.x `kaiser_fast`).

PARITY: PINNED IN TWO LAYERS.
(1) The container: `oracle/ref_class.py` imports the reference's own
`/root/reference/util_audio.py` unmodified and `tests/test_ref_class.py` asserts
`AudioOracle` == that class bit for bit over the producer loop of training.py:265-449;
golden vectors it produced (`tests/golden/ref_class_*.npz`) travel to the GPU box.
(2) The librosa / resampy layer underneath is a restatement (neither package is
installable here): pinned on the reference's own 24-bit FLAC outputs
(`subtraction_demo/<name>_test{,_guess,_sub}.flac`, decoded by `tests/flac_reader.py`):
`stft -> magphase -> ref_mag -> subtract -> istft` reproduces `_sub` to <= 2.5 LSB for
three mixes (tests/test_reference_pins.py), plus `torch.stft`, transformers'
`amplitude_to_db`, an analytic constant-Q response and the literal resampy loop
(tests/test_oracle_pins.py).  STILL UNPINNED against reference outputs, because none
exist: `cqt` (+ the resampy cascade) and `amplitude_to_db`.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import this package.  The product
(`amt-saga_b200/`) never does.
"""

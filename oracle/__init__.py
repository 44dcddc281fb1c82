"""CPU oracle for the AMT-SAGA feature hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, in plain numpy (float64 where the reference's stack is
float64), the algorithm that the reference executes for the hot path named in
BASELINE.json: `util_audio.audio_complete` (/root/reference/util_audio.py:32-527)
and the third-party routines it calls (librosa 0.6.3-era `stft`, `istft`,
`magphase`, `amplitude_to_db`, `cqt`; resampy 0.2.x `kaiser_fast`).

PARITY: PARTLY PINNED.  The reference's own 24-bit FLAC outputs
(`/root/reference/subtraction_demo/<name>_test{,_guess,_sub}.flac`, written by
`test_snippets.py:473-514`) are decoded by the test-only reader
`tests/flac_reader.py`; `stft -> magphase -> ref_mag -> subtract -> istft` of this
package reproduces the reference's `_sub` waveform from `_test` and `_guess` to
<= 2.5 LSB of 24 bits (0.32 LSB rms) for three mixes (tests/test_reference_pins.py,
tests/golden/ref_subtraction_pins.json).  STILL UNPINNED against reference
outputs, because none exist: `cqt` (+ the resampy cascade) and `amplitude_to_db`
(the reference only ships PNG plots of them); those rest on `torch.stft`,
transformers' `amplitude_to_db`, an analytic constant-Q response and an
independent polyphase resampler cross-check (tests/test_oracle_pins.py).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import this package.  The product
(`amt-saga_b200/`) never does.
"""

"""CPU oracle for the AMT-SAGA feature hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, in plain numpy (float64 where the reference's stack is
float64), the algorithm that the reference executes for the hot path named in
BASELINE.json: `util_audio.audio_complete` (/root/reference/util_audio.py:32-527)
and the third-party routines it calls (librosa 0.6.3-era `stft`, `istft`,
`magphase`, `amplitude_to_db`, `cqt`; resampy 0.2.x `kaiser_fast`).

PARITY UNPINNED: the reference ships no numerical golden vectors or asserting
tests for this path and neither librosa nor resampy is installed here (no
network), so the restatement cannot be compared with the reference's own
outputs.  What *is* pinned: frame arithmetic from the reference's FLAC fixture
lengths (tests/test_oracle_pins.py), `torch.stft`, transformers'
`amplitude_to_db`, an analytic constant-Q response, and an independent
polyphase resampler cross-check.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import this package.  The product
(`amt-saga_b200/`) never does.
"""

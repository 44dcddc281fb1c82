"""On-device synthetic audio of SURVEY.md section 8d ("piano-shaped": decaying
inharmonic partials, 5 ms attack, -60 dBFS noise floor, peak 0.9).  Used by the
bench / shard driver to fill HBM before the timed region; not part of the hot
path.  Note parameters come from a CPU generator seeded per clip (seed_base +
clip_id) so any rank can regenerate any clip."""
import math

import numpy as np
import torch


def piano_batch(clip_ids, n_samples, sr=44100, n_notes=24, seed_base=1234, device=None, chunk=50,
                pitch_range=(21, 108)):
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    clip_ids = list(clip_ids)
    out = torch.empty((len(clip_ids), n_samples), device=device, dtype=torch.float32)
    t = torch.arange(n_samples, device=device, dtype=torch.float32) / sr
    dur = n_samples / sr
    for c0 in range(0, len(clip_ids), chunk):
        ids = clip_ids[c0:c0 + chunk]
        prm = np.empty((len(ids), n_notes, 3), dtype=np.float64)
        for i, cid in enumerate(ids):
            rng = np.random.default_rng(seed_base + cid)
            prm[i, :, 0] = rng.uniform(0, max(0.9 * dur, 1e-3), n_notes)                  # onset
            prm[i, :, 1] = rng.integers(pitch_range[0], pitch_range[1] + 1, n_notes)      # midi pitch
            prm[i, :, 2] = rng.integers(30, 121, n_notes)                                 # velocity
        prm_d = torch.as_tensor(prm, device=device, dtype=torch.float32)
        y = torch.zeros((len(ids), n_samples), device=device, dtype=torch.float32)
        for k in range(n_notes):
            onset, pitch, vel = prm_d[:, k, 0:1], prm_d[:, k, 1:2], prm_d[:, k, 2:3]
            f0 = 440.0 * torch.pow(2.0, (pitch - 69.0) / 12.0)
            tau = 0.3 + 1.2 * (108.0 - pitch) / 87.0
            tt = (t.unsqueeze(0) - onset).clamp_min(0.0)
            env = torch.exp(-tt / tau) * (tt / 0.005).clamp_max(1.0) * (vel / 128.0) ** 2
            for h in range(1, 9):
                f = h * f0 * math.sqrt(1 + 1e-4 * h * h)
                y += torch.where(f < sr / 2, env / h, torch.zeros_like(env)) * torch.sin(2 * math.pi * f * tt)
        # noise floor: one generator state per CLIP (not per chunk), so that a clip is the same audio whichever
        # rank or chunk it is generated in (cfg5 compares checksums across 1/2/4/8-GPU partitions)
        g = torch.Generator(device=device)
        for i, cid in enumerate(ids):
            g.manual_seed(seed_base * 7919 + cid)
            y[i] += 1e-3 * torch.randn(n_samples, device=device, generator=g)
        y *= 0.9 / y.abs().amax(dim=1, keepdim=True).clamp_min(1e-12)
        out[c0:c0 + len(ids)] = y
    return out

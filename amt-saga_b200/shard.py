"""Multi-GPU driver: the path is embarrassingly parallel over clips / analysis
windows (that is how the reference parallelises it too: one MIDI file per
`Pool` worker, training.py:623-634), so ranks own contiguous blocks of window
ids and no collective runs on the data path.  A single all_gather collects the
reduced per-window results (post-subtraction ref_mag, checksums, counts).

One process per GPU (`torchrun`); backend nccl on GPUs, gloo in the CPU tests
of the host logic.
"""
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world_size):
    """Contiguous block partition: rank r owns [r*n/W, (r+1)*n/W) (first ranks get the remainder)."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError("bad rank/world_size")
    base, rem = divmod(n_items, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def chunks(start, stop, size):
    for s in range(start, stop, size):
        yield s, min(stop, s + size)


def gather_ragged(local, group=None):
    """All-gather 1-D tensors whose length differs per rank (block partition
    with remainder): pad to the max length, gather, trim."""
    world = dist.get_world_size(group)
    n = torch.tensor([local.numel()], device=local.device, dtype=torch.int64)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    buf = torch.zeros(max(sizes), device=local.device, dtype=local.dtype)
    buf[: local.numel()] = local
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    return torch.cat([o[:s] for o, s in zip(out, sizes)])


def run_sharded(n_windows, compute_fn, chunk_size=600, group=None):
    """Each rank runs `compute_fn(first_id, last_id) -> 1-D tensor of per-window
    results` over its block in chunks; results are gathered in window order on
    every rank.  With torch.distributed uninitialised this is a single shard."""
    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    else:
        rank, world = 0, 1
    lo, hi = shard_range(n_windows, rank, world)
    parts = [compute_fn(a, b) for a, b in chunks(lo, hi, chunk_size)]
    local = torch.cat(parts) if parts else torch.zeros(0)
    return gather_ragged(local, group) if world > 1 else local


def make_pipeline_compute(pipe, synth_fn, offsets_fn):
    """compute_fn for run_sharded backed by a WindowFeaturePipeline: windows are
    synthesised on device from their ids (nothing is read from disk), pushed
    through STFT + CQT + subtract/dB, and reduced to the post-subtraction
    ref_mag per window."""
    def compute(a, b):
        ids = list(range(a, b))
        if len(ids) != pipe.W:
            raise ValueError("chunk size must equal the pipeline's window count")
        wav, guess = synth_fn(ids)
        pipe.run(wav, guess, offsets_fn(ids))
        return pipe.ref.clone()
    return compute

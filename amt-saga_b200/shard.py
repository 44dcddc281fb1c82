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


def run_sharded(n_windows, compute_fn, chunk_size=600, group=None, device=None, dtype=torch.float32,
                require_full_chunks=False):
    """Each rank runs `compute_fn(first_id, last_id) -> 1-D tensor of per-window
    results` over its block in chunks; results are gathered in window order on
    every rank.  With torch.distributed uninitialised this is a single shard.

    `device` / `dtype`: where an EMPTY shard's placeholder lives (a rank that owns no window must
    still join the gather with a tensor of the right device and dtype).  `require_full_chunks`:
    validate ON EVERY RANK, before any compute, that every shard is a whole number of chunks --
    a compute_fn with preallocated buffers that raised on a ragged last chunk on some ranks only
    would leave the others hanging in the collective."""
    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    else:
        rank, world = 0, 1
    if require_full_chunks:
        for r in range(world):
            a, b = shard_range(n_windows, r, world)
            if (b - a) % chunk_size:
                raise ValueError("shard of rank %d has %d windows: not a multiple of chunk_size %d "
                                 "(pad the corpus or pick a divisor)" % (r, b - a, chunk_size))
    lo, hi = shard_range(n_windows, rank, world)
    parts = [compute_fn(a, b) for a, b in chunks(lo, hi, chunk_size)]
    if parts:
        local = torch.cat(parts)
    else:
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() and \
                world > 1 and dist.get_backend(group) == "nccl" else torch.device("cpu")
        local = torch.zeros(0, device=device, dtype=dtype)
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

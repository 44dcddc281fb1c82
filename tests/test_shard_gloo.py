"""Host logic of the multi-GPU driver on CPU: world_size-2 gloo processes,
block partition with remainder, ragged gather in window order."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import amt_saga_b200  # noqa: F401
from amt_saga_b200 import shard


def test_shard_range_covers_everything_once():
    for n in (0, 1, 7, 600, 72000, 72001):
        for w in (1, 2, 3, 8):
            got = []
            for r in range(w):
                a, b = shard.shard_range(n, r, w)
                assert 0 <= a <= b <= n
                got += list(range(a, b))
            assert got == list(range(n))
    with pytest.raises(ValueError):
        shard.shard_range(10, 2, 2)


def _worker(rank, world, port, n_windows, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    calls = []

    def compute(a, b):
        calls.append((a, b))
        return torch.arange(a, b, dtype=torch.float32) * 2.0 + 1.0

    res = shard.run_sharded(n_windows, compute, chunk_size=4)
    torch.save({"res": res, "calls": calls}, os.path.join(out_dir, "r%d.pt" % rank))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_windows", [10, 11, 1])   # 1: rank 1 owns an EMPTY shard and still joins the gather
def test_run_sharded_two_ranks_gloo(tmp_path, n_windows):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, n_windows, str(tmp_path)), nprocs=2, join=True)
    expect = torch.arange(n_windows, dtype=torch.float32) * 2.0 + 1.0
    seen = []
    for r in range(2):
        d = torch.load(os.path.join(str(tmp_path), "r%d.pt" % r))
        assert torch.equal(d["res"], expect)          # every rank holds all windows, in order
        seen += [i for a, b in d["calls"] for i in range(a, b)]
        assert all(b - a <= 4 for a, b in d["calls"])
    assert sorted(seen) == list(range(n_windows))     # no window computed twice / skipped


def test_run_sharded_validates_chunking_on_every_rank_before_computing():
    """A ragged last chunk must be refused before any rank computes (otherwise the ranks whose shard is
    fine would hang in the gather while another raised)."""
    ran = []
    with pytest.raises(ValueError, match="not a multiple of chunk_size"):
        shard.run_sharded(10, lambda a, b: ran.append((a, b)), chunk_size=4, require_full_chunks=True)
    assert ran == []
    out = shard.run_sharded(8, lambda a, b: torch.arange(a, b, dtype=torch.float32), chunk_size=4,
                            require_full_chunks=True)
    assert torch.equal(out, torch.arange(8, dtype=torch.float32))
    assert shard.run_sharded(0, lambda a, b: None).numel() == 0

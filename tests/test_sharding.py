"""Host-side sharding logic: unit tests + a world-size-2 gloo run on CPU (no GPU needed)."""
import os
import sys

import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_round_robin_covers_everything_once():
    from vlgae_b200.sharding import shard_indices

    for n in (0, 1, 7, 128, 1000):
        for w in (1, 2, 3, 8):
            seen = torch.cat([shard_indices(n, r, w) for r in range(w)])
            assert sorted(seen.tolist()) == list(range(n))
    with pytest.raises(ValueError):
        shard_indices(4, 2, 2)


def test_length_sorted_batches_stay_balanced():
    from vlgae_b200.sharding import imbalance

    g = torch.Generator().manual_seed(4)
    L = torch.randint(4, 41, (1024,), generator=g).sort(descending=True).values  # cfg4: global batch 1024
    for w in (2, 4, 8):
        assert imbalance(L, w) < 1.02
    # contiguous blocks of a sorted batch would be badly unbalanced; round-robin is not
    blocks = torch.stack([((L[r * 128:(r + 1) * 128] + 1.0) ** 3).sum() for r in range(8)])
    assert float(blocks.max() / blocks.mean()) > 2.0


def _worker(rank, world, port, tmp):
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    from vlgae_b200.sharding import gather_heads, shard_batch

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    B, N = 11, 6
    g = torch.Generator().manual_seed(0)
    dec = torch.randn(B, N, 2, 2, 2, generator=g)
    attach = torch.randn(B, N, N, 2, generator=g)
    lengths = torch.randint(1, N, (B,), generator=g).sort(descending=True).values
    d, a, L, idx = shard_batch(dec, attach, lengths, rank, world)
    assert torch.equal(d, dec[idx]) and torch.equal(a, attach[idx]) and torch.equal(L, lengths[idx])
    # stand-in for the per-rank parse: heads[b, c] = 100 * sentence_id + c
    local = (100 * idx).unsqueeze(1) + torch.arange(N).unsqueeze(0)
    full = gather_heads(local, B)
    want = (100 * torch.arange(B)).unsqueeze(1) + torch.arange(N).unsqueeze(0)
    ok = torch.equal(full, want)
    torch.save({"ok": ok, "n": int(idx.numel())}, os.path.join(tmp, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo(tmp_path):
    import socket

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    res = [torch.load(os.path.join(str(tmp_path), f"r{r}.pt")) for r in range(2)]
    assert all(r["ok"] for r in res)
    assert sorted(r["n"] for r in res) == [5, 6]

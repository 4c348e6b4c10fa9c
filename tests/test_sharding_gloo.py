"""N>1 host logic on CPU: world_size-2 gloo run of the batch shard + gather + row-0-stop merge."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from starvector_b200.parallel import all_gather_generated, merge_generated, shard_range

PAD, STOP = 99, [7, 8]


def test_shard_range_covers_batch():
    for B in (1, 2, 7, 8, 64):
        for W in (1, 2, 4, 8):
            spans = [shard_range(B, r, W) for r in range(W)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(W - 1))


def _single_process_rectangle(rows, stop):
    """What one HF process returns for independent rows `rows` (lists that end where each row would stop)."""
    n = max(len(r) for r in rows)
    for s in range(len(stop), len(rows[0]) + 1):
        if rows[0][s - len(stop):s] == stop:
            n = s
            break
    return torch.tensor([r[:n] + [PAD] * (n - len(r[:n])) for r in rows], dtype=torch.int32)


def test_merge_matches_single_process_semantics():
    rows = [[1, 2, 7, 8], [3, 3, 3, 3, 3, 3], [4, 4], [5, 5, 5, 5, 5]]          # row 0 hits the stop at step 4
    want = _single_process_rectangle(rows, STOP)
    r0 = torch.tensor([[1, 2, 7, 8], [3, 3, 3, 3]], dtype=torch.int32)          # rank 0 stopped by row 0
    r1 = torch.tensor([[4, 4, PAD, PAD, PAD], [5, 5, 5, 5, 5]], dtype=torch.int32)   # rank 1 ran to its own end
    assert torch.equal(merge_generated([r0, r1], STOP, PAD), want)
    # no stop: global length is the longest rank, shorter ranks padded
    r0b = torch.tensor([[1, 2, 3]], dtype=torch.int32)
    r1b = torch.tensor([[4, 4, 4, 4, 4]], dtype=torch.int32)
    got = merge_generated([r0b, r1b], STOP, PAD)
    assert got.tolist() == [[1, 2, 3, PAD, PAD], [4, 4, 4, 4, 4]]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    B, max_new = 5, 6
    lo, hi = shard_range(B, rank, world)
    rows = {0: [[1, 2, 7, 8], [3, 3, 3, 3], [6, 6, 6, 6]], 1: [[4, 4, PAD, PAD, PAD], [5, 5, 5, 5, 5]]}[rank]
    local = torch.tensor(rows, dtype=torch.int32)
    assert local.shape[0] == hi - lo
    out = all_gather_generated(local, max_new, STOP, PAD, B)
    q.put((rank, out.tolist()))
    dist.destroy_process_group()


def test_world2_gloo_gather():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = dict(q.get(timeout=120) for _ in range(2))
    [p.join(60) for p in procs]
    want = [[1, 2, 7, 8], [3, 3, 3, 3], [6, 6, 6, 6], [4, 4, PAD, PAD], [5, 5, 5, 5]]
    assert res[0] == want and res[1] == want

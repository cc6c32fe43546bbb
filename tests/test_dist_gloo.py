"""world_size-2 gloo test (CPU) of the multi-GPU host logic: row / column / upper-packed shards -> all-gather -> full matrix,
and the exactly-once ownership of the L-shaped shards."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sfb_b200.device import gather_row_slabs, shard_rows
    n = 37
    rng = np.random.default_rng(3)
    M = rng.random((n, n))                       # M[i, i']
    ell = np.repeat(np.arange(8), [3, 5, 4, 6, 2, 7, 5, 5])
    cost = (10.0 - ell) * np.ones(n)
    ranges = shard_rows(cost, ell, world)
    lo, hi = ranges[rank]
    slab = torch.from_numpy(np.ascontiguousarray(M[lo:hi, :].T))   # (nout, rows) like DevicePipeline
    full = gather_row_slabs(slab, ranges, n)
    ok = np.array_equal(full.numpy().T, M)
    # column slabs: in-place uneven all-gather
    from sfb_b200.device import allgather_col_slabs
    cr = shard_rows(cost, ell, world)
    fullc = torch.zeros((n, n), dtype=torch.float64)
    clo, chi = cr[rank]
    fullc[clo:chi] = torch.from_numpy(np.ascontiguousarray(M.T[clo:chi]))
    allgather_col_slabs(fullc, cr)
    ok = ok and np.array_equal(fullc.numpy().T, M)
    # upper-packed slabs: column j keeps rows [0, rend(j)); uneven in-place all-gather of the 1-D buffer
    from sfb_b200.device import allgather_packed_slabs
    first = np.concatenate([[0], np.cumsum(np.bincount(ell))])
    rend = first[ell + 1]
    off = np.concatenate([[0], np.cumsum(rend)])
    packed = torch.zeros(int(off[-1]), dtype=torch.float64)
    for j in range(clo, chi):
        packed[off[j]:off[j + 1]] = torch.from_numpy(M[:rend[j], j].copy())
    allgather_packed_slabs(packed, [(int(off[l]), int(off[h])) for l, h in cr])
    ref = np.concatenate([M[:rend[j], j] for j in range(n)])
    ok = ok and np.array_equal(packed.numpy(), ref)
    # L-shaped shards: rank g owns the upper-packed columns of its range and, mirrored from them, its rows of the part below
    # the block diagonal; summed over the ranks every element of M is owned exactly once and the pieces rebuild M (here a
    # symmetric kernel with f = 1, so the mirror image of M[r, j] is M[j, r] itself)
    S = M + M.T
    owned = torch.zeros((n, n), dtype=torch.float64)
    rebuilt = torch.zeros((n, n), dtype=torch.float64)
    rows = np.zeros((chi, chi - clo))                        # rows[r, j - lo] = S[j, r] for l(r) < l(j)
    for j in range(clo, chi):
        owned[:rend[j], j] += 1                              # packed column j
        rebuilt[:rend[j], j] = torch.from_numpy(S[:rend[j], j].copy())
        below = ell[:chi] < ell[j]
        rows[np.flatnonzero(below), j - clo] = S[np.flatnonzero(below), j]      # what cmix_mirror_rows_kernel writes
        owned[j, np.flatnonzero(below)] += 1                 # mirrored row j
        rebuilt[j, np.flatnonzero(below)] = torch.from_numpy(rows[np.flatnonzero(below), j - clo].copy())
    dist.all_reduce(owned)
    dist.all_reduce(rebuilt)
    ok = ok and bool((owned == 1).all()) and np.array_equal(rebuilt.numpy(), S)
    out = torch.tensor([1.0 if ok else 0.0])
    dist.all_reduce(out, op=dist.ReduceOp.MIN)
    if rank == 0:
        ret.put((float(out.item()), ranges))
    dist.destroy_process_group()


def test_row_shard_allgather_world2():
    ctx = mp.get_context("spawn")
    ret = ctx.SimpleQueue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    ok, ranges = ret.get()
    assert ok == 1.0 and ranges[0][1] == ranges[1][0] and ranges[0][1] not in (0, 37)

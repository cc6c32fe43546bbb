"""Measure ways of replicating column slabs of a 3.75 GB matrix across the GPUs of one node.
torchrun --nproc-per-node N tools/exchange_probe.py"""
import os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
n = 21638
full = torch.zeros(n * n, dtype=torch.float64, device="cuda")
cols = [n * g // world for g in range(world + 1)]
lo, hi = cols[rank], cols[rank + 1]
def timeit(fn, name, reps=4):
    for _ in range(2): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); dist.barrier()
    dt = (time.perf_counter() - t0) / reps
    t = torch.tensor([dt], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0: print(f"{name:34s} {t.item()*1e3:8.2f} ms  ({8*n*n*(world-1)/world/t.item()/1e9:7.1f} GB/s inbound per GPU)", flush=True)
# (a) NCCL all_gather_into_tensor, equal shards (pad to equal)
m = max(cols[g + 1] - cols[g] for g in range(world))
padded = torch.zeros(world * m * n, dtype=torch.float64, device="cuda")
mine = torch.ones(m * n, dtype=torch.float64, device="cuda")
timeit(lambda: dist.all_gather_into_tensor(padded, mine), "nccl all_gather_into_tensor")
# (b) grouped send/recv, uneven, in place
def sendrecv():
    ops = []
    for g in range(world):
        if g == rank: continue
        ops.append(dist.P2POp(dist.isend, full[lo * n:hi * n], g))
        ops.append(dist.P2POp(dist.irecv, full[cols[g] * n:cols[g + 1] * n], g))
    for w in dist.batch_isend_irecv(ops): w.wait()
timeit(sendrecv, "nccl batch_isend_irecv (in place)")
# (c) broadcasts in place
def bcasts():
    for g in range(world):
        dist.broadcast(full[cols[g] * n:cols[g + 1] * n], src=g)
timeit(bcasts, "nccl broadcast x world (in place)")
# (d) copy engines: IPC peer copies
from sfb_b200.device import PeerMatrix
from sfb_b200 import _lib
import ctypes as C
lib = _lib.load(); _lib.check(lib.sfb_set_device(rank))
pm = PeerMatrix(n)
def dma():
    _lib.check(lib.sfb_push_cols_to_peers(pm.ptr, pm.peer_array, len(pm.peer_ptrs), lo, hi, n, C.c_void_p(torch.cuda.current_stream().cuda_stream)))
timeit(dma, "cudaMemcpyAsync peer x(world-1)")
pm.close()
dist.destroy_process_group()

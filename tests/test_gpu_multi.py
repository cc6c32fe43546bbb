"""Multi-GPU parity (needs >= 2 CUDA devices; run with `gpurun --gpus 2`): shell-sharded stage 1 + row-sharded
stage 2/3 + NCCL all-gather must reproduce the single-GPU matrix (to rounding: narrower column tiles are used)."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    import warnings
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    import torch.distributed as dist
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import sfb_b200 as sfb
    from sfb_b200.device import DevicePipeline
    a = sfb.AnlmModes(0.03, 500.0, 1000.0)
    c = sfb.ClnnModes(a)
    wm = sfb.ConfigurationSpaceModes(a, 21)          # 21 shells: uneven split, padded shards
    rng = np.random.default_rng(7)
    win = rng.random((wm.nr, wm.npix))
    win[:, ::4] = 0
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        G = sfb.rsdrgnlr(a, wm)
    pipe = DevicePipeline(wm, c, G)
    d_win = torch.from_numpy(np.ascontiguousarray(win.T)).cuda()
    single = pipe.power_win_mix(d_win).clone()
    alm_single = pipe.alm.clone()
    full, ranges = pipe.power_win_mix_sharded(d_win)
    rel = lambda x, y: float((x - y).norm() / y.norm())
    ok = rel(pipe.alm, alm_single) < 1e-13 and rel(full, single) < 1e-13
    ag, _ = pipe.power_win_mix_allgather(d_win)
    ok = ok and rel(ag, single) < 1e-13
    pk, _ = pipe.power_win_mix_allgather_packed(d_win)      # default path: upper-packed slabs + unpack/mirror
    ok = ok and rel(pk, single) < 1e-13
    from sfb_b200.device import PeerBuffer
    pb = PeerBuffer(int(pipe.packed_offsets()[-1]))
    for kw in (dict(), dict(div2Lp1=True, interchange_NN=True)):
        pl, _ = pipe.power_win_mix_allgather_packed(d_win, peer_packed=pb, **kw)   # exchange fused into the expansion
        ref = pipe.power_win_mix_rows(0, pipe.nout, **kw)
        ok = ok and rel(pl, ref) < 1e-13
    pb.close()
    from sfb_b200.device import PeerMatrix
    pm = PeerMatrix(pipe.nout)
    for mode in ("cols", "dma", "stores"):
        pm.tensor.zero_()
        torch.cuda.synchronize()
        dist.barrier()
        fused, _ = pipe.power_win_mix_fused(d_win, pm, mode=mode)
        ok = ok and rel(fused, single) < 1e-13
    pm.close()
    flag = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        ret.put((float(flag.item()), ranges, float((full - single).abs().max())))
    pipe.close()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_sharded_equals_single_gpu():
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 4)
    ctx = mp.get_context("spawn")
    ret = ctx.SimpleQueue()
    port = 29600 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    ok, ranges, err = ret.get()
    assert ok == 1.0, (ranges, err)

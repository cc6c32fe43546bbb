"""Multi-GPU parity (needs >= 2 CUDA devices; run with `gpurun --gpus 2`).

1. one process per GPU (torch.distributed / NCCL): shell-sharded stage 1 + sharded stage 2/3 + all-gather variants must
   reproduce the single-GPU matrix (to rounding: narrower column tiles are used);
2. ONE process, sfb_set_devices(n): the host-pointer C ABI itself shards over the GPUs and must return the same
   arrays as with one device (this is what the Julia drop-in uses)."""
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


def _set_devices_case(sfb, ndev, nr=21):
    import warnings
    rng = np.random.default_rng(17)
    a = sfb.AnlmModes(0.03, 500.0, 1000.0)
    c = sfb.ClnnModes(a)
    wm = sfb.ConfigurationSpaceModes(a, nr)              # 21 shells: uneven shell shards; 9 shells on 4+ GPUs: an empty one
    win = np.asfortranarray(rng.random((wm.nr, wm.npix)))
    win[:, ::4] = 0
    win2 = np.asfortranarray(rng.random((wm.nr, wm.npix)))
    res = {}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        sfb.set_devices(ndev)
        try:
            res["M"] = sfb.power_win_mix(win, wm, c)
            res["M_flags"] = sfb.power_win_mix(win, wm, c, div2Lp1=True, interchange_NN=True, lnn_min=7)
            res["M_cross"] = sfb.power_win_mix(win, win2, wm, c)
            res["Wr"] = sfb.calc_Wr_lm(win, 2 * a.lmax, a.nside)
            res["Wr_fast"] = sfb.calc_Wr_lm(win, 2 * a.lmax, a.nside, layout=1)
            res["Wr_up"] = sfb.calc_Wr_lm(np.asfortranarray(win2[:, :12 * 8 * 8]), 2 * a.lmax, a.nside)  # nside 8 -> udgrade
            wt, v = sfb.bandpower_binning_weights(c, dl=3)
            bc = sfb.ClnnBinnedModes(wt, v, c)
            res["N"] = sfb.power_win_mix(win, wt, v, wm, bc)
            res["wM"] = sfb.power_win_mix(win, wt, None, wm, bc)
            res["Mv"] = sfb.power_win_mix(win, None, v, wm, bc, div2Lp1=True)
            out = sfb.pinned_empty(res["M"].shape)
            res["M_pinned"] = sfb.power_win_mix(win, wm, c, out=out).copy()
        finally:
            sfb.set_devices(1)
    return res


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_set_devices_host_api_equals_single_gpu():
    import sfb_b200 as sfb
    from conftest import relerr
    bad = []
    for nr in (21, 9):
        ref = _set_devices_case(sfb, 1, nr)
        for ndev in sorted({2, min(torch.cuda.device_count(), 4), min(torch.cuda.device_count(), 8)}):
            got = _set_devices_case(sfb, ndev, nr)
            assert sfb.get_devices() == 1
            for k in ref:
                assert got[k].shape == ref[k].shape, k
                e = relerr(got[k], ref[k])
                print(f"set_devices({ndev}) nr={nr} {k}: rel err {e:.3e}")
                if not e < 1e-13:
                    bad.append((ndev, nr, k, e))
    assert not bad, bad
    with pytest.raises(sfb._lib.SFBError, match="devices requested|must be in 1..8"):
        sfb.set_devices(torch.cuda.device_count() + 1)


def test_set_devices_one_is_the_default_and_pinned_buffers_work():
    # runs on a 1-GPU box: sfb_set_devices(1), page-locked result buffer, same numbers as a pageable one
    import warnings

    import sfb_b200 as sfb
    from conftest import relerr
    assert sfb.get_devices() == 1
    sfb.set_devices(1)
    a = sfb.AnlmModes(2, 4, 500.0, 1000.0)
    c = sfb.ClnnModes(a)
    wm = sfb.ConfigurationSpaceModes(a, 24)
    win = np.random.default_rng(2).random((wm.nr, wm.npix))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        M = sfb.power_win_mix(win, wm, c)
        out = sfb.pinned_empty(M.shape)
        M2 = sfb.power_win_mix(win, wm, c, out=out)
    assert M2 is out and out.flags.f_contiguous and relerr(M2, M) == 0.0

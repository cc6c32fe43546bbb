"""Device-resident pipeline and multi-GPU row sharding.

torch is plumbing only (device memory, streams, `torch.distributed`); every kernel is in libsfb_b200.so.

The coupling matrix shards by output rows (l,n,n'): each (i,i') element is an independent closure call in the
reference (src/windows.jl:717-738).  Rank g computes the contiguous row range [lo_g, hi_g) chosen on l-block
boundaries by a cost prefix sum, as a compact (hi_g-lo_g) x nout column-major slab, and an NCCL all-gather
assembles the full matrix on every rank (the reference's only analogue is the pmap gather, src/windows.jl:841-861).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .modes import getlmsize


def _torch():
    import torch
    return torch


def shard_shells(nr, world):
    """Contiguous shell ranges [lo, hi) per rank for stage 1 (shells are independent, src/windows.jl:531-535)."""
    base = -(-nr // world)
    return [(min(g * base, nr), min((g + 1) * base, nr)) for g in range(world)]


def shard_rows(costs, ell_of_row, world):
    """Contiguous row ranges [lo, hi) per rank, cut only where `ell_of_row` changes, balancing `costs`.
    Returns a list of `world` (lo, hi) tuples covering [0, n)."""
    costs = np.asarray(costs, dtype=np.float64)
    n = costs.size
    ell = np.asarray(ell_of_row)
    cuts = np.concatenate([[0], np.flatnonzero(np.diff(ell)) + 1, [n]]).astype(np.int64)
    cum = np.concatenate([[0.0], np.cumsum(costs)])
    total = cum[-1]
    bounds = [0]
    for g in range(1, world):
        best = int(cuts[np.argmin(np.abs(cum[cuts] - total * g / world))])
        bounds.append(max(best, bounds[-1]))
    bounds.append(n)
    return [(bounds[g], bounds[g + 1]) for g in range(world)]


class DevicePipeline:
    """Plans + device buffers for repeated power_win_mix calls on one GPU (one process per GPU)."""

    def __init__(self, wmodes, cmodes, G, nside_win=None, lnn_min=1, device=None):
        torch = _torch()
        self.lib = _lib.load()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        torch.cuda.set_device(self.device)
        _lib.check(self.lib.sfb_set_device(self.device.index))
        amodes = cmodes.amodes
        self.amodes, self.cmodes, self.wmodes = amodes, cmodes, wmodes
        self.nr = wmodes.nr
        self.LMAX = 2 * amodes.lmax
        self.lmsize = getlmsize(self.LMAX)
        self.nside_in = wmodes.nside if nside_win is None else nside_win
        self.lnn = np.asfortranarray(cmodes.lnn, dtype=np.int64)
        self.lnnsize = self.lnn.shape[1]
        self.nout = self.lnnsize - lnn_min + 1
        G = np.asfortranarray(G, dtype=np.float64)
        self._sht = C.c_void_p()
        self._cmix = C.c_void_p()
        _lib.check(self.lib.sfb_sht_plan_create(C.byref(self._sht), self.nside_in, amodes.nside, self.LMAX, self.nr))
        _lib.check(self.lib.sfb_cmix_plan_create(C.byref(self._cmix), _lib.ptr(self.lnn), self.lnnsize, lnn_min,
                                                 _lib.ptr(G), self.nr, amodes.nmax, amodes.lmax))
        nalm = self.lib.sfb_sht_alm_doubles(self._sht)
        self.alm = torch.empty(nalm, dtype=torch.float64, device=self.device)
        costs = np.zeros(self.nout)
        _lib.check(self.lib.sfb_cmix_row_costs(self._cmix, _lib.ptr(costs), self.nout))
        self.row_costs = costs
        ccosts = np.zeros(self.nout)
        _lib.check(self.lib.sfb_cmix_col_costs(self._cmix, _lib.ptr(ccosts), self.nout))
        self.col_costs = ccosts
        self.ell_of_row = self.lnn[0, lnn_min - 1:]

    def close(self):
        if self._sht:
            self.lib.sfb_sht_plan_destroy(self._sht)
            self._sht = C.c_void_p()
        if self._cmix:
            self.lib.sfb_cmix_plan_destroy(self._cmix)
            self._cmix = C.c_void_p()
        if getattr(self, "_sht_shard", None):
            self.lib.sfb_sht_plan_destroy(self._sht_shard)
            self._sht_shard = None
        if getattr(self, "_alm_peer", None):
            for pb in self._alm_peer:
                pb.close()
            self._alm_peer = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(_torch().cuda.current_stream().cuda_stream)

    # ---- stage 1 -------------------------------------------------------------------------------
    def calc_wr_lm(self, d_win, niter=3):
        """d_win: torch float64 CUDA tensor holding the Julia (nr, npix) array in memory order, i.e. shape
        (npix, nr) contiguous.  Fills and returns the planar W_lm(r) buffer."""
        assert d_win.is_cuda and d_win.dtype == _torch().float64 and d_win.is_contiguous()
        assert d_win.shape == (12 * self.nside_in ** 2, self.nr)
        _lib.check(self.lib.sfb_calc_wr_lm_dev(self._sht, d_win.data_ptr(), self.nr, niter, self.alm.data_ptr(),
                                               self._stream()))
        return self.alm

    def wr_lm_complex(self, layout=0):
        torch = _torch()
        out = torch.empty((self.lmsize, self.nr), dtype=torch.complex128, device=self.device)
        _lib.check(self.lib.sfb_alm_to_complex_dev(self._sht, self.alm.data_ptr(), layout, out.data_ptr(),
                                                   self._stream()))
        return out  # out.T is the Julia (nr, lmsize) matrix

    def calc_wr_lm_sharded(self, d_win, group=None, niter=3):
        """Stage 1 with the shells partitioned over the ranks of `group`, then an all-gather of the planar
        W_lm(r) shards (cfg5: 74 MB in total) so that every rank can build all W_{L1}."""
        torch = _torch()
        import torch.distributed as dist
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        if world == 1:
            return self.calc_wr_lm(d_win, niter)
        rank = dist.get_rank(group)
        ranges = shard_shells(self.nr, world)
        lo, hi = ranges[rank]
        cnt = hi - lo
        nrp_g = 8 * (-(-max(h - l for l, h in ranges) // 8))
        nrp_mine = 8 * (-(-max(cnt, 1) // 8))
        import os
        peer = os.environ.get("SFB_ALM_GATHER", "peer") == "peer" and world <= 8
        if getattr(self, "_sht_shard", None) is None:
            self._sht_shard = C.c_void_p()
            _lib.check(self.lib.sfb_sht_plan_create(C.byref(self._sht_shard), self.nside_in, self.amodes.nside,
                                                    self.LMAX, max(cnt, 1)))
            self._alm_shard = torch.zeros(self.lmsize * 2 * nrp_mine, dtype=torch.float64, device=self.device)
            self._alm_recv = None if peer else torch.empty(world * self.lmsize * 2 * nrp_g, dtype=torch.float64,
                                                           device=self.device)
            # peer gather: every rank's shard lives in an IPC-mapped buffer (two of them, alternating between calls) and
            # the placement kernel reads the peers' shards straight over NVLink — no NCCL all-gather, no staging copy
            self._alm_peer = [PeerBuffer(self.lmsize * 2 * nrp_mine, group) for _ in range(2)] if peer else None
            for pb in self._alm_peer or []:
                pb.tensor.zero_()
            self._alm_peer_k = 0
        if peer:
            return self._calc_wr_lm_peer(d_win, group, niter, ranges, lo, cnt, nrp_mine)
        if cnt > 0:
            _lib.check(self.lib.sfb_calc_wr_lm_dev(self._sht_shard, d_win.data_ptr() + 8 * lo, self.nr, niter,
                                                   self._alm_shard.data_ptr(), self._stream()))
        send = self._alm_shard
        if nrp_mine != nrp_g:
            send = torch.nn.functional.pad(send.view(self.lmsize, 2, nrp_mine), (0, nrp_g - nrp_mine)).contiguous().view(-1)
        dist.all_gather_into_tensor(self._alm_recv, send, group=group)
        # placement of the shards into the planar [lm][re,im][nrp] buffer: one library kernel
        if getattr(self, "_gather_args", None) is None:
            per = self.lmsize * 2 * nrp_g
            self._gather_args = ((C.c_void_p * world)(*[self._alm_recv.data_ptr() + 8 * g * per for g in range(world)]),
                                 np.asarray([r[0] for r in ranges] + [self.nr], dtype=np.int64),
                                 np.full(world, nrp_g, dtype=np.int64))
        ptrs, bounds, strides = self._gather_args
        _lib.check(self.lib.sfb_alm_gather_shards_dev(ptrs, _lib.ptr(bounds), _lib.ptr(strides), world, self.LMAX, self.nr,
                                                      self.alm.data_ptr(), self._stream()))
        return self.alm

    def _calc_wr_lm_peer(self, d_win, group, niter, ranges, lo, cnt, nrp_mine):
        """Stage 1 on this rank's shells into its IPC-mapped shard, a stream-ordered barrier, then ONE kernel that places
        all shards (the peers' read over NVLink) into the planar W_lm(r) buffer.  The shard buffers alternate between
        calls: a rank may start its next stage 1 while a slower peer still reads the previous shard, and by the time a
        buffer comes round again every rank has passed the barrier of the call in between, i.e. finished reading it."""
        world = len(ranges)
        b = self._alm_peer_k & 1
        self._alm_peer_k += 1
        pb = self._alm_peer[b]
        if cnt > 0:
            _lib.check(self.lib.sfb_calc_wr_lm_dev(self._sht_shard, d_win.data_ptr() + 8 * lo, self.nr, niter,
                                                   pb.ptr.value, self._stream()))
        stream_barrier(self, group)
        if getattr(self, "_peer_gather_args", None) is None:
            self._peer_gather_args = {}
        if b not in self._peer_gather_args:
            strides = np.asarray([8 * (-(-max(h - l, 1) // 8)) for l, h in ranges], dtype=np.int64)
            self._peer_gather_args[b] = ((C.c_void_p * world)(*pb.rank_ptrs),
                                         np.asarray([r[0] for r in ranges] + [self.nr], dtype=np.int64), strides)
        ptrs, bounds, strides = self._peer_gather_args[b]
        _lib.check(self.lib.sfb_alm_gather_shards_dev(ptrs, _lib.ptr(bounds), _lib.ptr(strides), world, self.LMAX, self.nr,
                                                      self.alm.data_ptr(), self._stream()))
        return self.alm

    # ---- stage 2+3 -----------------------------------------------------------------------------
    def power_win_mix_rows(self, lo, hi, out=None, alm2=None, div2Lp1=False, interchange_NN=False):
        """Rows [lo, hi) of M as a torch tensor of shape (nout, hi-lo) (memory = column-major (hi-lo) x nout)."""
        torch = _torch()
        if out is None:
            out = torch.empty((self.nout, hi - lo), dtype=torch.float64, device=self.device)
        a2 = self.alm if alm2 is None else alm2
        _lib.check(self.lib.sfb_power_win_mix_dev(self._cmix, self.alm.data_ptr(), a2.data_ptr(), int(div2Lp1),
                                                  int(interchange_NN), lo, hi, out.data_ptr(), hi - lo,
                                                  self._stream()))
        return out

    def power_win_mix_cols(self, lo, hi, out=None, alm2=None, div2Lp1=False, interchange_NN=False):
        """Columns [lo, hi) of M (all rows): a contiguous slab of the column-major matrix, returned as a torch
        tensor of shape (hi-lo, nout)."""
        torch = _torch()
        if out is None:
            out = torch.empty((hi - lo, self.nout), dtype=torch.float64, device=self.device)
        a2 = self.alm if alm2 is None else alm2
        _lib.check(self.lib.sfb_power_win_mix_block_dev(self._cmix, self.alm.data_ptr(), a2.data_ptr(), int(div2Lp1),
                                                        int(interchange_NN), 0, self.nout, lo, hi, out.data_ptr(),
                                                        self.nout, self._stream()))
        return out

    def power_win_mix(self, d_win, **kw):
        self.calc_wr_lm(d_win)
        return self.power_win_mix_rows(0, self.nout, **kw)

    # ---- multi-GPU -----------------------------------------------------------------------------
    def power_win_mix_allgather(self, d_win, full=None, group=None, div2Lp1=False, interchange_NN=False):
        """The multi-GPU path: stage 1 shell-sharded, stage 2+3 sharded over the COLUMN index (L,N,N') with a cost
        prefix sum, each rank writing its contiguous column slab straight into its copy of the full matrix, then
        an in-place uneven NCCL all-gather (grouped send/recv over NVLink; no padding, no placement pass).
        Returns (full, ranges): `full` has shape (nout, nout) and holds Mᵀ in C order, i.e. Julia's column-major M."""
        torch = _torch()
        import torch.distributed as dist
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        if full is None:
            full = torch.empty((self.nout, self.nout), dtype=torch.float64, device=self.device)
        self.calc_wr_lm_sharded(d_win, group)
        ranges = shard_rows(self.col_costs, self.ell_of_row, world)
        lo, hi = ranges[rank]
        if hi > lo:
            self.power_win_mix_cols(lo, hi, out=full[lo:hi], div2Lp1=div2Lp1, interchange_NN=interchange_NN)
        if world > 1:
            allgather_col_slabs(full, ranges, group)
        return full, ranges

    # ---- upper-packed exchange (default multi-GPU path) -----------------------------------------------
    def packed_offsets(self):
        """Element offset of every output column in upper-packed storage (nout + 1 entries, host int64)."""
        if getattr(self, "_packed_off", None) is None:
            off = np.zeros(self.nout + 1, dtype=np.int64)
            _lib.check(self.lib.sfb_cmix_packed_offsets(self._cmix, off.ctypes.data_as(C.c_void_p), self.nout + 1))
            self._packed_off = off
            cu = np.zeros(self.nout)
            _lib.check(self.lib.sfb_cmix_col_costs_upper(self._cmix, _lib.ptr(cu), self.nout))
            self.col_costs_upper = cu
        return self._packed_off

    def packed_shard_ranges(self, world, balance=None):
        """Column ranges (L-block aligned) of the upper-packed path.  Every rank's time is its own compute plus serving
        its slab to the world-1 readers, so the split balances  flops / (29 TFLOP/s) + (world-1) bytes / (700 GB/s)
        (rates measured at cfg4: register-Z kernel, NVLink peer loads); balance="cost" uses the flop model alone."""
        import os
        off = self.packed_offsets()
        balance = balance or os.environ.get("SFB_SHARD_BALANCE", "mixed")
        w = self.col_costs_upper / 29e12
        if balance == "mixed" and world > 1:
            w = w + (world - 1) * 8.0 * np.diff(off) / 700e9
        if balance == "lshard":
            # L-shaped shards: index j also owns its row below the block diagonal, mirrored locally at HBM speed
            # (read + write of the elements in front of j's own l-block; ~4.8 TB/s measured for the mirror passes)
            ell = np.asarray(self.ell_of_row)
            cuts = np.concatenate([[0], np.flatnonzero(np.diff(ell)) + 1, [ell.size]])
            lstart = np.repeat(cuts[:-1], np.diff(cuts)).astype(np.float64)
            w = w + 16.0 * lstart / 4.8e12
        return shard_rows(w, self.ell_of_row, world)

    def power_win_mix_upper_packed(self, lo, hi, packed, div2Lp1=False, interchange_NN=False):
        """Blocks with l <= L of the columns [lo, hi), written into `packed` (the whole upper-packed buffer)."""
        _lib.check(self.lib.sfb_power_win_mix_upper_packed_dev(self._cmix, self.alm.data_ptr(), int(div2Lp1),
                                                               int(interchange_NN), lo, hi, packed.data_ptr(),
                                                               self._stream()))
        return packed

    def power_win_mix_lshard(self, lo, hi, packed=None, rows=None, div2Lp1=False, interchange_NN=False, packed_slab=None):
        """"L-shaped" shard of the L-block aligned index range [lo, hi): the blocks with l <= L of the columns [lo, hi)
        in upper-packed storage (`packed`, the whole packed buffer) plus, mirrored locally from them, the rows [lo, hi) of
        the part of M below the block diagonal: rows[r, j - lo] = M[j, r] for l(r) < l(j)  (tensor of shape (hi, hi - lo)).
        Over the ranges of `packed_shard_ranges` every element of M is formed exactly once (no redundant flops, no
        exchange): element (i, i') lives where max(l_i, l_i') falls."""
        torch = _torch()
        off = self.packed_offsets()
        if packed_slab is not None:
            # only this range's slab [off[lo], off[hi]) of the packed buffer exists: address it through a virtual base
            assert packed_slab.is_contiguous() and packed_slab.numel() >= int(off[hi] - off[lo])
            base = packed_slab.data_ptr() - 8 * int(off[lo])
        else:
            base = packed.data_ptr()
        if rows is None:
            rows = torch.zeros((hi, hi - lo), dtype=torch.float64, device=self.device)
        assert rows.is_contiguous() and rows.shape[0] >= hi and rows.shape[1] == hi - lo
        if hi > lo:
            _lib.check(self.lib.sfb_power_win_mix_upper_packed_dev(self._cmix, self.alm.data_ptr(), int(div2Lp1),
                                                                   int(interchange_NN), lo, hi, base, self._stream()))
            _lib.check(self.lib.sfb_cmix_mirror_rows_dev(self._cmix, base, lo, hi, int(div2Lp1), int(interchange_NN),
                                                         rows.data_ptr(), hi - lo, self._stream()))
        return (packed if packed_slab is None else packed_slab), rows

    def lshard_assemble_host(self, packed, rows_of_range, ranges):
        """Full matrix (numpy, M[i, j]) from the pieces of `power_win_mix_lshard` — for checks."""
        off = self.packed_offsets()
        n = self.nout
        ell = np.asarray(self.ell_of_row)
        P = packed.cpu().numpy()
        M = np.full((n, n), np.nan)
        lend = np.zeros(n, dtype=np.int64)      # rend(j): end of j's own l-block
        cuts = np.concatenate([[0], np.flatnonzero(np.diff(ell)) + 1, [n]])
        for a, b in zip(cuts[:-1], cuts[1:]):
            lend[a:b] = b
        for j in range(n):
            M[:lend[j], j] = P[off[j]:off[j] + lend[j]]
        for (lo, hi), rows in zip(ranges, rows_of_range):
            R = rows.cpu().numpy()
            for j in range(lo, hi):
                below = ell[:hi] < ell[j]                 # columns r with l(r) < l(j)
                M[j, :hi][below] = R[:hi, j - lo][below]
        return M

    def unpack_mirror(self, packed, full=None, div2Lp1=False, interchange_NN=False):
        torch = _torch()
        if full is None:
            full = torch.empty((self.nout, self.nout), dtype=torch.float64, device=self.device)
        _lib.check(self.lib.sfb_cmix_unpack_mirror_dev(self._cmix, packed.data_ptr(), int(div2Lp1), int(interchange_NN),
                                                       full.data_ptr(), self.nout, self._stream()))
        return full

    def unpack_mirror_pull(self, peer_packed, ranges, full=None, div2Lp1=False, interchange_NN=False):
        """Fused exchange + expansion: every column of the packed matrix is read from its owner's buffer over NVLink
        (peer-mapped memory) while the local full matrix is written.  The caller orders this after all ranks' block
        kernels (`stream_barrier`)."""
        torch = _torch()
        if full is None:
            full = torch.empty((self.nout, self.nout), dtype=torch.float64, device=self.device)
        pb = peer_packed
        bases = (C.c_void_p * pb.world)(*pb.rank_ptrs)
        bounds = np.asarray([r[0] for r in ranges] + [ranges[-1][1]], dtype=np.int64)
        _lib.check(self.lib.sfb_cmix_unpack_mirror_peers_dev(self._cmix, bases, bounds.ctypes.data_as(C.c_void_p), pb.world,
                                                             pb.rank, int(div2Lp1), int(interchange_NN), full.data_ptr(),
                                                             self.nout, self._stream()))
        return full

    def power_win_mix_allgather_packed(self, d_win, full=None, packed=None, group=None, div2Lp1=False,
                                       interchange_NN=False, peer_packed=None):
        """Multi-GPU path for the auto-correlation matrix: stage 1 shell-sharded; stage 2+3 sharded over the column
        index (L,N,N') forming only the blocks with l <= L, written in upper-packed storage (half the bytes of the
        matrix); then every rank expands the packed matrix into the full one (direct half + mirror image).
        With `peer_packed` (a PeerBuffer of packed_offsets()[-1] doubles) the exchange is fused into the expansion
        kernel, which pulls each column from its owner over NVLink; otherwise the packed slabs are all-gathered in
        place with NCCL first.  Returns (full, ranges); `full` holds Mᵀ in C order."""
        torch = _torch()
        import torch.distributed as dist
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        off = self.packed_offsets()
        if peer_packed is not None:
            packed = peer_packed.tensor
        elif packed is None:
            packed = torch.empty(int(off[-1]), dtype=torch.float64, device=self.device)
        self.calc_wr_lm_sharded(d_win, group)
        ranges = self.packed_shard_ranges(world)
        lo, hi = ranges[rank]
        if hi > lo:
            self.power_win_mix_upper_packed(lo, hi, packed, div2Lp1=div2Lp1, interchange_NN=interchange_NN)
        if peer_packed is not None and world > 1:
            stream_barrier(self, group)      # every rank's slab is complete before anyone pulls it
            full = self.unpack_mirror_pull(peer_packed, ranges, full, div2Lp1=div2Lp1, interchange_NN=interchange_NN)
            stream_barrier(self, group)      # ... and stays untouched until every rank has pulled it
            return full, ranges
        if world > 1:
            allgather_packed_slabs(packed, [(int(off[l]), int(off[h])) for l, h in ranges], group)
        full = self.unpack_mirror(packed, full, div2Lp1=div2Lp1, interchange_NN=interchange_NN)
        return full, ranges

    def power_win_mix_sharded(self, d_win, group=None, gather=True, **kw):
        """Row-sharded coupling matrix over the ranks of `group`; with gather=True every rank returns the
        full matrix as a (nout, nout) tensor holding Mᵀ in C order (= M in Julia's column-major order)."""
        torch = _torch()
        import torch.distributed as dist
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.calc_wr_lm_sharded(d_win, group)
        ranges = shard_rows(self.row_costs, self.ell_of_row, world)
        lo, hi = ranges[rank]
        slab = self.power_win_mix_rows(lo, hi, **kw)  # (nout, hi-lo)
        if world == 1 or not gather:
            return slab, ranges
        return gather_row_slabs(slab, ranges, self.nout, group), ranges


def allgather_col_slabs(full, ranges, group=None):
    """In-place uneven all-gather of contiguous column slabs: rank g owns full[lo_g:hi_g] (rows of the C-ordered
    tensor = columns of Julia's matrix); grouped NCCL send/recv fills every other slab."""
    import torch.distributed as dist
    rank = dist.get_rank(group)
    lo, hi = ranges[rank]
    ops = []
    for g, (l, h) in enumerate(ranges):
        if g == rank or h <= l:
            continue
        if hi > lo:
            ops.append(dist.P2POp(dist.isend, full[lo:hi], g, group))
        ops.append(dist.P2POp(dist.irecv, full[l:h], g, group))
    if not ops:
        return
    for w in dist.batch_isend_irecv(ops):
        w.wait()


def stream_barrier(pipe, group=None):
    """Stream-ordered cross-rank barrier (a one-element NCCL all-reduce on the current stream, no host sync)."""
    import torch.distributed as dist
    if getattr(pipe, "_bar", None) is None:
        pipe._bar = _torch().zeros(1, dtype=_torch().float32, device=pipe.device)
    dist.all_reduce(pipe._bar, group=group)


def allgather_packed_slabs(packed, spans, group=None):
    """In-place uneven all-gather of the 1-D packed buffer: rank g owns packed[spans[g][0]:spans[g][1]]."""
    import torch.distributed as dist
    rank = dist.get_rank(group)
    lo, hi = spans[rank]
    ops = []
    for g, (l, h) in enumerate(spans):
        if g == rank or h <= l:
            continue
        if hi > lo:
            ops.append(dist.P2POp(dist.isend, packed[lo:hi], g, group))
        ops.append(dist.P2POp(dist.irecv, packed[l:h], g, group))
    if not ops:
        return
    for w in dist.batch_isend_irecv(ops):
        w.wait()


class _DevArray:
    """__cuda_array_interface__ holder so torch can view a library-owned device buffer without a copy."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


class PeerBuffer:
    """A device buffer of `nelem` doubles on every rank, IPC-mapped into all peers of the node: rank_ptrs[g] is the
    address of rank g's buffer in THIS process (own buffer included)."""

    def __init__(self, nelem, group=None):
        torch = _torch()
        import torch.distributed as dist
        self.lib = _lib.load()
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        if self.world > 8:
            raise ValueError("PeerBuffer supports at most 8 GPUs of one node")
        self.ptr = C.c_void_p()
        handle = C.create_string_buffer(64)
        _lib.check(self.lib.sfb_ipc_alloc(C.byref(self.ptr), 8 * int(nelem), handle))
        handles = [bytes(handle.raw)]
        if self.world > 1:
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(handle.raw), group=group)
        self.rank_ptrs, self._opened = [], []
        for g, h in enumerate(handles):
            if g == self.rank:
                self.rank_ptrs.append(self.ptr.value)
                continue
            q = C.c_void_p()
            _lib.check(self.lib.sfb_ipc_open(C.create_string_buffer(h, 64), C.byref(q)))
            self._opened.append(q)
            self.rank_ptrs.append(q.value)
        self.tensor = torch.as_tensor(_DevArray(self.ptr.value, (int(nelem),)),
                                      device=torch.device("cuda", torch.cuda.current_device()))

    def close(self):
        import torch.distributed as dist
        if self.ptr:
            _torch().cuda.synchronize()
            live = self.world > 1 and dist.is_initialized()     # at interpreter exit the group may be gone already
            if live:
                dist.barrier(self.group)
            for q in self._opened:
                self.lib.sfb_ipc_close(q)
            if live:
                dist.barrier(self.group)
            self.lib.sfb_ipc_free(self.ptr)
            self.ptr = C.c_void_p()
            self._opened = []


def gather_row_slabs(slab, ranges, nout, group=None):
    """All-gather variable-height row slabs into the full matrix.  `slab` has shape (nout, rows_g); the result
    has shape (nout, nout) with result[i', i] = M[i, i'] (C order), i.e. Julia's column-major M."""
    torch = _torch()
    import torch.distributed as dist
    world = dist.get_world_size(group)
    hmax = max(hi - lo for lo, hi in ranges)
    send = slab if slab.shape[1] == hmax else torch.nn.functional.pad(slab, (0, hmax - slab.shape[1]))
    send = send.contiguous()
    recv = torch.empty(world * nout * hmax, dtype=slab.dtype, device=slab.device)
    dist.all_gather_into_tensor(recv, send.view(-1), group=group)
    recv = recv.view(world, nout, hmax)
    full = torch.empty((nout, nout), dtype=slab.dtype, device=slab.device)
    for g, (lo, hi) in enumerate(ranges):
        full[:, lo:hi] = recv[g, :, :hi - lo]
    return full

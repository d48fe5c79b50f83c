"""Row-sharded block vectors and SpMM across GPUs.

One process per GPU (torch.distributed; NCCL on GPUs, gloo in the CPU tests).  The Morton-ordered block rows of the
(connection) Laplacian are split into contiguous ranges, one per rank.  Every rank keeps its rows of the matrix with
column indices renumbered to [local | halo], where the halo is the set of neighbour nodes owned by other ranks, sorted
by global id (hence grouped by owner).  One SpMM = pack the rows the peers need (a gather kernel) -> point-to-point
exchange straight into the halo slice of the extended block vector -> local SpMM.  Gram matrices / norms are summed
with all_reduce, after which every rank solves the identical small projected problem.

``HaloPlan`` is pure index logic on torch tensors (CPU or CUDA) so it is covered by the world_size-2 gloo tests.
"""

import numpy as np
import torch
import torch.distributed as dist


def partition_rows(indptr_host, world):
    """Contiguous row ranges balanced by stored blocks.  Returns bounds (world+1,) int64."""
    indptr_host = np.asarray(indptr_host, dtype=np.int64)
    n = indptr_host.size - 1
    nnz = indptr_host[-1]
    targets = (np.arange(1, world) * nnz) // world
    cuts = np.searchsorted(indptr_host, targets, side="left")
    bounds = np.concatenate([[0], cuts, [n]]).astype(np.int64)
    return np.maximum.accumulate(bounds)


class HaloPlan:
    """Local view of a row-sharded CSR pattern."""

    def __init__(self, indptr, indices, bounds, rank, group=None):
        """indptr / indices: the GLOBAL pattern (replicated) as torch tensors on the working device."""
        self.rank, self.world, self.group = int(rank), len(bounds) - 1, group
        self.bounds = [int(b) for b in bounds]
        r0, r1 = self.bounds[rank], self.bounds[rank + 1]
        self.r0, self.r1, self.n_loc = r0, r1, r1 - r0
        dev = indptr.device
        e0, e1 = int(indptr[r0].item()), int(indptr[r1].item())
        self.e0, self.e1 = e0, e1
        cols = indices[e0:e1].to(torch.int64)
        self.indptr_loc = (indptr[r0:r1 + 1] - e0).to(torch.int32).contiguous()
        outside = (cols < r0) | (cols >= r1)
        halo = torch.unique(cols[outside])                    # sorted ascending => grouped by owner rank
        self.halo_ids = halo
        self.n_halo = int(halo.numel())
        # renumber: local -> col - r0 ; halo -> n_loc + position in the sorted halo list
        pos = torch.searchsorted(halo, cols.clamp(min=0)) if self.n_halo else torch.zeros_like(cols)
        self.indices_loc = torch.where(outside, pos + self.n_loc, cols - r0).to(torch.int32).contiguous()
        # block rows that reference at least one halo column (they need the exchanged data; all others do not)
        rowlen = (self.indptr_loc[1:] - self.indptr_loc[:-1]).to(torch.int64)
        rowid = torch.repeat_interleave(torch.arange(self.n_loc, device=dev), rowlen)
        self.boundary_rows = torch.unique(rowid[outside]).to(torch.int32).contiguous()
        # keep every row's columns SORTED after the renumbering (halo nodes owned by lower ranks moved behind the local
        # ones): the row-group merge / MMA plans walk sorted column lists.  entry_perm reorders per-entry values alike.
        key = rowid * (self.n_loc + self.n_halo + 1) + self.indices_loc.to(torch.int64)
        self.entry_perm = torch.argsort(key)
        self.indices_loc = self.indices_loc[self.entry_perm].contiguous()
        # how many halo nodes come from each peer
        b = torch.tensor(self.bounds, dtype=torch.int64, device=dev)
        owner = torch.searchsorted(b, halo, right=True) - 1 if self.n_halo else halo
        self.recv_counts = [int((owner == q).sum().item()) for q in range(self.world)]
        self.send_counts = [0] * self.world
        self.send_ids = torch.zeros(0, dtype=torch.int32, device=dev)

    def local_values(self, vals_global):
        """This rank's per-entry values, in the (column-sorted) entry order of indices_loc."""
        return vals_global[self.e0:self.e1][self.entry_perm.to(vals_global.device)].contiguous()

    def exchange_requests(self):
        """Tell every owner which of its nodes this rank needs (one-time setup).  After this, ``send_ids`` lists the
        LOCAL node ids to pack, grouped by destination rank, and ``send_counts`` their number per peer."""
        dev = self.halo_ids.device
        if self.world == 1:
            return
        counts = torch.tensor(self.recv_counts, dtype=torch.int64, device=dev)
        gathered = [torch.zeros_like(counts) for _ in range(self.world)]
        dist.all_gather(gathered, counts, group=self.group)
        self.send_counts = [int(gathered[q][self.rank].item()) for q in range(self.world)]
        recv_bufs = [torch.empty(self.send_counts[q], dtype=torch.int64, device=dev) for q in range(self.world)]
        ops, off = [], 0
        for q in range(self.world):
            c = self.recv_counts[q]
            if q != self.rank and c:
                ops.append(dist.P2POp(dist.isend, self.halo_ids[off:off + c].contiguous(), q, group=self.group))
            off += c
            if q != self.rank and self.send_counts[q]:
                ops.append(dist.P2POp(dist.irecv, recv_bufs[q], q, group=self.group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        ids = torch.cat([recv_bufs[q] for q in range(self.world)]) if sum(self.send_counts) else torch.zeros(0, dtype=torch.int64, device=dev)
        self.send_ids = (ids - self.r0).to(torch.int32).contiguous()

    def exchange_async(self, ext, pack_fn, d):
        """Start the halo exchange (NCCL) and return a handle to wait on; None when there is nothing to do."""
        if self.world == 1 or (self.n_halo == 0 and self.send_ids.numel() == 0) or not ext.is_cuda:
            self.exchange(ext, pack_fn, d)
            return None
        sendbuf = pack_fn(ext[: self.n_loc * d], self.send_ids, d)
        return dist.all_to_all_single(ext[self.n_loc * d:], sendbuf,
                                      output_split_sizes=[c * d for c in self.recv_counts],
                                      input_split_sizes=[c * d for c in self.send_counts], group=self.group, async_op=True)

    def exchange(self, ext, pack_fn, d):
        """Fill the halo slice ext[n_loc*d:] of the extended block vector (rows x b, contiguous) from the peers.
        pack_fn(ext_local, send_ids, d) -> (n_send*d, b) contiguous tensor of the rows to ship."""
        if self.world == 1 or (self.n_halo == 0 and self.send_ids.numel() == 0):
            return
        sendbuf = pack_fn(ext[: self.n_loc * d], self.send_ids, d)
        if ext.is_cuda:
            # NCCL: one all-to-all with uneven splits (rows per peer) instead of 2*(P-1) point-to-point ops
            dist.all_to_all_single(ext[self.n_loc * d:], sendbuf,
                                   output_split_sizes=[c * d for c in self.recv_counts],
                                   input_split_sizes=[c * d for c in self.send_counts], group=self.group)
            return
        ops, soff, roff = [], 0, self.n_loc * d
        for q in range(self.world):
            sc, rc = self.send_counts[q] * d, self.recv_counts[q] * d
            if q != self.rank and sc:
                ops.append(dist.P2POp(dist.isend, sendbuf[soff:soff + sc], q, group=self.group))
            if q != self.rank and rc:
                ops.append(dist.P2POp(dist.irecv, ext[roff:roff + rc], q, group=self.group))
            soff += sc
            roff += rc
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()


class Comm:
    """Sum-reductions used by the eigensolver; a no-op for a single rank."""

    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0

    def allreduce_(self, t):
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def allgather_rows(self, local, counts):
        """Concatenate row blocks of every rank (counts = rows per rank)."""
        if self.world == 1:
            return local
        cmax = int(max(counts))
        pad = torch.zeros((cmax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        pad[: local.shape[0]] = local
        outs = [torch.empty_like(pad) for _ in counts]          # equal sizes: works on NCCL and gloo alike
        dist.all_gather(outs, pad, group=self.group)
        return torch.cat([o[: int(c)] for o, c in zip(outs, counts)], 0)


class _DevMem:
    """A raw device allocation exposed to torch through __cuda_array_interface__ (no copy, no ownership)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 3}


class PeerHalo:
    """Halo exchange over NVLink peer memory for one (ShardedBsr, ncols): the three extended block vectors and a flag array
    are allocated with rvgp_ipc_alloc, the 64-byte CUDA IPC handles are exchanged once through torch.distributed, and from
    then on a whole Chebyshev recurrence is ONE C call (rvgp_halo_cheb_filter_f64: signal / wait / pull / SpMM kernels per
    degree, no NCCL, no host round trip).  See csrc/halo.cu for the protocol."""

    def __init__(self, op, ncols, timeout_ms=20000):
        import ctypes
        from ._cabi import get_handle
        plan, d = op.plan, op.d
        import os
        self.op, self.ncols = op, int(ncols)
        dev = op.indptr.device
        self.h = h = get_handle(dev.index)
        lib = h.lib
        world, rank, group = plan.world, plan.rank, plan.group
        row_bytes = d * self.ncols * 8
        ebytes = max(256, (plan.n_loc + plan.n_halo) * row_bytes)
        # The shared buffers come from a process-wide POOL (_IpcPool below): cudaMalloc + cudaIpcOpenMemHandle on every peer
        # + the handle exchange cost tens of milliseconds per set, and a sharded create_data_object needs up to three sets.
        # A set (three extended block vectors + the flag array) is reused by later operators whenever it is large enough;
        # close() only returns it to the pool.  Sizes are agreed collectively (max over ranks) so that every rank takes the
        # same pool decision.
        self._set = _IpcPool.acquire(h, ebytes, world, rank, group, dev)
        self._released = False
        ebytes = self._set["cap"]
        self.E_ptrs = self._set["E_ptrs"]                             # [slot][rank] -> device address
        self.flag_ptrs = self._set["flag_ptrs"]
        nrows_ext = (plan.n_loc + plan.n_halo) * d
        self.E = [torch.as_tensor(_DevMem(self.E_ptrs[s][rank], ebytes), device=dev).view(torch.float64)[: nrows_ext * self.ncols]
                  .view(nrows_ext, self.ncols) for s in range(3)]
        # pull tables: address of every halo node's row in its owner's buffer
        b = torch.tensor(plan.bounds, dtype=torch.int64, device=dev)
        halo = plan.halo_ids
        owner = (torch.searchsorted(b, halo, right=True) - 1) if plan.n_halo else halo
        self.pull = []
        for s in range(3):
            base = torch.tensor(self.E_ptrs[s], dtype=torch.int64, device=dev)
            src = (base[owner] + (halo - b[owner]) * row_bytes) if plan.n_halo else torch.zeros(1, dtype=torch.int64, device=dev)
            self.pull.append(src.contiguous())
        peers = [q for q in range(world) if q != rank and (plan.recv_counts[q] > 0 or plan.send_counts[q] > 0)]
        self.peers = peers
        self.peer_slots = torch.tensor([self.flag_ptrs[q] + 8 * rank for q in peers] or [0], dtype=torch.int64, device=dev)
        self.wait_idx = torch.tensor(peers or [0], dtype=torch.int32, device=dev)
        self.err = torch.zeros(1, dtype=torch.int32, device=dev)
        self.epoch = self._set["epoch"]           # epochs only grow, also across the users of a pooled set (the flags persist)
        L = op.local
        mp = L.mma if (d == 2 and getattr(L, "mma", None) is not None and self.ncols % 16 == 0) else None
        pat = L.mma_pattern if (d == 1 and getattr(L, "mma_pattern", None) is not None and self.ncols % 32 == 0) else None
        self.kernel = "mma_native" if mp is not None else ("mma_native_pattern" if pat is not None else "gather")

        class Ctx(ctypes.Structure):
            _fields_ = [("n_loc", ctypes.c_int32), ("n_halo", ctypes.c_int32), ("d", ctypes.c_int32), ("ncols", ctypes.c_int32),
                        ("indptr", ctypes.c_void_p), ("indices", ctypes.c_void_p), ("vals", ctypes.c_void_p),
                        ("kptr", ctypes.c_void_p), ("kcols", ctypes.c_void_p), ("afrag", ctypes.c_void_p),
                        ("rotc", ctypes.c_int32), ("n_peers", ctypes.c_int32),
                        ("E", ctypes.c_void_p * 3), ("pull_src", ctypes.c_void_p * 3),
                        ("flags", ctypes.c_void_p), ("peer_slots", ctypes.c_void_p), ("wait_idx", ctypes.c_void_p),
                        ("err", ctypes.c_void_p), ("timeout_ms", ctypes.c_int32), ("n_interior", ctypes.c_int32),
                        ("glist_interior", ctypes.c_void_p), ("glist_boundary", ctypes.c_void_p),
                        ("n_boundary", ctypes.c_int32), ("reserved", ctypes.c_int32)]

        c = Ctx()
        c.n_loc, c.n_halo, c.d, c.ncols = plan.n_loc, plan.n_halo, d, self.ncols
        c.indptr, c.indices = L.indptr.data_ptr(), L.indices.data_ptr()
        c.vals = L.vals.data_ptr() if L.vals is not None else None
        if mp is not None:
            rotc = mp["rotc"]
            c.kptr = mp["kptr"].data_ptr()
            c.kcols = (mp["kcols_c"] if rotc else mp["kcols"]).data_ptr()
            c.afrag = (mp["afrag_c"] if rotc else mp["afrag"]).data_ptr()
            c.rotc = int(rotc)
        elif pat is not None:            # scalar unit-weight Laplacian on the MMA kernel (spmm_mma.cu AMODE 2): no values streamed
            c.kptr, c.kcols, c.afrag, c.rotc = pat["kptr"].data_ptr(), pat["kcols"].data_ptr(), pat["deg"].data_ptr(), 2
        if (mp is not None or pat is not None) and os.environ.get("RVGP_HALO_OVERLAP", "1") != "0":
            # interior / boundary row-group lists (groups of 4 block rows, the unit of the MMA kernel): the interior launch of a
            # step overlaps the neighbours' skew and the halo pull (csrc/halo.cu, halo_step)
            ng = (plan.n_loc + 3) // 4
            isb = torch.zeros(ng, dtype=torch.bool, device=dev)
            if plan.boundary_rows.numel():
                isb[(plan.boundary_rows.to(torch.int64) // 4)] = True
            self._g_int = torch.nonzero(~isb).reshape(-1).to(torch.int32).contiguous()
            self._g_bnd = torch.nonzero(isb).reshape(-1).to(torch.int32).contiguous()
            if self._g_int.numel() == 0:
                self._g_int = torch.zeros(1, dtype=torch.int32, device=dev)[:0]
            c.glist_interior = self._g_int.data_ptr() if self._g_int.numel() else self._g_bnd.data_ptr()
            c.glist_boundary = self._g_bnd.data_ptr() if self._g_bnd.numel() else self._g_int.data_ptr()
            c.n_interior, c.n_boundary = int(self._g_int.numel()), int(self._g_bnd.numel())
        c.n_peers = len(peers)
        for s in range(3):
            c.E[s] = self.E_ptrs[s][rank]
            c.pull_src[s] = self.pull[s].data_ptr()
        c.flags = self.flag_ptrs[rank]
        c.peer_slots, c.wait_idx, c.err = self.peer_slots.data_ptr(), self.wait_idx.data_ptr(), self.err.data_ptr()
        c.timeout_ms = int(timeout_ms)
        self.ctx = c
        self._cp = ctypes.pointer(c)
        torch.cuda.synchronize(dev)
        dist.barrier(group=group)                    # every rank has opened every handle before the first pull

    def cheb_filter(self, Vp, degree, lo_spec, lo_cut, hi):
        from ._cabi import I64, U64
        self.h.sync_stream()
        self.h.call("rvgp_halo_cheb_filter_f64", self._cp, U64(self.epoch), Vp, I64(Vp.stride(0)), int(degree), float(lo_spec),
                    float(lo_cut), float(hi))
        self.epoch += int(degree) + 1

    def spmm(self, X, Y):
        from ._cabi import I64, U64
        self.h.sync_stream()
        self.h.call("rvgp_halo_spmm_f64", self._cp, U64(self.epoch), X, I64(X.stride(0)), Y, I64(Y.stride(0)))
        self.epoch += 2

    def check(self):
        if int(self.err.item()) != 0:
            raise RuntimeError("peer halo exchange timed out waiting for a neighbour rank (rvgp_halo_ctx.err)")

    def close(self, collective=True):
        """Return the shared buffers to the pool (they stay mapped for the next operator; _IpcPool.release frees them).
        Every rank must call it at the same point of the program, like every other step of the SPMD pipeline."""
        if self._released:
            return
        self._released = True
        torch.cuda.synchronize(self.err.device)
        self._set["epoch"] = self.epoch + 2
        self._set["in_use"] = False


class _IpcPool:
    """Process-wide pool of CUDA-IPC shared buffer sets for PeerHalo.  One set = three extended block vectors of `cap` bytes
    + one flag array, allocated with rvgp_ipc_alloc on every rank and mapped into every peer (rvgp_ipc_open).  Sets are
    created collectively and in the same order on every rank, and `acquire` picks by a size agreed with one all-reduce(MAX),
    so all ranks always hold the same list and take the same decisions.  Memory is bounded by the largest operators seen;
    `release()` (collective; also run at interpreter exit, non-collectively) unmaps and frees everything."""
    sets = []

    @staticmethod
    def acquire(h, ebytes, world, rank, group, dev):
        t = torch.tensor([int(ebytes)], dtype=torch.int64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        need = int(t.item())
        for st in _IpcPool.sets:
            if not st["in_use"] and st["cap"] >= need and st["world"] == world and st["group"] is group:
                st["in_use"] = True
                return st
        st = _IpcPool._create(h, need, world, rank, group, dev)
        _IpcPool.sets.append(st)
        return st

    @staticmethod
    def _create(h, cap, world, rank, group, dev):
        import ctypes
        lib = h.lib
        sizes = [cap, cap, cap, 8 * max(world, 32)]
        own, opened = [], []
        # Failure-atomic: every rank runs the SAME collectives whatever fails locally.
        #   phase 1  local allocations only                     -> all_gather_object of (ok, handles)
        #   phase 2  open the peers' handles (only if all ok)   -> all_reduce(MIN) of ok
        # A failure on any rank makes every rank release what it holds and raise the same error.
        err, handles = None, []
        try:
            for nbytes in sizes:
                ptr = ctypes.c_void_p()
                hd = (ctypes.c_uint8 * 64)()
                rc = lib.rvgp_ipc_alloc(h._h, ctypes.c_int64(nbytes), ctypes.byref(ptr), hd)
                if rc != 0:
                    raise RuntimeError("rvgp_ipc_alloc: " + lib.rvgp_last_error(h._h).decode())
                own.append(ptr.value)
                handles.append(bytes(hd))
        except Exception as e:
            err = e
        gathered = [None] * world
        dist.all_gather_object(gathered, (err is None, handles), group=group)
        if not all(g[0] for g in gathered):
            _release_ipc(lib, h._h, opened, own)
            raise RuntimeError("peer halo setup failed on rank(s) %s%s" %
                               ([q for q, g in enumerate(gathered) if not g[0]], "" if err is None else ": %s" % (err,)))
        table = [[None] * world for _ in sizes]                      # [buffer][rank] -> device address
        try:
            for b_i in range(len(sizes)):
                for q in range(world):
                    if q == rank:
                        table[b_i][q] = own[b_i]
                        continue
                    pp = ctypes.c_void_p()
                    hq = (ctypes.c_uint8 * 64).from_buffer_copy(gathered[q][1][b_i])
                    rc = lib.rvgp_ipc_open(h._h, hq, ctypes.byref(pp))
                    if rc != 0:
                        raise RuntimeError("rvgp_ipc_open: " + lib.rvgp_last_error(h._h).decode())
                    opened.append(pp.value)
                    table[b_i][q] = pp.value
        except Exception as e:
            err = e
        ok = torch.tensor([0 if err is not None else 1], dtype=torch.int32, device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok.item()) != 1:
            for p in opened:
                lib.rvgp_ipc_close(h._h, ctypes.c_void_p(p))
            del opened[:]
            dist.barrier(group=group)                                 # peers unmap before owners free
            _release_ipc(lib, h._h, opened, own)
            raise RuntimeError("peer halo setup failed (cudaIpcOpenMemHandle)%s" % ("" if err is None else ": %s" % (err,)))
        torch.cuda.synchronize(dev)
        dist.barrier(group=group)                                     # every rank has opened every handle before the first pull
        return dict(cap=cap, world=world, group=group, E_ptrs=table[:3], flag_ptrs=table[3], epoch=1, in_use=True,
                    lib=lib, hh=h._h, opened=opened, own=own)

    @staticmethod
    def release(collective=True):
        """Unmap and free every pooled set.  collective=True: all ranks call it together (peers unmap, barrier, owners free)."""
        import ctypes
        sets, _IpcPool.sets = _IpcPool.sets, []
        if not sets:
            return
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        for st in sets:
            for p in st["opened"]:
                st["lib"].rvgp_ipc_close(st["hh"], ctypes.c_void_p(p))
            del st["opened"][:]
        if collective and dist.is_available() and dist.is_initialized():
            dist.barrier()
        for st in sets:
            _release_ipc(st["lib"], st["hh"], st["opened"], st["own"])


def release_ipc_pool(collective=True):
    """Free the pooled CUDA-IPC halo buffers of this process (call on every rank, e.g. before destroy_process_group)."""
    _IpcPool.release(collective=collective)


import atexit as _atexit
_atexit.register(lambda: _IpcPool.release(collective=False))


def _release_ipc(lib, hh, opened, own):
    import ctypes
    for p in opened:
        lib.rvgp_ipc_close(hh, ctypes.c_void_p(p))
    for p in own:
        lib.rvgp_ipc_free(hh, ctypes.c_void_p(p))
    del opened[:], own[:]


class ShardedBsr:
    """This rank's rows of a BSR matrix plus the halo machinery; duck-types eigensolver.BsrMatrix."""

    def __init__(self, plan, d, vals_loc, comm):
        """vals_loc: the blocks of this rank's entries in the plan's entry order, i.e. global_vals[plan.e0:plan.e1][plan.entry_perm]
        (``plan.local_values`` does that)."""
        from .eigensolver import BsrMatrix
        self.plan, self.d, self.comm = plan, int(d), comm
        self.local = BsrMatrix(plan.n_loc, d, plan.indptr_loc, plan.indices_loc, vals_loc)
        self.nbrows = plan.n_loc
        self.nrows = plan.n_loc * self.d                  # local rows
        self.ext_rows = (plan.n_loc + plan.n_halo) * self.d
        self.indptr, self.indices, self.vals = plan.indptr_loc, plan.indices_loc, vals_loc
        self.nnzb = self.local.nnzb
        self._bufs = {}
        self.row_offset = plan.r0 * self.d
        self._peer = {}                 # ncols -> PeerHalo (or None when CUDA IPC is unavailable)
        self.mma = None

    peer_halo = True                    # halo exchange by our own kernels over NVLink peer memory (csrc/halo.cu)

    def close(self, collective=True):
        """Release every PeerHalo of this operator (raw cudaMalloc / CUDA-IPC memory torch does not own) and the scratch
        block vectors.  Collective by default: call it on every rank at the same point."""
        for nc in sorted(self._peer):
            ph = self._peer[nc]
            if ph is not None:
                ph.close(collective=collective)
        self._peer.clear()
        self._bufs.clear()

    @property
    def spmm_kernel_name(self):
        ph = [p for p in self._peer.values() if p is not None]
        if ph:
            return ph[0].kernel + " + peer-memory halo"
        return "gather + nccl halo"

    def enable_mma(self, on=True, h=None):
        """FP64-MMA SpMM for the local rows (d == 2: compact-rotation plan; d == 1 pattern mode: L (x) I_2 plan); only used by
        the peer-memory path."""
        if self.d == 2:
            self.mma = self.local.enable_mma(on, h=h)
        elif self.d == 1 and self.vals is None:
            self.mma = self.local.enable_mma_pattern(on, h=h)
        else:
            self.mma = None
        return self.mma

    @property
    def mma_pattern(self):
        return self.local.mma_pattern

    def _peer_for(self, ncols):
        """PeerHalo for this panel width, built collectively on first use (all ranks take the same path)."""
        import os
        if not (self.peer_halo and self.indptr.is_cuda and self.plan.world > 1 and ncols % 2 == 0
                and os.environ.get("RVGP_PEER_HALO", "1") != "0"):
            return None
        if ncols not in self._peer:
            ok = torch.ones(1, dtype=torch.int32, device=self.indptr.device)
            ph = None
            try:
                ph = PeerHalo(self, ncols)
            except Exception as e:                      # no CUDA IPC between these devices: keep the NCCL exchange
                print("rvgp_b200: peer-memory halo exchange unavailable (%s); using NCCL all_to_all" % (e,))
                ok.zero_()
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.plan.group)
            self._peer[ncols] = ph if int(ok.item()) == 1 else None
        return self._peer[ncols]

    def spmm_bytes(self, ncols, fused=False):
        return self.local.spmm_bytes(ncols, fused)

    def _pack(self, ext_local, send_ids, d):
        n_send = int(send_ids.numel())
        key = ("pack", ext_local.shape[1])
        if key not in self._bufs:
            self._bufs[key] = torch.empty((n_send * d, ext_local.shape[1]), dtype=ext_local.dtype, device=ext_local.device)
        out = self._bufs[key]
        if n_send:
            from ._cabi import get_handle, I64
            h = get_handle(ext_local.device.index)
            h.call("rvgp_gather_rows_f64", I64(n_send * d), int(ext_local.shape[1]), ext_local, I64(ext_local.stride(0)),
                   send_ids, int(d), out, I64(out.stride(0)))
        return out

    def _ext(self, ncols, slot):
        key = (ncols, slot)
        if key not in self._bufs:
            self._bufs[key] = torch.zeros((self.ext_rows, ncols), dtype=torch.float64, device=self.indptr.device)
        return self._bufs[key]

    def _spmm_overlapped(self, E_in, E_out, alpha=1.0, beta=0.0, gamma=0.0, W=None, h=None):
        """E_out[:local] = alpha*A*E_in + beta*E_in + gamma*W with the halo exchange of E_in hidden behind the SpMM:
        the full launch runs while the exchange is in flight (its boundary rows read stale halo data and are
        discarded), then only the boundary rows are recomputed once the halo has landed."""
        from ._cabi import get_handle, I64
        h = h or get_handle(E_in.device.index)
        ncols = E_in.shape[1]
        nb = int(self.plan.boundary_rows.numel())
        can_overlap = (self.overlap and nb > 0 and ncols % 2 == 0 and nb * 4 < self.plan.n_loc)
        if not can_overlap:
            self.plan.exchange(E_in, self._pack, self.d)
            self.local.spmm(E_in, E_out[: self.nrows], alpha=alpha, beta=beta, gamma=gamma, W=W, h=h)
            return
        work = self.plan.exchange_async(E_in, self._pack, self.d)
        self.local.spmm(E_in, E_out[: self.nrows], alpha=alpha, beta=beta, gamma=gamma, W=W, h=h)
        if work is not None:
            work.wait()
        L = self.local
        Yv = E_out[: self.nrows]
        h.call("rvgp_bsr_spmm_rows_f64", L.nbrows, L.d, L.indptr, L.indices, L.vals, E_in, I64(E_in.stride(0)), W,
               I64(W.stride(0) if W is not None else 0), Yv, I64(Yv.stride(0)), int(ncols), float(alpha), float(beta),
               float(gamma), self.plan.boundary_rows, nb)

    overlap = True

    def matmat(self, X, out=None, h=None):
        if out is None:
            out = torch.empty_like(X)
        for c0 in range(0, X.shape[1], 64):
            c1 = min(X.shape[1], c0 + 64)
            ph = self._peer_for(c1 - c0) if (c1 - c0) == 64 else None
            if ph is not None:
                ph.spmm(X[:, c0:c1], out[:, c0:c1])
                if c1 == X.shape[1]:
                    ph.check()
                continue
            E = self._ext(c1 - c0, 0)
            E[: self.nrows].copy_(X[:, c0:c1])
            tmp = self._ext(c1 - c0, 1)
            self._spmm_overlapped(E, tmp, h=h)
            out[:, c0:c1].copy_(tmp[: self.nrows])
        return out

    def cheb_filter(self, Vp, w0, w1, ncols, degree, lo_spec, lo_cut, hi, h=None, w2=None):
        """Same recurrence as rvgp_cheb_filter_f64, one halo exchange + one fused SpMM launch per degree."""
        if degree <= 0:
            return
        ph = self._peer_for(ncols) if ncols >= 16 else None
        if ph is not None:
            ph.cheb_filter(Vp, degree, lo_spec, lo_cut, hi)
            return
        E = [self._ext(ncols, s) for s in range(3)]
        E[0][: self.nrows].copy_(Vp)
        e, c = 0.5 * (hi - lo_cut), 0.5 * (hi + lo_cut)
        sigma1 = e / (lo_spec - c)
        tau, sigma = 2.0 / sigma1, sigma1
        self._spmm_overlapped(E[0], E[1], alpha=sigma1 / e, beta=-c * sigma1 / e, h=h)
        prev, cur = 0, 1
        for _ in range(2, degree + 1):
            sn = 1.0 / (tau - sigma)
            nxt = 3 - prev - cur
            self._spmm_overlapped(E[cur], E[nxt], alpha=2.0 * sn / e, beta=-2.0 * sn * c / e, gamma=-sigma * sn,
                                  W=E[prev][: self.nrows], h=h)
            sigma, prev, cur = sn, cur, nxt
        Vp.copy_(E[cur][: self.nrows])

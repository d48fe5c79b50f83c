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
        # how many halo nodes come from each peer
        b = torch.tensor(self.bounds, dtype=torch.int64, device=dev)
        owner = torch.searchsorted(b, halo, right=True) - 1 if self.n_halo else halo
        self.recv_counts = [int((owner == q).sum().item()) for q in range(self.world)]
        self.send_counts = [0] * self.world
        self.send_ids = torch.zeros(0, dtype=torch.int32, device=dev)

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


class ShardedBsr:
    """This rank's rows of a BSR matrix plus the halo machinery; duck-types eigensolver.BsrMatrix."""

    def __init__(self, plan, d, vals_loc, comm):
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
            E = self._ext(c1 - c0, 0)
            E[: self.nrows].copy_(X[:, c0:c1])
            tmp = self._ext(c1 - c0, 1)
            self._spmm_overlapped(E, tmp, h=h)
            out[:, c0:c1].copy_(tmp[: self.nrows])
        return out

    def cheb_filter(self, Vp, w0, w1, ncols, degree, lo_spec, lo_cut, hi, h=None):
        """Same recurrence as rvgp_cheb_filter_f64, one halo exchange + one fused SpMM launch per degree."""
        if degree <= 0:
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

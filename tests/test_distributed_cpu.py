"""world_size-2 (and 3) gloo tests on CPU for the N>1 host logic: row partition, halo plan, request exchange and
halo exchange of the row-sharded SpMM (rvgp_b200/distributed.py).  The local SpMM itself is emulated with SciPy here;
the CUDA kernel is the same one the single-GPU tests cover."""
import os
import socket

import numpy as np
import pytest
import scipy.sparse as sp
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.conftest import load_golden


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, case, d, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from rvgp_b200.distributed import HaloPlan, partition_rows, Comm
        g = load_golden(case)
        n = g["X"].shape[0]
        indptr, indices = g["indptr"], g["indices"]
        if d == 1:
            A = sp.csr_matrix((g["L_data"], g["L_indices"], g["L_indptr"]), shape=(n, n))
            blocks = g["L_data"].reshape(-1, 1, 1)
        else:
            A = sp.bsr_matrix((g["Lc_data"], g["Lc_indices"], g["Lc_indptr"]), shape=(n * d, n * d)).tocsr()
            blocks = g["Lc_data"]
        bounds = partition_rows(indptr, world)
        assert bounds[0] == 0 and bounds[-1] == n and np.all(np.diff(bounds) > 0)
        plan = HaloPlan(torch.from_numpy(indptr), torch.from_numpy(indices), bounds, rank)
        plan.exchange_requests()
        assert sum(plan.recv_counts) == plan.n_halo and plan.recv_counts[rank] == 0
        # every local row keeps sorted columns after the [local | halo] renumbering (the MMA / merge plans rely on it)
        ipl, ixl = plan.indptr_loc.numpy(), plan.indices_loc.numpy()
        assert all(np.all(np.diff(ixl[ipl[i]:ipl[i + 1]]) > 0) for i in range(plan.n_loc))
        # local matrix in [local | halo] numbering
        Aloc = sp.bsr_matrix((plan.local_values(torch.from_numpy(blocks)).numpy(), plan.indices_loc.numpy(), plan.indptr_loc.numpy()),
                             shape=(plan.n_loc * d, (plan.n_loc + plan.n_halo) * d))
        rng = np.random.default_rng(0)
        X = rng.normal(size=(n * d, 5))                               # same global vector on every rank
        ext = torch.zeros(((plan.n_loc + plan.n_halo) * d, 5), dtype=torch.float64)
        ext[: plan.n_loc * d] = torch.from_numpy(X[plan.r0 * d: plan.r1 * d])

        def pack(ext_local, send_ids, dd):
            rows = (send_ids.to(torch.int64)[:, None] * dd + torch.arange(dd)[None, :]).reshape(-1)
            return ext_local[rows].contiguous()

        plan.exchange(ext, pack, d)
        halo_rows = (plan.halo_ids[:, None] * d + torch.arange(d)[None, :]).reshape(-1).numpy()
        assert np.array_equal(ext[plan.n_loc * d:].numpy(), X[halo_rows])      # halo filled with the owners' rows
        Yloc = Aloc @ ext.numpy()
        ref = (A @ X)[plan.r0 * d: plan.r1 * d]
        assert np.abs(Yloc - ref).max() < 1e-12
        # reductions: Gram of row shards == global Gram
        comm = Comm()
        G = torch.from_numpy(X[plan.r0 * d: plan.r1 * d].T @ X[plan.r0 * d: plan.r1 * d])
        comm.allreduce_(G)
        assert np.abs(G.numpy() - X.T @ X).max() < 1e-9
        full = comm.allgather_rows(torch.from_numpy(Yloc), [int(bounds[r + 1] - bounds[r]) * d for r in range(world)])
        assert np.abs(full.numpy() - A @ X).max() < 1e-12
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, "FAIL: " + traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,case,d", [(2, "torus_n600_k20", 2), (2, "sphere_n2000_k50", 1), (3, "flat3torus_R6_n900_k24", 3)])
def test_sharded_halo_spmm_gloo(world, case, d):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, case, d, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res


def test_partition_balanced():
    from rvgp_b200.distributed import partition_rows
    g = load_golden("sphere_n2000_k50")
    b = partition_rows(g["indptr"], 8)
    nnz = np.diff(g["indptr"][b])
    assert len(b) == 9 and nnz.max() - nnz.min() <= 2 * np.diff(g["indptr"]).max()


def _worker_gp_pieces(rank, world, port, q):
    """The host-side pieces of the round-2 sharding: contiguous rank shares of an index list (row-sharded GP fit / transform),
    the int32 row all-gather of the sharded geodesic stage, the bitwise OR of the per-rank flag words, and the data-parallel
    identity the sharded fit rests on: sum over ranks of Phi_r^T [Phi_r y_r] == Phi^T [Phi y]."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from rvgp_b200.distributed import Comm
        from rvgp_b200.main import _rank_chunk
        comm = Comm()
        for n_items in (0, 1, 7, 100, 101):
            a, b, counts = _rank_chunk(n_items, comm)
            assert sum(counts) == n_items and counts[rank] == b - a and a == sum(counts[:rank])
            assert max(counts) - min(counts) <= 1
        n, K = 103, 5
        seq = torch.arange(n * (K + 1), dtype=torch.int32).reshape(n, K + 1)
        s0, s1 = (n * rank) // world, (n * (rank + 1)) // world
        rows = [(n * (r + 1)) // world - (n * r) // world for r in range(world)]
        full = comm.allgather_rows(seq[s0:s1].clone(), rows)
        assert full.dtype == torch.int32 and torch.equal(full, seq)
        flags = torch.tensor([1 if rank == 0 else 4], dtype=torch.int32)
        bits = torch.stack([(flags >> i) & 1 for i in range(3)]).reshape(3).to(torch.int32)
        comm.allreduce_(bits)
        word = int(((bits[0] > 0).to(torch.int32) | ((bits[1] > 0).to(torch.int32) << 1) | ((bits[2] > 0).to(torch.int32) << 2)).item())
        assert word == (5 if world > 1 else 1)
        rng = np.random.default_rng(0)
        Phi, y = rng.normal(size=(200, 6)), rng.normal(size=(200, 1))
        a, b, _ = _rank_chunk(200, comm)
        XY = np.hstack([Phi[a:b], y[a:b]])
        G = torch.from_numpy(XY.T @ XY)
        comm.allreduce_(G)
        XYf = np.hstack([Phi, y])
        assert np.abs(G.numpy() - XYf.T @ XYf).max() < 1e-10
        q.put((rank, "ok"))
    except Exception:  # pragma: no cover
        import traceback
        q.put((rank, "FAIL: " + traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_gp_and_geodesic_host_pieces_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_gp_pieces, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res


def _worker_krylov(rank, world, port, paired, fast_accept, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["RVGP_KRYLOV_FAST_ACCEPT"] = fast_accept
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import scipy.sparse.linalg as spla
        from tests import fake_cabi, test_krylov_host as TK
        from rvgp_b200.distributed import Comm
        from rvgp_b200.krylov import krylov_eigenpairs
        mpatch = pytest.MonkeyPatch()
        fake_cabi.install_eigensolver(mpatch)
        L, hi = TK._laplacian(1500, seed=1)
        n = L.shape[0]
        if paired:
            # complex-Hermitian "magnetic" Laplacian in real 2x2-block storage (the construction of test_krylov_host.py)
            rng = np.random.default_rng(0)
            Lc = sp.coo_matrix(L)
            phi, data = {}, np.empty(Lc.nnz, dtype=np.complex128)
            for e, (i, j, v) in enumerate(zip(Lc.row, Lc.col, Lc.data)):
                if i == j:
                    data[e] = v
                else:
                    key = (min(i, j), max(i, j))
                    if key not in phi:
                        phi[key] = rng.uniform(-0.3, 0.3)
                    data[e] = v * np.exp(1j * (phi[key] if i < j else -phi[key]))
            Hc = sp.csr_matrix((data, (Lc.row, Lc.col)), shape=(n, n))
            M = sp.bmat([[Hc.real, -Hc.imag], [Hc.imag, Hc.real]]).tocsr()
            perm = np.empty(2 * n, dtype=np.int64)
            perm[0::2] = np.arange(n); perm[1::2] = n + np.arange(n)
            M = M[perm][:, perm].tocsr()
            d = 2
        else:
            M, d = L, 1
        k = 30
        ref = np.sort(spla.eigsh(M, k=70, which="SM", return_eigenvectors=False))
        nodes = [n * r // world for r in range(world + 1)]          # node boundaries: a complex entry (2 rows) never straddles ranks
        r0, r1 = nodes[rank] * d, nodes[rank + 1] * d
        counts = [(nodes[r + 1] - nodes[r]) * d for r in range(world)]
        comm = Comm()
        A = fake_cabi.FakeShardedBsr(M, r0, r1, counts, comm, d=d)
        st = {}
        evals, evecs = krylov_eigenpairs(A, k, hi, cut=1.05 * ref[int(1.5 * k) + 16], lam_k=ref[k - 1], paired=paired,
                                         block=16, stats=st, comm=comm)
        ev = evals.numpy()
        U = comm.allgather_rows(evecs.contiguous(), counts).numpy()  # global eigenvectors, row blocks in rank order
        assert np.allclose(ev, ref[:k], rtol=1e-8, atol=1e-10), np.abs(ev - ref[:k]).max()
        res = np.linalg.norm(M @ U - U * ev, axis=0)
        assert res.max() <= 1e-12 * hi * 1.0001, res.max()
        assert np.abs(U.T @ U - np.eye(k)).max() < 1e-9
        assert st["converged"] and st["world"] == world
        assert st["final_rr_outer"] == (0 if fast_accept == "1" else 1), st["final_rr_outer"]
        mpatch.undo()
        q.put((rank, "ok"))
    except Exception:  # pragma: no cover
        import traceback
        q.put((rank, "FAIL: " + traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("paired,fast_accept", [(False, "0"), (True, "0"), (False, "1"), (True, "1")])
def test_sharded_krylov_eigensolver_gloo(paired, fast_accept):
    """The filtered block Lanczos solver with its block vectors ROW-SHARDED over 2 ranks (all-reduced Gram matrices, Rayleigh
    quotients and residuals; real and paired algebra; both hand-over paths) against ARPACK on the global matrix."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_krylov, args=(r, world, port, paired, fast_accept, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res

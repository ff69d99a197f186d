"""world_size-2 and -3 gloo tests (CPU) of the multi-GPU host logic in pydca_b200/distributed.py:
shard arithmetic and parameter-slice layout, both branches of the reweighting counts logic,
the gather of the parameter slices, the replicated-x decomposition + all-reduce of the
fallback path, and the OWNER-SHARDED decomposition (which conditional's half of which block
goes to which rank) with a point-to-point exchange standing in for the peer-memory stores.
The per-shard numbers come from the CPU oracle (no CUDA here)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions():
    from pydca_b200.distributed import shard_range
    for total in (1, 7, 8, 150, 500, 5000):
        for world in (1, 2, 3, 4, 8):
            parts = [shard_range(total, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == total
            assert all(parts[k][1] == parts[k + 1][0] for k in range(world - 1))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1


def test_balanced_site_bounds_cover_and_order():
    from pydca_b200.distributed import balanced_site_bounds
    for L in (2, 3, 9, 150, 500):
        for world in (1, 2, 3, 4, 8):
            if world > L:
                with pytest.raises(ValueError):
                    balanced_site_bounds(L, world)
                continue
            for rho in (0.0, 0.5, 2.0):
                b = balanced_site_bounds(L, world, rho)
                assert len(b) == world + 1 and b[0] == 0 and b[-1] == L
                assert all(b[k + 1] > b[k] for k in range(world))
    # rho = 0 is the plain balanced split; rho > 0 gives low ranks (which own more blocks) fewer sites
    sizes = np.diff(balanced_site_bounds(150, 8, 0.0))
    assert sizes.max() - sizes.min() <= 1
    b = balanced_site_bounds(150, 8, 0.5)
    assert b[1] - b[0] < b[8] - b[7]
    assert balanced_site_bounds(150, 8) == balanced_site_bounds(150, 8, 0.0)


def test_shard_layout_partitions_the_parameter_vector():
    """dca_shard_layout (C, host-only) against the reference's index arithmetic
    (plmdca_numerics.cpp:226, :338-342): the owners' slices tile the vector exactly, and
    scatter_full_vector inverts local_slice."""
    from pydca_b200 import _lib
    from pydca_b200.distributed import balanced_site_bounds, scatter_full_vector
    for L, q in ((7, 5), (150, 21), (33, 21)):
        P = L * q + L * (L - 1) // 2 * q * q
        pair = lambda i, j: L * (L - 1) // 2 - (L - i) * (L - i - 1) // 2 + j - i - 1
        for world in (1, 2, 3, min(8, L)):
            b = balanced_site_bounds(L, world)
            lay = [_lib.shard_layout(L, q, b[r], b[r + 1]) for r in range(world)]
            assert lay[0][0] == 0 and lay[0][2] == L * q
            for r in range(world):
                fo, fc, co, cc = lay[r]
                assert fo == b[r] * q and fc == (b[r + 1] - b[r]) * q
                if b[r] + 1 < L:
                    assert co == L * q + pair(b[r], b[r] + 1) * q * q
                nblocks = sum(L - 1 - i for i in range(b[r], b[r + 1]))
                assert cc == nblocks * q * q
                if r + 1 < world:
                    assert fo + fc == lay[r + 1][0] and co + cc == lay[r + 1][2]
            assert lay[-1][0] + lay[-1][1] == L * q and lay[-1][2] + lay[-1][3] == P
            x = np.arange(P, dtype=np.float32)
            slices = [np.concatenate([x[fo:fo + fc], x[co:co + cc]]) for fo, fc, co, cc in lay]
            assert (scatter_full_vector(slices, lay, P) == x).all()
    with pytest.raises(_lib.DcaBackendError):
        _lib.shard_layout(10, 21, 4, 4)


def test_owner_sharded_index_model():
    """Index arithmetic of the fused exchange kernels (csrc/plm_eval.cu: plm_unpack_push_kernel,
    plm_export_kernel, plm_assemble_owner_kernel) restated in numpy for several shard layouts:
    after the pushes every rank's W holds exactly the columns of the full symmetric coupling
    matrix it needs, and after the exports every owner assembles
    dJ_ij(a,b) = G[(j,b)][(i,a)] + G[(i,a)][(j,b)] for its blocks from its own G and its R buffer."""
    from pydca_b200 import _lib
    rng = np.random.default_rng(5)
    for L, q, bounds in ((7, 5, [0, 3, 7]), (9, 3, [0, 1, 8, 9]), (12, 4, [0, 3, 6, 9, 12]), (6, 5, [0, 6])):
        Lq, qq = L * q, q * q
        world = len(bounds) - 1
        pair = lambda i, j: L * (L - 1) // 2 - (L - i) * (L - i - 1) // 2 + j - i - 1
        owner = lambda s: max(r for r in range(world) if bounds[r] <= s)
        P = Lq + L * (L - 1) // 2 * qq
        x = rng.normal(size=P)
        Jfull = np.zeros((Lq, Lq))                                   # [(i,a)][(j,b)] = J_ij(a,b), symmetric
        for i in range(L):
            for j in range(i + 1, L):
                blk = x[Lq + pair(i, j) * qq: Lq + (pair(i, j) + 1) * qq].reshape(q, q)
                Jfull[i * q:(i + 1) * q, j * q:(j + 1) * q] = blk
                Jfull[j * q:(j + 1) * q, i * q:(i + 1) * q] = blk.T
        # --- unpack + push: every rank walks its LOCAL slice
        W = [np.full((Lq, Lq), np.nan) for _ in range(world)]
        for r in range(world):
            lo, hi = bounds[r], bounds[r + 1]
            fo, fc, co, cc = _lib.shard_layout(L, q, lo, hi)
            xl = np.concatenate([x[fo:fo + fc], x[co:co + cc]])
            nq, pair0 = (hi - lo) * q, pair(lo, lo + 1) if lo + 1 < L else L * (L - 1) // 2
            for i in range(lo, hi):
                for j in range(i + 1, L):
                    blk = xl[nq + (pair(i, j) - pair0) * qq: nq + (pair(i, j) - pair0 + 1) * qq].reshape(q, q)
                    W[r][j * q:(j + 1) * q, i * q:(i + 1) * q] = blk.T             # my columns (i,a), rows (j,b)
                    W[owner(j)][i * q:(i + 1) * q, j * q:(j + 1) * q] = blk        # owner(j)'s columns (j,b), rows (i,a)
        for r in range(world):
            lo, hi = bounds[r], bounds[r + 1]
            for i in range(lo, hi):
                cols = slice(i * q, (i + 1) * q)
                rows = np.r_[0:i * q, (i + 1) * q:Lq]                               # every j != i
                assert np.array_equal(W[r][rows, cols], Jfull[rows, cols])
        # --- export + assemble: G[(j,b)][(i,a)] is produced by owner(i) (columns of site i)
        G = rng.normal(size=(Lq, Lq))
        max_sites = max(bounds[r + 1] - bounds[r] for r in range(world))
        R = [np.full((max_sites * q, Lq), np.nan) for _ in range(world)]
        for r in range(world):
            lo, hi = bounds[r], bounds[r + 1]
            for row in range(lo * q):                                               # rows (i,a) with i below my range
                o = owner(row // q)
                R[o][row - bounds[o] * q, lo * q:hi * q] = G[row, lo * q:hi * q]
        for r in range(world):
            lo, hi = bounds[r], bounds[r + 1]
            for i in range(lo, hi):
                for j in range(i + 1, L):
                    u = G[j * q:(j + 1) * q, i * q:(i + 1) * q].T                   # my conditional: G[(j,b)][(i,a)]
                    if j < hi:
                        v = G[i * q:(i + 1) * q, j * q:(j + 1) * q]
                    else:
                        v = R[r][(i - lo) * q:(i - lo + 1) * q, j * q:(j + 1) * q]
                    want = G[j * q:(j + 1) * q, i * q:(i + 1) * q].T + G[i * q:(i + 1) * q, j * q:(j + 1) * q]
                    assert np.array_equal(u + v, want)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import oracle
    from pydca_b200.distributed import allgather_counts, shard_range
    from pydca_b200.synthetic import synthetic_msa

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    P = oracle.port()
    P.set_num_threads(2)
    N, L, q = 240, 14, 21
    msa = synthetic_msa(N, L, q, seed=21)

    # row-sharded reweighting: local rows against all rows, then one collective
    lo, hi = shard_range(N, rank, world)
    full = P.reweight_counts(msa, 0.8)
    local = np.zeros(N, dtype=np.uint32)
    local[lo:hi] = full[lo:hi]           # what dca_weights(rows=(lo,hi)) returns on the GPU path
    got = allgather_counts(local, (lo, hi), N)
    assert (got == full).all()

    # sharded_counts(): shallow alignments are computed redundantly (no collective), deep ones are
    # row-sharded + summed; the device kernel is replaced by the oracle here (no CUDA on this box)
    from pydca_b200 import _lib, distributed as dd
    calls = []

    def fake_counts(m, seqid, compare_f64=False, rows=None):
        calls.append(rows)
        c = P.reweight_counts(m, seqid, f64=compare_f64)
        if rows is None:
            return c
        out = np.zeros_like(c)
        out[rows[0]:rows[1]] = c[rows[0]:rows[1]]
        return out
    def fake_part(m, seqid, part, nparts, compare_f64=False, want_ms=False):
        """numpy model of the symmetric kernel (csrc/reweight.cu, sym = 1): 128-row tiles dealt
        cyclically; the diagonal block counts rows only, tiles right of it count both rows."""
        calls.append(("part", part, nparts))
        n, l = m.shape
        thr = [k for k in range(l + 1) if np.float32(k) / np.float32(l) > np.float32(seqid)]
        tmin = thr[0] if thr else l + 1
        out = np.zeros(n, dtype=np.uint32)
        for i0 in range(128 * part, n, 128 * nparts):
            rows = m[i0:i0 + 128]
            ident = (rows[:, None, :] == m[None, i0:, :]).sum(axis=2)        # j >= i0 only
            hit = ident >= tmin
            out[i0:i0 + rows.shape[0]] += hit.sum(axis=1).astype(np.uint32)
            out[i0 + 128:] += hit[:, 128:].sum(axis=0).astype(np.uint32)     # mirrored counts
        return out
    real, real_part = _lib.sequence_counts, _lib.sequence_counts_part
    _lib.sequence_counts, _lib.sequence_counts_part = fake_counts, fake_part
    try:
        assert (dd.sharded_counts(msa, 0.8) == full).all() and calls == [None]
        assert (dd.sharded_counts(msa, 0.8, force_shard=True, symmetric=False) == full).all() and calls[-1] == (lo, hi)
        assert (dd.sharded_counts(msa, 0.8, force_shard=True) == full).all() and calls[-1] == ("part", rank, world)
        big = synthetic_msa(700, 9, q, seed=5)                               # 6 row tiles over `world` parts
        assert (dd.sharded_counts(big, 0.8, force_shard=True) == P.reweight_counts(big, 0.8)).all()
    finally:
        _lib.sequence_counts, _lib.sequence_counts_part = real, real_part

    # site-sharded objective: each rank's data term, all-reduce, regulariser added once
    w = P.plm_weights(full)
    lam = 0.2 * (L - 1)
    rng = np.random.default_rng(4)
    x = (P.plm_init(msa, w, q) + rng.normal(0, 0.1, L * q + L * (L - 1) // 2 * q * q)).astype(np.float32)
    f_reg, g_reg = P.plm_gradient(msa, w, q, lam, lam, x, sites=(0, 0))
    slo, shi = shard_range(L, rank, world)
    f_loc, g_loc = P.plm_gradient(msa, w, q, lam, lam, x, sites=(slo, shi))
    g = torch.from_numpy((g_loc.astype(np.float64) - g_reg).copy())
    s = torch.tensor([f_loc - f_reg], dtype=torch.float64)
    dist.all_reduce(g)
    dist.all_reduce(s)
    f_full, g_full = P.plm_gradient_f64(msa, w, q, lam, lam, x)
    g_tot = g.numpy() + g_reg
    f_tot = float(s[0]) + f_reg
    assert np.abs(g_tot - g_full).max() <= 1e-5 * np.abs(g_full).max()
    assert abs(f_tot - f_full) <= 5e-5 * abs(f_full)
    # owner-sharded parameters: every rank holds its slice; ShardedPlm.gather() rebuilds the
    # full vector with one all_gather_into_tensor (the GPU path does the same over NCCL)
    from pydca_b200 import _lib
    from pydca_b200.distributed import ShardedPlm, balanced_site_bounds
    bounds = balanced_site_bounds(L, world)
    sp = ShardedPlm.__new__(ShardedPlm)
    sp.group, sp.rank, sp.world, sp.bounds = None, rank, world, bounds
    sp.layouts = [_lib.shard_layout(L, q, bounds[r], bounds[r + 1]) for r in range(world)]

    class _Obj:
        P = x.size
        Ploc = sp.layouts[rank][1] + sp.layouts[rank][3]
    sp.obj = _Obj()
    fo, fc, co, cc = sp.layouts[rank]
    mine = np.concatenate([x[fo:fo + fc], x[co:co + cc]])
    assert (sp.gather(mine) == x).all()
    # owner-sharded evaluation, decomposition only (the CUDA kernels do this over peer memory):
    # rank r evaluates the conditionals of its sites; what they contribute to blocks J_ij owned by
    # another rank (i outside my range, j inside) is "exported" to owner(i); every owner adds what
    # it receives to its own part and the regulariser of its slice -> its slice of the full gradient
    nq = L * q
    qq = q * q

    def block_range(i_lo, i_hi):          # coupling elements of the blocks (i, j > i), i in [i_lo, i_hi)
        pair = lambda i: L * (L - 1) // 2 - (L - i) * (L - i - 1) // 2
        return nq + pair(i_lo) * qq, nq + pair(i_hi) * qq
    slo, shi = bounds[rank], bounds[rank + 1]
    f_mine, g_mine = P.plm_gradient(msa, w, q, lam, lam, x, sites=(slo, shi))
    data = g_mine.astype(np.float64) - g_reg                  # data term of my sites' conditionals
    sends = []
    for r in range(world):
        c0, c1 = block_range(bounds[r], bounds[r + 1])
        part = data[c0:c1].copy()
        if r == rank:
            part[:] = 0.0                                     # my own blocks stay here
        sends.append(torch.from_numpy(part))
    recvs = [torch.zeros(block_range(slo, shi)[1] - block_range(slo, shi)[0], dtype=torch.float64)
             for _ in range(world)]
    # all-to-all by point-to-point exchanges (gloo has no all_to_all for uneven CPU tensors everywhere)
    for r in range(world):
        if r == rank:
            continue
        if rank < r:
            dist.send(sends[r], dst=r); dist.recv(recvs[r], src=r)
        else:
            dist.recv(recvs[r], src=r); dist.send(sends[r], dst=r)
    c0, c1 = block_range(slo, shi)
    own_c = data[c0:c1] + sum(t.numpy() for t in recvs) + g_reg[c0:c1]
    own_f = data[slo * q:shi * q] + g_reg[slo * q:shi * q]
    lay = sp.layouts[rank]
    assert (lay[0], lay[1], lay[2], lay[3]) == (slo * q, (shi - slo) * q, c0, c1 - c0)
    want = np.concatenate([g_full[slo * q:shi * q], g_full[c0:c1]])
    got_slice = np.concatenate([own_f, own_c])
    assert np.abs(got_slice - want).max() <= 1e-5 * np.abs(g_full).max()
    # the fields of my sites never leave the rank; blocks below my range got nothing from me twice
    assert all(float(t.abs().sum()) == 0.0 for r, t in enumerate(recvs) if r < rank)
    # asymmetric estimator: every rank holds its sites' columns of the directed couplings; one
    # all-gather + symmetrisation (the oracle's stands in for the CUDA kernel) gives the reference vector
    from oracle import asym_oracle
    from pydca_b200.distributed import AsymmetricPlm
    rngA = np.random.default_rng(9)
    W_all = rngA.normal(size=(L * q, L * q)).astype(np.float32)
    h_all = rngA.normal(size=L * q).astype(np.float32)
    ap = AsymmetricPlm.__new__(AsymmetricPlm)
    ap.group, ap.rank, ap.world, ap.L, ap.q = None, rank, world, L, q
    ap.bounds = [shard_range(L, r, world)[0] for r in range(world)] + [L]
    alo, ahi = ap.bounds[rank], ap.bounds[rank + 1]
    ld_a = ap._ld(rank)
    mine_a = np.zeros((L * q + 1, ld_a), dtype=np.float32)
    mine_a[0, :(ahi - alo) * q] = h_all[alo * q:ahi * q]
    mine_a[1:, :(ahi - alo) * q] = W_all[:, alo * q:ahi * q]
    got_a = ap.gather(mine_a.ravel(), pack=lambda W, h, l_, q_: asym_oracle.symmetrise(h, W, l_, q_))
    assert np.abs(got_a - asym_oracle.symmetrise(h_all, W_all, L, q)).max() == 0.0
    dist.barrier()
    dist.destroy_process_group()
    open(os.path.join(out_dir, "ok%d" % rank), "w").write("ok")


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_collectives_and_owner_sharding(tmp_path, world):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / ("ok%d" % r)).exists() for r in range(world))

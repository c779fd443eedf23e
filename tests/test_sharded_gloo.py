"""World-size-2 test of the sharded path on CPU over gloo: routing by owner, all-to-all exchange, per-shard application,
SpMV slice all-gather.  The per-rank structure is the oracle (checker); the union of the shards must equal ONE global
oracle matrix that received the same batches."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleBackend:
    """Two oracle matrices per rank: one stands for the column-major shard, one for the row-major shard."""

    def __init__(self):
        from oracle import oracle as O
        self.O = O
        self.cm = O.Matrix(fill_mode=False)
        self.rm = O.Matrix(fill_mode=False)

    def set_batch(self, which, inkeys, partkeys, vals):
        ik, pk, v = inkeys.numpy(), partkeys.numpy(), vals.numpy()
        if which == 0:
            self.cm.set_batch_policy(ik, pk, v)      # rows = in-array keys, cols = partition keys
        else:
            self.rm.set_batch_policy(pk, ik, v)      # rows = partition keys, cols = in-array keys

    def spmv_range(self, trans, x, y_slice, key_lo, key_hi):
        M = self.cm if trans else self.rm
        m, n = M.size
        xs = x.numpy()
        ny = key_hi - 1
        if trans:
            xx = np.zeros(max(m, 1))
            xx[: min(m, len(xs))] = xs[: min(m, len(xs))]
            y = M.mul_dense(xx, ny, trans=True)
        else:
            xx = np.zeros(max(n, 1))
            xx[: min(n, len(xs))] = xs[: min(n, len(xs))]
            y = M.mul_dense(xx, ny)
        y_slice.copy_(torch.from_numpy(y[key_lo - 1:key_hi - 1].copy()))


def _worker(rank, world, port, m, n, seed, q, skew=False):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import dsa_b200  # noqa: F401
        from dsa_b200.sharded import ShardedMatrix, even_splitters, owner_of, sampled_splitters
        from oracle import oracle as O

        rng = np.random.default_rng(seed)  # same stream on both ranks: every rank knows the whole global batch

        def draw(hi, size):
            if not skew:
                return rng.integers(1, hi + 1, size)
            return np.minimum(hi, np.floor(hi ** rng.random(size)).astype(np.int64))   # log-uniform: half the keys below sqrt(hi)

        if skew:   # splitters from sampled quantiles: each rank contributes a sample of the keys it is about to submit
            own = np.random.default_rng([seed, rank])
            rs = sampled_splitters(np.minimum(m, np.floor(m ** own.random(400)).astype(np.int64)), m, world)
            cs = sampled_splitters(np.minimum(n, np.floor(n ** own.random(400)).astype(np.int64)), n, world)
            assert rs[0] == 1 and rs[-1] == m + 1 and cs[0] == 1 and cs[-1] == n + 1 and len(rs) == world + 1
            assert rs[1] < m // 3 and cs[1] < n // 3, (rs, cs)          # the cut follows the mass, not the key range
            A = ShardedMatrix(m, n, OracleBackend(), row_split=rs, col_split=cs)
        else:
            A = ShardedMatrix(m, n, OracleBackend())
        G = O.Matrix(fill_mode=False)     # the global matrix, replicated as the checker
        rounds = []
        for rnd in range(4):
            nb = 3000
            I, J = draw(m, nb), draw(n, nb)
            V = np.where(rng.random(nb) < 0.3, 0.0, rng.integers(1, 9, nb).astype(float))
            # The global batch in its global op order is rank-major, arrival within a rank: rank r's share is the r-th contiguous
            # slice.  The same (i, j) is written by several ranks (keys are drawn from a small range): the last writer in THAT
            # order must win, exactly like the replicated checker that applies the whole batch in order.
            mine = slice(rank * nb // world, (rank + 1) * nb // world)
            share = (torch.from_numpy(I[mine].copy()), torch.from_numpy(J[mine].copy()), torch.from_numpy(V[mine].copy()))
            x = torch.from_numpy(rng.integers(0, 4, n).astype(float))
            xt = torch.from_numpy(rng.integers(0, 4, m).astype(float))
            rounds.append((I, J, V, share, x, xt))
        for s, (I, J, V, share, x, xt) in enumerate(rounds):
            G.set_batch_policy(I, J, V)
            A.set_batch(*share)
            y = A.spmv(x).numpy()
            assert np.array_equal(y, G.mul_dense(x.numpy(), m)), "A*x differs"
            yt = A.spmv(xt, trans=True).numpy()
            assert np.array_equal(yt, G.mul_dense(xt.numpy(), n, trans=True)), "A'*x differs"
        # shard contents == global contents restricted to the shard
        lo, hi = A.my_cols()
        rr, cc = np.meshgrid(np.arange(1, m + 1), np.arange(lo, hi), indexing="ij")
        assert np.array_equal(A.local.cm.get_many(rr.ravel(), cc.ravel()), G.get_many(rr.ravel(), cc.ravel()))
        lo, hi = A.my_rows()
        rr, cc = np.meshgrid(np.arange(lo, hi), np.arange(1, n + 1), indexing="ij")
        assert np.array_equal(A.local.rm.get_many(rr.ravel(), cc.ravel(), which=1), G.get_many(rr.ravel(), cc.ravel(), which=1))
        # every key landed on its owner only
        e = A.local.cm.export(0)
        ck = e["col_keys"][e["col_live"] == 1]
        assert np.all(owner_of(ck, A.col_split) == rank)
        if skew:   # balance of the routed work: equal key ranges would send ~85 % of a log-uniform key stream to rank 0
            allJ = np.concatenate([r[1] for r in rounds])
            share = np.bincount(owner_of(allJ, A.col_split), minlength=world) / len(allJ)
            even = np.bincount(owner_of(allJ, even_splitters(n, world)), minlength=world) / len(allJ)
            assert share.max() <= 0.65 and even.max() >= 0.8, (share, even)
        A.close()
        q.put((rank, "ok"))
    except Exception as ex:  # pragma: no cover
        import traceback
        q.put((rank, "FAIL: " + traceback.format_exc()))
        raise ex
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("skew", [False, True])
def test_sharded_two_ranks_gloo(skew):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 90, 70, 123, q, skew)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_splitters_and_owner():
    sys.path.insert(0, ROOT)
    import dsa_b200  # noqa: F401
    from dsa_b200.sharded import even_splitters, owner_of
    s = even_splitters(100_000, 8)
    assert s[0] == 1 and s[-1] == 100_001 and len(s) == 9
    assert owner_of([1, 12_500, 12_501, 100_000], s).tolist() == [0, 0, 1, 7]
    s = even_splitters(10, 4)
    assert owner_of(np.arange(1, 11), s).tolist() == [0, 0, 0, 1, 1, 1, 2, 2, 2, 3]

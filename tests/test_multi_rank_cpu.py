"""world_size-2 test of the N>1 path on CPU (gloo): the row partition and the shard replication of
spsparse_b200/dist.py, with the oracle standing in for the device kernels.  Property checked: the
ranks' row blocks of C, concatenated in rank order, equal the single-process product."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, m, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    from spsparse_b200 import gen
    from spsparse_b200.dist import replicate_csr_finish, replicate_csr_start, replicate_start, replicate_wait, row_range
    orc = O.port()
    r0, r1 = row_range(m, rank, world)
    a, b, w = gen.banded(5, m, r0, r1), gen.banded(6, m, r0, r1), gen.vector(7, m)
    Ac = orc.consolidate(O.Coo(*a), (0, 1))
    Bc = orc.consolidate(O.Coo(*b), (0, 1))
    local = [torch.from_numpy(Bc.idx[0].copy()), torch.from_numpy(Bc.idx[1].copy()), torch.from_numpy(Bc.val.copy())]
    fulls, works, sizes = replicate_start(local, rank, world)
    replicate_wait(works)
    Bf = O.Coo((m, m), [fulls[0].numpy(), fulls[1].numpy()], fulls[2].numpy(), None)
    # rows of the gathered B must be globally sorted: shards are row ranges in rank order
    key = Bf.idx[0].astype(np.int64) * m + Bf.idx[1]
    assert np.all(np.diff(key) > 0)
    # compressed-form replication (what bench.py uses): local row pointers + cols + vals
    lp = np.searchsorted(Bc.idx[0], np.arange(r0, r1)).astype(np.int32)
    st = replicate_csr_start(torch.from_numpy(lp), local[1], local[2], rank, world)
    ptr, cols, vals, total = replicate_csr_finish(st)
    assert total == Bf.n and np.array_equal(cols.numpy(), Bf.idx[1]) and np.array_equal(vals.numpy(), Bf.val)
    assert np.array_equal(ptr.numpy(), np.searchsorted(Bf.idx[0], np.arange(m + 1)))
    Cb = orc.multiply_mm(1.0, None, Ac, ".", O.Coo(w[0], w[1], w[2], (0,)), Bf, ".", None)
    # pruned replication: only the rows of B inside the hull of this block's column support are fetched.  The
    # transport (peer memory on the GPU) is emulated by slicing the gathered shards; plan and assembly are real.
    from spsparse_b200.dist import assemble_pruned_ptr, plan_pulls
    roffs, eoffs = st["roffs"], st["eoffs"]
    need = (int(Ac.idx[1].min()), int(Ac.idx[1].max()))
    pulls = plan_pulls(roffs, *need)
    assert 0 < sum(b_ - a_ for _, a_, b_ in pulls) <= (r1 - r0) + 4  # own shard + a halo of 2 rows either side
    gptr = ptr.numpy().astype(np.int64)
    chunks, e_los, counts, pc, pv = [], [], [], [], []
    for g, a_, b_ in pulls:
        lo, hi = gptr[roffs[g] + a_], gptr[roffs[g] + b_]
        chunks.append(torch.from_numpy((gptr[roffs[g] + a_:roffs[g] + b_] - eoffs[g]).astype(np.int32)))  # peer-local pointers
        e_los.append(int(lo - eoffs[g])); counts.append(int(hi - lo))
        pc.append(cols.numpy()[lo:hi]); pv.append(vals.numpy()[lo:hi])
    pptr, ptotal = assemble_pruned_ptr(m, roffs, pulls, chunks, e_los, counts, "cpu")
    pptr = pptr.numpy()
    assert ptotal == sum(counts) and np.all(np.diff(pptr) >= 0) and pptr[-1] == ptotal
    rows_p = np.repeat(np.arange(m), np.diff(pptr))
    Bp = O.Coo((m, m), [rows_p.astype(np.int32), np.concatenate(pc)], np.concatenate(pv), None)
    Cp = orc.multiply_mm(1.0, None, Ac, ".", O.Coo(w[0], w[1], w[2], (0,)), Bp, ".", None)
    assert np.array_equal(Cp.idx[0], Cb.idx[0]) and np.array_equal(Cp.idx[1], Cb.idx[1]) and np.array_equal(Cp.val, Cb.val)
    assert plan_pulls(roffs, 0, m - 1) == [(g, 0, int(roffs[g + 1] - roffs[g])) for g in range(world)]  # global support: whole shards
    np.savez(os.path.join(out_dir, f"c{rank}.npz"), i=Cb.idx[0], k=Cb.idx[1], v=Cb.val, sizes=np.array(sizes))
    dist.barrier()
    dist.destroy_process_group()


def test_row_partition_matches_single_process(tmp_path):
    sys.path.insert(0, ROOT)
    from oracle import oracle as O
    from spsparse_b200 import gen
    from spsparse_b200.dist import row_range
    m, world = 3001, 2
    assert [row_range(10, r, 3) for r in range(3)] == [(0, 3), (3, 6), (6, 10)]
    port = 29000 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, m, str(tmp_path)), nprocs=world, join=True)
    parts = [np.load(os.path.join(str(tmp_path), f"c{r}.npz")) for r in range(world)]
    got_i = np.concatenate([p["i"] for p in parts]); got_k = np.concatenate([p["k"] for p in parts])
    got_v = np.concatenate([p["v"] for p in parts])
    orc = O.port()
    a, b, w = gen.banded(5, m, 0, m), gen.banded(6, m, 0, m), gen.vector(7, m)
    # the full-size generator scrambles over all 5m slots, the shards scramble inside their own block:
    # different insertion orders, same matrices -- consolidate() makes them identical.
    want = orc.multiply_mm(1.0, None, O.Coo(*a), ".", O.Coo(w[0], w[1], w[2], (0,)), O.Coo(*b), ".", None)
    assert np.array_equal(got_i, want.idx[0]) and np.array_equal(got_k, want.idx[1])
    assert np.array_equal(got_v, want.val)
    assert int(parts[0]["sizes"].sum()) == 5 * m - 6  # B shards together hold the consolidated B


def test_pruned_plan_and_assembly_random():
    """plan_pulls + assemble_pruned_ptr on random shard sizes and need ranges (pure functions, no process group): every
    needed row comes back with exactly its entries, every other row is empty, and the pointer is monotone."""
    sys.path.insert(0, ROOT)
    from spsparse_b200.dist import assemble_pruned_ptr, plan_pulls
    rng = np.random.default_rng(123)
    for trial in range(200):
        world = int(rng.integers(1, 7))
        rows = [int(rng.integers(0, 40)) for _ in range(world)]          # some shards may be empty
        roffs = np.concatenate([[0], np.cumsum(rows)]).astype(np.int64)
        m = int(roffs[-1])
        if m == 0:
            continue
        row_len = rng.integers(0, 6, m)                                  # entries per global row
        # per-shard local pointers (with the sentinel) and entry payloads = (global row, running number)
        lptr, payload = [], []
        for g in range(world):
            ln = row_len[roffs[g]:roffs[g + 1]]
            lptr.append(np.concatenate([[0], np.cumsum(ln)]).astype(np.int32))
            payload.append(np.repeat(np.arange(roffs[g], roffs[g + 1]), ln))
        lo = int(rng.integers(0, m)); hi = int(rng.integers(lo, m))
        if trial % 10 == 0:
            lo, hi = 0, m - 1
        pulls = plan_pulls(roffs, lo, hi)
        assert sum(b - a for _, a, b in pulls) == hi - lo + 1
        chunks = [torch.from_numpy(lptr[g][a:b].copy()) for g, a, b in pulls]
        e_los = [int(lptr[g][a]) for g, a, b in pulls]
        counts = [int(lptr[g][b] - lptr[g][a]) for g, a, b in pulls]
        ptr, total = assemble_pruned_ptr(m, roffs, pulls, chunks, e_los, counts, "cpu")
        ptr = ptr.numpy().astype(np.int64)
        data = np.concatenate([payload[g][lptr[g][a]:lptr[g][b]] for g, a, b in pulls]) if pulls else np.empty(0, np.int64)
        assert total == len(data) == ptr[-1] and np.all(np.diff(ptr) >= 0) and ptr[0] == 0
        for r in range(m):
            got = data[ptr[r]:ptr[r + 1]]
            if lo <= r <= hi:
                assert len(got) == row_len[r] and np.all(got == r)
            else:
                assert len(got) == 0


def test_rowpart_pull_plan_emulation():
    """The plan arithmetic of k_rp_pull (tools/emulate_rowpart_pull.py restates it): for random shard sizes and hulls, every
    row of the hull comes back with exactly its entries through the absolute-row pointer, and nothing outside is written."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from emulate_rowpart_pull import pull
    rng = np.random.default_rng(321)
    for trial in range(300):
        world = int(rng.integers(1, 9))
        rows = [int(rng.integers(0, 30)) for _ in range(world)]
        row_lo = np.concatenate([[0], np.cumsum(rows)]).astype(np.int64)
        m = int(row_lo[-1])
        if m == 0:
            continue
        row_len = rng.integers(0, 5, m)
        ptrs, cols, vals = [], [], []
        for g in range(world):
            ln = row_len[row_lo[g]:row_lo[g + 1]]
            ptrs.append(np.concatenate([[0], np.cumsum(ln)]).astype(np.int64))
            cols.append(np.repeat(np.arange(row_lo[g], row_lo[g + 1]), ln).astype(np.int32))   # payload = the entry's global row
            vals.append(rng.random(int(ln.sum())))
        lo = int(rng.integers(0, m)); hi = int(rng.integers(lo, m))
        if trial % 7 == 0:
            lo, hi = 0, m - 1
        if trial % 11 == 0:
            lo, hi = 1, 0                                                                        # empty block of A: nothing fetched
        g_ptr, g_cols, g_vals, total = pull(row_lo, ptrs, cols, vals, lo, hi, m)
        if lo > hi:
            assert total == 0 and np.all(g_ptr == -1)
            continue
        assert total == int(row_len[lo:hi + 1].sum()) and g_ptr[lo] == 0 and g_ptr[hi + 1] == total
        assert np.all(g_ptr[:lo] == -1) and np.all(g_ptr[hi + 2:] == -1)
        for j in range(lo, hi + 1):
            seg = g_cols[g_ptr[j]:g_ptr[j + 1]]
            assert len(seg) == row_len[j] and np.all(seg == j)
        # the in-place variant, as every rank would run it: own shard untouched, halos in the slack around it
        from emulate_rowpart_pull import pull_in_place
        for rank in range(world):
            slack = int(rng.integers(0, 40))
            res = pull_in_place(row_lo, ptrs, cols, vals, lo, hi, m, rank, slack)
            need_lo = int(row_len[lo:max(lo, min(hi + 1, row_lo[rank]))].sum())
            need_hi = int(row_len[max(lo, min(hi + 1, row_lo[rank + 1])):hi + 1].sum())
            if need_lo > slack or need_hi > slack:
                assert res is None
                continue
            p2, bc, bv = res
            assert np.all(p2[:lo] == -1) and np.all(p2[hi + 2:] == -1)
            for j in range(lo, hi + 1):
                seg = bc[p2[j]:p2[j + 1]]
                assert len(seg) == row_len[j] and np.all(seg == j), (trial, rank, j)


def _handles_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from spsparse_b200.dist import HANDLE_BYTES, exchange_handles
    mine = bytes([(rank * 37 + i) % 256 for i in range(HANDLE_BYTES)])
    allh = exchange_handles(mine, rank, world)
    ok = all(allh[g * HANDLE_BYTES:(g + 1) * HANDLE_BYTES] == bytes([(g * 37 + i) % 256 for i in range(HANDLE_BYTES)]) for g in range(world))
    open(os.path.join(out_dir, f"h{rank}.txt"), "w").write("ok" if ok and len(allh) == world * HANDLE_BYTES else "bad")
    dist.barrier()
    dist.destroy_process_group()


def test_handle_exchange_gloo(tmp_path):
    """The one host-side step of the row-partitioned multiply behind the C ABI: every rank ends up with every rank's 64-byte
    buffer handle, in rank order (gloo, world size 2; on the GPU box the same call runs over NCCL)."""
    world = 2
    port = 31000 + os.getpid() % 2000
    mp.spawn(_handles_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert [open(os.path.join(str(tmp_path), f"h{r}.txt")).read() for r in range(world)] == ["ok"] * world

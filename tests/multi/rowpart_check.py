"""Multi-GPU parity check of the row-partitioned multiply behind the C ABI (spb_rowpart_*), run under torchrun:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi/rowpart_check.py

Every rank holds the raw entries of its rows of A and of B; the ranks' results, concatenated in rank order, must be the
CPU oracle's product of the whole matrices bit for bit (structure, order, values) -- with only the hull of rows fetched and
with all of B fetched, over several steps (the hand-shake counters), with scale vectors, duplicates, zeros, every policy,
and with a rank whose block is empty."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    import spsparse_b200 as sp
    from spsparse_b200 import gen
    from spsparse_b200.dist import RowPartition, row_range
    from oracle import oracle as O
    import _cases

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = sp.Context(local)
    orc = O.port()
    fails = []

    def block(c, lo, hi):
        pick = (c.idx[0] >= lo) & (c.idx[0] < hi)
        return O.Coo(c.shape, [c.idx[0][pick], c.idx[1][pick]], c.val[pick])

    def up(c):
        return sp.CooArray.from_host(ctx, c.shape, c.idx, c.val, c.sort_order) if c is not None else None

    def run_case(name, rp, m, A, B, si, sj, sk, Cst, pol, zn, fetch_all):
        r0, r1 = row_range(m, rank, world)               # this rank's rows of B (the inner index)
        a0, a1 = row_range(A.shape[0], rank, world)      # ... and its block of A's rows (any tiling of A's rows will do)
        hs = [up(x) for x in (si, block(A, a0, a1), sj, block(B, r0, r1), sk)]
        Cm, st = rp.multiply(Cst, hs[0], hs[1], hs[2], hs[3], hs[4], pol, zn, fetch_all)
        idx, val = Cm.to_host()
        for h in hs + [Cm]:
            if h is not None:
                h.free()
        parts = [None] * world
        dist.all_gather_object(parts, (idx[0], idx[1], val, int(st.rows_fetched), int(st.entries_fetched)))
        if rank == 0:
            gi = np.concatenate([p[0] for p in parts]); gk = np.concatenate([p[1] for p in parts]); gv = np.concatenate([p[2] for p in parts])
            want = orc.multiply_mm(Cst, si, A, ".", sj, B, ".", sk, pol, zn)
            ok = _cases.same_coo(O.Coo(want.shape, [gi, gk], gv), want)
            print(f"[rowpart_check] {name}: fetch_all={fetch_all} nnz={len(gv)} rows fetched per rank {[p[3] for p in parts]} -> {'ok' if ok else 'MISMATCH'}", flush=True)
            if not ok:
                fails.append(name)
            return [p[3] for p in parts]
        return None

    # 1. banded family, several steps on one partition object (the step counters), hull and fetch-all
    m = 20000
    A = O.Coo(*gen.banded(0x5EED0005, m, 0, m)); B = O.Coo(*gen.banded(0x5EED0015, m, 0, m))
    wv = gen.vector(0x5EED0025, m)
    w = O.Coo(wv[0], wv[1], wv[2], (0,))
    rp = RowPartition(ctx, rank, world, m, 5 * (m // world + 1) + 16)
    for step in range(4):
        fetched = run_case(f"banded step {step}", rp, m, A, B, None, w, None, 1.0, O.ADD, 0, fetch_all=(step == 2))
        if rank == 0 and step != 2 and world > 1:
            assert max(fetched) <= m // world + 1 + 4, fetched       # own shard + a 2-row halo on either side
        if rank == 0 and step == 2:
            assert min(fetched) == m, fetched
    rp.close()
    # 2. general matrices: random entries, duplicates, zeros / NaNs, scale vectors, every policy
    rng = np.random.default_rng(7)
    for s in range(6):
        m, nk = int(rng.integers(50, 3000)), int(rng.integers(1, 2000))
        flav = ["pos", "int", "mixed"][s % 3]
        na, nb = int(rng.integers(0, 8 * m)), int(rng.integers(1, 6 * m))
        A = O.Coo((m + 3, m), [rng.integers(0, m + 3, na), rng.integers(0, m, na)], _cases._values(rng, na, flav))
        B = O.Coo((m, nk), [rng.integers(0, m, nb), rng.integers(0, nk, nb)], _cases._values(rng, nb, flav))
        if s == 4:   # the last rank's block of A is empty
            lo, _ = row_range(m + 3, world - 1, world)
            A = block(A, 0, lo)
        si = _cases._sparse_vec(rng, m + 3, flav) if s % 2 else None
        sj = _cases._sparse_vec(rng, m, flav) if s % 3 else None
        sk = _cases._sparse_vec(rng, nk, flav) if s % 2 == 0 else None
        rp = RowPartition(ctx, rank, world, m, nb + 16)
        for fa in (False, True):
            run_case(f"general {s}", rp, m, A, B, si, sj, sk, [1.0, 2.5, -3.0][s % 3], _cases.POLICIES[s % 3], s % 2, fa)
        rp.close()
    ok = torch.tensor([0 if fails else 1], device="cuda")
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    ctx.close()
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("[rowpart_check] " + ("ALL OK" if not fails else f"FAILED: {fails}"), flush=True)
    sys.exit(0 if int(ok.item()) == 1 else 1)


if __name__ == "__main__":
    main()

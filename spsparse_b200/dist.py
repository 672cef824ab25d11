"""Multi-GPU plumbing for the row-partitioned multiply: A is split by contiguous
row ranges, B is sharded the same way (by its row = the inner index), consolidated locally, and then
replicated once on every rank -- or, when a rank's block of A only references an interval of inner indices
(plan_pulls), just that interval of B's rows is fetched.  Concatenating the ranks' consolidated blocks of C
in rank order IS the reference's row-major order, so nothing is exchanged after the multiply.

B travels in compressed form: per entry only its column and value (12 bytes instead of the 16 of COO),
plus one local row pointer per row; the row-index array is redundant once the pointers exist.

Device-agnostic on purpose: bench.py drives it with CUDA tensors over NCCL (NVLink), the CPU tests
with gloo.  The reference itself has no counterpart (it is single-process)."""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def row_range(m: int, rank: int, world: int) -> tuple[int, int]:
    """Rows [r0, r1) owned by `rank`."""
    return rank * m // world, (rank + 1) * m // world


HANDLE_BYTES = 64


def exchange_handles(handle: bytes, rank: int, world: int) -> bytes:
    """All-gathers one fixed-size opaque handle per rank (rank order) over torch.distributed -- gloo or NCCL alike."""
    assert len(handle) == HANDLE_BYTES
    if world == 1:
        return handle
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    mine = torch.tensor(list(handle), dtype=torch.uint8, device=dev)
    allh = torch.empty(world * HANDLE_BYTES, dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(allh, mine)
    return bytes(allh.cpu().tolist())


class RowPartition:
    """The row-partitioned multiply behind the C ABI (spb_rowpart_*, include/spsparse_b200.h): one process per GPU, rank r
    owns rows [r*m/world, (r+1)*m/world) of A and of B.  The library publishes / fetches the shards of B through
    peer-mapped memory itself; this class only carries the one-off exchange of the 64-byte buffer handles."""

    def __init__(self, ctx, rank: int, world: int, m: int, cap_entries: int):
        import ctypes as C
        from ._lib import check, vp
        self.ctx, self.rank, self.world, self.m = ctx, rank, world, m
        row_lo = (C.c_uint64 * (world + 1))(*[g * m // world for g in range(world + 1)])
        h = vp()
        check(ctx.lib.spb_rowpart_create(ctx.h, rank, world, row_lo, int(cap_entries), C.byref(h)))
        self.h = h
        buf = (C.c_ubyte * HANDLE_BYTES)()
        check(ctx.lib.spb_rowpart_handle(self.h, buf, HANDLE_BYTES))
        allh = exchange_handles(bytes(buf), rank, world)
        allbuf = (C.c_ubyte * len(allh)).from_buffer_copy(allh)
        check(ctx.lib.spb_rowpart_attach(self.h, allbuf, HANDLE_BYTES))
        if world > 1:
            dist.barrier()   # every rank has mapped every buffer before anybody's first step

    def multiply(self, Cst, scalei, A_block, scalej, B_shard, scalek, duplicate_policy=1, zero_nan=False, fetch_all=False):
        """-> (this rank's rows of C, RowpartStats).  Collective: every rank calls it once per step."""
        import ctypes as C
        from ._lib import RowpartStats, check, vp
        from .coo import CooArray, _h
        out, st = vp(), RowpartStats()
        check(self.ctx.lib.spb_rowpart_multiply(self.h, float(Cst), _h(scalei), A_block.h, _h(scalej), B_shard.h, _h(scalek),
                                                int(duplicate_policy), int(bool(zero_nan)), int(bool(fetch_all)), C.byref(out), C.byref(st)))
        return CooArray(self.ctx, out), st

    def close(self):
        if self.h:
            self.ctx.lib.spb_rowpart_destroy(self.h)
            self.h = None


def _gather_uneven(full: torch.Tensor, offs, local: torch.Tensor, rank: int, world: int):
    """Each rank's `local` lands in full[offs[g]:offs[g+1]] on every rank.  Returns pending works.
    NCCL: one grouped call (torch coalesces the per-shard broadcasts into a single launch, so all links
    are busy at once); other backends: one broadcast per shard."""
    views = [full[int(offs[g]):int(offs[g + 1])] for g in range(world)]
    if dist.get_backend() == "nccl":
        return [dist.all_gather(views, local, async_op=True)]
    views[rank].copy_(local)
    return [dist.broadcast(views[g], src=g, async_op=True) for g in range(world) if views[g].numel()]


def replicate_start(local: list[torch.Tensor], rank: int, world: int):
    """Starts replicating per-rank shards of plain arrays (all of one length on a rank; lengths may differ
    between ranks) straight into their slices of the full arrays -- no padding, no compaction copy.
    Returns (full_arrays, pending_works, sizes)."""
    dev = local[0].device
    n_local = int(local[0].shape[0])
    sizes_t = torch.zeros(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(sizes_t, torch.tensor([n_local], dtype=torch.int64, device=dev))
    sizes = [int(x) for x in sizes_t.tolist()]
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    fulls, works = [], []
    for t in local:
        full = torch.empty(int(offs[-1]), dtype=t.dtype, device=dev)
        works += _gather_uneven(full, offs, t, rank, world)
        fulls.append(full)
    return fulls, works, sizes


def replicate_csr_start(local_ptr: torch.Tensor, cols: torch.Tensor, vals: torch.Tensor, rank: int, world: int):
    """Starts replicating a row-sharded consolidated matrix in compressed form.
      local_ptr[i]  offset, inside this rank's shard, of the first entry of its i-th row   (int32, rows_local)
      cols, vals    column and value of every entry of the shard                            (n_local)
    Returns a state for replicate_csr_finish."""
    dev = cols.device
    meta = torch.zeros(2 * world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(meta, torch.tensor([int(cols.shape[0]), int(local_ptr.shape[0])], dtype=torch.int64, device=dev))
    meta = meta.view(world, 2).tolist()
    eoffs = np.concatenate([[0], np.cumsum([m[0] for m in meta])]).astype(np.int64)
    roffs = np.concatenate([[0], np.cumsum([m[1] for m in meta])]).astype(np.int64)
    full_ptr = torch.empty(int(roffs[-1]) + 1, dtype=torch.int32, device=dev)
    full_cols = torch.empty(int(eoffs[-1]), dtype=cols.dtype, device=dev)
    full_vals = torch.empty(int(eoffs[-1]), dtype=vals.dtype, device=dev)
    works = _gather_uneven(full_ptr[:-1], roffs, local_ptr, rank, world)
    works += _gather_uneven(full_cols, eoffs, cols, rank, world)
    works += _gather_uneven(full_vals, eoffs, vals, rank, world)
    return dict(ptr=full_ptr, cols=full_cols, vals=full_vals, works=works, eoffs=eoffs, roffs=roffs, world=world)


def replicate_csr_finish(st):
    """Waits for the transfers and turns the gathered local row pointers into global ones.
    Returns (ptr int32[rows+1], cols, vals, nnz)."""
    for w in st["works"]:
        w.wait()
    if st.get("event") is not None:
        torch.cuda.current_stream().wait_event(st["event"])
    if st.get("final"):  # pruned pull: the pointer is already global
        return st["ptr"], st["cols"], st["vals"], st["total"]
    ptr, eoffs, roffs = st["ptr"], st["eoffs"], st["roffs"]
    for g in range(1, st["world"]):
        if roffs[g + 1] > roffs[g] and eoffs[g]:
            ptr[int(roffs[g]):int(roffs[g + 1])] += int(eoffs[g])
    ptr[-1] = int(eoffs[-1])
    return ptr, st["cols"], st["vals"], int(eoffs[-1])


def plan_pulls(roffs, need_lo: int, need_hi: int):
    """Which rows of which peer a rank must fetch when its block of A only references the inner indices
    need_lo..need_hi (inclusive; the interval hull of its column support).  roffs[g]..roffs[g+1] are the global
    rows of peer g's shard.  Returns [(g, a, b)]: local rows [a, b) of peer g, in rank (= row) order.
    A block whose columns span everything gets every shard whole -- the plain replicate; a banded block gets its
    own shard plus a halo."""
    out = []
    for g in range(len(roffs) - 1):
        lo, hi = max(int(need_lo), int(roffs[g])), min(int(need_hi) + 1, int(roffs[g + 1]))
        if lo < hi:
            out.append((g, lo - int(roffs[g]), hi - int(roffs[g])))
    return out


def assemble_pruned_ptr(rows_total: int, roffs, pulls, ptr_chunks, e_los, counts, device):
    """Global row pointer (int32[rows_total + 1]) over the concatenation of the pulled entry ranges: rows that
    were not fetched are empty.  ptr_chunks[i] = peer rows' LOCAL pointers for pulls[i]; e_los[i] = local offset
    of the first pulled entry of that peer; counts[i] = entries pulled from it."""
    ptr = torch.empty(rows_total + 1, dtype=torch.int32, device=device)
    if not pulls:
        ptr.zero_()
        return ptr, 0
    g0, a0, _ = pulls[0]
    gl, _, bl = pulls[-1]
    first, end = int(roffs[g0]) + a0, int(roffs[gl]) + bl
    ptr[:first] = 0
    pos = 0
    for (g, a, b), chunk, e_lo, cnt in zip(pulls, ptr_chunks, e_los, counts):
        torch.add(chunk, int(pos) - int(e_lo), out=ptr[int(roffs[g]) + a:int(roffs[g]) + b])
        pos += int(cnt)
    ptr[end:] = pos
    return ptr, pos


class PeerReplicator:
    """Replicates the row-sharded compressed B by PULLING the peers' shards out of symmetric (peer-mapped)
    memory with device-to-device copies over NVLink.  Copy engines move the data, so unlike an NCCL
    collective it needs no SMs or shared memory and really overlaps consolidate(A) (measured on 4xB200: the
    grouped NCCL all-gather got ~200 GB/s while the radix passes held every SM; the same bytes pulled by copy
    engines arrive at ~680 GB/s).  Falls back to the NCCL path (replicate_csr_*) when symmetric memory is
    unavailable (CPU/gloo, old torch)."""

    def __init__(self, rank: int, world: int, cap_entries: int, cap_rows: int, device):
        import torch.distributed._symmetric_memory as symm
        self.rank, self.world = rank, world
        caps = torch.tensor([cap_entries, cap_rows], dtype=torch.int64, device=device)
        dist.all_reduce(caps, op=dist.ReduceOp.MAX)  # symmetric buffers have one size on every rank
        self.cap_entries, self.cap_rows = (int(x) for x in caps.tolist())
        name = dist.group.WORLD.group_name
        self.bufs, self.hdls = [], []
        for n, dt in ((self.cap_rows + 1, torch.int32), (self.cap_entries, torch.int32), (self.cap_entries, torch.float64)):
            t = symm.empty(max(n, 1), dtype=dt, device=device)
            self.bufs.append(t)
            self.hdls.append(symm.rendezvous(t, name))
        self.peers = [[h.get_buffer(g, (b.shape[0],), b.dtype) for g in range(world)] for h, b in zip(self.hdls, self.bufs)]
        # per-step metadata travels through peer memory too: a NCCL collective launched while the radix sort holds
        # every SM waits for the tail of the running pass (milliseconds); these are plain loads on a high-priority stream
        self.meta_buf = symm.empty(4, dtype=torch.int64, device=device)
        self.meta_hdl = symm.rendezvous(self.meta_buf, name)
        self.meta_peers = [self.meta_hdl.get_buffer(g, (4,), torch.int64) for g in range(world)]
        self.side = torch.cuda.Stream(device=device, priority=-1)  # tiny kernels + copies: ahead of the sort's blocks
        self.done = torch.cuda.Event()

    def start(self, local_ptr: torch.Tensor, cols: torch.Tensor, vals: torch.Tensor, need=None):
        """need = (lo, hi): inclusive range of inner indices this rank's block of A references, or None for
        everything.  Only the rows of B inside it are fetched (plan_pulls), and only the rows some peer asked for
        are published."""
        dev, rank, world = cols.device, self.rank, self.world
        n_local, rows_local = int(cols.shape[0]), int(local_ptr.shape[0])
        assert n_local <= self.cap_entries and rows_local <= self.cap_rows
        lo, hi = need if need is not None else (0, (1 << 62))
        self.meta_buf.copy_(torch.tensor([n_local, rows_local, lo, hi], dtype=torch.int64), non_blocking=False)
        self.meta_hdl.barrier(channel=0)         # every rank has written its numbers
        meta = torch.stack(self.meta_peers).cpu().tolist()
        self.meta_hdl.barrier(channel=1)         # ... and every rank has read them (the buffer is rewritten next step)
        if need is not None:
            return self._start_pruned(meta, local_ptr, cols, vals, dev)
        # publish this rank's shard (previous step's readers are past their "done" barrier, see below)
        self.bufs[0][:rows_local].copy_(local_ptr)
        self.bufs[1][:n_local].copy_(cols)
        self.bufs[2][:n_local].copy_(vals)
        eoffs = np.concatenate([[0], np.cumsum([m[0] for m in meta])]).astype(np.int64)
        roffs = np.concatenate([[0], np.cumsum([m[1] for m in meta])]).astype(np.int64)
        full = [torch.empty(int(roffs[-1]) + 1, dtype=torch.int32, device=dev),
                torch.empty(int(eoffs[-1]), dtype=torch.int32, device=dev),
                torch.empty(int(eoffs[-1]), dtype=torch.float64, device=dev)]
        cur = torch.cuda.current_stream()
        self.hdls[0].barrier(channel=0)          # every rank has published
        self.side.wait_stream(cur)
        with torch.cuda.stream(self.side):
            for k in range(world):               # own shard first, then round the ring: spreads the load over the peers
                g = (rank + k) % world
                for a, offs in ((0, roffs), (1, eoffs), (2, eoffs)):
                    cnt = int(offs[g + 1] - offs[g])
                    if cnt:
                        full[a][int(offs[g]):int(offs[g + 1])].copy_(self.peers[a][g][:cnt], non_blocking=True)
            self.hdls[0].barrier(channel=1)      # every rank has finished pulling: buffers may be overwritten
            self.done.record(self.side)
        for t in full:
            t.record_stream(self.side)
        return dict(ptr=full[0], cols=full[1], vals=full[2], works=[], eoffs=eoffs, roffs=roffs, world=world, event=self.done)


def _peer_start_pruned(self, meta, local_ptr, cols, vals, dev):
    import os
    import time
    trace = bool(os.environ.get("SPB_DIST_TRACE"))
    t0 = time.perf_counter()
    marks = []
    rank, world = self.rank, self.world
    n_local, rows_local = int(cols.shape[0]), int(local_ptr.shape[0])
    roffs = np.concatenate([[0], np.cumsum([m[1] for m in meta])]).astype(np.int64)
    pulls = plan_pulls(roffs, meta[rank][2], meta[rank][3])
    # publish the rows of this shard that some peer will fetch (their hull), at their usual offsets
    mine = [p for q in range(world) if q != rank for p in plan_pulls(roffs, meta[q][2], meta[q][3]) if p[0] == rank]
    own = [p for p in pulls if p[0] == rank]
    want = sorted({x for (_, a, b) in mine + own for x in (a, b)})  # local rows whose pointer value the host needs
    ptr_at = {}
    if want:
        inside = [x for x in want if x < rows_local]
        got = local_ptr[torch.tensor(inside, dtype=torch.int64, device=dev)].cpu().tolist() if inside else []
        ptr_at = dict(zip(inside, got))
        ptr_at[rows_local] = n_local
    if mine:
        u_lo, u_hi = min(a for _, a, _ in mine), max(b for _, _, b in mine)
        e0, e1 = int(ptr_at[u_lo]), int(ptr_at[u_hi])
        self.bufs[0][u_lo:u_hi].copy_(local_ptr[u_lo:u_hi])
        self.bufs[0][u_hi:u_hi + 1].fill_(e1)    # pointer one past the last published row
        self.bufs[1][e0:e1].copy_(cols[e0:e1])
        self.bufs[2][e0:e1].copy_(vals[e0:e1])
    marks.append(("published", time.perf_counter() - t0))
    cur = torch.cuda.current_stream()
    if trace:
        cur.synchronize()
        marks.append(("publish done", time.perf_counter() - t0))
    self.hdls[0].barrier(channel=0)              # every rank has published
    if trace:
        cur.synchronize()
        marks.append(("barrier0 done", time.perf_counter() - t0))
    side = cur if os.environ.get("SPB_DIST_ONE_STREAM") else self.side
    side.wait_stream(cur)
    with torch.cuda.stream(side):
        # where the wanted rows start and end inside each peer's shard: two pointers per peer, read from its memory
        remote = [(g, a, b) for (g, a, b) in pulls if g != rank]
        rem = torch.stack([self.peers[0][g][x] for (g, a, b) in remote for x in (a, b)]).cpu().tolist() if remote else []
        marks.append(("peer pointers read", time.perf_counter() - t0))
        ends = {}
        for i, (g, a, b) in enumerate(remote):
            ends[g] = (int(rem[2 * i]), int(rem[2 * i + 1]))
        for (g, a, b) in own:
            ends[g] = (int(ptr_at[a]), int(ptr_at[b]))
        e_los = [ends[g][0] for (g, _, _) in pulls]
        counts = [ends[g][1] - ends[g][0] for (g, _, _) in pulls]
        total = int(sum(counts))
        fc = torch.empty(max(total, 1), dtype=torch.int32, device=dev)
        fv = torch.empty(max(total, 1), dtype=torch.float64, device=dev)
        starts = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        order = sorted(range(len(pulls)), key=lambda i: (pulls[i][0] - rank) % world)  # own shard first, then round the ring
        chunks = []
        for i in order:
            g, a, b = pulls[i]
            src_c, src_v = (cols, vals) if g == rank else (self.peers[1][g], self.peers[2][g])
            if counts[i]:
                fc[int(starts[i]):int(starts[i + 1])].copy_(src_c[e_los[i]:e_los[i] + counts[i]], non_blocking=True)
                fv[int(starts[i]):int(starts[i + 1])].copy_(src_v[e_los[i]:e_los[i] + counts[i]], non_blocking=True)
        for (g, a, b) in pulls:
            chunks.append(local_ptr[a:b] if g == rank else self.peers[0][g][a:b])
        ptr, total = assemble_pruned_ptr(int(roffs[-1]), roffs, pulls, chunks, e_los, counts, dev)
        self.hdls[0].barrier(channel=1)          # every rank has finished pulling: buffers may be overwritten
        self.done.record(side)
        if trace:
            marks.append(("enqueued", time.perf_counter() - t0))
            self.done.synchronize()
            marks.append(("done", time.perf_counter() - t0))
            import sys
            print(f"[dist] rank {rank} pruned pull, ms since start: " + ", ".join(f"{k} {v * 1e3:.2f}" for k, v in marks), file=sys.stderr)
    for t in (ptr, fc, fv):
        t.record_stream(side)
    return dict(ptr=ptr, cols=fc, vals=fv, works=[], final=True, total=total, world=world, event=self.done,
                pulled_rows=int(sum(b - a for _, a, b in pulls)))


PeerReplicator._start_pruned = _peer_start_pruned


def replicate_wait(works) -> None:
    for w in works:
        w.wait()

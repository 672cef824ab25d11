"""Measures the ways of replicating a sharded array across the ranks of one box (run under torchrun)."""
import os
import sys
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n_total = int(float(sys.argv[1])) if len(sys.argv) > 1 else 500_000_000
n = n_total // world
x = torch.full((n,), float(rank), dtype=torch.float64, device="cuda")


def timeit(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


full = torch.empty(n * world, dtype=torch.float64, device="cuda")
rx_gb = n * 8 * (world - 1) / 1e9


def ag_equal():
    dist.all_gather_into_tensor(full, x)


def ag_uneven_group():
    views = [full[g * n:(g + 1) * n - (1 if g == 0 else 0)] for g in range(world)]
    dist.all_gather(views, x[:views[rank].numel()])


def bcasts():
    full[rank * n:(rank + 1) * n] = x
    ws = [dist.broadcast(full[g * n:(g + 1) * n], src=g, async_op=True) for g in range(world)]
    for w in ws:
        w.wait()


res = {"ag_equal": timeit(ag_equal), "ag_uneven_group": timeit(ag_uneven_group), "bcasts": timeit(bcasts)}
# peer copies through symmetric memory, if this torch build has it
try:
    import torch.distributed._symmetric_memory as symm
    buf = symm.empty(n, dtype=torch.float64, device="cuda")
    hdl = symm.rendezvous(buf, dist.group.WORLD.group_name)
    buf.copy_(x)
    peers = [hdl.get_buffer(g, (n,), torch.float64) for g in range(world)]

    def pull():
        hdl.barrier()
        for k in range(world):
            g = (rank + k) % world  # start with own shard, then walk the ring: spreads the load over the peers
            full[g * n:(g + 1) * n].copy_(peers[g])
        hdl.barrier()
    res["symm_pull"] = timeit(pull)
    ok = all(bool((full[g * n:(g + 1) * n] == float(g)).all()) for g in range(world))
    res["symm_ok"] = ok
except Exception as e:  # noqa: BLE001
    res["symm_error"] = repr(e)[:300]
if rank == 0:
    for k, v in res.items():
        if isinstance(v, float):
            print(f"{k:18s} {v:8.2f} ms   {rx_gb / (v * 1e-3):7.1f} GB/s received per rank")
        else:
            print(k, v)
dist.destroy_process_group()

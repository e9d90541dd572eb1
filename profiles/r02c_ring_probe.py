"""2-GPU probe (torchrun): where the time of the gathered one-sweep ring build goes -- the all-gather of the rows, the
build on the gathered matrix with the plan prepared across the ranks, and make_knn_ring as a whole."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from bench import synth_cells
from vivae_b200.device import knn_device
from vivae_b200.sharded import all_gather_rows, make_knn_ring, shard_bounds

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n, d, k = 1_000_000, 50, 50
x = torch.from_numpy(synth_cells(n, d)).to(dev)
q0, nq = shard_bounds(n, world, rank)
xl = x[q0:q0 + nq].contiguous()

def timed(fn, reps=5):
    fn(); torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return {"ms_events": e0.elapsed_time(e1) / reps, "ms_wall": 1e3 * (time.perf_counter() - t0) / reps}

res = {}
res["all_gather_rows"] = timed(lambda: all_gather_rows(xl, n))
xa = all_gather_rows(xl, n)
res["knn_on_gathered_with_group"] = timed(lambda: knn_device(xa, k, q0, nq, group=dist.group.WORLD))
res["knn_on_replicated_with_group"] = timed(lambda: knn_device(x, k, q0, nq, group=dist.group.WORLD))
res["knn_on_gathered_no_group"] = timed(lambda: knn_device(xa, k, q0, nq))
res["gather_then_knn"] = timed(lambda: knn_device(all_gather_rows(xl, n), k, q0, nq, group=dist.group.WORLD))
res["make_knn_ring_auto"] = timed(lambda: make_knn_ring(xl, n, k))
res["make_knn_ring_one_per_sweep"] = timed(lambda: make_knn_ring(xl, n, k, shards_per_sweep=1))
if rank == 0:
    print(json.dumps(res))
dist.barrier()
dist.destroy_process_group()

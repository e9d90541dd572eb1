"""Timing probe of the peer-memory smoother (run under torchrun): what its system-scope fences cost."""
import json, os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import synth_cells
from vivae_b200.device import knn_device, smooth_device
from vivae_b200.sharded import release_peer_smoothers, smooth_peer_sharded
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
x = torch.from_numpy(synth_cells(1_000_000, 50, seed=5)).to(dev)
idx, _ = knn_device(x, 50)
want = smooth_device(x, idx, k=50, coef=1.0, n_iter=1)
res = {}
for cfg in os.environ.get("MODES", "0:0").split(","):
    mode, batch = cfg.split(":")
    os.environ["VVB_PEER_DBG_FENCE"] = mode
    if batch == "0":
        os.environ.pop("VVB_PEER_BATCH", None)
    else:
        os.environ["VVB_PEER_BATCH"] = batch
    fn = lambda: smooth_peer_sharded(x, idx, k=50, coef=1.0, n_iter=1)
    for _ in range(3): got = fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 5], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res[f"dbg{mode}_batch{batch}"] = {"ms": round(float(t.item()), 3), "same": bool(torch.equal(got, want))}
release_peer_smoothers()
if rank == 0: print(json.dumps(res))
dist.barrier(); dist.destroy_process_group()

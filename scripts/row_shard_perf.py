"""One n-body world on N GPUs: replicas vs row shards over NCCL vs row shards through peer windows (N = 1024, 8192).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/row_shard_perf.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import elodin_b200 as el
from elodin_b200.executor import WORLD_POS
from elodin_b200.sharding import RowShardedWorld

rank, world_size, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
stream = torch.cuda.Stream()

def barrier():
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()

def max_over_ranks(x):
    t = torch.tensor([x], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); return float(t.item())

def world(n, seed):
    g = np.random.default_rng(seed)
    p = np.zeros((1, n, 7)); p[..., 3] = 1.0; p[..., 4:] = g.uniform(-30, 30, (1, n, 3))
    v = np.zeros((1, n, 6)); v[..., 3:] = g.normal(0, 1e-7, (1, n, 3))
    m = 10 ** g.uniform(-10, -3, (1, n)); m[:, 0] = 1.0
    I = np.zeros((1, n, 7)); I[..., :3] = m[..., None]; I[..., 6] = m
    return p, v, I

grav = lambda n: el.GravityEdges("softened", k_squared=2.9591220828e-4 / 86400.0 ** 2, softening=1e-10, edges=el.all_pairs_edges(n))
ev = lambda: torch.cuda.Event(enable_timing=True)
out = {}
for n, warm, ticks in ((1024, 20, 400), (2048, 10, 200), (4096, 10, 100), (8192, 5, 40)):
    w = world(n, 7 if n == 1024 else 11)
    ex = el.B200Exec(n, 1, 3600.0, None, [grav(n)], "rk4", "fast", device=local)
    ex.set_stream(stream.cuda_stream); ex.set_state(*w)
    with torch.cuda.stream(stream):
        ex.step(warm); barrier(); a, b = ev(), ev(); a.record(stream); ex.step(ticks); b.record(stream); barrier()
    ms = max_over_ranks(a.elapsed_time(b))
    ref = ex.download(WORLD_POS); ex.close()
    r = RowShardedWorld._both(torch, dist, el, stream, local, rank, world_size, barrier, max_over_ranks, w, grav, ref, warm, ticks)
    out[n] = {"replica_us": ms * 1e3 / ticks, "nccl": r["nccl"], "peer": r["peer"]}
    if rank == 0:
        print(n, "replica %.1f us" % (ms * 1e3 / ticks), {k: (round(r[k].get("us_per_tick", -1), 1), r[k].get("max_rel_diff_vs_replica"), r[k].get("unavailable")) for k in ("nccl", "peer")}, flush=True)
if rank == 0:
    print(json.dumps(out))
dist.destroy_process_group()

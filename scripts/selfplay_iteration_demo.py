"""One self-play iteration sharded over the ranks of a torchrun job (config c4 in miniature): rank 0's weights are broadcast
over NCCL, every rank plays its block of args['num_selfPlay_iterations'] games on its GPU, per-game digests are gathered.
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/selfplay_iteration_demo.py"""
import hashlib, os, sys
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sigma_zero_b200.network import policyNN
from sigma_zero_b200.train_RL import flatten_state_dict, selfplay_iteration, shard_of

local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
torch.manual_seed(100 + rank)                      # ranks start with different weights; the iteration begins with a broadcast
model = policyNN({}).eval()
args = {"C": 2, "num_searches": 24, "num_selfPlay_iterations": 24, "chess960": True}
games, counters = selfplay_iteration(model, args, seed=5, max_plies=8)
_, flat = flatten_state_dict(model.state_dict())
wsum = torch.tensor([float(flat.double().sum())], dtype=torch.float64, device="cuda")
sums = [torch.zeros_like(wsum) for _ in range(world)]
dist.all_gather(sums, wsum)
assert all(float(s) == float(sums[0]) for s in sums), "weights differ after the broadcast"
lo, hi = shard_of(args["num_selfPlay_iterations"], rank, world)
digest = [hashlib.sha1(("|".join(",".join(m.uci() + ":%r" % p for m, p in a.items()) for a in g["actions"])).encode()).hexdigest()[:12] for g in games]
out = [None] * world
dist.all_gather_object(out, (lo, hi, digest))
if rank == 0:
    for lo_, hi_, d in out:
        print("games [%d, %d): %s" % (lo_, hi_, " ".join(d)))
    print("world %d: %d games, %d plies each, weights identical on all ranks" % (world, sum(h - l for l, h, _ in out), counters["plies"]))
dist.destroy_process_group()

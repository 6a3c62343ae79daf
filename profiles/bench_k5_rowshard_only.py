"""bench.py's `k5_rowshard` leg alone (config 5 as worded: matrix row-sharded over the N GPUs beside the replicated mode).
Launch:  python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port P \
             profiles/bench_k5_rowshard_only.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch                                                    # noqa: E402
import torch.distributed as dist                                # noqa: E402
import adaptive_matrix_solver_b200 as pkg                       # noqa: E402
from adaptive_matrix_solver_b200 import _abi                    # noqa: E402
from adaptive_matrix_solver_b200.dist import Shard              # noqa: E402
import bench                                                    # noqa: E402

rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
eng = pkg.MausEngine(local)
eng.enable_row_sharding(rank, world)
shard = Shard(rank, world, dev if world > 1 else None, engine=eng)
out = bench.bench_k5_rowshard(pkg, eng, shard, torch, _abi, rank, world)
if rank == 0:
    print(json.dumps(out), flush=True)
shard.barrier()
eng.close()
if world > 1:
    dist.destroy_process_group()

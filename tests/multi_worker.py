"""One decomposition rank of a multi-GPU parity run (launched by test_gpu_multi.py, one
process per GPU).  Every process runs the full multi-rank CPU oracle in-process (it is
deterministic) and checks ITS rank of the CUDA path against it: per-cell and per-rank
particle counts bit-exact after exchange, E/B/J within 1e-12 relative L2."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from tests.parity_check import make_deck, run_case, run_rebalance_case  # noqa: E402,F401


def main():
    name = sys.argv[1]
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("gloo")

    def share_id(uid):
        ids = [uid]
        dist.broadcast_object_list(ids, 0)
        return ids[0]

    def allgather(obj):
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out

    if name == "rebalance2d":
        res = run_rebalance_case(rank, world, share_id, allgather)
    else:
        res = run_case(name, rank, world, share_id)
    dist.barrier()
    dist.destroy_process_group()
    print("RESULT " + json.dumps(res), flush=True)
    sys.exit(0 if res["ok"] else 1)


if __name__ == "__main__":
    main()

"""Multi-GPU parity script (run under torchrun on G GPUs): otgan_b200.train.parity_check -- the summed tower gradient,
distance and entropy of one critic step and one generator step computed by G ranks must equal the same step computed by a
single rank on the full batch (relative gates, see parity_check), and every rank must derive bitwise identical grad_ys /
distance / entropy from the gathered embeddings.  The same check runs inside `bench.py --gpus N` (key `mgpu_parity`).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port P tests/mgpu_parity.py
Prints "MGPU PARITY OK" on rank 0 (exit code 0) or the rows that failed (exit code 1).
"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from otgan_b200 import train as T  # noqa: E402


def main():
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    res = T.parity_check(world, rank, dev)
    if rank == 0:
        for r in res["rows"]:
            print(json.dumps(r))
        print("matching outputs bitwise identical on all ranks:", res["matching_bitwise"])
        print("MGPU PARITY OK (world %d)" % world if res["ok"] else "MGPU PARITY FAILED")
    dist.destroy_process_group()
    sys.exit(0 if res["ok"] else 1)


if __name__ == "__main__":
    main()

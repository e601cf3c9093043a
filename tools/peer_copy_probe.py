#!/usr/bin/env python
"""Probe: device-to-device copy rate through symmetric (peer-mapped) memory, push and pull, idle and under a
compute kernel that occupies every SM.  Run under torchrun with >= 2 ranks; rank 0 prints JSON."""
import json
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n = 115 * (1 << 20) // 4
    buf = symm.empty((world, n), dtype=torch.float32, device=dev)
    hdl = symm.rendezvous(buf, dist.group.WORLD)
    mine = torch.ones(n, dtype=torch.float32, device=dev)
    out = {}

    def timed(fn, reps=5):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    # busy kernel: a big matmul loop on another stream keeps the SMs occupied
    a = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
    busy_stream = torch.cuda.Stream(device=dev)

    def busy_on():
        with torch.cuda.stream(busy_stream):
            for _ in range(40):
                torch.mm(a, a)

    for label, busy in (("idle", False), ("busy", True)):
        dist.barrier()
        torch.cuda.synchronize()
        if rank == 1:
            peer0 = hdl.get_buffer(0, (n,), torch.float32, 1 * n)     # my place on rank 0
            if busy:
                busy_on()
            ms = timed(lambda: peer0.copy_(mine, non_blocking=True))
            out[f"push_1to0_{label}_GBps"] = n * 4 / ms / 1e6
            torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        if rank == 0:
            src1 = hdl.get_buffer(1, (n,), torch.float32, 0)
            dst = torch.empty(n, dtype=torch.float32, device=dev)
            if busy:
                busy_on()
            ms = timed(lambda: dst.copy_(src1, non_blocking=True))
            out[f"pull_0from1_{label}_GBps"] = n * 4 / ms / 1e6
            torch.cuda.synchronize()
        dist.barrier()
    # all ranks push to rank 0 at once
    dist.barrier()
    torch.cuda.synchronize()
    if rank != 0:
        peer0 = hdl.get_buffer(0, (n,), torch.float32, rank * n)
        ms = timed(lambda: peer0.copy_(mine, non_blocking=True))
        out["push_all_GBps_per_rank"] = n * 4 / ms / 1e6
    dist.barrier()
    res = [None] * world
    dist.all_gather_object(res, out)
    if rank == 0:
        print(json.dumps(res))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

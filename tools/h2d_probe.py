#!/usr/bin/env python
"""What the host can feed N GPUs at once: every rank copies the same-size page-locked buffer to its device, all ranks
concurrently, once from ordinary pinned memory and once from write-combined pinned memory.  Rank 0 prints one JSON line.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/h2d_probe.py
"""
from __future__ import annotations

import ctypes as C
import json
import os
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))


def main() -> None:
    import torch
    import torch.distributed as dist
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    rt = C.CDLL("libcudart.so.12")
    nbytes = 256 << 20
    dst = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    out = {}
    for name, flags in (("pinned", 0), ("write_combined", 4)):          # cudaHostAllocWriteCombined = 0x04
        p = C.c_void_p()
        assert rt.cudaHostAlloc(C.byref(p), C.c_size_t(nbytes), C.c_uint(flags)) == 0
        C.memset(p, 1, nbytes)
        reps = 8
        rt.cudaMemcpyAsync(C.c_void_p(dst.data_ptr()), p, C.c_size_t(nbytes), 1, None)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            rt.cudaMemcpyAsync(C.c_void_p(dst.data_ptr()), p, C.c_size_t(nbytes), 1, None)
        torch.cuda.synchronize()
        gbps = reps * nbytes / (time.perf_counter() - t0) / 1e9
        t = torch.tensor([gbps], dtype=torch.float64, device=dev)
        every = [torch.zeros_like(t) for _ in range(world)]
        if world > 1:
            dist.all_gather(every, t)
        else:
            every = [t]
        out[name] = [round(float(v.item()), 2) for v in every]
        rt.cudaFreeHost(p)
    if rank == 0:
        out["aggregate_GBps"] = {k: round(sum(v), 1) for k, v in out.items()}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

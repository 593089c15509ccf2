"""The platform ceiling of the end-to-end path: bare concurrent host<->device copies from pinned memory on all ranks at once
(what bench.py's e2e can at most reach: 24.9 MB of RGB per 4K frame must cross PCIe into host memory).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/pcie_ceiling.py
Prints one JSON line: aggregate and per-rank GB/s for D2H alone, H2D alone and both at once, with the ranks pinned to the CPUs
next to their GPU (like bench.py) and unpinned."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import bench


def run(pin):
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    aff = bench.bind_to_gpu_numa_node(local) if pin else "unpinned"
    n = 1 << 30
    host_a = torch.empty(n, dtype=torch.uint8).pin_memory()
    host_b = torch.empty(n, dtype=torch.uint8).pin_memory()
    host_a.fill_(1)
    dev_a = torch.empty(n, dtype=torch.uint8, device="cuda")
    dev_b = torch.empty(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    out = {}
    for mode in ("d2h", "h2d", "both"):
        for rep in range(2):  # first repetition warms up
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(4):
                if mode in ("d2h", "both"):
                    with torch.cuda.stream(s1):
                        host_b.copy_(dev_b, non_blocking=True)
                if mode in ("h2d", "both"):
                    with torch.cuda.stream(s2):
                        dev_a.copy_(host_a, non_blocking=True)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            dt = time.perf_counter() - t0
        t = torch.tensor([dt], device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        nbytes = 4 * n * (2 if mode == "both" else 1)
        out[mode + "_gbs_per_gpu"] = round(nbytes / float(t.item()) / 1e9, 1)
        out[mode + "_gbs_total"] = round(world * nbytes / float(t.item()) / 1e9, 1)
    return aff, out


def main():
    local, world = int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    res = {}
    for pin in (False, True):
        if not pin:
            os.environ["B200JPG_NO_AFFINITY"] = "1"
        else:
            os.environ.pop("B200JPG_NO_AFFINITY", None)
        aff, out = run(pin)
        res["pinned_to_gpu_numa_node" if pin else "unpinned"] = dict(out, affinity=aff)
    if int(os.environ.get("RANK", 0)) == 0:
        print(json.dumps({"n_gpus": world, "bytes_per_copy": 1 << 30, "frame_bytes_4k_rgb": 3840 * 2160 * 3, **res}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

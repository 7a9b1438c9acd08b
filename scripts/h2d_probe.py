"""torchrun diagnostic: what the box moves between host and its GPUs when all ranks copy at once.
Per rank: GPU NUMA node (sysfs), CPU affinity, pinned H2D / D2H bandwidth alone-ish and all together."""
import os
import time

import torch
import torch.distributed as dist


def numa_of_gpu(i):
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(i)
        bdf = pynvml.nvmlDeviceGetPciInfo(h).busId
        if isinstance(bdf, bytes):
            bdf = bdf.decode()
        bdf = bdf.lower()
        if len(bdf.split(":")[0]) == 8:
            bdf = bdf[4:]
        return open("/sys/bus/pci/devices/%s/numa_node" % bdf).read().strip(), bdf
    except Exception as e:  # noqa: BLE001
        return "?(%r)" % (e,), "?"


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    if os.environ.get("PROBE_BIND", "1") == "1":
        import bench

        nb = bench.bind_to_gpu_numa_node(local)
    else:
        nb = -1
    n = 512 << 20
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_in.fill_(1)
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def run(h2d, d2h, reps=6):
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            if h2d:
                with torch.cuda.stream(s1):
                    d.copy_(h_in, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h_out.copy_(d, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        return n * reps / dt / 1e9

    run(True, True, 2)
    a, b, cboth = run(True, False), run(False, True), run(True, True)
    node, bdf = numa_of_gpu(local)
    try:
        nodes = sorted(x for x in os.listdir("/sys/devices/system/node") if x.startswith("node"))
    except Exception:
        nodes = []
    msg = "[rank %d] gpu %s numa %s | cpus allowed %d (bound %s) of %d, host nodes %s | all ranks at once: H2D %.1f GB/s, D2H %.1f GB/s, both %.1f + %.1f GB/s" % (
        rank, bdf, node, len(os.sched_getaffinity(0)), nb, os.cpu_count(), nodes, a, b, cboth, cboth)
    out = [None] * world
    dist.all_gather_object(out, msg)
    if rank == 0:
        print("\n".join(out), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

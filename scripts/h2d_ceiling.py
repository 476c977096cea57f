"""Copy-only ceiling of the end-to-end path: every rank copies pinned 21.8 MB
blocks (the row range of a configs[1] frame that holds georeferenced pixels) host -> device with
cudaMemcpyAsync, all ranks at once; prints GB/s per GPU and the aggregate.  Run under torchrun with
1, 2, 4, 8 ranks on one box:

    for n in 1 2 4 8; do python -m torch.distributed.run --nproc-per-node $n --master-addr 127.0.0.1 \
        scripts/h2d_ceiling.py; done
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    if world > 1:
        from auromat_b200.parallel import bindToLocalCpus
        bindToLocalCpus(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nbytes = int(os.environ.get("H2D_BLOCK", 21_800_000))
    reps = 200
    host = [torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in range(4)]
    devb = [torch.empty(nbytes, dtype=torch.uint8, device="cuda") for _ in range(4)]
    stream = torch.cuda.Stream()
    out = {}
    for mode in ("h2d", "d2h", "both"):
        for _ in range(2):
            with torch.cuda.stream(stream):
                for i in range(8):
                    devb[i % 4].copy_(host[i % 4], non_blocking=True)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s2 = torch.cuda.Stream()
        e0.record(stream)
        with torch.cuda.stream(stream):
            for i in range(reps):
                if mode in ("h2d", "both"):
                    devb[i % 4].copy_(host[i % 4], non_blocking=True)
                elif mode == "d2h":
                    host[i % 4].copy_(devb[i % 4], non_blocking=True)
        if mode == "both":
            with torch.cuda.stream(s2):
                for i in range(reps):
                    host[(i + 2) % 4].copy_(devb[(i + 2) % 4], non_blocking=True)
            stream.wait_stream(s2)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        out[mode] = reps * nbytes / (ms.item() * 1e-3) / 1e9
    if rank == 0:
        print("ranks %d  block %.1f MB  per-GPU GB/s (slowest rank): H2D %.1f  D2H %.1f  H2D while D2H %.1f  |  "
              "aggregate H2D %.1f GB/s  =>  copy-only floor of the e2e step: %.3f ms/frame/GPU" % (
                  world, nbytes / 1e6, out["h2d"], out["d2h"], out["both"], out["h2d"] * world,
                  nbytes / out["h2d"] / 1e6))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

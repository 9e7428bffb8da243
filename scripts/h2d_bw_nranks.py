"""Pure host->device ingest probe under the same launch as the benchmark: N ranks (one per GPU), each copying the
byte count of one end-to-end step (64 stereo 1280x720 frames = 117 964 800 B) from pinned host memory to its GPU, all ranks
at once.  Reports per-rank and aggregate GB/s; rank 0 prints one JSON line.  This is the ceiling `bench.py`'s e2e leg is
compared with (`e2e.roofline`): if the aggregate here equals what the pipeline reaches, the limit is the host platform.

    python scripts/h2d_bw_nranks.py                                   # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        scripts/h2d_bw_nranks.py [--bytes B] [--iters K] [--streams S] [--bind 0|1]
"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bytes", type=int, default=64 * 2 * 1280 * 720)
    ap.add_argument("--iters", type=int, default=40)
    ap.add_argument("--streams", type=int, default=1, help="concurrent copy streams per rank (the step's bytes are split)")
    ap.add_argument("--bind", type=int, default=1, help="bind the rank to its GPU's NVML-local CPUs before allocating")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    placement = "default"
    if args.bind:
        from bench import bind_to_gpu_cpus
        placement = bind_to_gpu_cpus(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        from bench import stdout_to_stderr
        with stdout_to_stderr():
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
    n = args.bytes // args.streams
    host = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(args.streams)]
    for h in host:
        h.fill_(rank + 1)                                  # first touch on this rank's CPUs
    dev = [torch.empty(n, dtype=torch.uint8, device="cuda") for _ in range(args.streams)]
    streams = [torch.cuda.Stream() for _ in range(args.streams)]

    def run(iters):
        for _ in range(iters):
            for h, d, s in zip(host, dev, streams):
                with torch.cuda.stream(s):
                    d.copy_(h, non_blocking=True)

    run(5)
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0.record()
    run(args.iters)
    for s in streams:
        torch.cuda.current_stream().wait_stream(s)
    e1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    ms = max(e0.elapsed_time(e1), wall * 1e3 * 0.0)        # device time of this rank's copies
    mine = args.iters * n * args.streams / (ms * 1e-3) / 1e9
    t = torch.tensor([ms, mine], dtype=torch.float64, device="cuda")
    if dist:
        allv = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allv, t)
    else:
        allv = [t]
    if rank == 0:
        per_rank = [float(v[1]) for v in allv]
        slowest_ms = max(float(v[0]) for v in allv)
        agg = world * args.iters * n * args.streams / (slowest_ms * 1e-3) / 1e9
        print(json.dumps({"probe": "pinned host -> device, all ranks at once", "n_gpus": world, "bytes_per_copy": n * args.streams,
                          "iters": args.iters, "streams_per_rank": args.streams, "per_rank_gbs": [round(x, 2) for x in per_rank],
                          "aggregate_gbs": round(agg, 2), "ms_per_copy_slowest_rank": round(slowest_ms / args.iters, 4),
                          "host_placement": placement, "cpus_visible": len(os.sched_getaffinity(0))}))
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

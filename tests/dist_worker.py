"""world_size-2 worker for tests/test_dist_cpu.py (gloo, CPU): exercises the host-side logic bench.py uses for N > 1 --
per-rank chunk assignment (independent chunks, no data-path collective), barrier, max-over-ranks time, whole-job value."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    pcms = bench.make_inputs(rank, 2, seconds=0.5)               # this rank's chunks (seeded by rank: disjoint work)
    digest = float(sum(np.abs(p).sum() for p in pcms))
    dist.barrier()
    t0 = time.perf_counter()
    time.sleep(0.05 * (rank + 1))                                # ranks finish at different times
    dt_local = time.perf_counter() - t0
    dt = bench.max_over_ranks(dt_local, world, device="cpu")
    audio_s = 0.5 * 2 * world                                    # whole-job audio: every rank's chunks
    gathered = [None] * world
    dist.all_gather_object(gathered, {"rank": rank, "digest": digest, "dt_local": dt_local, "dt": dt})
    if rank == 0:
        print(json.dumps({"world": world, "ranks": gathered, "value": audio_s / dt}), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""BASELINE configs[2]: ONE long synthetic AP recording (385 ch x 30 kHz int16, 1 s chunks) on tmpfs, compressed by all
ranks into ONE .cbin / .ch (mtscomp_b200.sharding.write_sharded) and verified (every chunk decoded by the GPU Reader ==
source; a sample of chunks inflated by CPython zlib == the oracle's transform).  Rank 0 prints one JSON line.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \\
        tools/sharded_recording.py [--chunks 3600]

--chunks is the total asked for (3600 = the 1-hour recording of configs[2], 83 GB); it is reduced, and the line says so,
when host memory / tmpfs cannot hold the recording, its compressed copy and the staging buffers with a 2 x margin."""
import argparse
import json
import os
import shutil
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / 'tools'))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--chunks', type=int, default=3600)
    ap.add_argument('--zlib-sample', type=int, default=8)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', 0)))
    os.environ['MTSCOMP_B200_DEVICE'] = os.environ.get('LOCAL_RANK', '0')
    group = None                                             # (the default NCCL group, as in bench.py)
    local = int(os.environ.get('LOCAL_RANK', 0))
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    import bench_legs
    import mtscomp_b200  # noqa: F401
    from mtscomp_b200 import _native
    _native.default_codec(int(os.environ.get('LOCAL_RANK', 0)))
    chunk_bytes = 30000 * 385 * 2
    plan = [args.chunks]
    if rank == 0:
        avail = None
        for line in open('/proc/meminfo'):
            if line.startswith('MemAvailable'):
                avail = int(line.split()[1]) * 1024
        shm = shutil.disk_usage('/dev/shm').free if os.path.isdir('/dev/shm') else avail
        room = min(avail or shm, shm)
        per_chunk = chunk_bytes * 1.40                      # raw + compressed copy on tmpfs
        fixed = world * (3 << 30)                            # pinned staging, contexts
        fit = int((room / 2 - fixed) / per_chunk)
        plan = [max(world, min(args.chunks, fit) // world * world), avail, shm]
    if world > 1:
        dist.broadcast_object_list(plan, src=0, group=group)
    n_chunks = plan[0]
    rep = bench_legs.sharded_leg(rank, world, n_chunks // world, group=group, zlib_sample=args.zlib_sample, warm_chunks_per_rank=12)
    if rank == 0:
        rep['asked_chunks'] = args.chunks
        rep['host_mem_available_GB'] = (plan[1] or 0) / 1e9 if len(plan) > 1 else None
        rep['tmpfs_free_GB'] = (plan[2] or 0) / 1e9 if len(plan) > 2 else None
        rep['config'] = 'BASELINE configs[2]: 385 ch x 30 kHz int16 AP recording, 1 s chunks, %d ranks -> one .cbin/.ch' % world
        print(json.dumps(rep), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()

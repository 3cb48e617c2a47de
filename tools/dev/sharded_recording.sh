N=${1:-8}
C=${2:-3600}
free -g | head -2; df -h /dev/shm | tail -1; nproc
timeout 1700 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/sharded_recording.py --chunks $C > gpurun_out/r02_sharded_recording_n$N.log 2> gpurun_out/r02_sharded_recording_n$N.err
tail -c 1800 gpurun_out/r02_sharded_recording_n$N.log; tail -4 gpurun_out/r02_sharded_recording_n$N.err

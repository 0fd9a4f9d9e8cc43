#!/bin/sh
# Round-2 stage C (2 GPUs): cross-device parity (device tier, IPC pull), NVLink counters, bench at N=2, then the
# end-to-end rehearsal: real vLLM x unmodified router x unmodified harness at N=1,2 with a qps sweep, kv-aware
# routing, cross-replica pulls (device tier vs shared host pool), 1P+1D.
#   gpurun --gpus 2 --timeout 1500 -- tools/r2/stage_c.sh
set -u
out=gpurun_out/r2c
mkdir -p "$out"
df -h /dev/shm > "$out/host.txt" 2>&1; free -g >> "$out/host.txt" 2>&1; nvidia-smi topo -m >> "$out/host.txt" 2>&1
timeout 600 python -m pytest tests/test_gpu_device_tier.py tests/test_gpu_ipc.py -m gpu -q -rA --timeout 300 -p no:cacheprovider > "$out/pytest_2gpu.txt" 2>&1
tail -4 "$out/pytest_2gpu.txt"
M=gpu__time_duration.sum,nvlrx__bytes.sum,nvlrx__bytes_data_user.sum,nvltx__bytes.sum,nvltx__bytes_data_user.sum,dram__bytes_read.sum,dram__bytes_write.sum
ncu --metrics $M --clock-control none -k regex:kv_ --csv --log-file "$out/ncu_nvlink_tier.csv" python tools/prof_tier.py 2 > "$out/prof_tier.log" 2>&1
ncu --metrics $M --clock-control none -k regex:kv_bulk_copy --csv --log-file "$out/ncu_nvlink_pull.csv" python tools/prof_pull.py 2 > "$out/prof_pull.log" 2>&1
python tools/prof_tier.py 5 > "$out/tier_noprof.log" 2>&1; tail -2 "$out/tier_noprof.log"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 > "$out/bench_n2.json" 2> "$out/bench_n2.err"
tail -c 1500 "$out/bench_n2.json"
timeout 1300 python tools/e2e/run_scale.py --gpus 2 --seconds 35 --qps-per-replica 8 --users-per-replica 24 --qps-sweep 4,12,16 \
    --cpu-gb 20 --pd-requests 8 --log-dir "$out/scale2" 2>&1 | cut -c1-900
ls "$out" "$out/scale2" | head -60

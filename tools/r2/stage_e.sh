#!/bin/sh
# Round-2 stage E (8 GPUs): the headline metric — multi-round-QA p50 TTFT and tokens/s at 1/2/4/8 replicas behind the
# unmodified router + harness (configs[2]), kv-aware routing (configs[3]), cross-replica pulls device tier vs shared
# host pool, 4P+4D (configs[4]) — every engine started once (tools/e2e/run_scale.py).
#   gpurun --gpus 8 --timeout 1500 -- tools/r2/stage_e.sh
set -u
out=gpurun_out/r2e
mkdir -p "$out"
nvidia-smi topo -m > "$out/host.txt" 2>&1; df -h /dev/shm >> "$out/host.txt" 2>&1
timeout 1150 python tools/e2e/run_scale.py --gpus 8 --seconds 30 --qps-per-replica 8 --users-per-replica 24 --qps-sweep 16 \
    --cpu-gb 20 --pd-requests 16 --log-dir "$out/scale8" 2>&1 | cut -c1-1200
M=gpu__time_duration.sum,nvlrx__bytes.sum,nvlrx__bytes_data_user.sum,nvltx__bytes.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 120 ncu --metrics $M --clock-control none -k regex:kv_ --csv --log-file "$out/ncu_nvlink_tier.csv" python tools/prof_tier.py 2 > "$out/prof_tier.log" 2>&1
tail -3 "$out/prof_tier.log"
rm -f "$out"/scale8/vllm_none_[2-7].log      # keep the merge-back small: one none log and all kv logs are enough
ls "$out/scale8" | head -80

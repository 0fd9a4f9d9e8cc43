#!/bin/sh
# One gpurun call per stage of DESIGN.md §9 (run from the repo root):
#   gpurun --timeout 900           -- tools/round2.sh kernels     # 1 GPU
#   gpurun --gpus 2 --timeout 1500 -- tools/round2.sh multi2      # 2 GPUs
#   gpurun --gpus 2 --timeout 1500 -- tools/round2.sh routing2
#   gpurun --gpus 8 --timeout 1200 -- tools/round2.sh scale8      # 8 GPUs
# Everything is written under gpurun_out/r2/ (merged back by gpurun).
set -u
out=gpurun_out/r2
mkdir -p "$out"
stage="${1:-kernels}"
case "$stage" in
  kernels)
    python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee "$out/pytest_gpu.txt"   # no -x: collect every failure of the experimental paths in one go
    python tools/microbench.py --out "$out/microbench.json" > "$out/microbench.log" 2>&1
    if grep -q B200KV_FP8_2PASS production-stack_b200/csrc/b200kv_engine.cu; then   # experimental kernel present
      B200KV_FP8_2PASS=1 python tools/microbench.py --out "$out/microbench_2pass.json" > "$out/microbench_2pass.log" 2>&1
    fi
    python bench.py > "$out/bench_n1.json" 2> "$out/bench_n1.err"
    # BASELINE configs[1] with the UNMODIFIED harness (reference numbers so far came from tools/e2e/mrqa_driver.py)
    python tools/e2e/run_e2e.py --harness --harness-time 90 --modes none,b200kv,offload --log-dir "$out/e2e_harness" 2>&1 | cut -c1-500
    tail -c 400 "$out/bench_n1.json"
    ;;
  multi2)
    [ -f tests/test_gpu_device_tier.py ] && python -m pytest tests/test_gpu_device_tier.py -m gpu -q 2>&1 | tail -3 | tee "$out/pytest_tier.txt"
    modes=none,private,shared
    grep -q '"tier"' tools/e2e/run_multi.py 2>/dev/null || grep -q 'startswith("tier")' tools/e2e/run_multi.py && modes=none,private,shared,tier
    python tools/e2e/run_multi.py --replicas 2 --routing roundrobin --modes $modes --log-dir "$out/multi2_rr" 2>&1 | cut -c1-600
    ;;
  routing2)
    python tools/e2e/run_multi.py --replicas 2 --routing roundrobin --modes remote --log-dir "$out/multi2_remote" 2>&1 | cut -c1-600
    python tools/e2e/run_multi.py --replicas 2 --routing kvaware --modes none,private --log-dir "$out/multi2_kvaware" 2>&1 | cut -c1-600
    python tools/e2e/run_multi.py --replicas 2 --routing pd --modes private --user-history-prompt 7000 --shared-system-prompt 1000 \
        --max-model-len 12288 --num-users 8 --num-rounds 2 --qps 1 --log-dir "$out/multi2_pd" 2>&1 | cut -c1-600
    ;;
  scale8)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        bench.py --gpus 8 > "$out/bench_n8.json" 2> "$out/bench_n8.err"
    python tools/e2e/run_pd.py --prefill 4 --decode 4 --concurrency 4 --requests 16 --log-dir "$out/pd_4p4d" 2>&1 | cut -c1-500
    python tools/e2e/run_multi.py --replicas 8 --routing session --modes none,shared --num-users 64 --num-rounds 4 --qps 8 \
        --log-dir "$out/multi8_session" 2>&1 | cut -c1-600
    ;;
  *) echo "unknown stage $stage"; exit 2 ;;
esac

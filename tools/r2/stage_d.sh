#!/bin/sh
# Round-2 stage D (1 GPU): validate the step-batched ops and the reworked Q4 kernels before the 8-GPU session.
set -u
out=gpurun_out/r2d
mkdir -p "$out"
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_q4.py -m gpu -q --timeout 600 -p no:cacheprovider \
    -k "batch or q4 or refused or store_retrieve" > "$out/pytest_kernels.txt" 2>&1
tail -6 "$out/pytest_kernels.txt"
timeout 600 python -m pytest tests/test_gpu_vllm_connector.py -m gpu -q -rA -s --timeout 600 -p no:cacheprovider -k "eager or fp8" > "$out/pytest_vllm.txt" 2>&1
grep -E "PASSED|FAILED|ERROR|raw \(|fp8:|passed|failed" "$out/pytest_vllm.txt" | tail -6
python tools/microbench.py --no-torch-baseline --iters 30 --out "$out/microbench.json" > "$out/microbench.log" 2>&1
grep -E '"n_tok": 32768.*q4' "$out/microbench.log" | cut -c1-230

#!/bin/sh
# Round-2 stage B (1 GPU): the new FP8 / Q4 kernels (parity, sanitizer, microbench, ncu) and the connector inside vLLM.
#   gpurun --timeout 1500 -- tools/r2/stage_b.sh
set -u
out=gpurun_out/r2b
mkdir -p "$out"
export B200KV_VERBOSE=1
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_q4.py -m gpu -q -x --timeout 600 -p no:cacheprovider \
    -k "fp8 or q4 or refused or store_retrieve or layerwise or batch" > "$out/pytest_kernels.txt" 2>&1
tail -15 "$out/pytest_kernels.txt"
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider \
    -k "persistent and (L3H8D128 or L2H4D64) and (257 or 1024)" > "$out/sanitizer_memcheck.log" 2>&1
tail -3 "$out/sanitizer_memcheck.log"
timeout 600 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider \
    -k "persistent and L3H8D128 and 257" > "$out/sanitizer_racecheck.log" 2>&1
tail -3 "$out/sanitizer_racecheck.log"
python tools/microbench.py --no-torch-baseline --out "$out/microbench.json" > "$out/microbench.log" 2>&1
grep -E '"n_tok": 32768' "$out/microbench.log" | cut -c1-220
grep -E 'requests_per_step' "$out/microbench.log" | cut -c1-400
B200KV_FP8_TMAP_HND=1 python tools/microbench.py --no-torch-baseline --iters 30 --out "$out/microbench_tmap_hnd.json" > "$out/microbench_tmap_hnd.log" 2>&1
grep -E '"n_tok": 32768.*fp8' "$out/microbench_tmap_hnd.log" | cut -c1-220
PROF_HND=1 PROF_FORMATS=fp8,q4 ncu --set full --clock-control none --import-source on -k regex:kv_ -o "$out/prof_hnd" -f \
    python tools/prof_kernels.py 1 > "$out/prof_hnd.log" 2>&1
PROF_FORMATS=fp8 ncu --set full --clock-control none --import-source on -k regex:kv_fp8_store -o "$out/prof_nhd" -f \
    python tools/prof_kernels.py 1 > "$out/prof_nhd.log" 2>&1
if [ "${1:-}" != "novllm" ]; then
timeout 900 python -m pytest tests/test_gpu_vllm_connector.py -m gpu -q -rA -s --timeout 900 -p no:cacheprovider > "$out/pytest_vllm.txt" 2>&1
grep -E "PASSED|FAILED|ERROR|raw \(|fp8:|passed|failed" "$out/pytest_vllm.txt" | tail -15
fi
ls "$out"

#!/bin/sh
# Round-2 stage A (1 GPU): the whole GPU test suite (incl. the connector inside a real vLLM engine),
# kernel microbenchmarks (default and two-pass FP8 store), bench.py, ncu launch list + full captures.
#   gpurun --timeout 1500 -- tools/r2/stage_a.sh
set -u
out=gpurun_out/r2a
mkdir -p "$out"
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > "$out/gpu.txt" 2>&1
lscpu | grep -i -E "numa|socket|model name|^CPU\(s\)" > "$out/host.txt" 2>&1
nvidia-smi topo -m >> "$out/host.txt" 2>&1
timeout 1100 python -m pytest tests -m gpu -q -rA --timeout 900 -p no:cacheprovider > "$out/pytest_gpu.txt" 2>&1
tail -5 "$out/pytest_gpu.txt"
grep -E "^(PASSED|FAILED|ERROR)|passed|failed" "$out/pytest_gpu.txt" | grep -v PASSED | head -40
python tools/microbench.py --out "$out/microbench.json" > "$out/microbench.log" 2>&1
B200KV_FP8_2PASS=1 python tools/microbench.py --no-torch-baseline --out "$out/microbench_2pass.json" > "$out/microbench_2pass.log" 2>&1
python bench.py > "$out/bench_n1.json" 2> "$out/bench_n1.err"
tail -c 600 "$out/bench_n1.json"
python bench.py --impl reference --steps 5 --warmup 3 > "$out/bench_ref.json" 2> "$out/bench_ref.err"
# ncu: launch list of the bench command, then full captures of the FP8 / Q4 kernels (HND tiles = vLLM on B200)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$out/launches.csv" \
    python bench.py --steps 2 --warmup 3 --no-fp8 > "$out/bench_under_ncu.log" 2>&1
PROF_HND=1 PROF_FORMATS=fp8,q4 ncu --set full --clock-control none --import-source on -k regex:kv_ -o "$out/prof_fp8_hnd" -f \
    python tools/prof_kernels.py 1 > "$out/prof_fp8_hnd.log" 2>&1
PROF_HND=1 PROF_FORMATS=fp8 B200KV_FP8_2PASS=1 ncu --set full --clock-control none --import-source on -k regex:kv_fp8_store -o "$out/prof_fp8_2pass_hnd" -f \
    python tools/prof_kernels.py 1 > "$out/prof_fp8_2pass_hnd.log" 2>&1
PROF_FORMATS=fp8 ncu --set full --clock-control none --import-source on -k regex:kv_fp8_store -o "$out/prof_fp8_nhd" -f \
    python tools/prof_kernels.py 1 > "$out/prof_fp8_nhd.log" 2>&1
ls -la "$out"

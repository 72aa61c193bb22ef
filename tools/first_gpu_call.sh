#!/usr/bin/env bash
# First GPU call of a round, as ONE gpurun command (everything lands in gpurun_out/):
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/first_gpu_call.sh'
# 1. the hardware-verified GPU tests, then the never-run-on-hardware file on its own (XPASS / xfail per test with -rxX)
# 2. bench line (sync-free step) and the A/B against the device read-back of the layout meta words
# 3. ncu launch list of the bench command + loader / resident data-set timings
set -u
mkdir -p gpurun_out
./tests/native/check_new_kernels 2>&1 | tee gpurun_out/native_check.log
./tests/native/fused_step_harness 10000 300 2 2>&1 | tee gpurun_out/native_fused_step.log     # tile packing A/B, no Python
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_zz_first_run.py > gpurun_out/pytest_gpu.log 2>&1
echo "pytest(verified tiers) rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python -m pytest tests/test_gpu_zz_first_run.py -m gpu -q -rxX --runxfail > gpurun_out/pytest_first_run.log 2>&1
echo "pytest(first-run file, --runxfail) rc=$?"; tail -25 gpurun_out/pytest_first_run.log
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_sync_free.json 2> gpurun_out/bench_sync_free.err
echo "bench rc=$?"; cat gpurun_out/bench_sync_free.json
DMPNN_HOST_META=0 timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu --no-dataset > gpurun_out/bench_readback.json 2> gpurun_out/bench_readback.err
echo "bench(read-back) rc=$?"; cat gpurun_out/bench_readback.json
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu --no-dataset --no-pack > gpurun_out/bench_nopack.json 2> gpurun_out/bench_nopack.err
echo "bench(generator order, no tile packing) rc=$?"; cat gpurun_out/bench_nopack.json
(PACK=0 timeout 120 python tools/time_step.py; PACK=1 timeout 120 python tools/time_step.py) > gpurun_out/time_step_pack.log 2>&1
echo "fused step alone, generator vs tile-packed order:"; cat gpurun_out/time_step_pack.log
timeout 300 python tools/time_loader.py > gpurun_out/time_loader.log 2>&1; echo "time_loader rc=$?"; cat gpurun_out/time_loader.log
timeout 400 python tools/bench_configs.py C4 C2a C3 --steps 3 > gpurun_out/bench_configs.log 2>&1; echo "other configs rc=$?"; cat gpurun_out/bench_configs.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-dataset > gpurun_out/bench_under_ncu.log 2>&1
echo "ncu rc=$?"

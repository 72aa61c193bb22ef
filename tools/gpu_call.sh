#!/usr/bin/env bash
# One GPU call of round 2 as ONE gpurun command (everything lands in gpurun_out/<tag>/):
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_call.sh <tag> [steps...]'
# steps: tests bench launches ncu_fused ncu_gemm native  (default: tests bench launches)
set -u
TAG=${1:-call}; shift || true
STEPS=${*:-tests bench launches}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > "$OUT/gpu.txt" 2>&1
for s in $STEPS; do
  case $s in
    tests)
      timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -rfEX --durations=15 > "$OUT/pytest_gpu.log" 2>&1
      echo "pytest rc=$?"; tail -40 "$OUT/pytest_gpu.log" ;;
    testsx)
      timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider > "$OUT/pytest_gpu_x.log" 2>&1
      echo "pytest -x rc=$?"; tail -5 "$OUT/pytest_gpu_x.log" ;;
    smoke)
      python __graft_entry__.py --smoke > "$OUT/smoke.log" 2>&1; echo "smoke rc=$?"; tail -3 "$OUT/smoke.log" ;;
    bench)
      timeout 600 python bench.py --steps 20 --warmup 5 > "$OUT/bench.json" 2> "$OUT/bench.err"
      echo "bench rc=$?"; cat "$OUT/bench.json"; tail -5 "$OUT/bench.err" ;;
    launches)
      timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file "$OUT/launches.csv" \
        python bench.py --steps 2 --warmup 3 --no-cpu --no-dataset > "$OUT/bench_under_ncu.log" 2>&1
      echo "ncu launches rc=$?" ;;
    launches_c4)
      timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 400 --csv --log-file "$OUT/launches_c4.csv" \
        python bench.py --config C4 --steps 2 --warmup 3 --no-cpu --no-graph > "$OUT/bench_c4_under_ncu.log" 2>&1
      echo "ncu launches C4 rc=$?" ;;
    ncu_fused_c4)
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_bond_step_fused -s 12 -c 4 \
        -o "$OUT/atom_step" -f python bench.py --config C4 --steps 1 --warmup 3 --no-cpu > "$OUT/ncu_fused_c4.log" 2>&1
      echo "ncu fused C4 rc=$?" ;;
    ncu_fused)
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_bond_step_fused -s 12 -c 4 \
        -o "$OUT/fused_step" -f python bench.py --steps 1 --warmup 3 --no-cpu --no-dataset > "$OUT/ncu_fused.log" 2>&1
      echo "ncu fused rc=$?" ;;
    ncu_glue)
      timeout 900 ncu --set full --clock-control none --import-source on \
        -k "regex:k_act_bwd_v4|k_segment_sum_flat8|k_concat_bf16_v4|k_wgrad_reduce|k_segment_sum_v4|k_segment_bcast|k_tiles_chunk" -s 27 -c 10 \
        -o "$OUT/glue" -f python bench.py --steps 1 --warmup 3 --no-cpu --no-dataset > "$OUT/ncu_glue.log" 2>&1
      echo "ncu glue rc=$?" ;;
    ncu_gemm)
      timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_linear_tc|k_wgrad_tc" -s 18 -c 6 \
        -o "$OUT/gemm" -f python bench.py --steps 1 --warmup 3 --no-cpu --no-dataset > "$OUT/ncu_gemm.log" 2>&1
      echo "ncu gemm rc=$?" ;;
    x3)
      for k in linear_x3_vs_f64 transposed wgrad_x3_vs_f64 deterministic; do
        timeout 600 python -m pytest tests/test_gpu_x3.py -m gpu -q -x -p no:cacheprovider -k "$k" > "$OUT/pytest_x3_$k.log" 2>&1
        echo "x3[$k] rc=$?"; tail -4 "$OUT/pytest_x3_$k.log"
      done ;;
    x3san)
      timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_x3.py -m gpu -q -x -p no:cacheprovider \
        -k "linear_x3_vs_f64 and 300-300-300" > "$OUT/san_linear.log" 2>&1
      echo "san linear rc=$?"; grep -m 30 -A12 "Invalid\|Error\|error" "$OUT/san_linear.log" | head -80
      timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_x3.py -m gpu -q -x -p no:cacheprovider \
        -k "wgrad_x3_vs_f64 and 31-300-372" > "$OUT/san_wgrad.log" 2>&1
      echo "san wgrad rc=$?"; grep -m 30 -A12 "Invalid\|Error\|error" "$OUT/san_wgrad.log" | head -80 ;;
    tests_nox3)
      DMPNN_X3=0 timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -rfEX --deselect tests/test_gpu_x3.py > "$OUT/pytest_gpu_nox3.log" 2>&1
      echo "pytest (DMPNN_X3=0) rc=$?"; tail -30 "$OUT/pytest_gpu_nox3.log" ;;
    bench_c4)
      timeout 600 python bench.py --config C4 --steps 10 --warmup 3 > "$OUT/bench_c4.json" 2> "$OUT/bench_c4.err"
      echo "bench C4 rc=$?"; cat "$OUT/bench_c4.json"; tail -5 "$OUT/bench_c4.err" ;;
    bench_c3)
      timeout 900 python bench.py --config C3 --steps 3 --warmup 3 > "$OUT/bench_c3.json" 2> "$OUT/bench_c3.err"
      echo "bench C3 rc=$?"; cat "$OUT/bench_c3.json"; tail -5 "$OUT/bench_c3.err" ;;
    bench_c5)
      timeout 900 python bench.py --config C5 --steps 3 --warmup 3 > "$OUT/bench_c5.json" 2> "$OUT/bench_c5.err"
      echo "bench C5 rc=$?"; cat "$OUT/bench_c5.json"; tail -5 "$OUT/bench_c5.err" ;;
    variants)
      for d in chemprop_b200/lib/variants/*/; do
        tag=$(basename "$d")
        echo "== variant $tag"
        LD_LIBRARY_PATH="$d" timeout 120 ./tests/native/fused_step_harness 10000 300 1 2>&1 | tee "$OUT/variant_$tag.log" | tail -6
      done ;;
    trace)
      timeout 300 python tools/trace_step.py > "$OUT/trace_block0.log" 2>&1; echo "trace rc=$?"; cat "$OUT/trace_block0.log"
      DMPNN_TRACE_BLOCK=77 timeout 300 python tools/trace_step.py > "$OUT/trace_block77.log" 2>&1 ;;
    bench_n2)
      timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        bench.py --gpus 2 --steps 10 --warmup 3 > "$OUT/bench_n2.json" 2> "$OUT/bench_n2.err"
      echo "bench N=2 rc=$?"; cat "$OUT/bench_n2.json"; tail -8 "$OUT/bench_n2.err"
      timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
        bench.py --config C5 --gpus 2 --steps 3 --warmup 3 > "$OUT/bench_c5_n2.json" 2> "$OUT/bench_c5_n2.err"
      echo "bench C5 N=2 rc=$?"; cat "$OUT/bench_c5_n2.json"; tail -8 "$OUT/bench_c5_n2.err" ;;
    bench_ref)
      timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > "$OUT/bench_ref.json" 2> "$OUT/bench_ref.err"
      echo "bench reference rc=$?"; cat "$OUT/bench_ref.json"; tail -5 "$OUT/bench_ref.err" ;;
    exps)
      # timing experiments of the fused depth step: the library built with -DDMPNN_EXPERIMENTS=1 (tools/build_variants.sh -> variants/exp)
      L=chemprop_b200/lib/variants/exp
      for e in ${DMPNN_EXPS:-0 4096 8192 16384 32768 49152 65536 131072 196608 262144}; do
        echo "== DMPNN_EXP=$e"
        LD_LIBRARY_PATH=$L DMPNN_EXP=$e timeout 120 ./tests/native/fused_step_harness 10000 300 1 2>&1 | tee "$OUT/exp_$e.log" | tail -3
      done ;;
    prof_e2e)
      timeout 300 python tools/profile_e2e.py 30 > "$OUT/profile_e2e.log" 2>&1; echo "profile_e2e rc=$?"; head -60 "$OUT/profile_e2e.log" ;;
    native)
      ./tests/native/fused_step_harness 10000 300 2 2>&1 | tee "$OUT/native_fused_step.log" ;;
    *) echo "unknown step $s" ;;
  esac
done

#!/bin/bash
# Round-2 GPU session driver: tools/gpu_session.sh <stage> ...   (run on the GPU box through gpurun)
# Every stage writes under gpurun_out/; an ncu / sanitizer stage only runs after the same command exited 0 plainly.
set -u
mkdir -p gpurun_out
tag=${TAG:-r02}
for stage in "$@"; do
  echo "=== stage $stage ($(date +%T))"
  case $stage in
    tests)
      python -m pytest tests -m gpu -q -x --durations=8 > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/${tag}_pytest_gpu.log ;;
    tests_all)
      python -m pytest tests -m gpu -q --durations=8 --deselect tests/test_gpu_fullsize.py::test_c5_top_level_rows_and_fiedler_value > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/${tag}_pytest_gpu.log ;;
    bench_small)
      for w in c1 c2 c3; do
        python bench.py --workload $w --steps 5 --warmup 3 > gpurun_out/${tag}_bench_$w.json 2> gpurun_out/${tag}_bench_$w.err; echo "bench $w rc=$?"
      done ;;
    bench)
      python bench.py --steps 5 --warmup 3 > gpurun_out/${tag}_bench_c4.json 2> gpurun_out/${tag}_bench_c4.err; echo "bench c4 rc=$?"; head -c 600 gpurun_out/${tag}_bench_c4.json ;;
    bench32)
      python bench.py --steps 5 --warmup 3 --small-limit 32 --no-cpu-baseline > gpurun_out/${tag}_bench_c4_small32.json 2> gpurun_out/${tag}_bench_c4_small32.err; echo "bench c4 small32 rc=$?"; head -c 300 gpurun_out/${tag}_bench_c4_small32.json ;;
    launches)
      python bench.py --profile-step > gpurun_out/${tag}_profile_step.json 2> gpurun_out/${tag}_profile_step.err; rc=$?; echo "profile-step rc=$rc"; cat gpurun_out/${tag}_profile_step.json
      if [ $rc -eq 0 ]; then
        ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
            --log-file gpurun_out/${tag}_launches_step.csv python bench.py --profile-step > gpurun_out/${tag}_launches_step.out 2>&1; echo "ncu launches rc=$?"
        python tools/ncu_summary.py launches gpurun_out/${tag}_launches_step.csv > gpurun_out/${tag}_launches_step_summary.csv; head -30 gpurun_out/${tag}_launches_step_summary.csv
      fi ;;
    wavetrace)
      SCS_DRIVER_TRACE=1 python tools/driver_profile.py c4 > gpurun_out/${tag}_wavetrace.log 2>&1; echo "wavetrace rc=$?"; tail -45 gpurun_out/${tag}_wavetrace.log ;;
    ncufull)
      # one --set full capture per heavy kernel of the profiled step (bench.py --profile-step exited 0 in stage launches)
      for k in ${KERNELS:-matvec_rows pcg_rows_kernel small_batch_kernel}; do
        ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$k -c 3 -f \
            -o gpurun_out/${tag}_ncu_$k python bench.py --profile-step > gpurun_out/${tag}_ncu_$k.out 2>&1; echo "ncu $k rc=$?"
        ncu -i gpurun_out/${tag}_ncu_$k.ncu-rep --page raw --csv > gpurun_out/${tag}_ncu_${k}_raw.csv 2>/dev/null
        ncu -i gpurun_out/${tag}_ncu_$k.ncu-rep --page details > gpurun_out/${tag}_ncu_${k}_details.txt 2>/dev/null
      done
      python tools/ncu_summary.py traffic gpurun_out/${tag}_ncu_matvec_rows_raw.csv matvec_rows gpurun_out/${tag}_ncu_matvec_traffic.json ;;
    c5)
      python tools/run_workload.py c5 --repeat 2 --out gpurun_out/${tag}_c5_n1.json > gpurun_out/${tag}_c5_n1.log 2>&1; echo "c5 n1 rc=$?"; cut -c1-700 gpurun_out/${tag}_c5_n1.json ;;
    multi)
      # N = $GPUS ranks on one box: the bench line, c5 through the product path, tree-sharded W + all-reduce beside row sharding
      N=${GPUS:-2}
      python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/${tag}_bench_c4_n$N.json 2> gpurun_out/${tag}_bench_c4_n$N.err; echo "bench c4 n$N rc=$?"; head -c 400 gpurun_out/${tag}_bench_c4_n$N.json; tail -3 gpurun_out/${tag}_bench_c4_n$N.err
      [ -z "${SKIP_TREESHARD:-}" ] && python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/tree_sharded_allreduce.py c4 > gpurun_out/${tag}_treeshard_c4_n$N.json 2> gpurun_out/${tag}_treeshard_c4_n$N.err; echo "treeshard n$N rc=$?"; cut -c1-600 gpurun_out/${tag}_treeshard_c4_n$N.json
      if [ -z "${SKIP_C5:-}" ]; then
        python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/run_workload.py c5 --repeat 2 --out gpurun_out/${tag}_c5_n$N.json > gpurun_out/${tag}_c5_n$N.log 2>&1; echo "c5 n$N rc=$?"; cut -c1-700 gpurun_out/${tag}_c5_n$N.json; tail -3 gpurun_out/${tag}_c5_n$N.log
      fi ;;
    sanitize)
      python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; rc=$?; echo "smoke rc=$rc"; tail -2 gpurun_out/${tag}_smoke.log
      if [ $rc -eq 0 ]; then
        timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_sanitizer_memcheck_smoke.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/${tag}_sanitizer_memcheck_smoke.log
        timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_sanitizer_racecheck_smoke.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/${tag}_sanitizer_racecheck_smoke.log
      fi ;;
    parity)
      python tools/parity_report.py --out gpurun_out/${tag}_PARITY.json > gpurun_out/${tag}_parity.log 2>&1; echo "parity rc=$?"; tail -3 gpurun_out/${tag}_parity.log ;;
    c5test)
      python -m pytest tests/test_gpu_fullsize.py -q -k c5 -s > gpurun_out/${tag}_pytest_c5.log 2>&1; echo "c5 test rc=$?"; tail -5 gpurun_out/${tag}_pytest_c5.log ;;
    reference)
      python bench.py --impl reference --steps 1 --warmup 0 --reference-budget-s 1200 > gpurun_out/${tag}_bench_c4_reference.json 2> gpurun_out/${tag}_bench_c4_reference.err; echo "reference rc=$?"; head -c 400 gpurun_out/${tag}_bench_c4_reference.json ;;
    *) echo "unknown stage $stage" ;;
  esac
done
echo "=== done ($(date +%T))"

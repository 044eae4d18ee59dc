#!/bin/bash
# One GPU-box session that produces every number DESIGN.md / BASELINE.md quote (copied to profiles/ afterwards).
# usage (on the GPU box, repo root):  bash scripts/final_measurements.sh
set -u
out=gpurun_out/final
mkdir -p $out
python -m pytest tests -m gpu -q > $out/gputest.log 2>&1; tail -3 $out/gputest.log
python bench.py --steps 20 --warmup 5 > $out/bench.json 2> $out/bench.err
python bench.py --steps 200 --warmup 5 --no-traffic > $out/bench_200steps.json 2>> $out/bench.err
python bench.py --impl reference --steps 5 --warmup 2 > $out/bench_reference.json 2>> $out/bench.err
for cfg in C1 C2 C3 C4 C5; do
  python bench.py --config $cfg --steps 20 --warmup 5 --no-traffic > $out/bench_$cfg.json 2>> $out/bench.err
done
python bench.py --config C2 --trunc 4.0 --steps 20 --warmup 5 --no-traffic > $out/bench_C2_trunc4.json 2>> $out/bench.err
python scripts/timeline_probe.py > $out/pipeline_timeline_device.json 2>> $out/bench.err
python scripts/timeline_probe.py --host > $out/pipeline_timeline_host.json 2>> $out/bench.err
python scripts/submit_probe.py > $out/pipeline_submit.json 2>> $out/bench.err
# launch list (per-launch durations, serialised and cold: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches_bench_steps3.csv \
    python bench.py --steps 3 --warmup 1 --no-traffic > $out/ncu_bench.log 2>&1
# full captures of the kernels that matter: the one-launch sort, the fold, the long-run apply, the bundle order
for k in k_sort k_merge k_apply_long k_bundle_order; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 2 -o $out/ncu_$k -f \
      python bench.py --steps 3 --warmup 1 --no-traffic > $out/ncu_$k.log 2>&1
  ncu -i $out/ncu_$k.ncu-rep --page raw --csv > $out/${k}_ncu_full_raw.csv 2>/dev/null
  rm -f $out/ncu_$k.ncu-rep
done
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_sort_gpu.py tests/test_order_gpu.py -q -x \
    -k "not more_tiles and not 1048576" > $out/memcheck_sort_order.log 2>&1; tail -3 $out/memcheck_sort_order.log
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_tsdf_gpu.py -q -x -k "async or pool" \
    > $out/memcheck_tsdf.log 2>&1; tail -3 $out/memcheck_tsdf.log
ls -la $out | head -50

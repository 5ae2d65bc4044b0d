#!/bin/bash
# Round-2 evidence refresh (final kernels): tests, the bench lines of every BASELINE config, per-layer tables, ncu launch lists,
# sanitizer passes.  The `ncu --set full` captures of tools/evidence_r02.sh are kept (those kernels did not change).  Everything lands
# in gpurun_out/ev2/ ; tools/evidence_summary_r02.py turns it into profiles/*_r02.*.  No throughput number is taken under a profiler.
O=gpurun_out/ev2; mkdir -p $O
timeout 400 python -m pytest tests -m gpu -q > $O/pytest_gpu.txt 2>&1; tail -2 $O/pytest_gpu.txt
timeout 420 python bench.py > $O/bench.json 2> $O/bench.err; tail -c 300 $O/bench.json
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
for w in resnet3d50 r2plus1d34 nonlocal50 resnet18 biggan256 trn; do
  timeout 150 python bench.py --workload $w --no-cpu --no-biggan --no-others --steps 30 --warmup 5 --layers > $O/line_$w.json 2> $O/layers_$w.txt
done
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_resnet3d50.csv python tools/fwd_once.py resnet3d50 > $O/ncu_list1.log 2>&1
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_r2plus1d34.csv python tools/fwd_once.py r2plus1d34 > $O/ncu_list2.log 2>&1
bash tools/sanitize.sh all > $O/sanitize.txt 2>&1; tail -20 $O/sanitize.txt
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,power.limit --format=csv > $O/gpu.txt
ls -la $O | head -40

#!/bin/bash
# Round evidence on the GPU box: tests, bench lines, per-layer tables, ncu launch list and --set full captures.
# Everything lands in gpurun_out/ev/ ; tools/evidence_summary.py turns it into profiles/.
O=gpurun_out/ev; mkdir -p $O
python -m pytest tests -m gpu -q > $O/pytest_gpu.txt 2>&1
python bench.py > $O/bench.json 2> $O/bench.err
python bench.py --no-cpu --steps 50 --layers > /dev/null 2> $O/layers.txt
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
python tools/bench_others.py --layers > $O/others.jsonl 2> $O/others_layers.txt
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_all.csv \
    python bench.py --no-cpu --steps 2 --warmup 3 > $O/bench_under_ncu.txt 2>&1
N="ncu --set full --clock-control none --import-source on -s 3 -c 1"
$N -k regex:stemconv -o $O/ncu_stem python tools/conv_micro.py 32 3 16 224 224 64 7 7 7 1 2 2 2 > /dev/null 2>&1
$N -k regex:slabconv -o $O/ncu_slab64 python tools/conv_micro.py 32 64 8 56 56 64 3 3 3 1 1 1 2 > /dev/null 2>&1
$N -k regex:slabconv -o $O/ncu_slab128 python tools/conv_micro.py 32 128 4 28 28 128 3 3 3 1 1 1 2 > /dev/null 2>&1
$N -k regex:pgemm -o $O/ncu_pgemm_64_256 python tools/conv_micro.py 32 64 8 56 56 256 1 1 1 1 1 1 2 > /dev/null 2>&1
$N -k regex:slabconv -o $O/ncu_slab_r2p1d_144 python tools/conv_micro.py 16 64 16 28 28 144 1 3 3 1 1 1 2 > /dev/null 2>&1
$N -k regex:attention -o $O/ncu_attention python tools/att_micro.py 8 6272 256 256 2 > /dev/null 2>&1
for f in $O/ncu_*.ncu-rep; do ncu -i $f --page raw --csv > ${f%.ncu-rep}.raw.csv 2>/dev/null; done
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,power.limit --format=csv > $O/gpu.txt
ls -la $O

#!/bin/bash
# Round-2 evidence on the GPU box: tests, the bench lines of every BASELINE config, per-layer tables, the ncu launch list and
# --set full captures of the kernels the step is made of.  Everything lands in gpurun_out/ev2/ ; tools/evidence_summary_r02.py
# turns it into profiles/*_r02.*.  No throughput number is taken under a profiler.
O=gpurun_out/ev2; mkdir -p $O
python -m pytest tests -m gpu -q > $O/pytest_gpu.txt 2>&1; tail -2 $O/pytest_gpu.txt
python bench.py > $O/bench.json 2> $O/bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
for w in resnet3d50 r2plus1d34 nonlocal50 resnet18 biggan256; do
  python bench.py --workload $w --no-cpu --no-biggan --no-others --steps 30 --warmup 5 --layers > $O/line_$w.json 2> $O/layers_$w.txt
done
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_resnet3d50.csv python tools/fwd_once.py resnet3d50 > $O/ncu_list1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_r2plus1d34.csv python tools/fwd_once.py r2plus1d34 > $O/ncu_list2.log 2>&1
N="ncu --set full --clock-control none --import-source on -c 1"
$N -s 1 -k regex:stemconv -o $O/ncu_stem_poolw python tools/fwd_once.py resnet3d50 > /dev/null 2>&1
$N -s 3 -k regex:slabconv -o $O/ncu_slab64 python tools/conv_micro.py 32 64 8 56 56 64 3 3 3 1 1 1 2 > /dev/null 2>&1
$N -s 3 -k regex:slabconv -o $O/ncu_slab128 python tools/conv_micro.py 32 128 4 28 28 128 3 3 3 1 1 1 2 > /dev/null 2>&1
$N -s 3 -k regex:pgemm -o $O/ncu_pgemm_64_256 python tools/conv_micro.py 32 64 8 56 56 256 1 1 1 1 1 1 2 > /dev/null 2>&1
$N -s 3 -k regex:densem -o $O/ncu_densem_l4 python tools/conv_micro.py 16 512 2 4 4 1152 1 3 3 1 1 1 2 > /dev/null 2>&1
$N -s 3 -k regex:slabconv -o $O/ncu_slab_r2p1d_144 python tools/conv_micro.py 16 64 16 28 28 144 1 3 3 1 1 1 2 > /dev/null 2>&1
$N -s 3 -k regex:attention -o $O/ncu_attention python tools/att_micro.py 8 6272 256 256 2 > /dev/null 2>&1
$N -s 3 -k regex:pgemm -o $O/ncu_gan_conv1_bn1A python tools/gan_micro.py 64 256 128 128 64 1 0 1 2 1 > /dev/null 2>&1
$N -s 3 -k regex:slabconv -o $O/ncu_gan_conv64 python tools/gan_micro.py 64 64 256 256 64 3 0 1 2 > /dev/null 2>&1
for f in $O/ncu_*.ncu-rep; do ncu -i $f --page raw --csv > ${f%.ncu-rep}.raw.csv 2>/dev/null; done
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,power.limit --format=csv > $O/gpu.txt
rm -f $O/*.ncu-rep            # the raw CSV pages are what profiles/ keeps
ls -la $O | head -60

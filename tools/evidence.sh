#!/bin/bash
# Round evidence on the GPU box: tests, bench lines, per-layer tables, ncu launch lists and --set full captures.
# Everything lands in gpurun_out/ev/ ; tools/evidence_summary.py turns it into profiles/.
O=gpurun_out/ev; mkdir -p $O
python -m pytest tests -m gpu -q > $O/pytest_gpu.txt 2>&1
python bench.py > $O/bench.json 2> $O/bench.err
python bench.py --no-cpu --no-biggan --steps 50 --layers > /dev/null 2> $O/layers.txt
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
python bench.py --workload biggan256 --steps 20 --warmup 3 --layers > $O/biggan.json 2> $O/biggan_layers.txt
python bench.py --workload biggan256 --impl reference --steps 2 --warmup 1 > $O/biggan_reference.json 2> $O/biggan_reference.err
python tools/bench_others.py --layers > $O/others.jsonl 2> $O/others_layers.txt
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_all.csv \
    python bench.py --no-cpu --no-biggan --steps 2 --warmup 3 > $O/bench_under_ncu.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_biggan.csv \
    python bench.py --workload biggan256 --no-cpu --steps 1 --warmup 3 --batch 64 > $O/biggan_under_ncu.txt 2>&1
N="ncu --set full --clock-control none --import-source on -s 3 -c 1"
$N -k regex:stemconv -o $O/ncu_stem python tools/conv_micro.py 32 3 16 224 224 64 7 7 7 1 2 2 2 > /dev/null 2>&1
$N -k regex:slabconv -o $O/ncu_slab64 python tools/conv_micro.py 32 64 8 56 56 64 3 3 3 1 1 1 2 > /dev/null 2>&1
$N -k regex:slabconv -o $O/ncu_slab128 python tools/conv_micro.py 32 128 4 28 28 128 3 3 3 1 1 1 2 > /dev/null 2>&1
$N -k regex:pgemm -o $O/ncu_pgemm_64_256 python tools/conv_micro.py 32 64 8 56 56 256 1 1 1 1 1 1 2 > /dev/null 2>&1
$N -k regex:attention -o $O/ncu_attention python tools/att_micro.py 8 6272 256 256 2 > /dev/null 2>&1
# BigGAN-deep-256 kernels at the B=256 shapes (64 images per capture: same per-image work, a quarter of the replay time)
$N -k regex:slabconv -o $O/ncu_gan_conv64 python tools/gan_micro.py 64 64 256 256 64 3 0 1 2 > /dev/null 2>&1
$N -k regex:slabconv -o $O/ncu_gan_upconv64 python tools/gan_micro.py 64 64 128 128 64 3 1 1 2 > /dev/null 2>&1
$N -k regex:slabconv -o $O/ncu_gan_conv128 python tools/gan_micro.py 64 128 128 128 128 3 0 1 2 > /dev/null 2>&1
$N -k regex:pgemm -o $O/ncu_gan_pgemm_64_128 python tools/gan_micro.py 64 64 256 256 128 1 0 0 2 > /dev/null 2>&1
for f in $O/ncu_*.ncu-rep; do ncu -i $f --page raw --csv > ${f%.ncu-rep}.raw.csv 2>/dev/null; done
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,power.limit --format=csv > $O/gpu.txt
rm -f $O/*.ncu-rep            # the raw CSV pages are what profiles/ keeps; the reports would exceed the 64 MiB return limit
ls -la $O

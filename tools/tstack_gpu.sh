#!/bin/bash
# Temporal stack kernel on the GPU box: parity tests, single-layer sweep (never / always), the R(2+1)D-34 bench line.
O=gpurun_out/tstack; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "tstack or temporal or r2p1d" > $O/pytest_kernels.txt 2>&1; tail -12 $O/pytest_kernels.txt
timeout 200 python tools/conv_sweep.py tools/tstack_shapes.txt 0:0:0:-1:0 0:0:0:-1:1 > $O/sweep.txt 2>&1; cat $O/sweep.txt
timeout 300 python -m pytest tests/test_gpu_models.py -m gpu -q -k "r2plus1d" > $O/pytest_models.txt 2>&1; tail -4 $O/pytest_models.txt
timeout 200 python bench.py --workload r2plus1d34 --no-cpu --no-biggan --no-others --steps 30 --warmup 5 --layers > $O/line_r2plus1d34.json 2> $O/layers_r2plus1d34.txt
head -14 $O/layers_r2plus1d34.txt; tail -2 $O/layers_r2plus1d34.txt; python -c "
import json; d=json.loads(open('$O/line_r2plus1d34.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['parity'])"
B2_TSTACK=0 timeout 200 python bench.py --workload r2plus1d34 --no-cpu --no-biggan --no-others --steps 30 --warmup 5 > $O/line_r2plus1d34_off.json 2> /dev/null
python -c "
import json; d=json.loads(open('$O/line_r2plus1d34_off.json').read().strip().splitlines()[-1]); print('tstack off:', d['value'], d['ms_per_step'])"

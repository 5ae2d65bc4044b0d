#!/bin/bash
O=gpurun_out/x3; mkdir -p $O
python tools/conv_sweep.py tools/smallm_shapes.txt 0:0 1073741824:1 1073741824:2 1073741824:3 1073741824:4 1073741824:6 1073741824:8 > $O/sweep.txt 2>&1
cat $O/sweep.txt
B2_DENSEM_MAXM=0 python -m pytest tests/test_gpu_models.py tests/test_gpu_kernels.py -q -x --timeout 600 2>&1 | tail -5
B2_DENSEM_MAXM=0 python bench.py --steps 30 --warmup 5 --no-biggan --no-cpu --layers > $O/bench_poolw.json 2> $O/bench_poolw.err; tail -3 $O/bench_poolw.err
B2_DENSEM_MAXM=0 B2_STEM_POOLW=0 python bench.py --steps 30 --warmup 5 --no-biggan --no-cpu > $O/bench_nopoolw.json 2> $O/bench_nopoolw.err
python -c "
import json
for n in ('poolw','nopoolw'):
    d=json.load(open('$O/bench_%s.json'%n)); print(n, d['value'], d['ms_per_step'], d['parity'])"

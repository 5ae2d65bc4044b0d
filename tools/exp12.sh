#!/bin/bash
O=gpurun_out/x12; mkdir -p $O
python -m pytest tests/test_gpu_kernels.py -q -x --timeout 600 -k "slabts or slab_layer1 or temporal" 2>&1 | tail -6
for ts in 1 0; do echo "B2_SLABTS=$ts"; B2_SLABTS=$ts python tools/conv_sweep.py tools/ts_shapes.txt -1:0:0; done
for ts in 1 0; do B2_SLABTS=$ts python bench.py --steps 30 --warmup 5 --no-cpu --no-biggan --no-others > $O/r3d_ts$ts.json 2> $O/r3d_ts$ts.err; python -c "
import json; d=json.load(open('$O/r3d_ts$ts.json')); print('resnet3d50 slabts=$ts', round(d['value']), round(d['ms_per_step'],3), d['parity']['max_rel_err'])"; done

#!/bin/bash
O=gpurun_out/x5; mkdir -p $O
python tools/conv_sweep.py tools/layer_shapes.txt -1:0 > $O/sweep.txt 2>&1; cat $O/sweep.txt
python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -12 > $O/pytest.log; tail -4 $O/pytest.log
for w in r2plus1d34 resnet3d50 nonlocal50 resnet18; do python bench.py --workload $w --steps 20 --warmup 5 --no-cpu --no-biggan > $O/$w.json 2> $O/$w.err; python -c "
import json; d=json.load(open('$O/$w.json')); print('$w', round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), d['e2e'].get('uint8_frames_value'), d['parity']['max_rel_err'])"; done

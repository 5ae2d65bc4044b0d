#!/bin/bash
O=gpurun_out/x8; mkdir -p $O
python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -8 > $O/pytest.log; tail -3 $O/pytest.log
for w in resnet3d50 r2plus1d34 resnet18; do python bench.py --workload $w --steps 30 --warmup 5 --no-cpu --no-biggan --layers > $O/$w.json 2> $O/$w.err; python -c "
import json; d=json.load(open('$O/$w.json')); print('$w', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['parity']['max_rel_err'])"; done
head -3 $O/resnet3d50.err

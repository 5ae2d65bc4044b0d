#!/bin/bash
O=gpurun_out/x7; mkdir -p $O
python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -8 > $O/pytest.log; tail -3 $O/pytest.log
for pair in 1 0; do B2_STEM_PAIR=$pair python bench.py --workload r2plus1d34 --steps 20 --warmup 5 --no-cpu > $O/r2p1d_pair$pair.json 2> $O/r2p1d_pair$pair.err; python -c "
import json; d=json.load(open('$O/r2p1d_pair$pair.json')); print('r2plus1d34 pair=$pair', round(d['value']), round(d['ms_per_step'],3), d['parity']['max_rel_err'])"; done
python bench.py --steps 30 --warmup 5 --no-cpu --no-biggan --layers > $O/resnet3d50.json 2> $O/resnet3d50.err; python -c "
import json; d=json.load(open('$O/resnet3d50.json')); print('resnet3d50', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['parity']['max_rel_err'])"

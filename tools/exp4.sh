#!/bin/bash
O=gpurun_out/x4; mkdir -p $O
python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -15 > $O/pytest.log; tail -4 $O/pytest.log
python bench.py --workload r2plus1d34 --steps 20 --warmup 5 --no-cpu > $O/r2p1d.json 2> $O/r2p1d.err
python -c "
import json; d=json.load(open('$O/r2p1d.json')); print('r2plus1d34', d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e'].get('uint8_frames_value'), d['parity']['max_rel_err'])"
N="ncu --set full --clock-control none --import-source on -s 3 -c 1"
$N -k regex:slabconv -o $O/ncu_r2p1d_l1_temporal python tools/conv_micro.py 16 144 16 28 28 64 3 1 1 1 1 1 2 --residual > /dev/null 2>&1
$N -k regex:slabconv -o $O/ncu_r2p1d_l1_spatial python tools/conv_micro.py 16 64 16 28 28 144 1 3 3 1 1 1 2 > /dev/null 2>&1
$N -k regex:slabconv -o $O/ncu_r2p1d_l3_spatial python tools/conv_micro.py 16 256 4 7 7 576 1 3 3 1 1 1 2 > /dev/null 2>&1
ls -la $O

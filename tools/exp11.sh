#!/bin/bash
O=gpurun_out/x11; mkdir -p $O
python -m pytest tests/test_gpu_biggan.py -q -x --timeout 600 2>&1 | tail -8
for f in 1 0; do B2_GAN_FUSE_BN1=$f python bench.py --workload biggan256 --steps 20 --warmup 5 --no-cpu --layers > $O/gan_fuse$f.json 2> $O/gan_fuse$f.err; python -c "
import json; d=json.load(open('$O/gan_fuse$f.json')); print('biggan fuse=$f', round(d['value']), round(d['ms_per_step'],3), d['parity'])"; done
head -12 $O/gan_fuse1.err

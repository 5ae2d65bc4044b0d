#!/bin/bash
# dense-M (split-K cluster) kernel vs slab / persistent-GEMM kernels on the small-M layer shapes, with a numerics check each
O=gpurun_out/x1; mkdir -p $O
run() { B2_DENSEM_MAXM=$1 python tools/conv_micro.py ${@:2} 2>&1 | tail -1; }
for mm in 0 100000000; do
 echo "== B2_DENSEM_MAXM=$mm"
 run $mm 16 128 8 14 14 288 1 3 3 1 1 1 20 --check
 run $mm 16 288 8 14 14 128 3 1 1 1 1 1 20 --check --residual
 run $mm 16 256 4 7 7 576 1 3 3 1 1 1 20 --check
 run $mm 16 576 4 7 7 256 3 1 1 1 1 1 20 --check --residual
 run $mm 16 512 2 4 4 1152 1 3 3 1 1 1 20 --check
 run $mm 16 1152 2 4 4 512 3 1 1 1 1 1 20 --check --residual
 run $mm 16 256 8 14 14 921 1 3 3 1 2 2 20 --check
 run $mm 32 512 1 7 7 512 3 3 3 1 1 1 20 --check
 run $mm 32 512 2 14 14 512 3 3 3 2 2 2 20 --check
 run $mm 32 2048 1 7 7 512 1 1 1 1 1 1 20 --check
 run $mm 32 512 1 7 7 2048 1 1 1 1 1 1 20 --check --residual
 run $mm 32 256 2 14 14 256 3 3 3 1 1 1 20 --check
 run $mm 32 1024 2 14 14 256 1 1 1 1 1 1 20 --check
 run $mm 32 256 2 14 14 1024 1 1 1 1 1 1 20 --check --residual
 run $mm 2 64 16 28 28 144 1 3 3 1 1 1 20 --check
 run $mm 2 144 16 28 28 64 3 1 1 1 1 1 20 --check --residual
done > $O/smallm.txt 2>&1
echo "== split sweep (layer3 spatial r2plus1d, layer4 spatial)" >> $O/smallm.txt
for s in 1 2 3 6; do echo "S<=$s" >> $O/smallm.txt; B2_DENSEM_S=$s B2_DENSEM_MAXM=100000000 python tools/conv_micro.py 16 256 4 7 7 576 1 3 3 1 1 1 20 2>&1 | tail -1 >> $O/smallm.txt; B2_DENSEM_S=$s B2_DENSEM_MAXM=100000000 python tools/conv_micro.py 16 512 2 4 4 1152 1 3 3 1 1 1 20 2>&1 | tail -1 >> $O/smallm.txt; done
cat $O/smallm.txt
python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -15 > $O/pytest.log; tail -5 $O/pytest.log

#!/bin/bash
O=gpurun_out/x2; mkdir -p $O
python tools/conv_sweep.py tools/smallm_shapes.txt 0:0 1073741824:1 1073741824:2 1073741824:4 1073741824:0 > $O/sweep.txt 2>&1
cat $O/sweep.txt
B2_DENSEM_MAXM=0 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_r2p1d_slab.csv python tools/fwd_once.py r2plus1d34 > $O/ncu1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_r2p1d_dm.csv python tools/fwd_once.py r2plus1d34 > $O/ncu2.log 2>&1
tail -2 $O/ncu1.log $O/ncu2.log

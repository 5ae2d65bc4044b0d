O=gpurun_out/tstack_final; mkdir -p $O
timeout 100 python bench.py --workload r2plus1d34 --no-cpu --no-biggan --no-others --steps 30 --warmup 5 --layers > $O/line_r2plus1d34.json 2> $O/layers_r2plus1d34.txt
tail -c 400 $O/line_r2plus1d34.json; head -8 $O/layers_r2plus1d34.txt
timeout 40 python tools/conv_sweep.py tools/tstack_shapes.txt 0:0:0:-1:0 0:0:0:-1:1 > $O/sweep.txt 2>&1; cat $O/sweep.txt

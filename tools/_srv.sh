BMG_COLSTATS_SERVER=1 BMG_TIMING=1 timeout 300 python bench.py --no-cpu-baseline --steps 4 > gpurun_out/bench_srv.json 2> gpurun_out/bench_srv.err
grep "end()\|store from bed\|create:" gpurun_out/bench_srv.err | cut -c1-200

cd $GRAFT_REPO_ROOT
timeout 80 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2l_bench_n1.json 2> gpurun_out/r2l_bench.err; cut -c1-120 gpurun_out/r2l_bench_n1.json; tail -c 300 gpurun_out/r2l_bench.err

cd $GRAFT_REPO_ROOT
NG=$(nvidia-smi -L | wc -l)
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
if [ "$NG" = "4" ]; then
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r2g_pytest_multi.log; cat gpurun_out/r2g_pytest_multi.log
timeout 400 $TR --nproc-per-node 2 --master-port 29610 bench.py --gpus 2 --config aggregator --steps 5 --warmup 2 > gpurun_out/r2g_agg_n2.json 2> gpurun_out/r2g_agg_n2.err; cut -c1-160 gpurun_out/r2g_agg_n2.json
fi
timeout 400 $TR --nproc-per-node $NG --master-port 29611 bench.py --gpus $NG --config aggregator --steps 5 --warmup 2 --dump-timeline gpurun_out/r2g_agg_tl_n$NG.json > gpurun_out/r2g_agg_n$NG.json 2> gpurun_out/r2g_agg_n$NG.err; tail -c 200 gpurun_out/r2g_agg_n$NG.err; cut -c1-160 gpurun_out/r2g_agg_n$NG.json
timeout 400 $TR --nproc-per-node $NG --master-port 29613 bench.py --gpus $NG --config statetransition --steps 5 --warmup 2 > gpurun_out/r2g_st_n$NG.json 2> gpurun_out/r2g_st_n$NG.err; tail -c 200 gpurun_out/r2g_st_n$NG.err; cut -c1-160 gpurun_out/r2g_st_n$NG.json

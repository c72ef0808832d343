cd $GRAFT_REPO_ROOT
NG=$(nvidia-smi -L | wc -l)
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 500 $TR --nproc-per-node $NG --master-port 29611 bench.py --gpus $NG --config aggregator --steps 5 --warmup 2 --dump-timeline gpurun_out/r2e_agg_tl_n$NG.json > gpurun_out/r2e_agg_n$NG.json 2> gpurun_out/r2e_agg_n$NG.err; tail -c 300 gpurun_out/r2e_agg_n$NG.err; cut -c1-200 gpurun_out/r2e_agg_n$NG.json
timeout 500 $TR --nproc-per-node $NG --master-port 29613 bench.py --gpus $NG --config statetransition --steps 5 --warmup 2 --dump-timeline gpurun_out/r2e_st_tl_n$NG.json > gpurun_out/r2e_st_n$NG.json 2> gpurun_out/r2e_st_n$NG.err; tail -c 300 gpurun_out/r2e_st_n$NG.err; cut -c1-200 gpurun_out/r2e_st_n$NG.json

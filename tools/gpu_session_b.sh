set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name --format=csv | head -3
NG=$(nvidia-smi -L | wc -l)
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_prove.py tests/test_gpu_msm.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2b_pytest.log; cat gpurun_out/r2b_pytest.log
for N in 2 $NG; do
timeout 900 $TR --nproc-per-node $N --master-port 29611 bench.py --gpus $N --config aggregator --steps 4 --warmup 2 --dump-timeline gpurun_out/r2b_agg_tl_n$N.json > gpurun_out/r2b_agg_n$N.json 2> gpurun_out/r2b_agg_n$N.err; tail -c 400 gpurun_out/r2b_agg_n$N.err; cut -c1-250 gpurun_out/r2b_agg_n$N.json
done
timeout 900 $TR --nproc-per-node $NG --master-port 29612 bench.py --gpus $NG --config aggregator --steps 4 --warmup 2 --no-shard-quotient > gpurun_out/r2b_agg_n${NG}_unsharded.json 2> gpurun_out/r2b_agg_unsh.err; tail -c 400 gpurun_out/r2b_agg_unsh.err; cut -c1-250 gpurun_out/r2b_agg_n${NG}_unsharded.json
timeout 900 $TR --nproc-per-node $NG --master-port 29613 bench.py --gpus $NG --config statetransition --steps 4 --warmup 2 --dump-timeline gpurun_out/r2b_st_tl_n$NG.json > gpurun_out/r2b_st_n$NG.json 2> gpurun_out/r2b_st.err; tail -c 400 gpurun_out/r2b_st.err; cut -c1-250 gpurun_out/r2b_st_n$NG.json
timeout 600 python bench.py --config aggregator --steps 4 --warmup 2 --no-cpu-baseline --mode range-split --dump-timeline gpurun_out/r2b_agg_tl_n1.json > gpurun_out/r2b_agg_n1.json 2> gpurun_out/r2b_agg_n1.err; tail -c 400 gpurun_out/r2b_agg_n1.err; cut -c1-250 gpurun_out/r2b_agg_n1.json

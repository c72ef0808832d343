cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 80 python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 2 --master-port 29611 bench.py --gpus 2 --config aggregator --mode range-split --steps 4 --warmup 2 > $O/r2k_agg_n2.json 2> $O/r2k_agg_n2.err; cut -c1-250 $O/r2k_agg_n2.json; tail -c 300 $O/r2k_agg_n2.err

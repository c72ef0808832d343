set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,memory.total --format=csv
nproc
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2a_pytest.log; cat gpurun_out/r2a_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; tail -c 600 gpurun_out/r2a_bench.err; cut -c1-400 gpurun_out/r2a_bench.json
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2a_ref.json 2> gpurun_out/r2a_ref.err; tail -c 600 gpurun_out/r2a_ref.err; cut -c1-300 gpurun_out/r2a_ref.json
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2a_launches_proof.csv python tools/profile_proof.py > gpurun_out/r2a_profile_proof.log 2>&1; tail -2 gpurun_out/r2a_profile_proof.log
timeout 600 python tools/timeline.py --out gpurun_out/r2a_timeline.json > gpurun_out/r2a_timeline.log 2>&1; tail -30 gpurun_out/r2a_timeline.log
timeout 900 python bench.py --config statetransition --steps 2 --warmup 1 --no-cpu-baseline --no-roofline > gpurun_out/r2a_st.json 2> gpurun_out/r2a_st.err; tail -c 800 gpurun_out/r2a_st.err; cat gpurun_out/r2a_st.json
timeout 900 python bench.py --config aggregator --steps 2 --warmup 1 --no-cpu-baseline --no-roofline > gpurun_out/r2a_agg.json 2> gpurun_out/r2a_agg.err; tail -c 800 gpurun_out/r2a_agg.err; cat gpurun_out/r2a_agg.json

cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 > $O/r2f_pytest.log; cat $O/r2f_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > $O/r2f_bench_n1.json 2> $O/r2f_bench_n1.err; tail -c 200 $O/r2f_bench_n1.err; cut -c1-160 $O/r2f_bench_n1.json
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/r2f_reference.json 2> $O/r2f_reference.err; cut -c1-160 $O/r2f_reference.json
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2f_launches_proof.csv python tools/profile_proof.py > $O/r2f_profile_proof.log 2>&1; tail -1 $O/r2f_profile_proof.log | cut -c1-200
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2f_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-roofline --no-uniform > $O/r2f_bench_under_ncu.log 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"k_msm_accumulate|k_ntt_pass|k_ntt_fused|k_msm_wsum_level0" -c 12 -o $O/r2f_full python tools/profile_proof.py > $O/r2f_full.log 2>&1; tail -2 $O/r2f_full.log | cut -c1-200
timeout 300 python tools/timeline.py --out $O/r2f_timeline.json > $O/r2f_timeline.log 2>&1; tail -16 $O/r2f_timeline.log
timeout 600 python bench.py --config statetransition --steps 3 --warmup 1 --no-cpu-baseline --no-roofline > $O/r2f_st_n1.json 2> $O/r2f_st_n1.err; cut -c1-200 $O/r2f_st_n1.json
timeout 600 python bench.py --config aggregator --steps 3 --warmup 1 --no-cpu-baseline --no-roofline > $O/r2f_agg_n1.json 2> $O/r2f_agg_n1.err; cut -c1-200 $O/r2f_agg_n1.json
timeout 300 python bench.py --config blob --steps 5 --warmup 3 > $O/r2f_blob.json 2> $O/r2f_blob.err; cut -c1-200 $O/r2f_blob.json
timeout 400 python bench.py --impl reference --config aggregator --steps 1 --warmup 1 > $O/r2f_agg_reference.json 2> $O/r2f_agg_reference.err; cut -c1-200 $O/r2f_agg_reference.json
timeout 600 python tools/sweep.py --max-log 26 --cpu-max-log 20 > $O/r2f_sweep.log 2>&1; tail -3 $O/r2f_sweep.log | cut -c1-200; cp $O/sweep.json $O/r2f_sweep.json; cp $O/sweep.md $O/r2f_sweep.md

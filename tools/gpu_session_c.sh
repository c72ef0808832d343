cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2c_pytest.log; cat gpurun_out/r2c_pytest.log
timeout 400 compute-sanitizer --tool memcheck --print-limit 3 python -m pytest tests/test_gpu_msm.py tests/test_gpu_prove.py -m gpu -x -q -k "test_msm_random or (bit_exact and (bw6 or bn254-40))" 2>&1 | grep -v "Host Frame" | tail -12 > gpurun_out/r2c_sanitizer.log; cat gpurun_out/r2c_sanitizer.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; tail -c 300 gpurun_out/r2c_bench.err; cut -c1-200 gpurun_out/r2c_bench.json
timeout 300 python tools/pageable_probe.py > gpurun_out/r2c_pageable.log 2>&1; tail -5 gpurun_out/r2c_pageable.log

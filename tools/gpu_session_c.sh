cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r2c_pytest.log; cat gpurun_out/r2c_pytest.log
timeout 900 compute-sanitizer --tool memcheck --print-limit 2 python -m pytest tests/test_gpu_msm.py tests/test_gpu_prove.py tests/test_gpu_artifacts.py tests/test_gpu_field_ec.py tests/test_gpu_kzg.py tests/test_gpu_ntt.py -m gpu -q -k "not verify_against and not stride and not budget and not setup_mirror" 2>&1 | grep -v "Host Frame" | tail -6 > gpurun_out/r2c_sanitizer.log; cat gpurun_out/r2c_sanitizer.log

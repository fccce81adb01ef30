set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py tests/test_gpu_models.py -x -q -m gpu > gpurun_out/v35_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/v35_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/v35_bench.json 2> gpurun_out/v35_bench.err
UNO_B200_ROWGEMM_NO_PARITY=1 timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/v35_bench_noparity.json 2>> gpurun_out/v35_bench.err
timeout 300 python tools/sweep_spectral.py --mode levels --workload darcy --out gpurun_out/v35_levels_darcy.json > gpurun_out/v35_levels.log 2>&1
timeout 500 python tools/sweep_spectral.py --mode sweep --iters 5 --out gpurun_out/v35_sweep.json > gpurun_out/v35_sweep.log 2>&1
timeout 200 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/v35_launches_darcy.csv python tools/profile_step.py --workload darcy > gpurun_out/v35_ncu.log 2>&1
timeout 200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/v35_launches_ns3d.csv python tools/profile_step.py --workload ns3d >> gpurun_out/v35_ncu.log 2>&1
timeout 200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/v35_launches_ns2d.csv python tools/profile_step.py --workload ns2d >> gpurun_out/v35_ncu.log 2>&1
tail -3 gpurun_out/v35_pytest.log; cat gpurun_out/v35_bench.json gpurun_out/v35_bench_noparity.json | cut -c1-300

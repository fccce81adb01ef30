set -u
mkdir -p gpurun_out
TAG=r02_v92
for wl in darcy ns3d ns2d_ar; do
    timeout 300 ncu --profile-from-start off --clock-control none --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
        --log-file gpurun_out/${TAG}_launches_${wl}.csv python tools/profile_step.py --workload $wl > gpurun_out/${TAG}_launches_${wl}.log 2>&1
    python tools/ncu_summary.py gpurun_out/${TAG}_launches_${wl}.csv > gpurun_out/${TAG}_launches_${wl}_summary.txt 2>&1
    python tools/ncu_traffic.py gpurun_out/${TAG}_launches_${wl}.csv gpurun_out/r02_traffic_${wl}.json > /dev/null 2>&1
done
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
timeout 300 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log; tail -n 2 gpurun_out/${TAG}_pytest_gpu.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${TAG}_smoke.log; tail -n 2 gpurun_out/${TAG}_smoke.log

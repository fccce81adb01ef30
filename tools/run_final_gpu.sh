#!/bin/bash
# Round-end refresh on one GPU: the whole -m gpu suite, smoke(), the default bench line + reference arm, then the ncu launch lists
# (time + DRAM bytes per launch -> profiles/r02_traffic_<workload>.json tied to the source hash) of the three workloads.
#   gpurun --timeout 2400 -- 'TAG=r02_v55 bash tools/run_final_gpu.sh'
set -u
mkdir -p gpurun_out
TAG=${TAG:-r02_final}
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -n 3 gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${TAG}_smoke.log; tail -n 2 gpurun_out/${TAG}_smoke.log
T0=$SECONDS
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$? wall $((SECONDS - T0)) s"
python tools/show_bench.py gpurun_out/${TAG}_bench.json 2>/dev/null | grep -v "spectral \|sweep S" | head -70
timeout 400 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
cut -c1-260 gpurun_out/${TAG}_bench_reference.json
for wl in ${NCU_WORKLOADS-darcy ns3d ns2d_ar}; do
    timeout 600 ncu --profile-from-start off --clock-control none --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
        --log-file gpurun_out/${TAG}_launches_${wl}.csv python tools/profile_step.py --workload $wl > gpurun_out/${TAG}_launches_${wl}.log 2>&1
    python tools/ncu_summary.py gpurun_out/${TAG}_launches_${wl}.csv > gpurun_out/${TAG}_launches_${wl}_summary.txt 2>&1
    python tools/ncu_traffic.py gpurun_out/${TAG}_launches_${wl}.csv gpurun_out/r02_traffic_${wl}.json > /dev/null 2>&1
    head -n 6 gpurun_out/${TAG}_launches_${wl}_summary.txt
done

#!/bin/bash
# Profiling call: launch list of one step (time + DRAM bytes per launch) and one `--set full` capture of the kernels named in
# KERNELS, for WORKLOAD (darcy | ns2d | ns3d) with whatever UNO_B200_* switches are exported.  ncu replays every kernel ~40 times
# under --set full: keep the regex narrow and -c small.  Numbers printed by a run under ncu are never bench values.
#   gpurun --timeout 1200 -- 'WORKLOAD=darcy KERNELS="wgrad_tc|kpipe_kernel" TAG=r02_v45 bash tools/run_ncu_gpu.sh'
set -u
WORKLOAD=${WORKLOAD:-darcy}
KERNELS=${KERNELS:-"wgrad_tc|kpipe_kernel"}
TAG=${TAG:-r02}
COUNT=${COUNT:-8}
mkdir -p gpurun_out
timeout 600 ncu --profile-from-start off --clock-control none --csv \
    --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --log-file gpurun_out/${TAG}_launches_${WORKLOAD}.csv python tools/profile_step.py --workload $WORKLOAD > gpurun_out/${TAG}_launches_${WORKLOAD}.log 2>&1
python tools/ncu_summary.py gpurun_out/${TAG}_launches_${WORKLOAD}.csv > gpurun_out/${TAG}_launches_${WORKLOAD}_summary.txt 2>&1
python tools/ncu_traffic.py gpurun_out/${TAG}_launches_${WORKLOAD}.csv gpurun_out/${TAG}_traffic_${WORKLOAD}.json >> gpurun_out/${TAG}_launches_${WORKLOAD}.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k "regex:${KERNELS}" -c $COUNT \
    -f -o gpurun_out/${TAG}_full_${WORKLOAD} python tools/profile_step.py --workload $WORKLOAD > gpurun_out/${TAG}_full_${WORKLOAD}.log 2>&1
ncu -i gpurun_out/${TAG}_full_${WORKLOAD}.ncu-rep --page raw --csv > gpurun_out/${TAG}_full_${WORKLOAD}_raw.csv 2>/dev/null
head -12 gpurun_out/${TAG}_launches_${WORKLOAD}_summary.txt
ls -la gpurun_out | tail -8

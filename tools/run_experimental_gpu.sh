#!/bin/bash
# First GPU call for the opt-in kernels (tc_mid.cuh, tc_cmm.cuh, tc_kpipe.cuh row classes): parity against the oracle / the
# default kernels, then the Darcy, NS-2D and NS-3D benches with and without them.  Everything lands in gpurun_out/.
#   gpurun --timeout 1500 -- 'bash tools/run_experimental_gpu.sh'
# Each stage has its own timeout so that a hung kernel cannot hold the box until gpurun's limit.
set -u
mkdir -p gpurun_out
OUT=gpurun_out/experimental
: > $OUT.log
run() { echo "=== $*" >> $OUT.log; "$@" >> $OUT.log 2>&1; echo "=== exit $?" >> $OUT.log; }
export UNO_B200_EXPERIMENTAL=1
for t in test_instance_norm_big_cluster test_synthesis_16_warp_epilogue test_analysis_row_classes test_analysis_16_loader_warps test_spectral2d_experimental_tc test_spectral3d_experimental_tc test_experimental_tc_timing; do
    run timeout 300 python -m pytest tests/test_gpu_experimental.py -x -q -s -k $t
done
unset UNO_B200_EXPERIMENTAL
for wl in darcy ns2d ns3d; do
    for flags in "" "UNO_B200_MID_TC=1" "UNO_B200_CMM_TC=1" "UNO_B200_KPIPE_ALIGN=1" "UNO_B200_KPIPE_LW16=1" "UNO_B200_KPIPE_LW16=1 UNO_B200_KPIPE_ALIGN=1" "UNO_B200_ROWGEMM_EPI16=1" "UNO_B200_ROWGEMM_EPI16=2" "UNO_B200_NORM_BIG_CLUSTER=1" "UNO_B200_NORM_BIG_CLUSTER=1 UNO_B200_MID_TC=1 UNO_B200_CMM_TC=1 UNO_B200_KPIPE_ALIGN=1 UNO_B200_ROWGEMM_EPI16=1"; do
        echo "=== bench $wl [$flags]" >> $OUT.log
        env $flags timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline >> $OUT.log 2>&1
    done
done

# the regime where the default leading-axis transform and mode contraction lose to cuFFT (S = 64, most of the spectrum kept)
for flags in "" "UNO_B200_MID_TC=1 UNO_B200_CMM_TC=1"; do
    echo "=== sweep S=64 [$flags]" >> $OUT.log
    env $flags timeout 600 python tools/sweep_spectral.py --sizes 64,128 --modes 20,32 --channels 32,128 --iters 5 \
        --out "gpurun_out/sweep_s64_$(echo $flags | tr -c 'A-Z0-9_\n' '_').json" >> $OUT.log 2>&1
done

echo "=== wgrad probe" >> $OUT.log
timeout 300 python tools/wgrad_probe.py >> $OUT.log 2>&1
tail -5 $OUT.log

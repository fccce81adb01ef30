#!/bin/bash
# First GPU call for the opt-in kernels (README.md "Environment switches"): parity against the oracle / the default kernels,
# then the benches with and without them, the S=64 sweep and the weight-gradient probe.  Everything lands in gpurun_out/.
#   gpurun --timeout 2400 -- 'bash tools/run_experimental_gpu.sh'          (about 25 - 30 minutes of box time)
# Each stage has its own timeout so that a hung kernel cannot hold the box until gpurun's limit; a stage that fails or hangs
# does not stop the ones after it.  STAGES="tests bench" selects stages (default: all).
set -u
mkdir -p gpurun_out
OUT=gpurun_out/experimental
STAGES=${STAGES:-"tests bench sweep probe"}
: > $OUT.log
run() { echo "=== $*" >> $OUT.log; "$@" >> $OUT.log 2>&1; echo "=== exit $?" >> $OUT.log; }
bench() {   # bench <workload> <flags...>
    local wl=$1; shift
    echo "=== bench $wl [$*]" >> $OUT.log
    env "$@" timeout 240 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline 2>>$OUT.log | \
        python -c "import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    kb=d.get('kernel_breakdown') or {}
    print(json.dumps({'ms_per_step':d['ms_per_step'],'value':d['value'],'e2e':d['e2e']['value'],'top':{k:round(v['ms_per_step'],3) for k,v in list(kb.items())[:8]}}))
    print(json.dumps({'spectral_levels':d.get('spectral_levels')}))" >> $OUT.log 2>&1
}
ALL="UNO_B200_RS_SPLIT=1 UNO_B200_MID_TC=1 UNO_B200_CMM_TC=1 UNO_B200_KPIPE_ALIGN=1 UNO_B200_KPIPE_LW16=1 UNO_B200_ROWGEMM_EPI16=2 UNO_B200_NORM_BIG_CLUSTER=1"

if [[ $STAGES == *tests* ]]; then
    export UNO_B200_EXPERIMENTAL=1
    for t in test_resample_split_staging test_empty_batch_like_the_reference test_pointwise3d_fixed_mode_gpu test_instance_norm_big_cluster test_synthesis_16_warp_epilogue test_analysis_row_classes test_analysis_16_loader_warps \
             test_spectral2d_experimental_tc test_spectral3d_experimental_tc test_experimental_tc_timing; do
        run timeout 300 python -m pytest tests/test_gpu_experimental.py -x -q -s -k $t
    done
    unset UNO_B200_EXPERIMENTAL
fi
if [[ $STAGES == *bench* ]]; then
    for wl in darcy ns2d; do
        bench $wl UNO_NOFLAG=1
        for f in UNO_B200_RS_SPLIT=1 UNO_B200_MID_TC=1 UNO_B200_CMM_TC=1 UNO_B200_KPIPE_ALIGN=1 UNO_B200_KPIPE_LW16=1 UNO_B200_ROWGEMM_EPI16=1 UNO_B200_ROWGEMM_EPI16=2; do
            bench $wl $f
        done
        bench $wl $ALL
    done
    bench ns3d UNO_NOFLAG=1
    bench ns3d UNO_B200_NORM_BIG_CLUSTER=1
    bench ns3d UNO_B200_MID_TC=1
    bench ns3d $ALL
fi
if [[ $STAGES == *sweep* ]]; then
    # the regime where the default leading-axis transform and mode contraction lose to cuFFT (S = 64, most of the spectrum kept)
    echo "=== sweep S=64 [default]" >> $OUT.log
    timeout 420 python tools/sweep_spectral.py --sizes 64 --modes 20,32 --channels 32,128 --iters 3 --out gpurun_out/sweep_s64_default.json >> $OUT.log 2>&1
    echo "=== sweep S=64 [MID_TC CMM_TC]" >> $OUT.log
    UNO_B200_MID_TC=1 UNO_B200_CMM_TC=1 timeout 420 python tools/sweep_spectral.py --sizes 64 --modes 20,32 --channels 32,128 --iters 3 \
        --out gpurun_out/sweep_s64_tc.json >> $OUT.log 2>&1
fi
if [[ $STAGES == *probe* ]]; then
    run timeout 300 python tools/wgrad_probe.py
fi
grep -E "^=== |passed|failed|error|ms_per_step" $OUT.log | tail -60

#!/bin/bash
# One GPU call: the whole `-m gpu` suite, the bench line (all legs), then optional extras selected by STAGES.
#   gpurun --timeout 2400 -- 'STAGES="tests bench" bash tools/run_gpu_checks.sh'
# STAGES: tests | bench | benchlean | sanitizer | ncu (uses WORKLOAD / KERNELS / TAG of tools/run_ncu_gpu.sh)
set -u
mkdir -p gpurun_out
STAGES=${STAGES:-"tests bench"}
TAG=${TAG:-r02}
if [[ $STAGES == *tests* ]]; then
    timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/${TAG}_pytest_gpu.log 2>&1
    echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
    tail -5 gpurun_out/${TAG}_pytest_gpu.log
fi
if [[ $STAGES == *benchlean* ]]; then
    for wl in darcy ns2d ns3d ns2d_ar; do
        timeout 300 python bench.py --workload $wl --lean --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_${wl}.json 2> gpurun_out/${TAG}_bench_${wl}.err
        python tools/show_bench.py gpurun_out/${TAG}_bench_${wl}.json 2>/dev/null | head -30
    done
elif [[ $STAGES == *bench* ]]; then
    timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
    echo "bench rc=$?"; tail -3 gpurun_out/${TAG}_bench.err
    python tools/show_bench.py gpurun_out/${TAG}_bench.json 2>/dev/null | head -60
    timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
    cut -c1-400 gpurun_out/${TAG}_bench_reference.json
fi
if [[ $STAGES == *sanitizer* ]]; then
    # memcheck: every kernel variant, the pixel MLPs and the golden parity cases; racecheck (shared-memory hazards, slow): the
    # golden parity cases, the fan-out / overlap / tensor-core-core variants and the fused lift / projection kernels
    timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_variants.py tests/test_gpu_glue.py tests/test_gpu_parity.py -x -q \
        > gpurun_out/${TAG}_sanitizer_memcheck.log 2>&1
    echo "memcheck rc=$?" >> gpurun_out/${TAG}_sanitizer_memcheck.log
    timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_variants.py tests/test_gpu_glue.py -x -q \
        -k "golden or fanout or overlap or tc_core or test_project or test_lift or first_call or sector" > gpurun_out/${TAG}_sanitizer_racecheck.log 2>&1
    echo "racecheck rc=$?" >> gpurun_out/${TAG}_sanitizer_racecheck.log
    tail -n 4 gpurun_out/${TAG}_sanitizer_memcheck.log; tail -n 4 gpurun_out/${TAG}_sanitizer_racecheck.log
fi
if [[ $STAGES == *ncu* ]]; then
    bash tools/run_ncu_gpu.sh
fi

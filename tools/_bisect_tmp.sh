run() { echo "== $1: $(env $2 timeout 400 python -m pytest tests/test_gpu_models.py tests/test_gpu_fullsize.py -q -k "not sweep and not rollout_matches" 2>&1 | grep -E '^FAILED|passed|failed' | tr '\n' ' ' | cut -c1-300)"; }
for i in 1 2 3 4 5 6 7 8 9 10; do
  run "default $i" "X=0"
done

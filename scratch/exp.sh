timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for cfg in "APA_SPLIT=1" "APA_SPLIT=0"; do
  echo "== $cfg"
  env $cfg python bench.py --pairs 10000 --steps 3 --warmup 2 --e2e-steps 2 --cpu-sample 8 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['pairs_per_s'], d['e2e']['ms_per_step'], d['gpu_launches'])"
done

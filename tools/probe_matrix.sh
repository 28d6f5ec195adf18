for cv in 1024 2048; do
  echo "== cluster_views=$cv"
  for pair in "probes8192 1" "probes4096 2" "probes2048 4" "probes1024 8"; do set -- $pair
    ORZ_CLUSTER_VIEWS=$cv ORZ_VIEW_STRIDE=$2 python tools/step_time.py $1 2>&1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:(round(v['ms_median'],3), v['launches_per_step']) for k,v in d.items() if k!='lib'})"
  done
done
python tools/step_time.py castle1024 probes1024 castle256 sponza256

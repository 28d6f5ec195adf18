for v in 1 2 4; do
  echo "== views_per_cta=$v"
  for pair in "probes8192 1" "probes4096 2" "probes2048 4" "probes1024 8" "probes1024 1"; do set -- $pair
    ORZ_CLUSTER_VIEWS=8192 ORZ_VIEWS_PER_CTA=$v ORZ_VIEW_STRIDE=$2 python tools/step_time.py $1 2>&1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('stride $2', {k:(round(v['ms_median'],3), v['launches_per_step'], v['vis_checksum']%100000) for k,v in d.items() if k!='lib'})"
  done
done

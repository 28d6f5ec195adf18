for cuts in 250,500,750 150,500,850 100,400,800 200,550,900 100,450,900 333,667,1000 500,1000,1000 1000,1000,1000 50,350,750 120,420,780; do
  echo -n "cuts=$cuts  "
  ORZ_GROUP_CUTS=$cuts ORZ_VIEW_STRIDE=8 python tools/step_time.py probes1024 castle1024 sponza256 2>&1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:(round(v['ms_median'],3), v['vis_checksum']%100000) for k,v in d.items() if k!='lib'})"
done

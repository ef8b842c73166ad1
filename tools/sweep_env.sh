#!/bin/bash
# usage: sweep_env.sh VAR "v1 v2 ..." [bench args]  -- runs bench.py once per value of VAR, prints ms/step and stage times
var=$1; vals=$2; shift 2
for v in $vals; do
  env $var=$v python bench.py --steps 5 --no-e2e "$@" 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$var=$v', round(d['ms_per_step'],4), {k:round(x,4) for k,x in d['stage_ms'].items() if x})"
done

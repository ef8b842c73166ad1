MP3B_K1_MODE=chunk MP3B_K1_THREADS=512 timeout 600 python -m pytest -m gpu -x -q tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_poison.py 2>&1 | tail -2
for wl in cfg2 cfg3 cfg4; do
for nt in 256 512; do for c in 0 768 1024; do
  if [ $nt = 256 -a $c != 0 ]; then continue; fi
  MP3B_K1_THREADS=$nt MP3B_K1_CHUNK=$c python bench.py --steps 10 --no-e2e --no-sweep --no-cpu --workload $wl 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$wl threads=$nt chunk=$c', round(d['ms_per_step'],4), {k:round(x,4) for k,x in d['stage_ms'].items() if x})"
done; done; done

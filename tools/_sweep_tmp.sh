timeout 900 python -m pytest tests -m gpu -x -q tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_poison.py 2>&1 | tail -3
for wl in cfg2 cfg3 cfg4; do
for mode in sorted chunk; do
  MP3B_K1_MODE=$mode python bench.py --steps 20 --no-e2e --no-sweep --no-cpu --workload $wl 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$wl mode=$mode', round(d['ms_per_step'],4), {k:round(x,4) for k,x in d['stage_ms'].items() if x})"
done; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_huffman_sorted|k_hsort' -s 8 -c 4 -f -o gpurun_out/k1sorted3 python bench.py --steps 2 --warmup 1 --no-e2e --no-sweep --no-cpu > gpurun_out/k1sorted_ncu.log 2>&1

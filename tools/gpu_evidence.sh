#!/bin/bash
# One GPU call that produces the round's evidence: GPU tests, the bench line, the ncu launch list of the same
# bench command and one `ncu --set full` capture of the two dominant kernels.  Outputs under gpurun_out/<tag>_*.
# usage: tools/gpu_evidence.sh <tag> [skip-tests]
tag=${1:-r02}
out=gpurun_out
mkdir -p $out
if [ "$2" != "skip-tests" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1
  echo "pytest rc=$?" >> $out/${tag}_pytest.log
  tail -3 $out/${tag}_pytest.log
fi
timeout 600 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err
echo "bench rc=$?"
tail -c 600 $out/${tag}_bench.json
# launch list of a short bench run of the same workload (kernel share of the step; cold-cache, serialised)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-e2e --no-sweep --no-cpu > $out/${tag}_ncu_launches.log 2>&1
echo "ncu launches rc=$?"
# full capture: one launch each of the Huffman kernel and the back end (skip the warm-up launches)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_huffman|k_backend' -s 4 -c 2 \
  -f -o $out/${tag}_full python bench.py --steps 2 --warmup 1 --no-e2e --no-sweep --no-cpu > $out/${tag}_ncu_full.log 2>&1
echo "ncu full rc=$?"
ls -la $out

#!/bin/bash
# One gpurun call: tests -> smoke -> bench -> kernel timings -> ncu launch list -> ncu --set full of the hot kernels.
# Usage (from the repo root, on the GPU box):  bash tools/gpu_round.sh [tag]
TAG=${1:-r02}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee $O/status.txt
timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/status.txt
timeout 600 python bench.py > $O/bench.log 2>$O/bench.err; echo "bench rc=$?" | tee -a $O/status.txt
timeout 200 python tools/attn_time.py > $O/attn_time.log 2>&1
timeout 200 python tools/attn_trace.py > $O/attn_trace.log 2>&1
if [ "$2" != "noncu" ]; then
timeout 600 python bench.py --steps 1 --warmup 3 --no-sub > $O/plain_bench.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 160 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-sub > $O/ncu_launches.log 2>&1
echo "ncu launches rc=$?" | tee -a $O/status.txt
for k in gemm attn ln gather sam; do
  pat=$k; skip=0; cnt=12
  case $k in
    gemm) pat='regex:gemm_tcgen05'; skip=8; cnt=4;;
    attn) pat='regex:flash_attn'; skip=2; cnt=1;;
    ln) pat='regex:layernorm'; skip=2; cnt=1;;
    gather) pat='regex:g1_fused'; skip=4; cnt=2;;
    sam) pat='regex:flash_attn|attn_win14'; skip=2; cnt=2;;
  esac
  timeout 300 python tools/prof_kernels.py $k > $O/plain_$k.log 2>&1 &&
  timeout 900 ncu --set full --clock-control none --import-source on -k "$pat" -s $skip -c $cnt -f -o $O/${TAG}_prof_$k \
      python tools/prof_kernels.py $k > $O/ncu_$k.log 2>&1
  echo "ncu $k rc=$?" | tee -a $O/status.txt
  ncu -i $O/${TAG}_prof_$k.ncu-rep --page raw --csv > $O/${TAG}_prof_${k}_raw.csv 2>/dev/null
  ncu -i $O/${TAG}_prof_$k.ncu-rep --page source --csv > $O/${TAG}_prof_${k}_source.csv 2>/dev/null
  sz=$(stat -c %s $O/${TAG}_prof_$k.ncu-rep 2>/dev/null || echo 0)
  if [ "$sz" -gt 12000000 ]; then rm -f $O/${TAG}_prof_$k.ncu-rep; fi
done
fi
du -sh $O; tail -3 $O/pytest_gpu.log; cat $O/smoke.log | tail -2; cat $O/attn_time.log; cat $O/bench.log | cut -c1-600

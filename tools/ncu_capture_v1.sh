set -x
mkdir -p gpurun_out
export NOVIC_NO_GRAPHS=1
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 || exit 1
for spec in "EpiQKV:700" "attention_kernel:760" "EpiRow:1400" "EpiGelu:700" "EpiLogits:110"; do
  k=${spec%%:*}; skip=${spec##*:}
  ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip $skip -c 1 -o gpurun_out/prof_$k $CMD > gpurun_out/ncu_$k.log 2>&1
  echo "ncu $k rc=$?"
done
ls -la gpurun_out

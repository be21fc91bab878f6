# usage: bash tools/ncu_capture.sh <tag>   (run under gpurun)
# 1) plain run must pass, 2) launch list with device time per launch, 3) --set full of the dominant kernels (one launch each)
set -x
TAG=${1:-v5}
mkdir -p gpurun_out
export NOVIC_NO_GRAPHS=1
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_$TAG.log 2>&1 || { tail -5 gpurun_out/plain_$TAG.log; exit 1; }
# one decode = 1 prep + 1 prefix + 15 x 6 x (qkv, attn, outproj, ffn) + 15 x (logits, select) + finalize ~ 393 launches
ncu --metrics gpu__time_duration.sum --clock-control none -s 2400 -c 420 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
echo "ncu list rc=$?"
for spec in "attention_stream_kernel_t:320:attn" "gemm_kernel:350:qkv" "outproj_ffn_kernel:300:block"; do
  IFS=: read k skip name <<< "$spec"
  ncu --set full --clock-control none --import-source on -k regex:^$k\$ --launch-skip $skip -c 1 -o gpurun_out/prof_${TAG}_$name $CMD > gpurun_out/ncu_${TAG}_$name.log 2>&1
  echo "ncu $name rc=$?"
done
ls -la gpurun_out | head -30

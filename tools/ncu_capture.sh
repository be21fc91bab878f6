# usage: bash tools/ncu_capture.sh <tag>   (run under gpurun; needs the plain run to pass first)
set -x
TAG=${1:-v2}
mkdir -p gpurun_out
export NOVIC_NO_GRAPHS=1
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_$TAG.log 2>&1 || { tail -5 gpurun_out/plain_$TAG.log; exit 1; }
for spec in "gemm_kernel:693:qkv" "gemm_kernel:694:gelu" "gemm_kernel:701:logits" "gemm_rowln_kernel:644:outproj" "gemm_rowln_kernel:645:ffn2" "attention_bulk_kernel:320:attn"; do
  IFS=: read k skip name <<< "$spec"
  ncu --set full --clock-control none --import-source on -k regex:^$k\$ --launch-skip $skip -c 1 -o gpurun_out/prof_${TAG}_$name $CMD > gpurun_out/ncu_${TAG}_$name.log 2>&1
  echo "ncu $name rc=$?"
done
ls -la gpurun_out | head -30

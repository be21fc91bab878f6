# usage: bash tools/ncu_capture_r02d.sh <tag>   (under gpurun, one GPU).  The round-2 capture (tools/ncu_capture_r02.sh) for the state with the
# row-owner block kernel: 1) the plain run must pass, 2) launch list with the device time of every launch of one decode (cold-cache,
# serialised: compare SHARES), 3) ncu --set full of ONE decode-step launch (the step whose query sees 16 keys) of the block kernel and of the QKV kernel; the
# captures of the unchanged kernel classes (attention, logits, selection) are profiles/r02c_*.
set -x
TAG=${1:-r02d}
mkdir -p gpurun_out
CMD="python tools/ncu_target.py"
$CMD > gpurun_out/plain_$TAG.log 2>&1 || { tail -5 gpurun_out/plain_$TAG.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 340 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
echo "ncu list rc=$?"
for spec in "block_rows_kernel:234:block_outproj_ffn" "EpiQKV:252:qkv_gemm"; do
  IFS=: read k skip name <<< "$spec"
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:$k --launch-skip $skip -c 1 -o gpurun_out/prof_${TAG}_$name $CMD > gpurun_out/ncu_${TAG}_$name.log 2>&1
  echo "ncu $name rc=$?"
  ncu -i gpurun_out/prof_${TAG}_$name.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_${name}_raw.csv 2>/dev/null
  ncu -i gpurun_out/prof_${TAG}_$name.ncu-rep --page source --csv > gpurun_out/${TAG}_ncu_${name}_source.csv 2>/dev/null
  ncu -i gpurun_out/prof_${TAG}_$name.ncu-rep --page details > gpurun_out/${TAG}_ncu_${name}_details.txt 2>/dev/null
  rm -f gpurun_out/prof_${TAG}_$name.ncu-rep
done
ls -la gpurun_out | tail -12

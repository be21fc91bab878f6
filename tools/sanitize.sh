# usage: bash tools/sanitize.sh <tag>   (under gpurun, one GPU).  compute-sanitizer lanes of SURVEY.md section 5 over small invocations of
# every kernel class (tools/sanitize_target.py): memcheck, racecheck (shared-memory hazards; mbarrier / async-proxy hand-overs are
# tracked by the tool since CUDA 12), synccheck (barrier misuse).  Each lane has its own time limit; the plain runs go first.
# Logs land in gpurun_out/sanitize_<tag>_<tool>_<mode>.log (copied to profiles/ when they back a statement in DESIGN.md).
set -x
TAG=${1:-r02}
mkdir -p gpurun_out
SAN=/usr/local/cuda/bin/compute-sanitizer
for mode in decode train encoder; do
  timeout 300 python tools/sanitize_target.py $mode > gpurun_out/sanitize_${TAG}_plain_$mode.log 2>&1 || { tail -5 gpurun_out/sanitize_${TAG}_plain_$mode.log; echo "plain $mode FAILED"; }
done
for spec in "memcheck:decode" "racecheck:decode" "memcheck:train" "racecheck:train" "synccheck:decode" "memcheck:encoder" "synccheck:train" "racecheck:encoder"; do
  IFS=: read tool mode <<< "$spec"
  timeout ${SAN_LIMIT:-420} $SAN --tool $tool --print-limit 40 --log-file gpurun_out/sanitize_${TAG}_${tool}_$mode.log python tools/sanitize_target.py $mode > gpurun_out/sanitize_${TAG}_${tool}_${mode}_stdout.log 2>&1
  echo "$tool $mode rc=$?"
  tail -3 gpurun_out/sanitize_${TAG}_${tool}_$mode.log
done

# usage: bash tools/ncu_capture_train.sh <tag>   (run under gpurun, one GPU)
# DRAM bytes, duration and throughput percentages of the non-GEMM kernels of one training step (first step of tools/bench_train.py):
# 6 attention forward, 13 LayerNorm backward, 6 attention backward launches.
TAG=${1:-v1}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__occupancy_limit_registers,launch__occupancy_limit_shared_mem \
    --clock-control none -k 'regex:ln_bwd_kernel|attn_bwd_kernel|attention_tf_kernel|pos_grad_kernel|embed_bwd_kernel' -c 27 --csv --log-file gpurun_out/ncu_train_kernels_$TAG.csv \
    python tools/bench_train.py --steps 1 --warmup 1 > gpurun_out/ncu_train_kernels_$TAG.log 2>&1
echo "ncu rc=$?"
tail -3 gpurun_out/ncu_train_kernels_$TAG.log

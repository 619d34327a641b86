#!/bin/bash
# Evidence captures of the CURRENT build on the GPU box (run through gpurun; outputs in gpurun_out/, summaries are made on the
# CPU box with tools/step_metrics_summary.py / tools/ncu_summary.py and committed under profiles/).
#   1. per-launch metrics of one eager step (time, DRAM bytes, tensor-pipe activity, L2 bytes), ncu --clock-control none
#   2. ncu --set full of every tcgen05 slab-kernel launch of one step (per layer shape), with source correlation
#   3. ncu --set full of the attention / linear-attention kernels of one step
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum \
    --clock-control none --csv --log-file gpurun_out/r02_step_metrics.csv python tools/one_step.py 4 2 > gpurun_out/r02_step_metrics.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tc_slab_kernel -o gpurun_out/r02_slab_full -f \
    python tools/one_step.py 4 1 > gpurun_out/r02_slab_full.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:linattn|attention' -o gpurun_out/r02_attn_full -f \
    python tools/one_step.py 4 1 > gpurun_out/r02_attn_full.log 2>&1
ls -la gpurun_out/r02_*

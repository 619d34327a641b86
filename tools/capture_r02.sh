#!/bin/bash
# Evidence captures of the CURRENT build on the GPU box (run through gpurun).  Raw .ncu-rep files are converted to CSV on the
# box and deleted (gpurun_out/ is capped at 64 MiB); summaries are made on the CPU box and committed under profiles/.
#   1. per-launch metrics of one eager step (time, DRAM bytes, tensor-pipe activity, L2 bytes), ncu --clock-control none
#   2. ncu --set full of ONE launch per distinct tcgen05 layer shape (tools/prof_layers_r02.py)
#   3. ncu --set full of the attention / linear-attention kernels of one step
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum \
    --clock-control none --csv --log-file gpurun_out/r02_step_metrics.csv python tools/one_step.py 4 2 > gpurun_out/r02_step_metrics.log 2>&1
ncu --set full --clock-control none -k regex:tc_slab_kernel -o /tmp/r02_slab_full -f python tools/prof_layers_r02.py > gpurun_out/r02_slab_full.log 2>&1
ncu -i /tmp/r02_slab_full.ncu-rep --page raw --csv > gpurun_out/r02_slab_full_raw.csv 2>/dev/null
ncu --set full --clock-control none -k 'regex:linattn|attention' -o /tmp/r02_attn_full -f python tools/one_step.py 4 1 > gpurun_out/r02_attn_full.log 2>&1
ncu -i /tmp/r02_attn_full.ncu-rep --page raw --csv > gpurun_out/r02_attn_full_raw.csv 2>/dev/null
ls -la gpurun_out/r02_* /tmp/r02_*

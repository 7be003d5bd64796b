#!/bin/bash
mkdir -p gpurun_out
for mode in 0 1; do
DL4DS_TC_NO_T=$mode timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/t${mode}_launches.csv python scratch/time_layers.py tf32x3 > gpurun_out/t${mode}.log 2>&1
echo "== NO_T=$mode"; python scratch/summarize_seq.py gpurun_out/t${mode}_launches.csv | grep -E "conv_tc_fwd" | awk '{print $1, $(NF-2), $(NF-1)}' | uniq -c | head -40
done

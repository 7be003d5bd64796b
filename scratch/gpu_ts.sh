#!/bin/bash
mkdir -p gpurun_out
for mode in 0 1; do
DL4DS_TC_A_TMEM=$mode timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ts${mode}_launches.csv python scratch/time_layers.py tf32x3 > gpurun_out/ts${mode}.log 2>&1
echo "== A_TMEM=$mode"; python scratch/summarize_seq.py gpurun_out/ts${mode}_launches.csv | grep conv_tc_fwd | awk '{print $(NF-2), $(NF-1)}' | sort | uniq -c | sort -k3 -n | awk '{printf "%s x%s  ", $3, $1} END{print ""}'
done

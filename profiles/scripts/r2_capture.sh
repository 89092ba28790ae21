#!/bin/bash
# Run on a B200 box from the repo root (under gpurun): ncu launch list, DRAM bytes and one `--set full` capture per phase
# kernel of the resident headline batch (10 000 pairs, n = 100 k, e = 5 %, astarpa2_full, cost + CIGAR). TAG names the outputs.
TAG=${TAG:-r2}
KERNELS=${KERNELS:-"build pass trace"}
mkdir -p gpurun_out
AB_REPS=1 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:apa_ -c 12 --csv \
    --log-file gpurun_out/${TAG}_launches.csv python profiles/scripts/resident_run.py > gpurun_out/${TAG}_ncu_launch.log 2>&1
for k in $KERNELS; do
AB_REPS=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:apa_phase_${k} -c 1 -o gpurun_out/${k}_full_${TAG} -f \
    python profiles/scripts/resident_run.py > gpurun_out/${TAG}_ncu_$k.log 2>&1
done
# the reports exceed what gpurun carries back (64 MiB in total): summarise them here and keep the text
for k in $KERNELS; do
python profiles/summarize_ncu.py gpurun_out/${k}_full_${TAG}.ncu-rep > gpurun_out/${TAG}_${k}_kernel_ncu_full.txt 2>&1
python profiles/top_lines.py gpurun_out/${k}_full_${TAG}.ncu-rep 40 > gpurun_out/${TAG}_${k}_kernel_top_lines.txt 2>&1
done
ls -la gpurun_out/*_${TAG}.ncu-rep
[ -n "$KEEP_REPORTS" ] || rm -f gpurun_out/*_${TAG}.ncu-rep

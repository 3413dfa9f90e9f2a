#!/bin/bash
# Turns the .ncu-rep files a gpurun call brought back into the text summaries kept under profiles/:
#   tools/ncu_export.sh gpurun_out/r02_frame.ncu-rep profiles/r02_frame_kernels_ncu_full.txt
set -e
rep="$1"; out="$2"
ncu -i "$rep" --page raw --csv > "${rep%.ncu-rep}_raw.csv"
python profiles/summarize.py raw "${rep%.ncu-rep}_raw.csv" > "$out"
echo "$out"

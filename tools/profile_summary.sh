#!/bin/bash
# usage: tools/profile_summary.sh <file.ncu-rep> <samples-per-launch> <out.txt>   (run here, no GPU needed)
rep=$1; samples=$2; out=$3
{
  echo "# ncu --set full --clock-control none --import-source on  ($(basename $rep))"
  python tools/ncu_summary.py $rep
  ncu -i $rep --page source --csv 2>/dev/null > /tmp/_src.csv
  echo; echo "# SASS instruction mix (warp instructions executed; per-sample = x32/samples) and stall samples"
  python tools/ncu_mix.py /tmp/_src.csv $samples
} > $out

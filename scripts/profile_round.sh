#!/bin/bash
# Round profile capture (run under gpurun, ONE GPU).  Outputs land in gpurun_out/ (kept small: the .ncu-rep files are
# exported to CSV on the box and deleted) and are summarised into profiles/ by scripts/summarise_profiles.py.
R=${1:-r1}
O=gpurun_out
T=/tmp/ncu_$R
mkdir -p $T $O
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${R}_launches.csv python bench.py --steps 2 --warmup 1 > $O/${R}_launches_bench.log 2>&1
cap() {  # name kernel-regex skip script args...
  local name=$1 rx=$2 skip=$3; shift 3
  ncu --set full --import-source on --clock-control none -k regex:$rx -s $skip -c 1 -o $T/$name -f python "$@" > $O/${R}_$name.log 2>&1
  ncu -i $T/$name.ncu-rep --page raw --csv > $O/${R}_${name}_raw.csv 2>/dev/null
  ncu -i $T/$name.ncu-rep --page source --csv 2>/dev/null | gzip -9 > $O/${R}_${name}_src.csv.gz
  rm -f $T/$name.ncu-rep
}
cap nuts nuts_warp_kernel 0 scripts/ncu_target_radon.py
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/${R}_launches_logistic.csv python scripts/ncu_target3.py logistic 4 2 > $O/${R}_launches_logistic.log 2>&1
cap logistic logistic_fused_kernel 2 scripts/ncu_target3.py logistic 2 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/${R}_launches_mvgauss.csv python scripts/ncu_target3.py mvgauss 4 2 > $O/${R}_launches_mvgauss.log 2>&1
cap gemm gemm_nt_dmma_kernel 4 scripts/ncu_target3.py mvgauss 2 1
for f in $O/${R}_*.log; do echo "== $f"; tail -n 2 $f; done
du -sh $O

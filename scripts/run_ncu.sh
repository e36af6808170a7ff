#!/bin/bash
# Runs on the GPU box (under gpurun): one small `ncu --set full` report per kernel + CSV/text exports.
# usage: scripts/run_ncu.sh <tag> [kernels...]   kernels: chain k1n3 k1n8 k2n8 rqsfwd rqsinv launches
set -u
TAG=${1:-r1}; shift
WHAT=${@:-chain k1n3 k1n8 k2n8 rqsfwd rqsinv launches}
OUT=gpurun_out
mkdir -p $OUT
python scripts/profile_target.py all > $OUT/pt_plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/pt_plain_$TAG.log; exit 1; }
prof() {  # name kernel-regex target launch-skip
  ncu --set full --clock-control none --import-source on -k regex:"$2" -s $4 -c 1 -f -o $OUT/prof_${TAG}_$1 \
      python scripts/profile_target.py $3 > $OUT/ncu_${TAG}_$1.log 2>&1
  ncu -i $OUT/prof_${TAG}_$1.ncu-rep --page raw --csv > $OUT/raw_${TAG}_$1.csv 2>/dev/null
  ncu -i $OUT/prof_${TAG}_$1.ncu-rep --page details > $OUT/details_${TAG}_$1.txt 2>/dev/null
  ncu -i $OUT/prof_${TAG}_$1.ncu-rep --page source --csv > $OUT/source_${TAG}_$1.csv 2>/dev/null
  ls -la $OUT/prof_${TAG}_$1.ncu-rep
  # gpurun_out is capped at 64 MiB: keep the binary report only when small, the CSV/text exports always
  if [ $(stat -c %s $OUT/prof_${TAG}_$1.ncu-rep) -gt 9000000 ]; then rm -f $OUT/prof_${TAG}_$1.ncu-rep; fi
}
for w in $WHAT; do
  case $w in
    chain)  prof chain  "is_chain_kernel" chain 2 ;;
    flowA)  prof flowA  "flow_sample_fast_kernel" chain 2 ;;
    rqs64f) prof rqs64f "rqs_coop_kernel" rqs64 2 ;;
    rqs64b) prof rqs64b "rqs_backward_coop_kernel" rqs64 2 ;;
    flowB)  prof flowB  "rambo_is_kernel" chain 2 ;;
    k1n3)   prof k1n3   "rambo_forward_kernel" k1 1 ;;
    k1n8)   prof k1n8   "rambo_forward_kernel" k1 4 ;;
    k2n8)   prof k2n8   "rambo_inverse_stream_kernel" k2 1 ;;
    rqsfwd) prof rqsfwd "rqs_reg2_kernel" rqs 2 ;;
    rqsinv) prof rqsinv "rqs_reg2_kernel" rqs 3 ;;
    launches) ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_$TAG.csv python scripts/profile_target.py all > /dev/null 2>&1 ;;
  esac
done
du -sh $OUT

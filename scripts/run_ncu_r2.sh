#!/bin/bash
# Round-2 ncu captures (one GPU, under gpurun): `ncu --set full` of one launch per kernel + CSV/text exports.
# usage: scripts/run_ncu_r2.sh <tag> <name:kernel-regex:target:launch-skip> ...
set -u
TAG=$1; shift
OUT=gpurun_out
mkdir -p $OUT
for spec in "$@"; do
  IFS=: read -r name regex target skip <<< "$spec"
  python scripts/profile_target.py $target > $OUT/pt_plain_${TAG}_$name.log 2>&1 || { echo "plain run of $target failed"; tail -5 $OUT/pt_plain_${TAG}_$name.log; continue; }
  ncu --set full --clock-control none --import-source on -k regex:"$regex" -s $skip -c 1 -f -o $OUT/prof_${TAG}_$name \
      python scripts/profile_target.py $target > $OUT/ncu_${TAG}_$name.log 2>&1
  ncu -i $OUT/prof_${TAG}_$name.ncu-rep --page raw --csv > $OUT/raw_${TAG}_$name.csv 2>/dev/null
  ncu -i $OUT/prof_${TAG}_$name.ncu-rep --page details > $OUT/details_${TAG}_$name.txt 2>/dev/null
  ncu -i $OUT/prof_${TAG}_$name.ncu-rep --page source --csv > $OUT/source_${TAG}_$name.csv 2>/dev/null
  ls -la $OUT/prof_${TAG}_$name.ncu-rep
  if [ $(stat -c %s $OUT/prof_${TAG}_$name.ncu-rep) -gt 6000000 ]; then rm -f $OUT/prof_${TAG}_$name.ncu-rep; fi
done
du -sh $OUT

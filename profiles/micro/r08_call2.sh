#!/bin/bash
# r08 call 2: A/B of the decoupled walk kernel's thresholds (library variants built with -D..., loaded through ARV2_LIB)
OUT=gpurun_out
python -m pytest tests -m gpu -x -q -k "rerender or c4_ or fullsize_rerender" > $OUT/r08_tests_rr.log 2>&1; tail -3 $OUT/r08_tests_rr.log
: > $OUT/r08_ab.log
for d in audiorenderingv2_b200/lib audiorenderingv2_b200/lib_*; do
  [ -f $d/libarv2.so ] || continue
  LABEL="$(basename $d) c2" ARV2_LIB=$PWD/$d/libarv2.so python profiles/micro/rr_only.py >> $OUT/r08_ab.log 2>&1
done
LABEL="lib c4" RR_WORKLOAD=c4 RR_STEPS=10 python profiles/micro/rr_only.py >> $OUT/r08_ab.log 2>&1
cat $OUT/r08_ab.log | cut -c1-200
RR_STEPS=6 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum --clock-control none -k regex:"rr_" --csv --log-file $OUT/r08_rr_launches.csv python profiles/micro/rr_only.py > $OUT/r08_ncu1.log 2>&1

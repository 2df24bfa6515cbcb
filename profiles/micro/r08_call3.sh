#!/bin/bash
OUT=gpurun_out
python -m pytest tests -m gpu -x -q -k "rerender or c4_ or fullsize_rerender" > $OUT/r08_tests_rr.log 2>&1; tail -3 $OUT/r08_tests_rr.log
: > $OUT/r08_ab.log
LABEL="shared c2" python profiles/micro/rr_only.py >> $OUT/r08_ab.log 2>&1
LABEL="global-nodes c2" ARV2_RR_NO_SHARED=1 python profiles/micro/rr_only.py >> $OUT/r08_ab.log 2>&1
LABEL="shared m4 c2" ARV2_LIB=$PWD/audiorenderingv2_b200/lib_m4/libarv2.so python profiles/micro/rr_only.py >> $OUT/r08_ab.log 2>&1
LABEL="shared c4" RR_WORKLOAD=c4 RR_STEPS=10 python profiles/micro/rr_only.py >> $OUT/r08_ab.log 2>&1
cat $OUT/r08_ab.log | cut -c1-200
RR_STEPS=6 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"rr_" --csv --log-file $OUT/r08_rr_launches.csv python profiles/micro/rr_only.py > $OUT/r08_ncu1.log 2>&1

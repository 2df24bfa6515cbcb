#!/bin/bash
OUT=gpurun_out
python -m pytest tests -m gpu -x -q > $OUT/r08_tests.log 2>&1; tail -3 $OUT/r08_tests.log
: > $OUT/r08_ab.log
LABEL="c2" python profiles/micro/rr_only.py >> $OUT/r08_ab.log 2>&1
LABEL="c2 serial" ARV2_RR_SERIAL=1 python profiles/micro/rr_only.py >> $OUT/r08_ab.log 2>&1
LABEL="c4" RR_WORKLOAD=c4 RR_STEPS=10 python profiles/micro/rr_only.py >> $OUT/r08_ab.log 2>&1
cat $OUT/r08_ab.log | cut -c1-700
RR_STEPS=6 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"rr_|pc_" --csv --log-file $OUT/r08_rr_launches.csv python profiles/micro/rr_only.py > $OUT/r08_ncu1.log 2>&1
RR_WORKLOAD=c4 RR_STEPS=4 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum --clock-control none -k regex:"rr_" --csv --log-file $OUT/r08_rr_launches_c4.csv python profiles/micro/rr_only.py > $OUT/r08_ncu1b.log 2>&1

#!/bin/bash
# r08 call 1: packed path cache + mask/walk re-render: parity, timing (C2, C4), kernel times, one full capture
set -x
OUT=gpurun_out
python -m pytest tests -m gpu -x -q -k "rerender or c4_ or fullsize_rerender" > $OUT/r08_tests_rr.log 2>&1; tail -5 $OUT/r08_tests_rr.log
LABEL=c2_maskwalk python profiles/micro/rr_only.py > $OUT/r08_rr.log 2>&1
LABEL=c2_serial ARV2_RR_SERIAL=1 python profiles/micro/rr_only.py >> $OUT/r08_rr.log 2>&1
LABEL=c4_maskwalk RR_WORKLOAD=c4 RR_STEPS=10 python profiles/micro/rr_only.py >> $OUT/r08_rr.log 2>&1
LABEL=c4_serial RR_WORKLOAD=c4 RR_STEPS=10 ARV2_RR_SERIAL=1 python profiles/micro/rr_only.py >> $OUT/r08_rr.log 2>&1
cat $OUT/r08_rr.log
RR_STEPS=6 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum --clock-control none -k regex:"rr_|pc_|finalize" --csv --log-file $OUT/r08_rr_launches.csv python profiles/micro/rr_only.py > $OUT/r08_ncu1.log 2>&1
RR_STEPS=6 ncu --set full --import-source on --clock-control none -k regex:"rr_mask|rr_walk" -s 8 -c 2 -f -o $OUT/r08_rr python profiles/micro/rr_only.py > $OUT/r08_ncu2.log 2>&1
ls -la $OUT | tail -5

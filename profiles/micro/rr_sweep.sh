for s in 1 2 3 4 5; do LABEL=s$s ARV2_RR_SCAN_WARPS=$s timeout 90 python profiles/micro/rr_only.py 2>&1 | tail -1 | cut -c1-300; done
timeout 120 python profiles/micro/rr_check.py 2>&1 | tail -2

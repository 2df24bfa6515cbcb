#!/bin/bash
# r09 (5): whole GPU suite + the default bench line with the sweeps in place
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r09_tests.log 2>&1; tail -3 gpurun_out/r09_tests.log
python bench.py > gpurun_out/r09_bench_n1.json 2> gpurun_out/r09_bench_n1.err; tail -c 600 gpurun_out/r09_bench_n1.json

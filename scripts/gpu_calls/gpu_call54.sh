#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | grep -v "^$" | tail -5 > gpurun_out/r02_c54_tests.log; echo "tests rc=${PIPESTATUS[0]}"; tail -2 gpurun_out/r02_c54_tests.log
for i in 1 2; do timeout 300 python scripts/ab_fwd.py > gpurun_out/r02_c54_ab$i.log 2>&1; grep "^default" gpurun_out/r02_c54_ab$i.log | cut -c1-175; done

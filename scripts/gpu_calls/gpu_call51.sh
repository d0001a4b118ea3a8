#!/bin/bash
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02_c51_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02_c51_smoke.log
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke(); print('entry ok')" 2>&1 | tail -2

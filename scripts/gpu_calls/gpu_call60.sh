#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | grep -v "^$" | tail -2; done
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1

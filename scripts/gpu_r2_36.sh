#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -s 2>&1 | grep "\[parity\]\|passed\|failed\|FAILED\|Error" > gpurun_out/r2_36_parity.log; tail -45 gpurun_out/r2_36_parity.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3

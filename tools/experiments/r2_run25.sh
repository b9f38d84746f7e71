#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_x_pytest_all.log 2>&1; echo "all gpu tests rc=$?"; tail -3 gpurun_out/r2_x_pytest_all.log | cut -c1-300
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -1

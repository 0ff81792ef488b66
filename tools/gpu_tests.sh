#!/bin/bash
mkdir -p gpurun_out
timeout 3000 python -m pytest tests -m gpu -q > gpurun_out/${1:-r2}_pytest_all.log 2>&1
echo "all tests rc=$?" >> gpurun_out/${1:-r2}_pytest_all.log
tail -25 gpurun_out/${1:-r2}_pytest_all.log

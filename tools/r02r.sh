#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/r02r_pytest.log 2>&1
tail -6 gpurun_out/r02r_pytest.log

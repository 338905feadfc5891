#!/bin/bash
mkdir -p gpurun_out/c2
( time timeout 1800 python -m pytest tests -m gpu -q ) > gpurun_out/c2/pytest.log 2>&1
tail -15 gpurun_out/c2/pytest.log
bash tools/calls/sanitize.sh

#!/bin/bash
# helper: run the gpu test-suite on the box and keep the log
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.log

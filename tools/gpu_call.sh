#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -12 | tee gpurun_out/r2_tests.log

#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest -q -x -p no:cacheprovider tests/test_gpu_g_batchnorm.py -s > gpurun_out/bn_tests.log 2>&1; echo "bn tests rc=$?"; tail -n 40 gpurun_out/bn_tests.log | cut -c1-400

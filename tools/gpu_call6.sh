#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_api.py::test_multi_gpu_in_process_equals_single_gpu 2>&1 | tail -40 | tee gpurun_out/pytest_gpu_r02e.log

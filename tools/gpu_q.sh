#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_gemm.py -x -q -k tc_linear 2>&1 | tail -15
timeout 300 python -m pytest tests/test_gpu_model.py -x -q 2>&1 | tail -5

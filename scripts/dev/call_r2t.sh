#!/bin/bash
O=gpurun_out/r2t; mkdir -p $O
timeout 1200 python -m pytest tests/test_reference_goldens.py -m gpu -q --no-header -rf --tb=short > $O/pytest_goldens.log 2>&1; tail -n 80 $O/pytest_goldens.log

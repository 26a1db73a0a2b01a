#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_serving.py -m gpu -q -x 2>&1 | tail -4

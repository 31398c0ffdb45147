#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_block.py -m gpu -q -p no:cacheprovider -k "inference" 2>&1 | grep -E "Error|assert|kept|sizes|passed|failed" | head -20

#!/bin/bash
timeout 900 python -m pytest tests/test_large_shard.py -x -q -m gpu 2>&1 | tail -3

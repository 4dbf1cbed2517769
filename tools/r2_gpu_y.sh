#!/bin/bash
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/y_tests.log 2>&1; tail -8 gpurun_out/y_tests.log

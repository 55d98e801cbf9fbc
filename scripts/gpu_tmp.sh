#!/bin/bash
OUT=gpurun_out/${1:-tmp}
mkdir -p $OUT
timeout 900 python -m pytest tests/test_train_step.py -m gpu -q --timeout 300 > $OUT/pytest_train.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_train.log
tail -30 $OUT/pytest_train.log

#!/bin/bash
# the GPU suite, then whatever bench commands follow as arguments
T=${TAG:-r2t}
( timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) > gpurun_out/${T}_tests.log 2>&1
cat gpurun_out/${T}_tests.log

#!/bin/bash
# builds the tracing variant of the library on the GPU box and prints in-kernel timelines
# (TRACE_TOOL=tools/tc_trace.py with PCV_SPLIT=0 for the whole-tile kernel, tools/tc_trace_split.py with PCV_SPLIT=1)
make -C perceiver_io_b200/csrc clean >/dev/null; make -C perceiver_io_b200/csrc -j8 TRACE=1 2>&1 | grep -E "error" 
for d in ${DBGS:-0 1}; do echo "== PCV_DBG=$d"; PCV_DBG=$d PCV_TRACE=1 timeout 200 python ${TRACE_TOOL:-tools/tc_trace.py} 2>&1 | tail -${TAILN:-6}; done

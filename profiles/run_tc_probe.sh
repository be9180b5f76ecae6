#!/bin/bash
# bring-up matrix of the tcgen05 decode kernel: one process per configuration (a trap poisons the context)
out=gpurun_out/tc_probe.log
: > $out
run() { echo "=== $*" >> $out; timeout 180 python profiles/probe_tc.py "$@" >> $out 2>&1; echo "rc=$?" >> $out; }
run small 1 4
run small 3 3
run small 5 3
run small 20 3
run small 40 3
run wide 1 3
run nano 5 3
grep -E "===|TC-PROBE .*worst|rc=|Error|error|timed out|neutts_b200" $out | head -60
for a in "1 tc 1" "4 tc 0" "8 tc 0" "64 tc 0"; do timeout 300 python profiles/perf_tc.py $a 2>&1 | grep -vE "^\s*$" | grep -E "PERF|TIMELINE|avg us|fine|    \+|Error|error|neutts_b200" ; done | tee gpurun_out/perf_tc_4.log

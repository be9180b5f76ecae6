#!/bin/bash
# bring-up matrix of the tcgen05 decode kernel: one process per configuration (a trap poisons the context)
out=gpurun_out/tc_probe.log
: > $out
run() { echo "=== $*" >> $out; timeout 180 env "${ENVV[@]}" python profiles/probe_tc.py "$@" >> $out 2>&1; echo "rc=$?" >> $out; }
ENVV=(X=1); run small 1 3
ENVV=(NT_TC_FOLD=phase); run small 1 3
ENVV=(X=1); run small 3 3
run small 6 3
run small 12 3
run small 20 3
run small 40 3
run wide 1 3
run nano 1 3
run nano 5 3
grep -E "===|TC-PROBE|rc=|Error|error|timed out|neutts_b200" $out | head -120

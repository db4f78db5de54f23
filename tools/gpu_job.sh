#!/bin/bash
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --print-limit 20 python __graft_entry__.py smoke > gpurun_out/r2_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -15 gpurun_out/r2_memcheck.log | cut -c1-250

#!/bin/bash
mkdir -p gpurun_out
echo "no job"

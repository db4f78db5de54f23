"""Registers / spills per kernel from `nvcc -Xptxas -v` output.  usage: python tools/ptxas_regs.py file.cu [pattern]"""
import re, subprocess, sys
src = sys.argv[1]; pat = sys.argv[2] if len(sys.argv) > 2 else ""
r = subprocess.run(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
                    "-Xptxas", "-v", "-c", src, "-o", "/tmp/_ptxas_regs.o"], capture_output=True, text=True)
name = None; spill = ""
for line in r.stderr.splitlines():
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().replace("void (anonymous namespace)::", "").split("(CUtensor")[0][:88]
    if "spill" in line: spill = line.strip()
    m = re.search(r"Used (\d+) registers", line)
    if m and name and pat in name:
        print(f"{name:90s} regs={m.group(1):>4s}  {spill}")
if r.returncode: print(r.stderr[-3000:])

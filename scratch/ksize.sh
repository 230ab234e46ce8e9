#!/bin/bash
# kernel SASS instruction counts + registers/stack from the last build
cd /root/repo/celeritas_b200
cuobjdump -sass build/csrc_kernels.o 2>/dev/null > /tmp/kern_now.sass
grep -n "Function :" /tmp/kern_now.sass | awk -F: '{print $1, $3}' | awk 'NR>1{print int(($1-prev)/2), name} {prev=$1; name=$2}' | sort -rn | head -${1:-8}
python3 - <<'PY'
import re
t=open('/root/repo/celeritas_b200/build/csrc_kernels.ptxas.log').read()
for m in re.finditer(r"Compiling entry function '(\w+)'.*?\n(.*?)\n.*?Used (\d+) registers.*?(?:, (\d+) bytes cumulative stack size)?\n", t):
    name=m.group(1)
    if any(k in name for k in ('along_step','interact','pre_step','boundary','discrete','initialize_tracks')):
        print(name[8:40], 'regs', m.group(3), '|', m.group(2).strip())
PY

"""Static SASS instruction count per source region for one kernel of the built library.
usage: static_breakdown.py <kernel-substring> [bucket]"""
import collections, os, re, subprocess, sys
kern = sys.argv[1]; bucket = int(sys.argv[2]) if len(sys.argv) > 2 else 25
os.chdir('/tmp')
subprocess.run(['cuobjdump', '-xelf', 'all', '/root/repo/celeritas_b200/build/csrc_kernels.o'], capture_output=True)
txt = subprocess.run(['nvdisasm', '-g', '-c', '/tmp/kernels.sm_100a.cubin'], capture_output=True, text=True).stdout
cur_fn = None; cur = None
per = collections.defaultdict(collections.Counter)
for line in txt.splitlines():
    m = re.match(r'\s*\.section\s+\.text\.(\S+?),', line)
    if m: cur_fn = m.group(1); continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', line) and cur_fn and cur:
        per[cur_fn][cur] += 1
for fn, c in per.items():
    if kern in fn:
        tot = sum(c.values())
        byfile = collections.Counter()
        for (f, l), n in c.items(): byfile[f] += n
        print(fn[:70], tot, byfile.most_common(9))
        b = collections.Counter()
        for (f, l), n in c.items(): b[(f, l // bucket * bucket)] += n
        for k, v in b.most_common(45): print('   ', k, v)

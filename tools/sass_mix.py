"""Static instruction mix of the largest backward-branch loop of a kernel in an object file.
usage: sass_mix.py <obj> <kernel-substring>"""
import re, sys, subprocess, collections
obj, key = sys.argv[1], sys.argv[2]
txt = subprocess.run(['cuobjdump', '-sass', obj], capture_output=True, text=True).stdout
ins, on = [], False
for l in txt.splitlines():
    if 'Function :' in l:
        on = key in l
        continue
    if not on: continue
    m = re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(.*?)\s*;', l)
    if m: ins.append((int(m.group(1), 16), m.group(2)))
print('total static instrs', len(ins))
loops = []
for a, t in ins:
    m = re.search(r'BRA(?:\.\w+)*\s+(?:\w+,\s*)?0x([0-9a-f]+)', t)
    if m:
        tgt = int(m.group(1), 16)
        if tgt < a: loops.append((a - tgt, tgt, a))
loops.sort(reverse=True)
for span, lo, hi in loops[:6]: print('loop', hex(lo), hex(hi), span // 16, 'instrs')
def mix(lo, hi):
    c = collections.Counter()
    for a, t in ins:
        if lo <= a <= hi:
            c[re.sub(r'^@!?U?P\w+\s+', '', t).split()[0].split('.')[0]] += 1
    return c
if len(sys.argv) > 4:
    lo, hi = int(sys.argv[3], 16), int(sys.argv[4], 16)
    c = mix(lo, hi)
    print('range', hex(lo), hex(hi), 'instrs', sum(c.values()), 'FP64', sum(c[k] for k in ('DFMA', 'DMUL', 'DADD', 'DSETP')))
    print(c.most_common(30))

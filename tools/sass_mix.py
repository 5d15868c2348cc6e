"""instruction mix per kernel from `cuobjdump -sass` output: python tools/sass_mix.py file.sass [name-filter]"""
import re, collections, sys
txt = open(sys.argv[1]).read()
flt = sys.argv[2] if len(sys.argv) > 2 else ""
for f in re.split(r'Function : ', txt)[1:]:
    name = f.split('\n')[0]
    if flt not in name:
        continue
    ops = collections.Counter()
    for line in f.split('\n'):
        m = re.search(r'^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
        if m:
            ops[m.group(1)] += 1
    print(name, sum(ops.values()))
    print('   ', ops.most_common(16))

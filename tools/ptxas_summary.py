import sys,re
name=None
for l in sys.stdin:
    m=re.search(r"Compiling entry function '(.*?)\(", l)
    if m: name=m.group(1)
    if 'error' in l or 'warning' in l: print(l.strip())
    m2=re.search(r'(\d+) bytes stack frame, (\d+) bytes spill stores', l)
    if m2 and name and int(m2.group(2))>0: print('SPILL', m2.group(2), name)
    m=re.search(r'Used (\d+) registers', l)
    if m: print(m.group(1), name)

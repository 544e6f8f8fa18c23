import re, sys
txt = open('fdm_b200/csrc/%s.ptxas.log' % (sys.argv[1] if len(sys.argv) > 1 else 'lapl_cube')).read()
pat = sys.argv[2] if len(sys.argv) > 2 else 'pipe'
for m in re.finditer(r"Compiling entry function '(\S+)'.*?\n(?:.*\n)*?.*?(\d+) bytes spill stores.*\n.*Used (\d+) registers", txt):
    name = m.group(1)
    if re.search(pat, name):
        print(name[10:80], 'spill', m.group(2), 'regs', m.group(3))

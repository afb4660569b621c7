import re, sys
txt = open(sys.argv[1]).read()
for m in re.finditer(r"^E\s+(AssertionError.*|assert .*|RuntimeError.*|.*Error.*)$", txt, re.M):
    print(m.group(0)[:300])
print(txt.strip().splitlines()[-1])

"""Tabulate an A/B log of tools/ab_opts.py runs over several libraries ('^^ lib' closes a block): best kernel ms per scene and library,
and the change against the first library.  usage: python tools/ab_table.py <log>"""
import collections, re, sys
res = collections.OrderedDict(); cur = []
for l in open(sys.argv[1]):
    if l.startswith('^^'):
        res.setdefault(l.split()[1], []).append(cur); cur = []
    elif 'Msamples/s' in l:
        m = re.search(r'^(\w+).*?([\d.]+) ms\s+([\d.]+) Msamples', l); cur.append((m.group(1), float(m.group(2))))
scenes = [s for s, _ in list(res.values())[0][0]]
print("lib".ljust(24) + "".join(s.rjust(13) for s in scenes))
base = None
for lib, passes in res.items():
    best = [min(p[i][1] for p in passes) for i in range(len(scenes))]
    base = base or best
    print(lib.ljust(24) + "".join(f"{b:7.3f} {100 * (b / base[i] - 1):+4.1f}%".rjust(13) for i, b in enumerate(best)))

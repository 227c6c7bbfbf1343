"""Aggregate the warp-stall samples of an ncu source page by CUDA source line.

    ncu -i rep.ncu-rep --page source --csv > src.csv
    cuobjdump -xelf all lib/<file>.o && nvdisasm --print-line-info <file>.sm_100a.cubin > dis.txt
    python tools/ncu_lines.py src.csv dis.txt <source.cu> [kernel-substring] [top]
"""
import collections
import csv
import re
import sys


def main():
    src_csv, dis_txt, cu = sys.argv[1:4]
    ksub = sys.argv[4] if len(sys.argv) > 4 else ""
    top = int(sys.argv[5]) if len(sys.argv) > 5 else 30
    func = None
    line = None
    maps = {}
    for l in open(dis_txt):
        m = re.match(r'\s*\.section\s+\.text\.(\S+?),', l)
        if m:
            func = m.group(1); maps[func] = {}; line = None; continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            line = (m.group(1).split('/')[-1], int(m.group(2))); continue
        m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
        if m and func:
            maps[func][int(m.group(1), 16)] = line
    rows = list(csv.reader(open(src_csv)))
    blocks = []
    cur = None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {'name': r[1], 'rows': []}; blocks.append(cur); continue
        if cur is not None:
            cur['rows'].append(r)
    src = open(cu).read().splitlines()
    cuname = cu.split('/')[-1]
    for b in blocks:
        # match the mangled function by template arguments
        targs = re.findall(r'\((?:int|bool)\)(\d+)', b['name'])
        cands = [f for f in maps if ksub in f]
        def score(f):
            return sum(1 for t in targs if re.search(r'(Li|Lb)' + t + 'E', f))
        fn = max(cands, key=score)
        mp = maps[fn]
        hdr = b['rows'][0]; data = b['rows'][1:]
        isamp = hdr.index('# Samples'); iaddr = hdr.index('Address')
        stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
        base = int(data[0][iaddr], 16)
        agg = collections.Counter(); reasons = collections.defaultdict(collections.Counter); tot = 0
        for r in data:
            off = int(r[iaddr], 16) - base; n = int(r[isamp] or 0); tot += n
            k = mp.get(off); agg[k] += n
            for i in stall_cols:
                reasons[k][hdr[i][6:]] += int(r[i] or 0)
        print(b['name'][:90], 'samples', tot)
        for k, n in agg.most_common(top):
            s = src[k[1] - 1].strip()[:86] if k and k[0] == cuname else str(k)
            rs = ' '.join(f"{a}:{c}" for a, c in reasons[k].most_common(2))
            print(f"{n:6d} {100 * n / max(tot, 1):5.1f}% {k[1] if k else -1:4d} {s:86s} {rs}")


if __name__ == "__main__":
    main()

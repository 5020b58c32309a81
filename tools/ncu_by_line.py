"""Attributes the executed warp instructions of an ncu capture (--set full --import-source on) to
source regions: joins `ncu --page source --print-source sass --csv` with `nvdisasm -g` line info.
usage: python tools/ncu_by_line.py <rep.ncu-rep> <kernel substring> <cells per launch>"""
import collections, csv, io, os, re, subprocess, sys, tempfile

rep, ksub, cells = sys.argv[1], sys.argv[2], float(sys.argv[3])
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(root, 'pyrddlgym_b200', 'csrc', 'librddl_b200.so')
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', so], cwd=tmp, check=True, capture_output=True)
# the ahead-of-time library plus every run-time specialised kernel in the cubin cache
cubins = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith('.cubin')]
cache = os.path.join(os.path.dirname(so), 'jit_cache')
if os.path.isdir(cache):
    cubins += [os.path.join(cache, f) for f in sorted(os.listdir(cache)) if f.endswith('.cubin')]
seq = collections.defaultdict(list)
for cubin in cubins:
    dis = subprocess.run(['nvdisasm', '-g', '-c', cubin], capture_output=True, text=True).stdout
    fn = None; line = None
    for l in dis.splitlines():
        m = re.match(r'\s*\.section\s+\.text\.(\S+),', l)
        if m: fn = m.group(1) + '@' + os.path.basename(cubin); line = None; continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m: line = (m.group(1).split('/')[-1], int(m.group(2))); continue
        m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
        if m and fn: seq[fn].append((line, m.group(2)))
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
# split per kernel
blocks = out.split('"Kernel Name",')
for blk in blocks[1:]:
    name = blk.splitlines()[0]
    if ksub not in name: continue
    rows = list(csv.reader(io.StringIO('\n'.join(blk.splitlines()[1:]))))
    hdr = rows[0]; ie = hdr.index('Instructions Executed'); ns = hdr.index('# Samples')
    data = [r for r in rows[1:] if len(r) > ie and r[ie].isdigit()]
    cand = [k for k, v in seq.items() if len(v) == len(data) and 'plane' in k] or [k for k, v in seq.items() if len(v) == len(data)]
    if not cand:
        print('no SASS match for', name[:80], len(data)); continue
    ins = seq[cand[0]]
    tot = sum(int(r[ie]) for r in data)
    by = collections.Counter(); bys = collections.Counter()
    for (ln, _), r in zip(ins, data):
        by[ln] += int(r[ie]); bys[ln] += int(r[ns])
    stot = sum(bys.values()) or 1
    print(name[:90], '| total warp inst', tot, '| thread-instr per cell %.1f' % (tot * 32 / cells))
    src = {}
    order = bys if os.environ.get('SORT') == 'stall' else by
    for ln, _c in order.most_common(int(os.environ.get('TOP', '45'))):
        c = by[ln]
        t = ''
        if ln:
            path = os.path.join(root, 'pyrddlgym_b200', 'csrc', ln[0])
            if os.path.exists(path):
                src.setdefault(ln[0], open(path).read().splitlines())
                t = src[ln[0]][ln[1] - 1].strip()[:70]
        print('%-22s %6.2f instr/cell %5.1f%% stall %5.1f%% | %s' % (ln, c * 32 / cells, 100 * c / tot, 100 * bys[ln] / stot, t))
    break

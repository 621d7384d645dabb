"""Attribute ncu per-instruction samples to CUDA source lines (CPU box, no GPU).

    python tools/ncu_lines.py gpurun_out/prof_x.ncu-rep <kernel-substring> [cubin-stem]

Uses `nvdisasm -g` line info of the in-tree library (compile with -lineinfo).
"""
import collections
import csv
import glob
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def line_map(kernel_sub):
    tmp = tempfile.mkdtemp()
    subprocess.run(['cuobjdump', '-xelf', 'all', os.path.join(ROOT, 'glenet_b200/lib/libglenet_geom.so')], cwd=tmp, stdout=subprocess.DEVNULL)
    best = None
    for cub in glob.glob(os.path.join(tmp, '*.cubin')):
        dis = subprocess.run(['nvdisasm', '-g', '-c', cub], stdout=subprocess.PIPE, text=True).stdout
        cur_fn, cur_line, seq = None, None, []
        fns = {}
        for ln in dis.splitlines():
            m = re.match(r'\s*\.text\.(\S+):', ln)
            if m:
                cur_fn = m.group(1); fns[cur_fn] = []; continue
            m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
            if m:
                cur_line = (os.path.basename(m.group(1)), int(m.group(2)), 'inlined' in m.group(3)); continue
            m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
            if m and cur_fn:
                fns[cur_fn].append((int(m.group(1), 16), cur_line, m.group(2)))
        for fn, ins in fns.items():
            if kernel_sub in fn and ins:
                best = ins
    return best


def main():
    rep, ksub = sys.argv[1], sys.argv[2]
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    # several kernels may be in the report: split on "Kernel Name" rows
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == 'Kernel Name':
            cur = {'name': r[1], 'rows': []}; blocks.append(cur)
        elif cur is not None:
            cur['rows'].append(r)
    for b in blocks:
        if ksub not in b['name'].replace('glenet::', ''):
            continue
        hdr = b['rows'][0]
        si, ii, ti = hdr.index('# Samples'), hdr.index('Instructions Executed'), hdr.index('Thread Instructions Executed')
        data = b['rows'][1:]
        mangled = sys.argv[3] if len(sys.argv) > 3 else ksub.split('<')[0]
        lm = line_map(mangled)
        print(f"kernel {b['name'][:100]}: {len(data)} sass rows, {len(lm) if lm else 0} disasm rows")
        agg = collections.defaultdict(lambda: [0, 0, 0])
        for k, r in enumerate(data):
            key = lm[k][1] if lm and k < len(lm) else ('?', 0, False)
            a = agg[key]
            a[0] += int(r[si]); a[1] += int(r[ii]); a[2] += int(r[ti])
        ts, tin = sum(a[0] for a in agg.values()), sum(a[1] for a in agg.values())
        print(f"total samples {ts}, warp instr {tin}")
        for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:40]:
            print(f"  {key[0]}:{key[1]:<5} samples {a[0]:6d} ({100 * a[0] / max(ts, 1):5.1f}%)  winstr {a[1]:10d} ({100 * a[1] / max(tin, 1):5.1f}%)  thr/instr {a[2] / max(a[1], 1):5.1f}")
        break


if __name__ == '__main__':
    main()

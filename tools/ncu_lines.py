"""Per-source-line instruction counts and stall samples of one kernel from an ncu report (CPU box, no GPU).

    python tools/ncu_lines.py gpurun_out/prof_x.ncu-rep <kernel-substring> [top-n]

Uses ncu's own SASS<->CUDA correlation (`--page source --print-source cuda,sass`; the report must have been captured
with `--import-source on` from a library compiled with -lineinfo)."""
import collections
import csv
import os
import subprocess
import sys


def main():
    rep, ksub = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 45
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    cur_file, cur_fn, hdr = None, None, None
    agg = collections.OrderedDict()
    for r in rows:
        if not r:
            continue
        if r[0] == 'File Path':
            cur_file = os.path.basename(r[1]); continue
        if r[0] == 'Function Name':
            cur_fn = r[1]; continue
        if r[0] == 'Line No':
            hdr = r; continue
        if hdr is None or cur_fn is None or ksub not in cur_fn or not r[0].strip().isdigit():
            continue
        shift = len(r) - len(hdr)   # source text with quotes / commas is not escaped by ncu: count columns from the right
        si, ii, ti = (hdr.index(c) + shift for c in ('# Samples', 'Instructions Executed', 'Thread Instructions Executed'))
        key = (cur_file, int(r[0]))
        a = agg.setdefault(key, [0, 0, 0, r[1].strip()[:90]])
        try:
            a[0] += int(r[si]); a[1] += int(r[ii]); a[2] += int(r[ti])
        except ValueError:
            pass
    ts, tin = sum(a[0] for a in agg.values()), sum(a[1] for a in agg.values())
    print(f"kernel *{ksub}*: {len(agg)} source lines, {ts} stall samples, {tin} warp instructions")
    byfile = collections.Counter()
    for (f, _), a in agg.items():
        byfile[f] += a[1]
    print("warp instructions per file:", ", ".join(f"{f} {n} ({100 * n / max(tin, 1):.1f}%)" for f, n in byfile.most_common()))
    print(f"-- top {top} lines by warp instructions")
    for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"  {f}:{ln:<4} winstr {a[1]:10d} ({100 * a[1] / max(tin, 1):5.1f}%)  thr/instr {a[2] / max(a[1], 1):5.1f}  samples {a[0]:6d} ({100 * a[0] / max(ts, 1):5.1f}%)  | {a[3]}")
    print(f"-- top {top // 2} lines by stall samples")
    for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top // 2]:
        print(f"  {f}:{ln:<4} samples {a[0]:6d} ({100 * a[0] / max(ts, 1):5.1f}%)  winstr {a[1]:10d}  thr/instr {a[2] / max(a[1], 1):5.1f}  | {a[3]}")


if __name__ == '__main__':
    main()

"""Per-SASS-instruction memory cost of a kernel from an .ncu-rep (source page): shared-memory wavefronts vs the ideal count,
global sectors vs ideal, and the unit-level totals -- finds scalarised or bank-conflicting accesses that the stall samples
do not point at.    python tools/ncu_hotmem.py <report.ncu-rep> [kernel-substring] [top]"""
import csv, subprocess, sys

def main():
    path = sys.argv[1]; want = sys.argv[2] if len(sys.argv) > 2 else ""; top = int(sys.argv[3]) if len(sys.argv) > 3 else 14
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(raw.splitlines())); hdr = rows[0]
    names = [r[hdr.index("Kernel Name")] for r in rows[2:]]
    ids = [r[hdr.index("ID")] for r in rows[2:]]
    for kid, name, r in zip(ids, names, rows[2:]):
        if want not in name: continue
        g = lambda k: r[hdr.index(k)] if k in hdr else "?"
        print(f"== [{kid}] {name[:80]}  {g('gpu__time_duration.sum')} us  inst {g('smsp__inst_executed.sum')}  issue {g('smsp__issue_active.avg.pct_of_peak_sustained_active')}%  "
              f"l1tex {g('l1tex__throughput.avg.pct_of_peak_sustained_active')}%  lsu-wavefronts {g('l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed')}%  "
              f"shared-conflicts {g('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum')}  alu {g('sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active')}%  fma {g('sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active')}%  xu {g('sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active')}%")
    # source page: one block per profiled launch ("Kernel Name" row, header row, one row per SASS instruction)
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1] if len(r) > 1 else "", "hdr": None, "rows": []}; blocks.append(cur)
        elif cur is not None and r and r[0] == "Address":
            cur["hdr"] = r
        elif cur is not None and cur["hdr"] is not None:
            cur["rows"].append(r)
    for kid, blk in enumerate(blocks):
        if want not in blk["name"] or not blk["hdr"]: continue
        hd = blk["hdr"]; ix = {k: j for j, k in enumerate(hd)}
        def f(r, k):
            try: return float(r[ix[k]] or 0)
            except (KeyError, IndexError, ValueError): return 0.0
        sh, gl = [], []
        tot_sh = tot_ideal = 0
        for r in blk["rows"]:
            if len(r) < len(hd): continue
            w, ideal = f(r, "L1 Wavefronts Shared"), f(r, "L1 Wavefronts Shared Ideal")
            if w > 0: sh.append((w, ideal, f(r, "Instructions Executed"), f(r, "Avg. Predicated-On Threads Executed"), r[ix["Source"]].strip()[:70])); tot_sh += w; tot_ideal += ideal
            s_, si = f(r, "L2 Theoretical Sectors Global"), f(r, "L2 Theoretical Sectors Global Ideal")
            if s_ > 0: gl.append((s_, si, f(r, "Instructions Executed"), f(r, "Avg. Predicated-On Threads Executed"), r[ix["Source"]].strip()[:70]))
        print(f"-- [{kid}] {blk['name'][:60]}: shared wavefronts {tot_sh:.0f} (ideal {tot_ideal:.0f})")
        for x in sorted(sh, reverse=True)[:top]: print("   sh  %10.0f ideal %10.0f  inst %9.0f thr %4.1f  %s" % x)
        for x in sorted(gl, reverse=True)[:max(4, top // 2)]: print("   gl  %10.0f ideal %10.0f  inst %9.0f thr %4.1f  %s" % x)

if __name__ == "__main__":
    main()

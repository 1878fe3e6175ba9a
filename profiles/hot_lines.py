"""Per-source-line totals of a kernel from an ncu report: joins the SASS source page (stall samples,
instructions executed) with nvdisasm's line table of the in-tree library.
usage: python profiles/hot_lines.py <report.ncu-rep> <kernel regex> [top N]
(the regex selects the function in the library's line table by its MANGLED name, e.g. k_observeILi128ELi7ELi0; NCU_KRE, default the
same regex, selects the launch in the report by its base name, e.g. k_observe)"""
import csv
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "flatland-marl_b200", "csrc", "libflatland_b200.so")


def line_table(kernel_re):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, capture_output=True)
    cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
    table, cur, infn, line = {}, None, False, None
    for ln in txt.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
        if m:
            infn = re.search(kernel_re, m.group(1)) is not None
            cur = m.group(1)
            continue
        if not infn:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            line = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*?);", ln)
        if m and line:
            table[int(m.group(1), 16)] = line
    return table


def main():
    rep, kre = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + os.environ.get("NCU_KRE", kre)],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    ia, isamp, iex, ithr = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
    table = line_table(kre)
    base = None
    agg = {}
    tot_s = tot_e = 0
    for r in rows[hi + 1:]:
        if len(r) <= ithr or not r[ia].startswith("0x"):
            if r and r[0] == "Kernel Name":
                break   # only the first captured launch
            continue
        addr = int(r[ia], 16)
        if base is None:
            base = addr
        key = table.get(addr - base, ("?", 0))
        s, ex, th = int(r[isamp]), int(r[iex]), int(r[ithr])
        a = agg.setdefault(key, [0, 0, 0])
        a[0] += s; a[1] += ex; a[2] += th
        tot_s += s; tot_e += ex
    print("# %s  kernel /%s/: %d stall samples, %d warp instructions" % (os.path.basename(rep), kre, tot_s, tot_e))
    print("%-24s %8s %7s %12s %7s %6s" % ("file:line", "samples", "%", "warp-inst", "%", "thr/in"))
    for key, (s, ex, th) in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
        print("%-24s %8d %6.2f%% %12d %6.2f%% %6.1f" % ("%s:%d" % key, s, 100.0 * s / max(tot_s, 1), ex, 100.0 * ex / max(tot_e, 1), th / max(ex, 1)))
    print("\n# by warp instructions")
    for key, (s, ex, th) in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
        print("%-24s %8d %6.2f%% %12d %6.2f%% %6.1f" % ("%s:%d" % key, s, 100.0 * s / max(tot_s, 1), ex, 100.0 * ex / max(tot_e, 1), th / max(ex, 1)))
    print("\n# in source order (lines with at least 0.1 %% of the samples or instructions)")
    for key, (s, ex, th) in sorted(agg.items()):
        if s >= tot_s / 1000 or ex >= tot_e / 1000:
            print("%-24s %8d %6.2f%% %12d %6.2f%% %6.1f" % ("%s:%d" % key, s, 100.0 * s / max(tot_s, 1), ex, 100.0 * ex / max(tot_e, 1), th / max(ex, 1)))


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""Per-source-line instruction counts and stall samples of one kernel from an ncu report.

The ncu CLI prints metrics per SASS instruction only (`--page source --csv`); this joins that table with
`nvdisasm -g` line info of the same cubin (same instruction order) and sums per source line.

    python tools/ncu_lines.py gpurun_out/v6_full.ncu-rep blend_fwd_kernelILi1ELb0 [--so gs-evt_b200/libgsevt.so] [--top 40]
"""
import argparse
import collections
import csv
import glob
import io
import os
import re
import subprocess
import tempfile


def sass_lines(so, mangled_sub):
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=td, check=True, stdout=subprocess.DEVNULL)
        for cubin in glob.glob(os.path.join(td, "*.cubin")):
            txt = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout
            m = re.search(r"\.section\s+\.text\.(\S*%s\S*?),\"ax\"" % re.escape(mangled_sub), txt)
            if not m:
                continue
            name = m.group(1)
            body = txt[m.end():]
            nxt = body.find("\t.section")
            body = body[:nxt] if nxt >= 0 else body
            out, cur = [], (None, 0)
            for ln in body.splitlines():
                f = re.search(r'//## File "([^"]+)", line (\d+)', ln)
                if f:
                    cur = (os.path.basename(f.group(1)), int(f.group(2)))
                    continue
                i = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
                if i:
                    out.append((int(i.group(1), 16), i.group(2).strip(), cur))
            return name, out
    raise SystemExit(f"kernel matching {mangled_sub!r} not found in {so}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("kernel", help="substring of the mangled kernel name")
    ap.add_argument("--regex", default=None, help="ncu --kernel-name regex (default: derived from the substring)")
    ap.add_argument("--so", default=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gs-evt_b200", "libgsevt.so"))
    ap.add_argument("--top", type=int, default=45)
    ap.add_argument("--source-root", default=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gs-evt_b200", "csrc"))
    a = ap.parse_args()
    name, sass = sass_lines(a.so, a.kernel)
    rx = a.regex or re.match(r"_ZN\d+\w+?\d+([a-z_0-9]+?)(I|E)", name).group(1)
    csvtxt = subprocess.run(["ncu", "-i", a.report, "--page", "source", "--csv", "--kernel-name", f"regex:{rx}", "--launch-count", "1"],
                            capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(csvtxt)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    print("kernel:", rows[hi - 1][1] if hi else name)
    ci, cs, csrc = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
    body = []
    for r in rows[hi + 1:]:
        if r and r[0] == "Kernel Name":   # the report holds several launches of this kernel: keep the first
            break
        if len(r) > ci:
            body.append(r)
    if len(body) != len(sass):
        print(f"warning: {len(body)} profiled instructions vs {len(sass)} disassembled (different build?)")
    per = collections.defaultdict(lambda: [0, 0, 0])
    tot_i = tot_s = 0
    for r, (_, op, loc) in zip(body, sass):
        n, s = int(r[ci] or 0), int(r[cs] or 0)
        per[loc][0] += n; per[loc][1] += s; per[loc][2] += 1
        tot_i += n; tot_s += s
    print(f"total warp instructions {tot_i}, stall samples {tot_s}, SASS instructions {len(body)}")
    src_cache = {}

    def src(loc):
        f, l = loc
        if f is None:
            return ""
        p = os.path.join(a.source_root, f)
        if p not in src_cache:
            src_cache[p] = open(p).read().splitlines() if os.path.exists(p) else []
        L = src_cache[p]
        return L[l - 1].strip()[:100] if 0 < l <= len(L) else ""

    print(f"{'file:line':<22}{'inst%':>7}{'stall%':>8}{'#sass':>6}  source")
    for loc, (n, s, k) in sorted(per.items(), key=lambda kv: -kv[1][0])[:a.top]:
        print(f"{str(loc[0]) + ':' + str(loc[1]):<22}{100 * n / max(tot_i, 1):7.2f}{100 * s / max(tot_s, 1):8.2f}{k:6d}  {src(loc)}")


if __name__ == "__main__":
    main()

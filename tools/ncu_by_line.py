"""Aggregate an ncu report's SASS-level samples / instruction counts by CUDA source line.

  python tools/ncu_by_line.py <report.ncu-rep> <kernel regex> <object .o or .so> [top N]

ncu's CSV export of the source page is SASS-only; this joins it (by instruction order) with the line table that
`nvdisasm -g` prints for the same function, so that stall samples and executed instructions can be read per source
line here in the build container (no GUI).
"""
import csv
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict


def main():
    import os
    rep, pat, obj = sys.argv[1], sys.argv[2], os.path.abspath(sys.argv[3])
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + pat] + (["--launch-skip", os.environ["NCU_SKIP"]] if os.environ.get("NCU_SKIP") else []), capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    kname = rows[0][1]
    hdr = rows[1]
    ia, ism, isrc = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
    body = rows[2:]
    for i, r in enumerate(body):  # several launches matched: the report repeats the table per launch — take the first
        if r and r[0] == "Kernel Name":
            body = body[:i]
            break
    sass = [(r[isrc].strip(), int(r[ia]), int(r[ism])) for r in body if len(r) > ia and r[0].startswith("0x")]
    # line table
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", "all", obj], cwd=td, capture_output=True)
        import glob
        lines = None
        mangled = None
        for cubin in glob.glob(td + "/*.cubin"):
            txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
            # split per function
            cur, cur_line, funcs = None, None, {}
            for ln in txt.splitlines():
                m = re.match(r"\s*\.text\.(\S+):", ln)
                if m:
                    cur = m.group(1)
                    funcs[cur] = []
                    continue
                m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
                if m:
                    cur_line = (m.group(1).split("/")[-1], int(m.group(2)))
                    continue
                if cur and re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
                    funcs[cur].append(cur_line)
            for f, ls in funcs.items():
                dem = subprocess.run(["cu++filt", f], capture_output=True, text=True).stdout.strip()
                if re.search(pat, dem) and len(ls) == len(sass):
                    lines, mangled = ls, dem
        if lines is None:
            print("no function with", len(sass), "instructions matched", pat)
            return
    agg = defaultdict(lambda: [0, 0])
    for (src, n, sm), ln in zip(sass, lines):
        agg[ln][0] += n
        agg[ln][1] += sm
    tot_i = sum(v[0] for v in agg.values())
    tot_s = sum(v[1] for v in agg.values())
    print(f"# {kname}\n# {tot_i} warp instructions, {tot_s} samples")
    print("| file:line | instr % | samples % |")
    print("|---|---:|---:|")
    for ln, (n, sm) in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
        print(f"| {ln[0]}:{ln[1]} | {100 * n / tot_i:.1f} | {100 * sm / tot_s:.1f} |")


if __name__ == "__main__":
    main()

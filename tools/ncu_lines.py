"""Executed warp instructions and stall samples per SOURCE line of one kernel of an .ncu-rep: joins the SASS page of the
report with `nvdisasm -g` line info of the cubin inside libeffimvs.so (no GPU needed).

    python tools/ncu_lines.py gpurun_out/x.ncu-rep <kernel name substring> [cubin stem, default warp_tile] [top N]
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    rep, pat = sys.argv[1], sys.argv[2]
    stem = sys.argv[3] if len(sys.argv) > 3 else "warp_tile"
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "effi-mvs-plus_b200", "libeffimvs.so")], cwd=tmp, capture_output=True)
    cubin = os.path.join(tmp, stem + ".sm_100a.cubin")
    txt = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout.splitlines()
    start = next(i for i, ln in enumerate(txt) if ".section" in ln and ".text." in ln and pat in ln)
    ins, cur = [], None
    for ln in txt[start + 1:]:
        if ".section" in ln:
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+\S", ln):
            ins.append(cur)
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO("\n".join(ln for ln in out.splitlines() if not ln.startswith("==")))))
    blocks, cb = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cb = {"name": r[1] if len(r) > 1 else "", "rows": []}
            blocks.append(cb)
        elif cb is not None:
            cb["rows"].append(r)
    blk = next(b for b in blocks if pat.split("IL")[0] in b["name"].replace(" ", ""))
    h = blk["rows"][0]
    ia, ismp = h.index("Instructions Executed"), h.index("# Samples")
    body = [r for r in blk["rows"][1:] if len(r) > ia and r[ia].isdigit()]
    assert len(body) == len(ins), (len(body), len(ins), "the library on disk is not the one that was profiled")
    exe, smp = collections.Counter(), collections.Counter()
    for r, k in zip(body, ins):
        exe[k] += int(r[ia])
        smp[k] += int(r[ismp])
    tot, stot = sum(exe.values()), sum(smp.values())
    src = {}
    print("executed {} warp instructions, {} samples".format(tot, stot))
    for k, c in exe.most_common(top):
        f, ln = k if k else ("?", 0)
        if f not in src:
            p = os.path.join(ROOT, "effi-mvs-plus_b200", "csrc", f)
            src[f] = open(p).read().splitlines() if os.path.exists(p) else []
        text = src[f][ln - 1].strip()[:100] if 0 < ln <= len(src[f]) else ""
        print("%9d %5.1f%%  stall %5.1f%%  %s:%d  %s" % (c, 100.0 * c / tot, 100.0 * smp[k] / max(stot, 1), f, ln, text))


if __name__ == "__main__":
    main()
